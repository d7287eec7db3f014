"""Longer run of tests/test_host_emul_stability.py's drivers: the HOST build of the device stability routine against the oracle, both
semantics, both domains, many trajectories.   python scratch/host_emul_soak.py [envs per configuration]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import test_host_emul_stability as T  # noqa: E402
from harness import make_stream  # noqa: E402
from pct_oracle import OracleContinuous, OracleDiscrete, make_continuous_stream  # noqa: E402

L = T.lib.__wrapped__() if hasattr(T.lib, "__wrapped__") else None
if L is None:  # build / load like the fixture does
    import ctypes as C
    import subprocess
    inc = T._cuda_include()
    os.makedirs(os.path.dirname(T.OUT), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-w", "-I" + T.CSRC,
                           "-I" + os.path.join(ROOT, "include"), "-I" + inc, "-o", T.OUT, T.SRC])
    L = C.CDLL(T.OUT)
    dp = C.POINTER(C.c_double)
    L.sh_create.restype = C.c_void_p; L.sh_create.argtypes = [C.c_int] * 4
    for f in ("sh_destroy", "sh_reset", "sh_flags"):
        getattr(L, f).argtypes = [C.c_void_p]
    L.sh_virtual.argtypes = [C.c_void_p] + [C.c_int] * 5 + [C.c_double, C.POINTER(C.c_int)]
    L.sh_place.argtypes = [C.c_void_p] + [C.c_int] * 5 + [C.c_double]
    L.sh_set_alias.argtypes = [C.c_void_p, C.c_int]; L.sh_set_holder.argtypes = [C.c_void_p, C.c_int]
    L.shc_create.restype = C.c_void_p; L.shc_create.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double]
    for f in ("shc_destroy", "shc_reset", "shc_flags"):
        getattr(L, f).argtypes = [C.c_void_p]
    L.shc_set_alias.argtypes = [C.c_void_p, C.c_int]
    L.shc_virtual.argtypes = [C.c_void_p, dp, C.c_double]
    L.shc_place_row.argtypes = [C.c_void_p, dp, dp, C.c_double]

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
t0, checks = time.time(), 0
for alias in (False, True):
    for setting in (1, 3):
        for e in range(n):
            env = OracleDiscrete(setting, stream=make_stream(1234, e, 700, setting)); env.set_alias_mode(alias)
            nv, nr = T._drive(L, env, setting, (10, 10, 10), 80, 50, 4321, e, 300, alias=alias)
            checks += nv + nr
            envc = OracleContinuous(setting, stream=make_continuous_stream(1234, e, 700, setting)); envc.set_alias_mode(alias)
            checks += T._drive_c(L, envc, setting, (1.0, 1.0, 1.0), 4321, e, 250, alias=alias)
        print("alias=%s setting %d: %d envs per domain done, %d verdicts so far, %.0f s" % (alias, setting, n, checks, time.time() - t0), flush=True)
print("all verdicts agree:", checks)

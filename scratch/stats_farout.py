"""How many stability walks does the integer quick reject (rest_height_supports far_out) remove?  Host build, BASELINE streams."""
import ctypes as C, os, sys
sys.path[:0] = [os.path.join(os.path.dirname(__file__), "..", "oracle"), os.path.join(os.path.dirname(__file__), "..", "tests")]
import test_host_emul_stability as T
from harness import make_stream
from pct_oracle import OracleDiscrete
L = C.CDLL(T.OUT)
T.lib.__wrapped__ if hasattr(T.lib, "__wrapped__") else None
L.sh_create.restype = C.c_void_p; L.sh_create.argtypes = [C.c_int] * 4
for f in ("sh_destroy", "sh_reset", "sh_flags"): getattr(L, f).argtypes = [C.c_void_p]
L.sh_virtual.argtypes = [C.c_void_p] + [C.c_int] * 5 + [C.c_double, C.POINTER(C.c_int)]
L.sh_place.argtypes = [C.c_void_p] + [C.c_int] * 5 + [C.c_double]
L.sh_set_alias.argtypes = [C.c_void_p, C.c_int]; L.sh_set_holder.argtypes = [C.c_void_p, C.c_int]; L.sh_use_v2.argtypes = [C.c_int]
L.sh_use_v2(1)
for setting in (1, 3):
    nv = 0
    for env_id in range(8):
        env = OracleDiscrete(setting, stream=make_stream(1234, env_id, 700, setting))
        a, b = T._drive(L, env, setting, (10, 10, 10), 80, 50, 4321, env_id, 300)
        nv += a
    st = (C.c_longlong * 2)(); L.sh_stats(st)
    print("setting", setting, "virtual checks", nv, "far_out rejects", st[0], "walks", st[1])
    v = (C.c_longlong * 32)(); L.sh_stats_visits(v); v = list(v)
    print("  light visits k=0: %d k=1: %d | heavy: root %d placed %d | by k: %s | lstsq %d | heavy fail root/placed %d/%d | light fail root/placed %d/%d" % (
        v[0], v[1], v[2], v[3], {k: v[4 + k] for k in range(2, 9)}, v[14], v[15], v[16], v[17], v[18]))

"""Would forking the subtrees of a stability walk shorten the feasibility stage's tail?  Host build of the device sources, BASELINE streams: per step the most
expensive continuation walk over E envs, as total cost (what one lane runs today) and as the cost of its longest root-to-leaf path (what it would be if
every support subtree below a >= 2-support node ran on its own lane).  Cost units: one single-support visit = 1, placed >= 2-support visit 3.5, root 6, lstsq +10."""
import ctypes as C, os, sys
import numpy as np
sys.path[:0] = [os.path.join(os.path.dirname(__file__), "..", "oracle"), os.path.join(os.path.dirname(__file__), "..", "tests")]
import test_host_emul_stability as T
from harness import make_stream, policy_pick
from pct_oracle import OracleDiscrete
T._build() if hasattr(T, "_build") else None
L = C.CDLL(T.OUT)
L.sh_create.restype = C.c_void_p; L.sh_create.argtypes = [C.c_int] * 4
for f in ("sh_destroy", "sh_reset", "sh_flags"): getattr(L, f).argtypes = [C.c_void_p]
L.sh_virtual.argtypes = [C.c_void_p] + [C.c_int] * 5 + [C.c_double, C.POINTER(C.c_int)]
L.sh_place.argtypes = [C.c_void_p] + [C.c_int] * 5 + [C.c_double]
L.sh_set_alias.argtypes = [C.c_void_p, C.c_int]; L.sh_set_holder.argtypes = [C.c_void_p, C.c_int]; L.sh_use_v2.argtypes = [C.c_int]
L.sh_use_v2(int(sys.argv[3]) if len(sys.argv) > 3 else 3)
E, STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 96, int(sys.argv[2]) if len(sys.argv) > 2 else 150
setting = 1
tot = np.zeros((E, STEPS)); path = np.zeros((E, STEPS)); allw = []
out = (C.c_double * 3)()
for env_id in range(E):
    env = OracleDiscrete(setting, stream=make_stream(1234, env_id, 700, setting))
    h = L.sh_create(setting, 10, 10, 10); L.sh_set_holder(h, 80); L.sh_set_alias(h, 1); L.sh_reset(h)
    o = env.reset()
    for t in range(STEPS):
        cand, feas = env.candidates()
        for p, f in zip(cand, feas):
            if f < 0:
                continue
            mh = C.c_int()
            got = L.sh_virtual(h, int(p[3] - p[0]), int(p[4] - p[1]), int(p[5] - p[2]), int(p[0]), int(p[1]), env.next_den, C.byref(mh))
            assert got == f
            L.sh_last_path(out)
            if out[0] > 0:
                allw.append((out[0], out[1], mh.value, env_id, t))
                if out[0] > tot[env_id, t]: tot[env_id, t] = out[0]
                if out[1] > path[env_id, t]: path[env_id, t] = out[1]
        _, row = policy_pick(o, 80, 50, 4321, env_id, t)
        (x, y, z), lx, ly = T._placement(row, env.next_box)
        den = env.next_den
        o, _, done, _ = env.step(row)
        L.sh_place(h, x, y, z, lx, ly, den)
        if done:
            o = env.reset(); L.sh_reset(h)
    L.sh_destroy(h)
a = np.array(allw)
print("continuation walks %d: mean total %.1f, mean longest path %.1f" % (len(a), a[:, 0].mean(), a[:, 1].mean()))
for q in (0.5, 0.9, 0.99, 0.999, 0.9999, 1.0):
    print("  quantile %.4f: total %.1f  path %.1f" % (q, np.quantile(a[:, 0], q), np.quantile(a[:, 1], q)))
mt, mp = tot.max(axis=0), path.max(axis=0)
print("per step, max over %d envs: total mean %.1f (max %.1f) | longest path mean %.1f (max %.1f) | ratio of means %.2f" % (E, mt.mean(), mt.max(), mp.mean(), mp.max(), mt.mean() / mp.mean()))

# which walks are the longest?  cost by resting height of the candidate (the continuation kernel deals walks by this key)
print("resting height: walks per env-step, mean cost, 99.9 % cost, max cost")
for mh in range(1, 10):
    sel = a[a[:, 2] == mh]
    if len(sel):
        print("  mh %d: %.3f walks/env-step  mean %.1f  q999 %.1f  max %.1f" % (mh, len(sel) / (E * STEPS), sel[:, 0].mean(), np.quantile(sel[:, 0], 0.999), sel[:, 0].max()))
for thr in (15, 20, 25):
    sel = a[a[:, 0] >= thr]
    print("  walks with cost >= %d: %.4f per env-step; resting heights %s" % (thr, len(sel) / (E * STEPS), np.bincount(sel[:, 2].astype(int), minlength=10).tolist()))
if len(sys.argv) > 3 and sys.argv[3] == "4":
    L.sh_pieces.restype = C.c_longlong
    print("fork-join: %d pieces for %d continuation walks (%.2f per walk)" % (L.sh_pieces(), len(a), L.sh_pieces() / len(a)))

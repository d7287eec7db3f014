import sys, time
sys.path.insert(0, '/root/repo/oracle')
import numpy as np
import ref_shim, pct_oracle
from pct_oracle import OracleContinuous, policy_pick, make_continuous_stream
D, Cm = ref_shim.load_reference()

def run(setting, seed, envid, steps):
    stream = make_continuous_stream(seed, envid, steps+200, setting)
    ref = Cm.PackingContinuous(setting=setting, container_size=[1,1,1], item_set=[(0.1,0.1,0.1)], internal_node_holder=80, leaf_node_holder=50,
                               shuffle=False, sample_from_distribution=False, sample_left_bound=0.1, sample_right_bound=0.5)
    ref.size_minimum = 0.1; ref.space.low_bound = 0.1
    ref.box_creator = ref_shim.make_stream_creator(Cm, [tuple(float(v) for v in (r if setting==3 else r[:3])) for r in stream])
    ref.test = True
    orc = OracleContinuous(setting, stream=stream)
    o1 = ref.reset(); o2 = orc.reset()
    for t in range(steps):
        if not np.array_equal(o1, o2):
            d = np.where(o1 != o2)[0]
            print('OBS MISMATCH setting', setting, 'seed', seed, 'step', t, d[:10], o1[d[:10]], o2[d[:10]]); return False, orc.n_lstsq
        k, a = policy_pick(o1, 80, 50, seed, envid, t)
        o1, r1, d1, i1 = ref.step(a); o2, r2, d2, i2 = orc.step(a)
        if not np.array_equal(o1, o2) or r1 != r2 or d1 != d2 or i1 != i2:
            dd = np.where(o1 != o2)[0]
            print('STEP MISMATCH', setting, seed, t, r1, r2, d1, d2, i1, i2, dd[:8], o1[dd[:8]], o2[dd[:8]]); return False, orc.n_lstsq
        if d1:
            o1 = ref.reset(); o2 = orc.reset()
    return True, orc.n_lstsq

if __name__ == '__main__':
    setting = int(sys.argv[1]); nseeds = int(sys.argv[2]); steps = int(sys.argv[3]); base = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    t0 = time.time(); ok = 0; nl = 0
    for s in range(base, base+nseeds):
        r, n = run(setting, 1000+s, s, steps); ok += r; nl += n
    print('continuous setting', setting, 'ok', ok, '/', nseeds, 'lstsq calls', nl, 'time', time.time()-t0)

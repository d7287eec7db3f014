"""CPU, build container only: live lock-step of the C oracle against the unmodified Python reference
(every observation including terminal ones, reward, done, info).  Skipped where /root/reference is absent."""
import numpy as np
import pytest

import ref_shim
from harness import ITEM_SET, make_stream, policy_pick
from pct_oracle import OracleDiscrete

pytestmark = [pytest.mark.reference, pytest.mark.skipif(not ref_shim.reference_available(), reason="reference not mounted")]


@pytest.mark.parametrize("setting", [1, 2, 3])
def test_lockstep_with_reference(setting):
    D, _ = ref_shim.load_reference()
    seed, env_id, steps = 900 + setting, 3, 260
    stream = make_stream(seed, env_id, steps + 64, setting)
    ref = D.PackingDiscrete(setting=setting, container_size=[10, 10, 10], item_set=ITEM_SET, internal_node_holder=80,
                            leaf_node_holder=50, shuffle=False, LNES="EMS")
    ref.box_creator = ref_shim.make_stream_creator(D, [tuple(r) if setting == 3 else tuple(int(v) for v in r[:3]) for r in stream])
    ref.test = True
    orc = OracleDiscrete(setting, stream=stream)
    o1, o2 = ref.reset(), orc.reset()
    for t in range(steps):
        assert np.array_equal(o1, o2), t
        _, row = policy_pick(o1, 80, 50, seed, env_id, t)
        o1, r1, d1, i1 = ref.step(row)
        o2, r2, d2, i2 = orc.step(row)
        assert np.array_equal(o1, o2), "observation after step %d (done=%s)" % (t, d1)
        assert (r1, d1) == (r2, d2) and i1 == i2
        if d1:
            o1, o2 = ref.reset(), orc.reset()

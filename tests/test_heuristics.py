"""Heuristic baselines: the CPU restatement (oracle/pct_oracle_heuristics.py) against records of the
reference's unmodified heuristic.py (tests/golden/heur_s*.npz), and the batched CUDA kernel against the restatement."""
import glob
import os

import numpy as np
import pytest

import pct_oracle_heuristics as OH
from harness import ITEM_SET
from pct_oracle import OracleDiscrete

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "heur_s*.npz")))
RECORDED = ("LSAH", "OnlineBPH", "BR", "MACS", "DBL", "HM")


def golden(path, name):
    g = np.load(path)
    off = np.concatenate([[0], np.cumsum(g["len_" + name])])
    return int(g["setting"]), g["data"], [g["flat_" + name][off[i]:off[i + 1]].tolist() for i in range(len(off) - 1)]


def dataset_stream(data, first, count):
    """LoadBoxCreator order (episode k plays trajectory k+1) as ONE item stream with a sentinel after every trajectory"""
    rows = []
    for k in range(first, first + count):
        t = data[(k + 1) % len(data)]
        t = t if t.shape[1] == 4 else np.concatenate([t, np.ones((len(t), 1))], 1)
        rows += [t, [[100, 100, 100, 1.0]]]
    return np.concatenate(rows)


def test_golden_present():
    assert len(GOLDEN) == 3


@pytest.mark.parametrize("name", RECORDED)
@pytest.mark.parametrize("path", GOLDEN)
def test_restated_heuristics_replay_reference(path, name):
    setting, data, packed = golden(path, name)
    L = data.shape[1] + 1
    env = OracleDiscrete(setting, stream=dataset_stream(data, 0, len(packed) + 1))
    env.set_trajectory_length(L)
    rec = OH.run_episodes(name, env, len(packed), item_set=ITEM_SET)
    assert [r[2] for r in rec] == packed

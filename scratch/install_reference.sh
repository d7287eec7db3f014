#!/bin/bash
# Copies the UNMODIFIED reference tree into baseline/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU box with the snapshot, it never enters
# the history).  Needed on the box only by (1) tests/test_gpu_rollout_policy.py — the reference's own DRL_GAT network inside the device-resident rollout —
# and (2) scratch/shmem_baseline.py — the reference's own ShmemVecEnv timed on the B200 host's cores.  Both skip without it.
set -e
cd "$(dirname "$0")/.."
[ -d /root/reference ] || { echo "no /root/reference here"; exit 1; }
rm -rf baseline/_ref && mkdir -p baseline/_ref
cp -r /root/reference/*.py /root/reference/pct_envs /root/reference/wrapper baseline/_ref/
find baseline/_ref -name __pycache__ -type d -exec rm -rf {} +
echo "installed: $(find baseline/_ref -name '*.py' | wc -l) python files, $(du -sh baseline/_ref | cut -f1)"

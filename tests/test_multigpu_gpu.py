"""Multi-GPU parity with the REAL kernels: a torchrun-spawned 2-rank NCCL job
(tests/workers/mgpu_worker.py).  Skipped on a box with fewer than two GPUs; `gpurun --gpus 2` runs it."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1500)]
ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.parametrize('world', [2])
def test_ddp_step_fit_checkpoint_and_sharded_gallery(world, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip(f'needs {world} GPUs, found {torch.cuda.device_count()}')
    port = 29600 + os.getpid() % 300
    env = dict(os.environ, MGPU_TMP=str(tmp_path))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}', '--master-addr', '127.0.0.1',
           '--master-port', str(port), str(ROOT / 'tests' / 'workers' / 'mgpu_worker.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=1400)
    print(r.stdout[-4000:], r.stderr[-4000:])
    assert r.returncode == 0 and 'MGPU CHECK OK' in r.stdout

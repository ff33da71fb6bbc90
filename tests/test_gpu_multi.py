"""GPU, >= 2 devices: sharded MatMult / Krylov / rdm against the oracle, one process per GPU
(skipped on single-GPU boxes; run with `gpurun --gpus 2`)."""
import ctypes
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_count():
    from dynamite_b200 import _capi
    n = ctypes.c_int(0)
    _capi.lib().dnm_device_count(ctypes.byref(n))
    return n.value


@pytest.mark.parametrize('nproc', [2, 4, 8])
def test_sharded_over_gpus(nproc):
    if _device_count() < nproc:
        pytest.skip(f'needs {nproc} GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={nproc}',
           '--master-addr', '127.0.0.1', '--master-port', str(29533 + nproc),
           os.path.join(ROOT, 'tests', 'multi_gpu_worker.py')]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]

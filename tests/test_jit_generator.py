"""CPU (no GPU): the planner + CUDA source generator + NVRTC, through dnm_jit_dryrun.  Every case plans
a tiled MatMult the way rank `rank` of `nranks` would, generates the operator-specialised pass kernels
(csrc/jit.cu) and compiles them to an sm_100a cubin -- shapes and rank counts the GPU tests cannot
afford (L = 30..33, 8 ranks) included.  What the kernels COMPUTE is checked on the GPU
(tests/test_gpu_matmult.py::test_generated_kernels_vs_oracle, tests/multi_gpu_worker.py)."""
import ctypes as C
import re

import numpy as np
import pytest

from dynamite_b200 import _capi, msc_tools
from dynamite_b200.hamiltonians import build_hamiltonian
from dynamite_b200.subspaces import Full, Parity


def dryrun(name, L, sub=None, nranks=1, rank=0, tile_bits=0, far_bits=-1, pipeline=0, tune=-1):
    """`name`: a benchmark model of hamiltonians.py, or an Operator"""
    H = build_hamiltonian(name, L) if isinstance(name, str) else name
    H.reduce_msc()
    masks, offs = msc_tools.mask_offsets(H.msc)
    masks, offs = _capi.as_i64(masks), _capi.as_i64(offs)
    signs = _capi.as_i64(H.msc['signs'])
    coeffs = _capi.as_c128(H.msc['coeffs'])
    sub = sub if sub is not None else Full(L=L)
    cdata = sub._to_c()['data']
    cap = 1 << 22
    buf = C.create_string_buffer(cap)
    src_len, cubin = C.c_int64(), C.c_int64()
    nk, npass, nrem, npipe = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    rc = _capi.lib().dnm_jit_dryrun(masks.size, _capi.ip(masks), _capi.ip(offs), _capi.ip(signs), _capi.fp(coeffs),
                                    C.byref(cdata.desc), nranks, rank, tile_bits, far_bits, pipeline, tune, buf, cap,
                                    _capi.ip(np.ctypeslib.as_array(C.pointer(src_len), (1,))),
                                    _capi.ip(np.ctypeslib.as_array(C.pointer(cubin), (1,))),
                                    C.byref(nk), C.byref(npass), C.byref(nrem), C.byref(npipe))
    assert rc == 0, _capi.lib().dnm_last_error().decode()
    return dict(src=buf.value.decode(), cubin=cubin.value, kernels=nk.value, passes=npass.value, remote=nrem.value,
                pipelined=npipe.value)


def test_default_plan_L30_mbl_is_two_tma_staged_passes():
    r = dryrun('MBL', 30)
    assert r['passes'] == 2 and r['kernels'] == 2 and r['cubin'] > 0
    src = r['src']
    # writing pass: contiguous tile by cp.async.bulk; accumulating pass: tensor-map load and reduce-add
    assert 'cp.async.bulk.shared::cluster.global.mbarrier' in src
    assert re.search(r'cp\.async\.bulk\.tensor\.\dd\.shared::cluster\.global', src)
    assert re.search(r'cp\.reduce\.async\.bulk\.tensor\.\dd\.global\.shared::cta\.add', src)
    assert src.count(' FAR') >= 10          # the L2 window carries the masks that leave the 11-bit window


@pytest.mark.parametrize('name', ['MBL', 'heisenberg', 'long_range', 'ising', 'XX'])
@pytest.mark.parametrize('tune', [0, 1, 2, 3])
def test_every_autotuner_shape_compiles(name, tune):
    r = dryrun(name, 30, tune=tune)
    assert r['kernels'] == r['passes'] >= 2 and r['cubin'] > 0


@pytest.mark.parametrize('name,L', [('MBL', 31), ('MBL', 33), ('long_range', 33), ('heisenberg', 32), ('XX', 33)])
def test_sharded_plans_fold_the_remote_masks(name, L):
    nranks = 1 << (L - 30)
    for rank in sorted({0, 1, nranks - 1, nranks // 2 + 1 if nranks > 2 else 0}):
        r = dryrun(name, L, nranks=nranks, rank=rank)
        assert r['cubin'] > 0 and r['kernels'] == r['passes']
        assert r['remote'] >= L - 30, (name, L, rank, r['remote'])      # every cross-rank mask became a FAR group
        assert 'xs.p[' in r['src']                                       # ... read through the peer mappings


def test_parity_subspace_and_pipelined_variant_compile():
    r = dryrun('heisenberg', 31, sub=Parity('even', L=31))
    assert r['cubin'] > 0 and r['kernels'] == r['passes']
    p = dryrun('MBL', 30, tile_bits=11, far_bits=8, pipeline=1)
    assert p['cubin'] > 0 and p['pipelined'] == p['kernels'] >= 1
    assert 'mbar_wait(&full[b], phase)' in p['src']
    q = dryrun('MBL', 30, tile_bits=12, far_bits=6, pipeline=1)          # 64 KB tiles: the box needs a 5-D tensor map
    assert q['cubin'] > 0


def test_non_lean_operator_generates_nothing():
    r = dryrun('SYK', 12, sub=Parity('even', L=12))
    assert r['kernels'] == 0 and r['cubin'] == 0


def test_committed_samples_are_what_the_generator_emits():
    """csrc/generated_samples/*.cu are the sources that were measured on the B200 (profiles/r02_*): a change to
    the generator that alters them has to be re-measured, so it must show up here first."""
    import os
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'dynamite_b200', 'csrc',
                        'generated_samples')
    assert dryrun('MBL', 30)['src'] == open(os.path.join(root, 'mbl_L30_default.cu')).read()
    assert dryrun('MBL', 30, pipeline=1)['src'] == open(os.path.join(root, 'mbl_L30_pipelined_tma.cu')).read()

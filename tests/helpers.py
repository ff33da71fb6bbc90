"""Shared test helpers: golden fixtures, and building the SAME subspace / MSC
for the oracle (checker) and for the product (dynamite_b200) from one spec."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, 'golden')

_cases = None


def golden_cases():
    global _cases
    if _cases is None:
        data = np.load(os.path.join(GOLDEN, 'reference_cases.npz'))
        with open(os.path.join(GOLDEN, 'reference_cases.json')) as f:
            meta = json.load(f)
        _cases = {}
        for c in meta['cases']:
            tag = c['tag']
            d = dict(c)
            for k in data.files:
                if k.startswith(tag + '.'):
                    d[k[len(tag) + 1:]] = data[k]
            _cases[tag] = d
    return _cases


def kats():
    with open(os.path.join(GOLDEN, 'reference_kats.json')) as f:
        return json.load(f)


def case_terms(case):
    return list(zip(case['msc_masks'].tolist(), case['msc_signs'].tolist(), case['msc_coeffs'].tolist()))


def product_subspace(spec):
    """dynamite_b200 subspace object from a spec dict"""
    from dynamite_b200 import subspaces as S
    t = spec['type']
    if t == 'full':
        return S.Full(L=spec['L'])
    if t == 'parity':
        return S.Parity(spec['space'], L=spec['L'])
    if t == 'spinconserve':
        return S.SpinConserve(spec['L'], spec['k'])
    if t == 'explicit':
        return S.Explicit(spec['states'], L=spec['L'])
    raise ValueError(t)


def product_mat(terms, left_spec, right_spec, xparity=False, precompute_diag=False):
    """Build a device matrix through the _backend mirror from raw MSC terms."""
    from dynamite_b200 import msc_tools
    from dynamite_b200._backend import bpetsc
    msc = msc_tools.make_msc(terms)
    masks, offs = msc_tools.mask_offsets(msc)
    left, right = product_subspace(left_spec), product_subspace(right_spec)
    mat = bpetsc.build_mat(masks=masks, mask_offsets=offs, signs=np.ascontiguousarray(msc['signs']),
                           coeffs=np.ascontiguousarray(msc['coeffs']), left_subspace=left._to_c(),
                           right_subspace=right._to_c(), xparity=xparity, shell=True, gpu=True)
    if precompute_diag:
        bpetsc.precompute_diagonal(mat)
    return mat


def device_mult(mat, x):
    from dynamite_b200.petsc import Vec
    m, n = mat.getSize()
    xv, yv = Vec(n), Vec(m)
    xv[0:n] = x
    mat.mult(xv, yv)
    y = yv[0:m]
    xv.destroy()
    yv.destroy()
    return y


def rand_state(n, seed):
    R = np.random.RandomState(seed)
    v = R.standard_normal(n) + 1j * R.standard_normal(n)
    return v / np.linalg.norm(v)


def rel_err(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)

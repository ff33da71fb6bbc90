"""
``State``: a state vector in HBM tied to a subspace (host mirror of reference
``states.py``; the ``vec`` attribute is a :class:`dynamite_b200.petsc.Vec`).
"""
from os import urandom

import numpy as np

from . import config, validate
from .petsc import COMM_WORLD, Vec


class UninitializedError(RuntimeError):
    pass


import pickle as _pickle


class _SubspaceUnpickler(_pickle.Unpickler):
    """Metadata written by the reference names ``dynamite.subspaces.<Class>``; files written here name
    ``dynamite_b200.subspaces.<Class>``.  Both resolve to this package's classes."""

    def find_class(self, module, name):
        if module in ('dynamite.subspaces', 'dynamite_b200.subspaces'):
            from . import subspaces
            return getattr(subspaces, name)
        return super().find_class(module, name)


class State:
    def __init__(self, state=None, subspace=None, L=None, seed=None):
        self._vec = None
        self._initialized = False
        if subspace is None:
            subspace = config.subspace.copy() if config._subspace is not None else None
        if subspace is None:
            from .subspaces import Full
            subspace = Full()
        self._subspace = subspace
        if L is None:
            L = config.L
        if L is not None:
            self.L = L
        if state is not None:
            if state == 'random':
                self.set_random(seed=seed)
            elif state == 'uniform':
                self.set_uniform()
            else:
                self.set_product(state)

    # ---- bookkeeping ---------------------------------------------------------
    @property
    def L(self):
        return self._subspace.L

    @L.setter
    def L(self, value):
        self._subspace.L = validate.L(value)

    @property
    def subspace(self):
        return self._subspace

    @property
    def vec(self):
        """device vector, created on first use with the subspace's dimension"""
        if self._vec is None:
            if self.L is None:
                raise ValueError('Must set state size before building vector (set L or pass a subspace with L)')
            config._initialize()
            self._vec = Vec(self.subspace.get_dimension())
        return self._vec

    @property
    def initialized(self):
        return self._initialized

    def set_initialized(self):
        self._initialized = True

    def assert_initialized(self):
        if not self._initialized:
            raise UninitializedError('State vector has not been initialized. Set its value with one of the '
                                     '"set_" methods or the "state" argument to the constructor.')

    def copy(self, result=None):
        if result is None:
            result = State(subspace=self.subspace)
        elif result.subspace != self.subspace:
            raise ValueError('subspace of state and result must match')
        self.vec.copy(result.vec)
        result._initialized = self._initialized
        return result

    # ---- initial values -----------------------------------------------------------
    @classmethod
    def str_to_state(cls, s, L):
        """'DUDU..' / '0101..' (leftmost character = spin 0) or int -> state integer"""
        if isinstance(s, str):
            if len(s) != L:
                raise ValueError('state string must have length L')
            if set(s) <= {'U', 'D'}:
                down = 'D'
            elif set(s) <= {'0', '1'}:
                down = '1'
            else:
                raise ValueError('state string can only contain characters U and D, or 0 and 1')
            return sum(1 << i for i, ch in enumerate(s) if ch == down)
        state = int(s)
        if state >> L != 0 or state < 0:
            raise ValueError(f'value (binary: {bin(state)[2:]}) does not correspond to a valid state of length L')
        return state

    def set_product(self, s):
        if self.L is None and isinstance(s, str):
            self.L = len(s)
        idx = self.subspace.state_to_idx(self.str_to_state(s, self.L))
        if idx == -1:
            raise ValueError('Provided initial state not in requested subspace.')
        self.vec.set(0)
        a, b = self.vec.getOwnershipRange()
        if a <= idx < b:
            self.vec[int(idx)] = 1
        self.set_initialized()

    def set_uniform(self):
        self.vec.set(1 / np.sqrt(self.subspace.get_dimension()))
        self.set_initialized()

    def set_random(self, seed=None, normalize=True):
        """Gaussian random state; per-rank stream ``RandomState((seed + rank) % 2**32)`` with the
        real parts drawn before the imaginary parts (reference ``states.py:272-318``)."""
        a, b = self.vec.getOwnershipRange()
        if seed is None:
            seed = int.from_bytes(urandom(4), 'big', signed=False)
        R = np.random.RandomState((seed + COMM_WORLD.rank) % 2**32)
        n = b - a
        block = 1 << 22
        if n <= block:
            self.vec[a:b] = R.standard_normal(n) + 1j * R.standard_normal(n)
        else:
            # same stream as one big draw: all real parts first, then all imaginary parts
            vals = np.empty(n, dtype=np.complex128)
            vals.real = R.standard_normal(n)
            vals.imag = R.standard_normal(n)
            self.vec[a:b] = vals
        if normalize:
            self.vec.normalize()
        self.set_initialized()

    def set_all_by_function(self, val_fn, vectorize=False):
        a, b = self.vec.getOwnershipRange()
        states = self.subspace.idx_to_state(np.arange(a, b))
        if vectorize:
            vals = val_fn(states)
        else:
            vals = np.array([val_fn(int(s)) for s in states], dtype=np.complex128)
        self.vec[a:b] = vals
        self.set_initialized()

    def project(self, index, value):
        """Project spin ``index`` onto ``value`` (0 or 1) and renormalise (reference ``states.py:364-401``)."""
        if not self.subspace.product_state_basis:
            raise ValueError('projection only implemented for product state subspaces')
        if value not in (0, 1):
            raise ValueError('value must be 0 or 1')
        a, b = self.vec.getOwnershipRange()
        states = self.subspace.idx_to_state(np.arange(a, b))
        vals = self.vec[a:b]
        vals[((states >> index) & 1) != value] = 0
        self.vec[a:b] = vals
        nrm = self.vec.norm()
        if nrm == 0:
            raise ValueError('projection leaves the zero vector')
        self.vec.scale(1 / nrm)

    # ---- host copies ---------------------------------------------------------------
    def to_numpy(self, to_all=False):
        """The state as a numpy array (reference ``states.py:703-740``): on one rank the whole vector.
        Sharded: with ``to_all`` every rank gets the whole vector (each rank's block is read through the
        peer mappings of the Vec), otherwise rank 0 gets it and the other ranks get ``None``."""
        self.assert_initialized()
        if COMM_WORLD.size == 1:
            return self.vec.getArray()
        from . import _capi
        COMM_WORLD.barrier()       # every rank's block is complete
        full = None
        if to_all or COMM_WORLD.rank == 0:
            full = np.empty(len(self), dtype=np.complex128)
            _capi.check(_capi.lib().dnm_vec_get_host_global(self.vec.handle, 0, full.size, _capi.fp(full)))
        COMM_WORLD.barrier()       # nobody changes its block while others still read it
        return full

    # ---- checkpoint / resume ---------------------------------------------------------
    _VEC_CLASSID = 1211214   # PETSc VEC_FILE_CLASSID

    def save(self, fname, indices=None):
        """``<fname>.vec`` in PETSc's binary Vec layout -- class id and length as big-endian PetscInt,
        then interleaved big-endian complex128 -- and ``<fname>.metadata`` = the pickled subspace, as
        reference ``states.py:627-652``.  ``indices`` = 32 or 64 selects the PetscInt width of the
        header; by default 32 bits (what a default PETSc build, and the reference's own
        ``petsc_config/complex-opt.py``, reads and writes) unless the length needs 64.
        Multi-rank: rank 0 writes the metadata and the header, every rank its own block."""
        import pickle
        self.assert_initialized()
        n = len(self)
        if indices is None:
            indices = 32 if n < 2**31 else 64
        if indices not in (32, 64) or (indices == 32 and n >= 2**31):
            raise ValueError('indices must be 32 or 64 (and 64 for vectors of 2^31 entries or more)')
        hdr = np.array([self._VEC_CLASSID, n], dtype='>i4' if indices == 32 else '>i8').tobytes()
        rank = COMM_WORLD.rank
        if rank == 0:
            with open(fname + '.metadata', 'wb') as f:
                pickle.dump(self.subspace, f)
            with open(fname + '.vec', 'wb') as f:
                f.write(hdr)
                f.truncate(len(hdr) + 16 * n)
        if COMM_WORLD.size > 1:
            COMM_WORLD.barrier()
        first, last = self.vec.getOwnershipRange()
        with open(fname + '.vec', 'r+b') as f:
            f.seek(len(hdr) + 16 * first)
            block = 1 << 22
            for a in range(first, last, block):
                b = min(last, a + block)
                f.write(self.vec[a:b].astype('>c16').tobytes())
        if COMM_WORLD.size > 1:
            COMM_WORLD.barrier()

    @classmethod
    def from_file(cls, fname):
        """inverse of :meth:`save` (uses ``pickle``: do not load untrusted files).  Accepts headers
        written with 32-bit or 64-bit PetscInt; every rank reads its own block."""
        import pickle
        with open(fname + '.metadata', 'rb') as f:
            subspace = _SubspaceUnpickler(f).load()
        with open(fname + '.vec', 'rb') as f:
            head = f.read(4)
            if len(head) == 4 and int(np.frombuffer(head, dtype='>i4')[0]) == cls._VEC_CLASSID:
                n = int(np.frombuffer(f.read(4), dtype='>i4')[0])
                hlen = 8
            else:
                rest = f.read(12)
                if len(rest) != 12:
                    raise RuntimeError('corrupt data encountered when loading state from file')
                classid, n = (int(v) for v in np.frombuffer(head + rest, dtype='>i8'))
                hlen = 16
                if classid != cls._VEC_CLASSID:
                    raise RuntimeError('corrupt data encountered when loading state from file')
            if subspace.get_dimension() != n:
                raise RuntimeError('corrupt data encountered when loading state from file')
            rtn = cls(subspace=subspace)
            first, last = rtn.vec.getOwnershipRange()
            f.seek(hlen + 16 * first)
            block = 1 << 22
            for a in range(first, last, block):
                b = min(last, a + block)
                rtn.vec[a:b] = np.frombuffer(f.read(16 * (b - a)), dtype='>c16').astype(np.complex128)
        rtn.set_initialized()
        return rtn

    # ---- algebra --------------------------------------------------------------------
    def dot(self, x):
        """<self|x> (conjugate-linear in self)"""
        return x.vec.dot(self.vec)

    def norm(self):
        return self.vec.norm()

    def normalize(self):
        self.vec.normalize()

    def scale(self, c):
        self.vec.scale(c)

    def axpy(self, alpha, x):
        self.vec.axpy(alpha, x.vec)

    def scale_and_sum(self, alpha, beta, x):
        """self = alpha*self + beta*x"""
        self.vec.axpby(beta, alpha, x.vec)

    def __imul__(self, c):
        self.scale(c)
        return self

    def __mul__(self, c):
        out = self.copy()
        out.scale(c)
        return out

    __rmul__ = __mul__

    def __itruediv__(self, c):
        self.scale(1 / c)
        return self

    def __iadd__(self, x):
        if isinstance(x, State):
            self.axpy(1, x)
        else:           # a number: added to every amplitude (reference states.py:799-807)
            self.assert_initialized()
            self.vec.shift(x)
        return self

    def __add__(self, x):
        out = self.copy()
        out += x
        return out

    def __isub__(self, x):
        self.axpy(-1, x)
        return self

    def __sub__(self, x):
        out = self.copy()
        out -= x
        return out

    def __len__(self):
        return self.subspace.get_dimension()

    def __repr__(self):
        return f'State(subspace={self.subspace!r}, initialized={self._initialized})'

"""
Subspace classes (host mirror of reference ``subspaces.py``): thin objects that
hold the parameters of a subspace and delegate dimension / rank / unrank to the
backend (``_backend.bsubspace`` -> C ABI), so there is exactly one
implementation of the maps -- the one the CUDA kernels use.
"""
import math
from zlib import crc32

import numpy as np

from . import config, validate
from ._backend import bsubspace
from .msc_tools import combine_and_sort, dnm_int_t, parity


def _as_index_array(val):
    single = not hasattr(val, '__len__')
    arr = np.ascontiguousarray(np.asarray(val, dtype=dnm_int_t).reshape(-1))
    return single, arr


class Subspace:
    """Base class (reference ``subspaces.py:20-171``)."""

    _chksum = None
    _product_state_basis = True

    def __eq__(self, other):
        if other is self:
            return True
        if not isinstance(other, Subspace):
            raise ValueError('Cannot compare Subspace to non-Subspace type')
        if self.L is None:
            raise ValueError('Cannot evaluate equality of subspaces before setting L')
        if self.get_dimension() != other.get_dimension():
            return False
        return self.get_checksum() == other.get_checksum()

    def identical(self, other):
        return hash(self) == hash(other)

    @property
    def L(self):
        return self._L

    @L.setter
    def L(self, value):
        if self.L is not None and value != self.L:
            raise AttributeError('Cannot change L for a subspace after it is set')
        self._L = self.check_L(validate.L(value))

    def check_L(self, value):
        return value

    @property
    def product_state_basis(self):
        return self._product_state_basis

    def copy(self):
        from copy import deepcopy
        return deepcopy(self)

    def get_checksum(self):
        """crc32 over the idx->state map in blocks of 2^14 (reference ``subspaces.py:88-102``)."""
        if self._chksum is None:
            chk = 0
            dim = self.get_dimension()
            for start in range(0, dim, 1 << 14):
                stop = min(start + (1 << 14), dim)
                chk = crc32(self.idx_to_state(np.arange(start, stop)), chk)
            self._chksum = chk
        return self._chksum

    def get_dimension(self):
        return self._get_dimension()

    def idx_to_state(self, idx):
        single, arr = _as_index_array(idx)
        dim = self.get_dimension()
        if arr.size and (arr.min() < 0 or arr.max() >= dim):
            bad = arr[(arr < 0) | (arr >= dim)]
            what = f'Index {bad[0]}' if bad.size == 1 else f'Indices {bad}'
            raise ValueError(f'{what} out of bounds for subspace of dimension {dim}')
        out = self._idx_to_state(arr)
        return out[0] if single else out

    def state_to_idx(self, state):
        single, arr = _as_index_array(state)
        out = self._state_to_idx(arr)
        return out[0] if single else out


class _ProductStateSubspace(Subspace):
    _enum = None
    _suffix = None

    def __init__(self, L=None):
        self._L = None
        if L is None:
            L = config.L
        if L is not None:
            self.L = L

    def _get_dimension(self):
        return getattr(bsubspace, 'get_dimension_' + self._suffix)(self._get_cdata())

    def _idx_to_state(self, idx):
        return getattr(bsubspace, 'idx_to_state_' + self._suffix)(idx, self._get_cdata())

    def _state_to_idx(self, state):
        return getattr(bsubspace, 'state_to_idx_' + self._suffix)(state, self._get_cdata())

    def _require_L(self):
        if self.L is None:
            raise ValueError('L has not been set for this subspace')

    def _to_c(self):
        return {'type': self._enum, 'data': self._get_cdata()}


class Full(_ProductStateSubspace):
    _enum = bsubspace.SubspaceType.FULL
    _suffix = 'Full'

    def __eq__(self, other):
        if isinstance(other, Full):
            return other.L == self.L
        return super().__eq__(other)

    def __hash__(self):
        return hash((self._enum, self.L))

    def __repr__(self):
        return 'Full()' if self.L is None else f'Full(L={self.L})'

    def _get_cdata(self):
        self._require_L()
        return bsubspace.CFull(self.L)


class Parity(_ProductStateSubspace):
    """States with an even (space 0 / 'even') or odd (1 / 'odd') number of 1 bits."""
    _enum = bsubspace.SubspaceType.PARITY
    _suffix = 'Parity'

    def __init__(self, space, L=None):
        super().__init__(L)
        if space in (0, 'even'):
            self._space = 0
        elif space in (1, 'odd'):
            self._space = 1
        else:
            raise ValueError(f'Invalid parity space "{space}" (valid choices are 0, 1, "even", or "odd")')

    @property
    def space(self):
        return self._space

    def __hash__(self):
        return hash((self._enum, self.L, self.space))

    def __repr__(self):
        arg = "'even'" if self.space == 0 else "'odd'"
        return f'Parity({arg})' if self.L is None else f'Parity({arg}, L={self.L})'

    def _get_cdata(self):
        self._require_L()
        return bsubspace.CParity(self.L, self.space)


class SpinConserve(_ProductStateSubspace):
    """States with exactly ``k`` set bits (total magnetisation sector)."""
    _enum = bsubspace.SubspaceType.SPIN_CONSERVE
    _suffix = 'SpinConserve'

    def __init__(self, L, k, spinflip=None):
        super().__init__(L=L)
        if not 0 <= k <= self.L:
            raise ValueError('k must be between 0 and L')
        if spinflip is not None:
            raise DeprecationWarning('spinflip argument has been deprecated; use the XParity class instead.')
        self._k = int(k)
        # nchoosek[kk, n] = C(n, kk); layout fixed by the backend (subspaces.py:341-352)
        self._nchoosek = np.array([[math.comb(n, kk) for n in range(self.L + 1)]
                                   for kk in range(self._k + 1)], dtype=dnm_int_t)

    @property
    def k(self):
        return self._k

    def __hash__(self):
        return hash((self._enum, self.L, self.k))

    def __repr__(self):
        return f'SpinConserve(L={self.L}, k={self.k})'

    def _get_cdata(self):
        self._require_L()
        return bsubspace.CSpinConserve(self.L, self.k, self._nchoosek)


class Explicit(_ProductStateSubspace):
    """Subspace given as an explicit list of product states."""
    _enum = bsubspace.SubspaceType.EXPLICIT
    _suffix = 'Explicit'

    def __init__(self, state_list, L=None):
        self.state_map = np.ascontiguousarray(state_list, dtype=dnm_int_t)
        if np.all(self.state_map[:-1] <= self.state_map[1:]):
            self.rmap_indices = np.array([-1], dtype=dnm_int_t)   # sentinel: already sorted
            self.rmap_states = self.state_map
        else:
            self.rmap_indices = np.ascontiguousarray(np.argsort(self.state_map, kind='stable'), dtype=dnm_int_t)
            self.rmap_states = np.ascontiguousarray(self.state_map[self.rmap_indices])
        if np.any(self.rmap_states[1:] == self.rmap_states[:-1]):
            raise ValueError('values in state_list must be unique')
        super().__init__(L=L)

    def check_L(self, value):
        if int(self.rmap_states[-1]) >> value != 0:
            raise ValueError('State in subspace has more spins than provided')
        return value

    def __hash__(self):
        return hash((self._enum, self.get_checksum()))

    def __repr__(self):
        n = len(self.state_map)
        shown = list(self.state_map) if n < 1000 else list(self.state_map[:3]) + ['...'] + list(self.state_map[-3:])
        L = self.L if self.L is not None else int(self.rmap_states[-1]).bit_length()
        body = ', '.join(x if isinstance(x, str) else '0b' + bin(int(x))[2:].zfill(L) for x in shown)
        return f'Explicit([{body}]' + (f', L={self.L})' if self.L is not None else ')')

    def _get_cdata(self):
        self._require_L()
        return bsubspace.CExplicit(self.L, self.state_map, self.rmap_indices, self.rmap_states)


class Auto(Explicit):
    """The connected component of ``state`` under the operator ``H`` (breadth-first
    search in the backend, reference ``subspaces.py:466-529``)."""

    def __init__(self, H, state, size_guess=None, sort=True):
        from .states import State
        H.establish_L()
        self._repr_args = f'H={H!r}, state={state!r}'
        self.state = State.str_to_state(state, H.L)
        if size_guess is None:
            size_guess = 2 ** H.L
        state_map = np.empty(size_guess, dtype=dnm_int_t)
        H.reduce_msc()
        dim = bsubspace.compute_rcm(H.msc['masks'], H.msc['signs'], H.msc['coeffs'],
                                    state_map, self.state, H.L)
        state_map = state_map[:dim]
        if sort:
            state_map.sort()
        else:
            state_map = state_map[::-1]   # reverse Cuthill-McKee order
        Explicit.__init__(self, state_map, L=H.L)

    def __repr__(self):
        return f'Auto({self._repr_args})'


class XParity(Subspace):
    """Parity in the X basis, optionally on top of a product-state ``parent``
    (reference ``subspaces.py:532-795``).  Basis states are
    ``|c> +- |complement(c)>``, represented by the member with spin L-1 = 0; the
    backend sees the parent's maps on the first half of the indices."""

    _product_state_basis = False

    def __init__(self, parent=None, sector='+', L=None):
        if parent is None:
            parent = Full()
        self._parent = parent
        if L is not None:
            self.parent.L = L
        self._validate_parent(parent)
        if sector in ('+', +1):
            self._sector = +1
        elif sector in ('-', -1):
            self._sector = -1
        else:
            raise ValueError('invalid value for sector')

    @classmethod
    def _validate_parent(cls, parent):
        if not parent.product_state_basis:
            raise ValueError('parent must be a product state subspace')
        if isinstance(parent, Full):
            return
        if parent.L is None:
            raise ValueError('L must be set for the parent subspace')
        if isinstance(parent, Parity):
            if parent.L % 2 == 0:
                return
            raise ValueError('Parity is only compatible with XParity when L is even')
        if isinstance(parent, SpinConserve):
            if parent.L == 2 * parent.k:
                return
            raise ValueError('SpinConserve is only compatible with XParity when k=L/2')
        dim = parent.get_dimension()
        if dim % 2:
            raise ValueError('parent subspace must have even dimension')
        flip = (1 << parent.L) - 1
        for start in range(0, dim // 2, 1024):
            block = parent.idx_to_state(np.arange(start, min(start + 1024, dim // 2)))
            if np.count_nonzero(block >> (parent.L - 1)):
                raise ValueError('first dim/2 basis states must have spin L-1 up (0 in integer notation)')
            if np.any(parent.state_to_idx(block ^ flip) == -1):
                raise ValueError('the complement of every state in subspace (all spins flipped) '
                                 'must also be in subspace')

    @property
    def parent(self):
        return self._parent

    @property
    def sector(self):
        return self._sector

    def reduce_msc(self, msc, check_conserves=False):
        """Fold an operator onto the representatives: drop terms that anticommute
        with prod(sigma_x) (odd sign-mask parity), complement the flip mask of
        terms that would leave the representative half, and for the '-' sector
        negate those (reference ``subspaces.py:633-674``)."""
        msc = msc.copy()
        keep = parity(msc['signs']) == 0
        conserved = bool(np.all(keep))
        msc = msc[keep]
        leaves = (msc['masks'] >> (self.L - 1)) != 0
        msc['masks'][leaves] ^= (1 << self.L) - 1
        if self.sector == -1:
            msc['coeffs'][leaves] *= -1
        msc = combine_and_sort(msc)
        return (msc, conserved) if check_conserves else msc

    def convert_state(self, state):
        """XParity <-> parent conversion (reference ``subspaces.py:676-762``), done
        through host copies of the local blocks."""
        from .petsc import COMM_WORLD
        from .states import State
        state.assert_initialized()
        if COMM_WORLD.size > 1:
            # the conversion pairs index i with the index of the globally flipped state, which lives on
            # another rank's shard; only the single-process form exists here
            raise NotImplementedError('XParity.convert_state works on unsharded vectors only in this backend')
        flip = (1 << self.L) - 1
        half = self.get_dimension()
        src = state.to_numpy()
        if state.subspace is self:
            out = State(subspace=self.parent)
            rep = self.idx_to_state(np.arange(half))
            vals = np.empty(2 * half, dtype=np.complex128)
            vals[:half] = src
            vals[self.parent.state_to_idx(rep ^ flip)] = self.sector * src
        elif state.subspace is self.parent:
            out = State(subspace=self)
            second = self.parent.idx_to_state(np.arange(half, 2 * half))
            vals = src[:half].copy()
            vals[self.state_to_idx(second ^ flip)] += self.sector * src[half:]
        else:
            raise ValueError('subspace of input state must be this XParity subspace or its parent')
        out.vec[0:vals.size] = vals / np.sqrt(2)
        out.set_initialized()
        return out

    def __hash__(self):
        return hash(('XParity', self.sector, self.parent))

    def __repr__(self):
        return f'XParity({self.parent!r}, sector={self.sector:+d})'

    @property
    def _L(self):
        return self.parent.L

    @_L.setter
    def _L(self, value):
        self.parent.L = value

    def _get_dimension(self):
        return self.parent.get_dimension() // 2

    def _idx_to_state(self, idx):
        return self.parent.idx_to_state(idx)

    def _state_to_idx(self, state):
        if np.count_nonzero(state >> (self.L - 1)):
            raise ValueError('invalid state')
        return self.parent.state_to_idx(state)

    def _to_c(self):
        return self.parent._to_c()

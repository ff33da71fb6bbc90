"""
``Operator``: symbolic sums of Pauli strings kept in MSC form, and the calls
that put them on the GPU (host mirror of the parts of reference
``operators.py`` that lead into the shell-matrix path: ``build_mat``
``:570-631``, ``dot`` ``:1063-1108``, ``conserves`` ``:382-423``,
``infinity_norm`` ``:206-224``, ``evolve``/``eigsolve``).

String/LaTeX representations and save/load are not part of this path and are
reduced to a plain ``repr``.
"""
import numpy as np

from . import config, msc_tools, validate
from .msc_tools import msc_dtype


class Operator:
    def __init__(self, msc=None, L=None, name=None):
        self._msc = msc_tools.make_msc([] if msc is None else msc).copy()
        self._is_reduced = False
        self._L = None
        self._subspaces = []
        self._mats = {}
        self._shell = True
        self._precompute_diagonal = True
        self._allow_projection = False
        self._name = name
        if L is None:
            L = config.L
        if L is not None:
            self.L = L

    # ---- basic properties -----------------------------------------------------
    @property
    def L(self):
        return self._L

    @L.setter
    def L(self, value):
        value = validate.L(value)
        if value < self.max_spin_idx + 1:
            raise ValueError(f'Cannot set L smaller than one plus the largest spin index '
                             f'on which the operator has support (max_spin_idx = {self.max_spin_idx})')
        for left, right in self._subspaces:
            left.L = value
            right.L = value
        self._L = value

    def establish_L(self):
        """If L is unset, use the smallest chain that holds the operator."""
        if self._L is None:
            self.L = self.max_spin_idx + 1

    @property
    def max_spin_idx(self):
        return msc_tools.max_spin_idx(self._msc)

    @property
    def msc(self):
        return self._msc

    @msc.setter
    def msc(self, value):
        self._msc = msc_tools.make_msc(value)
        self._is_reduced = False
        self.destroy_mat()

    @property
    def is_reduced(self):
        return self._is_reduced

    def reduce_msc(self):
        if not self._is_reduced:
            self._msc = msc_tools.combine_and_sort(self._msc)
            self._is_reduced = True

    @property
    def nterms(self):
        self.reduce_msc()
        return int(self._msc.size)

    @property
    def nnz(self):
        return msc_tools.nnz(self._msc)

    @property
    def dim(self):
        return (self.left_subspace.get_dimension(), self.right_subspace.get_dimension())

    @property
    def density(self):
        return self.nnz / self.dim[1]

    def is_hermitian(self):
        self.reduce_msc()
        return msc_tools.is_hermitian(self._msc)

    def copy(self):
        rtn = Operator(msc=self._msc, L=self._L, name=self._name)
        rtn._is_reduced = self._is_reduced
        rtn._shell = self._shell
        rtn._precompute_diagonal = self._precompute_diagonal
        rtn._allow_projection = self._allow_projection
        rtn._subspaces = [(l.copy(), r.copy()) for l, r in self._subspaces]
        return rtn

    # ---- options ----------------------------------------------------------------
    @property
    def shell(self):
        return self._shell

    @shell.setter
    def shell(self, value):
        if not value:
            raise ValueError('dynamite_b200 only provides shell (matrix-free) matrices.')
        self._shell = True

    @property
    def precompute_diagonal(self):
        """Cache the real diagonal (8 bytes/row) at build time so each multiply
        skips the diagonal term loop (reference ``operators.py:249-270``)."""
        return self._precompute_diagonal

    @precompute_diagonal.setter
    def precompute_diagonal(self, value):
        if bool(value) != self._precompute_diagonal:
            self.destroy_mat()
        self._precompute_diagonal = bool(value)

    @property
    def allow_projection(self):
        return self._allow_projection

    @allow_projection.setter
    def allow_projection(self, value):
        self._allow_projection = bool(value)

    # ---- subspaces ----------------------------------------------------------------
    def get_subspace_list(self):
        if not self._subspaces:
            space = config.subspace.copy()
            if self.L is not None and space.L is None:
                space.L = self.L
            self._subspaces = [(space, space)]
        return self._subspaces

    @property
    def left_subspace(self):
        return self.get_subspace_list()[-1][0]

    @property
    def right_subspace(self):
        return self.get_subspace_list()[-1][1]

    @property
    def subspace(self):
        if self.left_subspace != self.right_subspace:
            raise ValueError('Left and right subspaces are different for this operator. '
                             'use Operator.left_subspace and Operator.right_subspace to '
                             'access them individually.')
        return self.left_subspace

    @subspace.setter
    def subspace(self, value):
        self.add_subspace(value, value)

    def add_subspace(self, left, right=None):
        from .subspaces import Subspace
        if right is None:
            right = left
        elif left is not right and not (left.product_state_basis and right.product_state_basis):
            raise ValueError('subspaces must be the same object if either is not a product state basis')
        for s in (left, right):
            if not isinstance(s, Subspace):
                raise ValueError('subspace can only be set to objects of Subspace type')
        if self.L is None:
            if left.L is not None:
                self.L = left.L
            elif right.L is not None:
                self.L = right.L
        if self.L is not None:
            for s in (left, right):
                if s.L is None:
                    s.L = self.L
                elif s.L != self.L:
                    raise ValueError('operator and subspaces must all have the same spin chain length L')
        if not self.has_subspace(left, right):
            self._subspaces = [p for p in self._subspaces] + [(left, right)]
        else:
            # move the existing pair to the end so it becomes the default
            pair = next(p for p in self._subspaces if p[0].identical(left) and p[1].identical(right))
            self._subspaces.remove(pair)
            self._subspaces.append(pair)

    def has_subspace(self, left, right=None):
        if right is None:
            right = left
        return any(left.identical(l) and right.identical(r) for l, r in self._subspaces or self.get_subspace_list())

    def conserves(self, left, right=None):
        """Does the operator map ``right`` into ``left``? (reference ``operators.py:382-423``)"""
        from .subspaces import XParity
        self.establish_L()
        if right is None:
            right = left
        if not (left.product_state_basis and right.product_state_basis) and left is not right:
            raise ValueError('if left or right subspace is not a product state basis, '
                             'they must be the same object')
        left.L = self.L
        right.L = self.L
        self.reduce_msc()
        if not left.product_state_basis:
            msc, conserved = left.reduce_msc(self.msc, check_conserves=True)
            if not conserved:
                return False
        else:
            msc = self.msc
        if msc.size == 0:
            return True
        masks, offsets = msc_tools.mask_offsets(msc)
        config._initialize()
        from ._backend import bpetsc
        return bpetsc.check_conserves(masks=masks, mask_offsets=offsets,
                                      signs=np.ascontiguousarray(msc['signs']),
                                      coeffs=np.ascontiguousarray(msc['coeffs']),
                                      left_subspace=left._to_c(), right_subspace=right._to_c(),
                                      xparity=isinstance(left, XParity))

    # ---- the device matrix ------------------------------------------------------------
    def get_mat(self, subspaces=None):
        if subspaces is None:
            subspaces = (self.left_subspace, self.right_subspace)
        if subspaces not in self._mats:
            self.build_mat(subspaces)
        return self._mats[subspaces]

    def build_mat(self, subspaces=None):
        """Build the shell matrix on the GPU (reference ``operators.py:570-631``)."""
        from .subspaces import XParity
        if subspaces is None:
            subspaces = (self.left_subspace, self.right_subspace)
        if not self.has_subspace(*subspaces):
            raise ValueError('Attempted to build matrix for a subspace that has not been added to the operator.')
        config._initialize()
        from ._backend import bpetsc
        self.establish_L()
        self.reduce_msc()
        msc = self.msc if subspaces[0].product_state_basis else subspaces[0].reduce_msc(self.msc)
        if not self.allow_projection and not self.conserves(*subspaces):
            raise ValueError("Constructing the operator's matrix on this subspace yields a projection "
                             '(e.g. subspace is not conserved by the operator). If this behavior is '
                             'desired, set the Operator.allow_projection parameter to True.')
        if not msc_tools.is_hermitian(msc):
            raise ValueError('Building non-Hermitian matrices currently not supported.')
        if msc.size == 0:
            msc = msc_tools.make_msc([(0, 0, 0.0)])   # the zero operator still needs one mask
        masks, offsets = msc_tools.mask_offsets(msc)
        mat = bpetsc.build_mat(masks=masks, mask_offsets=offsets,
                               signs=np.ascontiguousarray(msc['signs']),
                               coeffs=np.ascontiguousarray(msc['coeffs']),
                               left_subspace=subspaces[0]._to_c(), right_subspace=subspaces[1]._to_c(),
                               xparity=isinstance(subspaces[0], XParity),
                               shell=self.shell, gpu=config.gpu)
        if self.shell and self.precompute_diagonal and subspaces[0] == subspaces[1] and masks[0] == 0:
            bpetsc.precompute_diagonal(mat)
        old = self._mats.pop(subspaces, None)
        if old is not None:
            old.destroy()
        self._mats[subspaces] = mat

    def destroy_mat(self, subspaces=None):
        keys = [subspaces] if subspaces is not None else list(self._mats)
        for k in keys:
            mat = self._mats.pop(k, None)
            if mat is not None:
                mat.destroy()

    def infinity_norm(self, subspaces=None):
        from .petsc import NormType
        return self.get_mat(subspaces=subspaces).norm(NormType.INFINITY)

    def to_numpy(self, subspaces=None, sparse=True):
        """Host matrix straight from the MSC definition (for checks; not used by the path)."""
        self.establish_L()
        if subspaces is None:
            subspaces = (self.left_subspace, self.right_subspace)
        self.reduce_msc()
        left, right = subspaces
        msc = self.msc if left.product_state_basis else left.reduce_msc(self.msc)
        return msc_tools.msc_to_numpy(msc, (left.get_dimension(), right.get_dimension()),
                                      left.idx_to_state, right.state_to_idx, sparse=sparse)

    # ---- the hot path entry points -------------------------------------------------------
    def dot(self, x, result=None):
        """y = A x (reference ``operators.py:1063-1108``)."""
        from .states import State
        x.assert_initialized()
        self.establish_L()
        right = x.subspace
        matches = [(l, r) for l, r in self.get_subspace_list() if r.identical(right)]
        if not matches:
            raise ValueError('No operator subspace found that matches input vector subspace. '
                             'Try adding the subspace with the Operator.add_subspace method.')
        if result is None:
            if len(matches) != 1:
                raise ValueError('Ambiguous subspace for result vector. Pass a state with the desired '
                                 'subspace as the "result" option to Operator.dot.')
            left = matches[0][0]
            result = State(L=left.L, subspace=left)
        else:
            left = result.subspace
        pair = next(((l, r) for l, r in matches if l.identical(left)), None)
        if pair is None:
            raise ValueError('Subspaces of matrix and result vector do not match.')
        self.get_mat(subspaces=pair).mult(x.vec, result.vec)
        result.set_initialized()
        return result

    def create_states(self):
        """(bra, ket): uninitialised states in this operator's left and right subspaces
        (reference ``operators.py:760-773``)."""
        from .states import State
        self.establish_L()
        return State(subspace=self.left_subspace), State(subspace=self.right_subspace)

    def expectation(self, state, tmp_state=None):
        """<state| A |state> as a float (operators here are Hermitian, so the imaginary part is
        dropped; reference ``operators.py:775-796``).  The product and the inner product run on the
        device -- only the final scalar crosses to the host -- so chains such as
        ``A.expectation(H.evolve(psi, t))`` (the OTOC loops of examples/scripts/SYK/run_syk.py) never
        move a state vector off the GPU.  ``tmp_state`` is optional scratch in the left subspace."""
        if tmp_state is None:
            tmp_state = self.dot(state)
        else:
            self.dot(state, result=tmp_state)
        return state.dot(tmp_state).real

    def evolve(self, state, t, **kwargs):
        from .computations import evolve
        return evolve(self, state, t, **kwargs)

    def eigsolve(self, **kwargs):
        from .computations import eigsolve
        return eigsolve(self, **kwargs)

    # ---- algebra on the host -------------------------------------------------------------
    def get_shifted_msc(self, shift, wrap_idx=None):
        return msc_tools.shift(self._msc, shift, wrap_idx)

    def scale(self, x):
        """in-place scalar multiplication"""
        self._msc = self._msc.copy()
        self._msc['coeffs'] *= x
        self.destroy_mat()
        return self

    def _like(self, msc):
        rtn = Operator(msc=msc)
        L = self._L
        if L is not None and rtn.max_spin_idx < L:
            rtn.L = L
        return rtn

    def __add__(self, other):
        if not isinstance(other, Operator):
            other = other * identity()
        return self._like(msc_tools.msc_sum([self._msc, other._msc]))

    __radd__ = __add__

    def __sub__(self, other):
        return self + (-1) * other

    def __neg__(self):
        return (-1) * self

    def __mul__(self, other):
        if isinstance(other, Operator):
            return self._like(msc_tools.msc_product([self._msc, other._msc]))
        from .states import State
        if isinstance(other, State):
            return self.dot(other)
        return self.__rmul__(other)

    def __rmul__(self, scalar):
        msc = self._msc.copy()
        msc['coeffs'] *= scalar
        return self._like(msc)

    def __eq__(self, other):
        if not isinstance(other, Operator):
            return NotImplemented
        self.reduce_msc()
        other.reduce_msc()
        return self._msc.size == other._msc.size and bool(np.all(self._msc == other._msc))

    __hash__ = None

    def __repr__(self):
        return self._name or f'Operator(<{self._msc.size} terms>)'

    def __len__(self):
        return self._msc.size


def sigmax(i=0):
    r""":math:`\sigma^x_i` = flip bit i."""
    i = validate.spin_index(i)
    return Operator(msc=[(1 << i, 0, 1)], name=f'sigmax({i})')


def sigmay(i=0):
    r""":math:`\sigma^y_i` = i * X_i Z_i."""
    i = validate.spin_index(i)
    return Operator(msc=[(1 << i, 1 << i, 1j)], name=f'sigmay({i})')


def sigmaz(i=0):
    r""":math:`\sigma^z_i` = sign of bit i."""
    i = validate.spin_index(i)
    return Operator(msc=[(0, 1 << i, 1)], name=f'sigmaz({i})')


def sigma_plus(i=0):
    return sigmax(i) + 1j * sigmay(i)


def sigma_minus(i=0):
    return sigmax(i) - 1j * sigmay(i)


def identity():
    return Operator(msc=[(0, 0, 1)], name='identity()')


def zero():
    return Operator(msc=[], name='zero()')


def op_sum(terms, nshow=3):
    return Operator(msc=msc_tools.msc_sum([t.msc for t in terms]))


def op_product(terms):
    mscs = [t.msc for t in terms]
    if not mscs:
        return identity()
    return Operator(msc=msc_tools.msc_product(mscs))


def index_sum(op, size=None, start=0, boundary='open'):
    """Translate ``op`` along the chain and add the copies (reference ``operators.py:1253-1331``)."""
    if size is None:
        if op.L is None:
            raise ValueError('Must specify index_sum size with either the "size" argument '
                             'or by setting Operator.L (possibly through config.L).')
        size = op.L
    size = validate.L(size)
    if boundary == 'open':
        stop = start + size - op.max_spin_idx
        if stop <= start:
            raise ValueError("requested size %d for sum operator's support smaller than "
                             "summand's support %d; impossible to satisfy" % (size, op.max_spin_idx))
        wrap = None
    elif boundary == 'closed':
        if start != 0:
            raise ValueError('cannot set start != 0 for closed boundary conditions.')
        stop = start + size
        wrap = stop
    else:
        raise ValueError("invalid value for argument 'boundary' (can be 'open' or 'closed')")
    return Operator(msc=msc_tools.msc_sum([op.get_shifted_msc(i, wrap) for i in range(start, stop)]))


def index_product(op, size=None, start=0):
    if size is None:
        if op.L is None:
            raise ValueError('Must specify index_product size with either the "size" argument '
                             'or by setting Operator.L (possibly through config.L).')
        size = op.L
    if size == 0:
        return identity()
    size = validate.L(size)
    stop = start + size - op.max_spin_idx
    return Operator(msc=msc_tools.msc_product([op.get_shifted_msc(i) for i in range(start, stop)]))

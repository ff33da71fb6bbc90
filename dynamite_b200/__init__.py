"""
dynamite_b200 -- B200-native backend for dynamite's matrix-free MSC shell path.

The package mirrors the part of dynamite's Python interface that sits on the
hot path (``Operator.dot / evolve / eigsolve``, ``State``, the subspace
classes, ``computations``) on top of a thin C ABI (``include/dynamite_b200.h``,
``libdynamite_b200.so``: hand-written sm_100a CUDA).  There is no CPU
fallback: computing anything needs the built library and a B200.

``config`` mirrors ``dynamite.config`` (reference ``__init__.py:12-227``) for
the options that matter to this path: ``L``, ``shell``, ``subspace``, ``gpu``.
"""

__version__ = '0.1.0'


class _Config:
    """Package-wide configuration (reference ``dynamite/__init__.py:12-227``)."""

    def __init__(self):
        self.initialized = False
        self._L = None
        self._shell = True      # this backend only has shell matrices
        self._subspace = None
        self._gpu = True
        self._device = None

    def initialize(self, slepc_args=None, version_check=True, gpu=None, device=None):
        """Bind the process to its GPU.  ``slepc_args`` is accepted for source
        compatibility and ignored (there is no PETSc/SLEPc underneath)."""
        if self.initialized:
            raise RuntimeError('dynamite_b200.config.initialize() can only be called once.')
        self._initialize(slepc_args, version_check, gpu, device)

    def _initialize(self, slepc_args=None, version_check=True, gpu=None, device=None):
        if self.initialized:
            return
        if gpu is False:
            raise RuntimeError('dynamite_b200 has no CPU path: gpu=False is not available.')
        from . import _capi
        self._device = _capi.ensure_gpu(device)
        self.initialized = True

    @property
    def L(self):
        return self._L

    @L.setter
    def L(self, value):
        if value is not None:
            from . import validate
            value = validate.L(value)
        self._L = value

    @property
    def shell(self):
        return self._shell

    @shell.setter
    def shell(self, value):
        if not value:
            raise ValueError('dynamite_b200 only provides shell (matrix-free) matrices.')
        self._shell = True

    @property
    def subspace(self):
        if self._subspace is None:
            from .subspaces import Full
            self._subspace = Full()
        return self._subspace

    @subspace.setter
    def subspace(self, value):
        self._subspace = value

    @property
    def gpu(self):
        return self._gpu


config = _Config()

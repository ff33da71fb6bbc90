"""
Module substitution that binds an UNMODIFIED dynamite checkout to this backend
(INTEGRATION.md section 3): ``install()`` registers stand-ins for
``petsc4py``, ``slepc4py`` and ``dynamite._backend.*`` in ``sys.modules``;
``import dynamite`` afterwards picks them up.
"""
import sys
import types


def install():
    import dynamite_b200
    from . import _backend, petsc, slepc
    from ._backend import bbuild, bpetsc, bsubspace

    p4 = types.ModuleType('petsc4py')
    p4.PETSc = petsc
    p4.init = lambda *a, **k: None
    s4 = types.ModuleType('slepc4py')
    s4.SLEPc = slepc
    # dynamite calls slepc4py.init(args) exactly once, from config._initialize (__init__.py:141)
    s4.init = lambda *a, **k: dynamite_b200.config._initialize()
    sys.modules.update({
        'petsc4py': p4, 'petsc4py.PETSc': petsc, 'slepc4py': s4, 'slepc4py.SLEPc': slepc,
        'dynamite._backend': _backend, 'dynamite._backend.bbuild': bbuild,
        'dynamite._backend.bsubspace': bsubspace, 'dynamite._backend.bpetsc': bpetsc,
    })

"""Build-time facts of the backend (reference ``_backend/bbuild.pyx:7-44``)."""
import numpy as np

from .. import __version__, _capi

# int64 indices and complex128 scalars are fixed in this backend
dnm_int_t = np.int64


def complex_enabled():
    return True


def have_gpu_shell():
    """True when the CUDA library is built (``bbuild.pyx`` have_gpu_shell)."""
    try:
        _capi.lib()
        return True
    except ImportError:
        return False


def petsc_initialized():
    return bool(_capi.lib().dnm_have_gpu())


def get_build_version():
    return __version__


def get_build_commit():
    return 'unknown'


def get_build_branch():
    return 'unknown'

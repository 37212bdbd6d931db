"""porespy_b200 -- B200-native (sm_100a) implementation of PoreSpy's hot path:
exact 3-D EDT (`edt.edt`) -> `filters.local_thickness` / `filters.porosimetry`
(+ `filters.trim_disconnected_blobs`), behind the reference's own signatures.

    import porespy_b200 as psb
    lt = psb.filters.local_thickness(im, sizes=25)          # same array PoreSpy returns
    psb.install()                                           # or: accelerate an unmodified PoreSpy
"""
from . import _lib
from . import edt as edt_module
from . import filters
from . import generators
from . import metrics
from . import simulations
from . import sizemap
from .edt import edt, edtsq
from .filters import local_thickness, porosimetry, trim_disconnected_blobs
from .sizemap import IndexMap, local_thickness_index, porosimetry_index
from .patch import install, uninstall
from ._device import pinned_empty, to_pinned

__version__ = "0.1.0"
__all__ = ["edt", "edtsq", "filters", "generators", "metrics", "simulations", "sizemap", "IndexMap", "local_thickness_index",
           "porosimetry_index", "local_thickness", "porosimetry",
           "trim_disconnected_blobs", "install", "uninstall", "build", "pinned_empty", "to_pinned"]


def build(force=False, verbose=False):
    """Compile the CUDA library in-tree (nvcc, sm_100a)."""
    return _lib.build(force=force, verbose=verbose)

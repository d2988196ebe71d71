"""sequitr_b200 -- B200-native implementation of Sequitr's data-parallel hot path.

UNet segmentation inference (``networks.unet``), loss-weight maps (``pipeline``,
``weightmap``) and the label-and-localise post-process (``utils.CentroidWriter``)
behind the reference's own Python entry points, executed by hand-written sm_100a
CUDA kernels through the C ABI in ``include/sequitr_b200.h``.  There is no CPU
fallback: without the built library and a B200 every compute call raises.
"""
from . import _lib

__version__ = '0.1'


def require_gpu(device=None):
    """Load the CUDA library and create the handle for ``device`` (raises loudly)."""
    _lib.load()
    return _lib.handle(device)


def library_path():
    return _lib.LIB_PATH

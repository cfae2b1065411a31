"""cacaoengine_b200 -- B200-native libcacaoimage / cubetool pixel hot path.

Python mirror of the reference's image interface (libs/image/include/libcacaoimage.hpp) over the
C ABI in include/cacao_img.h.  The compute lives in hand-written sm_100a CUDA kernels inside
lib/libcacaoimage_b200.so; this package only marshals buffers.  No CPU fallback exists.
"""
from ._capi import (CacaoError, FMT_F16, FMT_F32, FMT_U8, FMT_U16, MEM_DEVICE, MEM_HOST, MODE_FIXED, MODE_REF_EXACT, lib)  # noqa: F401
from .image import (ChangeChannelLayout, Convert, Convert16To8BitColor, EquirectToCubemap, Flip, Image, ImageEx, Layout, ToRGB8)  # noqa: F401
from . import device  # noqa: F401

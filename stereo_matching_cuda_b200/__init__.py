"""stereo_matching_cuda_b200 -- B200-native local stereo pipeline behind a C ABI.

The product is `libstereo_b200.so` (host C++ driving hand-written sm_100a CUDA kernels,
include/stereo_b200.h).  This package is the thin host-side mirror of the reference's
stage functions (rgb_to_grayscale, compute_cost, integral, filter, compute_guided_filter,
winner_take_all/dispSelect, detect_occlusion, fill_occlusion) plus the fused pipeline.
There is no CPU path: importing works anywhere, calling anything needs a B200.
"""
from .api import (  # noqa: F401
    BOX_SAT,
    BOX_SLIDING,
    GUIDE_GRAY,
    GUIDE_RGB,
    Context,
    Params,
    StereoB200Error,
    lib_path,
    load_library,
)

__all__ = ["Context", "Params", "StereoB200Error", "load_library", "lib_path", "BOX_SAT", "BOX_SLIDING", "GUIDE_GRAY",
           "GUIDE_RGB"]

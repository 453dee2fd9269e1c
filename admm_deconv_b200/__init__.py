"""admm_deconv_b200 -- B200-native (sm_100a) drop-in for the unrolled ADMM-TV deconvolution layer of
georgegrosu1/admm-deconv (src/layers/deconv_admm.jl + src/ops/ops.jl).

Only this one hot path is implemented.  The arithmetic lives in
``libadmmtv.so`` (C ABI: include/admmtv.h); this package is the host-side mirror of the
reference's operator / layer interface.  Importing the package does not need a GPU; calling it
does -- there is no CPU fallback.
"""
from ._lib import AdmmTvError, AdmmTvLib, Desc, load, make_desc  # noqa: F401
from . import dist, host, staging  # noqa: F401
from .layers import ADMMDeconv, ADMMDeconvF1, ADMMDeconvF2, ADMMDeconvF3, ADMMParallel, Admm  # noqa: F401
from .losses import gmsd, gmsd_loss, ssim, ssim_loss, ssim_loss_fast  # noqa: F401
from .ops import (admm_layer_call, from_julia, to_julia, tvd_fft, tvd_fft_gpu, tvd_fft_grouped,  # noqa: F401
                  tvd_fft_host)

__all__ = [
    "ADMMDeconv", "ADMMDeconvF1", "ADMMDeconvF2", "ADMMDeconvF3", "ADMMParallel", "Admm",
    "tvd_fft", "tvd_fft_gpu", "tvd_fft_host", "tvd_fft_grouped", "admm_layer_call", "to_julia", "from_julia",
    "gmsd", "gmsd_loss", "ssim", "ssim_loss", "ssim_loss_fast",
    "load", "make_desc", "Desc", "AdmmTvLib", "AdmmTvError",
]

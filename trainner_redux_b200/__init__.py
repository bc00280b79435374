"""trainner_redux_b200 — B200 (sm_100a) implementation of traiNNer-redux's on-the-fly
second-order degradation path, behind the reference's own Python API.

Modules mirror the reference files they replace:
    img_process_util  <- traiNNer/utils/img_process_util.py   (filter2d, USMSharp)
    diffjpeg          <- traiNNer/utils/diffjpeg.py           (DiffJPEG)
    degradations      <- traiNNer/data/degradations.py        (*_pt noise, resize_pt)
    transforms        <- traiNNer/data/transforms.py          (paired_random_crop)
    realesrgan_feed   <- traiNNer/models/realesrgan_model.py  (feed_data, pair pool)
The kernels live in csrc/ and are reached through the C ABI in include/otf_b200.h.
"""

from .degradations import (  # noqa: F401
    PhiloxState,
    add_gaussian_noise_pt,
    add_poisson_noise_pt,
    generate_gaussian_noise_pt,
    generate_poisson_noise_pt,
    random_add_gaussian_noise_pt,
    random_add_poisson_noise_pt,
    random_generate_gaussian_noise_pt,
    random_generate_poisson_noise_pt,
    resize_pt,
)
from .diffjpeg import DiffJPEG  # noqa: F401
from .img_process_util import USMSharp, filter2d  # noqa: F401
from .transforms import paired_random_crop  # noqa: F401

__all__ = [
    "DiffJPEG", "PhiloxState", "USMSharp", "filter2d", "paired_random_crop", "resize_pt",
    "add_gaussian_noise_pt", "add_poisson_noise_pt", "generate_gaussian_noise_pt", "generate_poisson_noise_pt",
    "random_add_gaussian_noise_pt", "random_add_poisson_noise_pt", "random_generate_gaussian_noise_pt",
    "random_generate_poisson_noise_pt",
]

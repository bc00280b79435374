"""ctypes binding of libotf_b200.so (the C ABI declared in include/otf_b200.h).

PyTorch is used only as the owner of device memory and streams: every call below
passes raw device pointers, extents and the current CUDA stream handle.  There is
no CPU fallback — a missing library or a non-CUDA tensor raises.
"""

from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Any

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OTF_LIB_PATH") or os.path.join(_HERE, "libotf_b200.so")  # override: A/B builds in profiles/experiments
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "otf_b200.h")

OTF_OK = 0
ABI_VERSION = 6  # 6: otf_f32_to_u8; 5: otf_libjpeg_roundtrip_f32 + OP_LIBJPEG; 4: prefetcher upload step (otf_upload_async + events); 3: OtfStage.f2 + the fork-extra ops of the stage executor; 2: device-side Philox offset word / crop offsets, OtfStage.p4
RESIZE_BILINEAR_AA, RESIZE_BICUBIC_AA, RESIZE_AREA, RESIZE_NEAREST_EXACT, RESIZE_BICUBIC, RESIZE_NEAREST, RESIZE_LANCZOS = range(7)
WARP_LENS, WARP_SHUTTER, WARP_CHROMA = range(3)
TAPS_NONE, TAPS_OVERSHARPEN = 0, 1
NOISE_CLIP, NOISE_ROUNDS, NOISE_FIELD_ONLY, NOISE_RAW_FIELD = 1, 2, 4, 16

_p, _i, _i64, _u64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float

# name -> (restype, argtypes); mirrors include/otf_b200.h one to one
SIGNATURES: dict[str, tuple[Any, list[Any]]] = {
    "otf_abi_version": (_i, []),
    "otf_last_error": (C.c_char_p, []),
    "otf_device_cc": (_i, []),
    "otf_filter2d_scratch_words": (_i64, [_i]),
    "otf_filter2d_analyse_f32": (_i, [_p, _i, _i, _i, _p, _p]),
    "otf_filter2d_f32": (_i, [_p, _i, _i, _i, _i, _p, _i, _i, _p, _i, _p, _p]),
    "otf_sepconv_reflect_f32": (_i, [_p, _i, _i, _i, _p, _i, _i, _p, _p]),
    "otf_usm_workspace_bytes": (_i64, [_i, _i, _i]),
    "otf_usm_launch_count": (_i, [_i, _i, _i]),
    "otf_usm_sharp_f32": (_i, [_p, _i, _i, _i, _p, _i, _f, _f, _p, _i64, _p, _p]),
    "otf_resize_workspace_bytes": (_i64, [_i, _i, _i, _i, _i]),
    "otf_resize_tables_f32": (_i, [_i, _i, _i, _i, _i, _p, _i64, _p]),
    "otf_resize_f32": (_i, [_p, _i, _i, _i, _p, _i, _i, _i, _i, _p, _i64, _i, _p]),
    "otf_resize_gauss_f32": (_i, [_p, _i, _i, _i, _i, _p, _i, _i, _i, _i, _p, _i64, _i, _p, _p, _u64, _u64, _p, _i, _p]),
    "otf_gaussian_noise_f32": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _p, _u64, _u64, _p, _i, _p, _p]),
    "otf_philox_normal_f32": (_i, [_p, _i64, _u64, _u64, _p]),
    "otf_philox_uniform_f32": (_i, [_p, _i64, _u64, _u64, _p]),
    "otf_poisson_tables_bytes": (_i64, []),
    "otf_poisson_build_tables": (_i, [_p, _p]),
    "otf_poisson_noise_f32": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _p, _u64, _u64, _p, _i, _p, _p, _p, _p, _p, _p, _p]),
    "otf_philox_poisson_f32": (_i, [_p, _p, _i64, _u64, _u64, _p]),
    "otf_quality_to_factor_f32": (_i, [_p, _i, _p]),
    "otf_diffjpeg_f32": (_i, [_p, _i, _i, _i, _p, _f, _i, _i, _i, _i, _p, _p]),
    "otf_diffjpeg_crop_pair_f32": (_i, [_p, _i, _i, _i, _p, _f, _i, _i, _i, _p, _i, _i, _i, _i, _p, _i, _i, _p, _p, _p]),
    "otf_clamp_round_f32": (_i, [_p, _i64, _p, _p]),
    "otf_crop_pair_f32": (_i, [_p, _i, _i, _i, _p, _i, _i, _i, _i, _p, _i, _i, _i, _p, _p, _p]),
    "otf_u8_to_f32": (_i, [_p, _i64, _p, _p]),
    "otf_f32_to_u8": (_i, [_p, _i64, _p, _p]),
    "otf_libjpeg_workspace_bytes": (_i64, [_i, _i, _i]),
    "otf_libjpeg_roundtrip_f32": (_i, [_p, _i, _i, _i, _i, _p, _i64, _p, _p]),
    "otf_event_create": (_i, [_p]),
    "otf_event_destroy": (_i, [_p]),
    "otf_event_query": (_i, [_p, _p]),
    "otf_stream_wait_event": (_i, [_p, _p]),
    "otf_upload_async": (_i, [_i, _p, _p, _p, _p, _p, _p, _p]),
    "otf_download_async": (_i, [_p, _p, _u64, _p, _p, _p, _p]),
    "otf_event_synchronize": (_i, [_p]),
    "otf_copy_strided_f32": (_i, [_p, _p, _i, _i, _i, _i, _p, _p]),
    "otf_synth_kernels_f32": (_i, [_p, _i, _p, _p]),
    "otf_gather_slots_f32": (_i, [_p, _p, _i, _i64, _p, _p]),
    "otf_scatter_slots_f32": (_i, [_p, _p, _i, _i64, _p, _p]),
    "otf_pool_exchange_f32": (_i, [_p, _p, _p, _i, _i64, _i64, _p, _p, _p, _p, _p]),
    "otf_mixup_f32": (_i, [_p, _p, _i, _i64, _f, _f, _p, _p]),
    "otf_copy_box_f32": (_i, [_p, _i, _i, _i, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p]),
    "otf_warp_f32": (_i, [_p, _i, _i, _i, _i, _i, _f, _p, _p]),
    "otf_taps_zero_f32": (_i, [_p, _i, _i, _i, _i, _p, _i, _f, _p, _p]),
    "otf_channel_gain_f32": (_i, [_p, _i, _i, _i64, _f, _f, _f, _i, _p, _p]),
    "otf_sensor_noise_f32": (_i, [_p, _i64, _f, _p, _u64, _u64, _p, _p]),
    "otf_trunc8_f32": (_i, [_p, _i64, _p, _p]),
    "otf_demosaic_f32": (_i, [_p, _i, _i, _i, _p, _p]),
    "otf_run_stages_workspace_bytes": (_i64, [_i, _i, _i, _i, _p, _i]),
    "otf_run_stages_f32": (_i, [_p, _i, _i, _i, _i, _p, _i, _p, _i64, _p, _p, _p]),
    "otf_run_stages_launches": (_i, []),
}

(OP_ANALYSE, OP_FILTER2D, OP_USM, OP_SEPCONV, OP_RESIZE, OP_GAUSS, OP_POISSON, OP_JPEG, OP_CLAMP_ROUND,
 OP_CROP_PAIR, OP_WARP, OP_TAPS_ZERO, OP_GAIN, OP_SENSOR, OP_DEMOSAIC, OP_TRUNC8, OP_LIBJPEG) = range(17)


class Stage(C.Structure):
    """``OtfStage`` of include/otf_b200.h (one step of otf_run_stages_f32)."""

    _fields_ = [("op", C.c_int32), ("mode", C.c_int32), ("oh", C.c_int32), ("ow", C.c_int32), ("n", C.c_int32),
                ("kb", C.c_int32), ("K", C.c_int32), ("flags", C.c_int32), ("f0", C.c_float), ("f1", C.c_float),
                ("seed", C.c_uint64), ("offset", C.c_uint64), ("p0", C.c_void_p), ("p1", C.c_void_p), ("p2", C.c_void_p),
                ("p3", C.c_void_p), ("dst", C.c_void_p), ("p4", C.c_void_p), ("f2", C.c_float), ("reserved", C.c_int32)]


_lib: C.CDLL | None = None
_lock = threading.Lock()
launch_count = 0  # kernels enqueued through this binding (bench.py reports it)

# launches behind each entry point (see the .cu files)
_LAUNCHES = {
    "otf_filter2d_analyse_f32": 1,  # analysis + launch order, one launch
    "otf_filter2d_f32": 2,  # kernel analysis + blocked kernel (callers pass the exact count)
    "otf_sepconv_reflect_f32": 1,
    "otf_usm_sharp_f32": 4,
    "otf_resize_tables_f32": 1,
    "otf_resize_f32": 2,  # weight tables + resampler
    "otf_resize_gauss_f32": 2,
    "otf_gaussian_noise_f32": 1,
    "otf_philox_normal_f32": 1,
    "otf_philox_uniform_f32": 1,
    "otf_poisson_noise_f32": 2,
    "otf_poisson_build_tables": 1,
    "otf_philox_poisson_f32": 1,
    "otf_quality_to_factor_f32": 1,
    "otf_diffjpeg_f32": 1,
    "otf_diffjpeg_crop_pair_f32": 1,
    "otf_clamp_round_f32": 1,
    "otf_crop_pair_f32": 1,
    "otf_u8_to_f32": 1,
    "otf_copy_strided_f32": 1,
    "otf_synth_kernels_f32": 1,
    "otf_gather_slots_f32": 1,
    "otf_scatter_slots_f32": 1,
    "otf_pool_exchange_f32": 1,
    "otf_mixup_f32": 1,
    "otf_copy_box_f32": 1,
    "otf_warp_f32": 1,
    "otf_taps_zero_f32": 1,
    "otf_channel_gain_f32": 1,
    "otf_sensor_noise_f32": 1,
    "otf_trunc8_f32": 1,
    "otf_demosaic_f32": 1,
    "otf_libjpeg_roundtrip_f32": 2,
    "otf_f32_to_u8": 1,
}


class OtfError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the shared library (once). Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "or `make -C trainner_redux_b200/csrc` (nvcc, sm_100a). There is no CPU fallback."
                )
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            if lib.otf_abi_version() != ABI_VERSION:
                raise ImportError("libotf_b200.so ABI version mismatch")
            _lib = lib
    return _lib


def last_error() -> str:
    return load().otf_last_error().decode("utf-8", "replace")


def call(name: str, *args: Any, launches: int | None = None) -> None:
    """Invoke an int-returning entry point; raise OtfError with the library's message.
    `launches` overrides the table above when the number of kernels depends on the arguments."""
    global launch_count
    rc = getattr(load(), name)(*args)
    if rc != OTF_OK:
        raise OtfError(f"{name} failed ({rc}): {last_error()}")
    launch_count += _LAUNCHES.get(name, 0) if launches is None else launches


def require_cuda(*tensors: torch.Tensor | None) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "trainner_redux_b200 runs on CUDA tensors only (sm_100a kernels, no CPU fallback); "
                f"got a tensor on {t.device}"
            )


def ptr(t: torch.Tensor | None) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


try:  # the raw handle without building a torch.cuda.Stream object (~10x cheaper; this runs once per launch)
    _raw_stream = torch._C._cuda_getCurrentRawStream
    _cur_device = torch._C._cuda_getDevice
except AttributeError:  # pragma: no cover - other torch builds
    _raw_stream = _cur_device = None


def stream() -> C.c_void_p:
    if _raw_stream is not None:
        return C.c_void_p(_raw_stream(_cur_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dense_f32(t: torch.Tensor) -> torch.Tensor:
    """Dense NCHW fp32 view of a 4-D CUDA tensor; strided inputs (channels_last, slices)
    are normalised by the library's own copy kernel, not by ATen."""
    require_cuda(t)
    if t.dtype != torch.float32:
        raise TypeError(f"fp32 only, got {t.dtype}")
    if t.is_contiguous():
        return t
    if t.dim() != 4:
        return t.contiguous()
    b, c, h, w = t.shape
    out = torch.empty((b, c, h, w), dtype=torch.float32, device=t.device)
    strides = (C.c_int64 * 4)(*t.stride())
    call("otf_copy_strided_f32", ptr(t), strides, b, c, h, w, ptr(out), stream())
    return out

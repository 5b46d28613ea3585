"""ctypes binding of libsqdet_b200.so (include/sqdet_b200.h).

There is NO CPU fallback: if the shared library is missing, or a tensor is not on a CUDA
device, the call raises.  torch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# SQD_LIB_PATH: developer override (the profiling build libsqdet_b200_trace.so made with SQD_BUILD_TRACE=1)
LIB_PATH = os.environ.get("SQD_LIB_PATH") or os.path.join(_HERE, "libsqdet_b200.so")

LAYOUT_NCHW, LAYOUT_NHWC, LAYOUT_SPLIT_NHWC = 0, 1, 2
CONV_TCGEN05_F16X3, CONV_SIMT_FP32, CONV_TCGEN05_F16X3_1CTA = 0, 1, 2

_lib = None
_lock = threading.Lock()

_vp, _i, _d, _sz = C.c_void_p, C.c_int, C.c_double, C.c_size_t

# name -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "sqd_abi_version": (_i, []),
    "sqd_last_error": (C.c_char_p, []),
    "sqd_set_option": (_i, [C.c_char_p, _i]),
    "sqd_get_option": (_i, [C.c_char_p, C.POINTER(C.c_int)]),
    "sqd_convdet_packed_weight_bytes": (_sz, [_i, _i]),
    "sqd_convdet_pack_weights": (_i, [_vp, _i, _i, _vp, _vp]),
    "sqd_convdet_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i, _i]),
    "sqd_convdet_forward": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _sz, _i, _vp]),
    "sqd_convdet_status": (_i, [_vp, _vp]),
    "sqd_convdet_split_bytes": (_sz, [_i, _i, _i, _i]),
    "sqd_convdet_split_features": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "sqd_decode_scores": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sqd_topk_nms": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _d, _d, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sqd_detect_workspace_bytes": (_sz, [_i, _i]),
    "sqd_detect_from_pred": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _d, _d, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "sqd_head_detect_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i, _i]),
    "sqd_head_detect_status_offset": (_sz, [_i, _i, _i, _i]),
    "sqd_head_detect_fused": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _d, _d,
                                   _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "sqd_head_detect_profile_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "sqd_head_detect_profile": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _d, _d,
                                     _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, C.POINTER(C.c_float)]),
    "sqd_head_detect_host_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i, _i, _i, _i]),
    "sqd_head_detect_host_result_layout": (_i, [_i, _i, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "sqd_head_detect_host": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _d, _d,
                                  _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _vp, _vp, _i]),
    "sqd_match_anchors": (_i, [_vp, _vp, _i, _i, _vp, _i, _vp, _vp, _vp]),
    "sqd_build_targets": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "sqd_loss_workspace_bytes": (_sz, [_i, _i]),
    "sqd_loss_fwd_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, C.POINTER(C.c_float), _vp, _vp, _vp, _vp, _sz, _vp]),
    "sqd_boxes_postprocess": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "sqd_convdet_dgrad_packed_bytes": (_sz, [_i, _i]),
    "sqd_convdet_dgrad_pack_weights": (_i, [_vp, _i, _i, _vp, _vp]),
    "sqd_convdet_dgrad_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "sqd_convdet_dgrad": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "sqd_convdet_wgrad_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "sqd_convdet_wgrad": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "sqd_convdet_wgrad_tc_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "sqd_convdet_wgrad_tc": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "sqd_convdet_wgrad_tc_status": (_i, [_vp, _vp]),
    "sqd_convdet_bias_grad": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "sqd_pack_results": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "sqd_format_kitti": (C.c_longlong, [_vp, _vp, _i, _i, C.POINTER(C.c_char_p), _i, _vp, _sz, _vp]),
    "sqd_preprocess": (_i, [_vp, _i, _i, _i, _i, C.POINTER(C.c_float), C.POINTER(C.c_float), _i, _i, _vp, _vp]),
}


class SqdError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built: the product path has no
    fallback -- run `python -c "import __graft_entry__ as g; g.build()"` first."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise SqdError(f"{LIB_PATH} is missing: build the CUDA library first (there is no CPU fallback)")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().sqd_last_error().decode("utf-8", "replace")
        raise SqdError(f"{what} failed (code {rc}): {msg}")


class option:
    """Context manager: run a block with a developer option of the library set (sqd_set_option), then restore it.
    `with _lib.option("SQD_SPLIT_TWO_PASS", 1): ...` -- how the tests select an alternative route."""

    def __init__(self, name, value):
        self.name, self.value = name.encode(), int(value)

    def __enter__(self):
        lib = load()
        old = C.c_int(0)
        check(lib.sqd_get_option(self.name, C.byref(old)), "sqd_get_option")
        self.old = old.value
        check(lib.sqd_set_option(self.name, self.value), "sqd_set_option")
        return self

    def __exit__(self, *exc):
        check(load().sqd_set_option(self.name, self.old), "sqd_set_option")
        return False


def ptr(t):
    """Raw device pointer of a CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise SqdError("sqdet_b200 kernels need CUDA tensors (no CPU fallback)")
    if not t.is_contiguous():
        raise SqdError("sqdet_b200 kernels need contiguous tensors")
    return C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Workspace:
    """Grow-only per-(device, tag) scratch buffers, so steady-state calls never allocate."""

    def __init__(self):
        self._bufs = {}

    def get(self, tag, nbytes, device):
        key = (tag, torch.device(device).index)
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self._bufs[key] = buf
        return buf

    def clear(self):
        """Drop every scratch buffer (after a one-off large call; the next call re-allocates what it needs)."""
        self._bufs.clear()


_tls = threading.local()


def workspace() -> Workspace:
    """One Workspace per host thread (DataParallel calls forward from one thread per GPU)."""
    ws = getattr(_tls, "ws", None)
    if ws is None:
        ws = _tls.ws = Workspace()
    return ws

"""ctypes binding of libfacerec_b200.so (the C ABI declared in include/facerec_b200.h).

There is no CPU fallback: if the shared library is missing or no sm_100 GPU is present the
import of a context fails loudly.  The library is built in-tree by facerec_b200/build.py.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfacerec_b200.so")

FRB_RETINA_MNET025 = 0
FRB_RETINA_R50 = 1
FRB_IRV1 = 2

FRB_CONDENSED_F32 = 0
FRB_CONDENSED_F64 = 1
FRB_SQUARE_F32 = 2
FRB_SQUARE_SYM_F32 = 3

FRB_LAYOUT_NHWC8 = 0
FRB_LAYOUT_S2D16 = 1

FRB_OPT_BF16 = 0
FRB_OPT_BN_CAP = 1
FRB_OPT_FORCE_GATHER = 2
FRB_OPT_HALO = 3
FRB_OPT_TRACE = 4
FRB_OPT_KBPS2 = 5
FRB_OPT_LAYER_TIMES = 6
FRB_OPT_MAX_STAGES = 7
FRB_OPT_SKIP_LOADS = 8
FRB_OPT_DUAL_ISSUE = 9
FRB_OPT_PDL = 10
FRB_OPT_WEIGHT_SLICE = 11
FRB_OPT_MPAIR = 12
FRB_OPT_GRAPHS = 13
FRB_OPT_CTA2 = 14
FRB_OPT_FUSE_TAIL = 15
FRB_OPT_FUSE_DWPW = 16
FRB_OPT_FUSE_POOL = 17
FRB_OPT_FUSE_TAIL2 = 18

_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t

# name -> (restype, argtypes); mirrors include/facerec_b200.h one to one
SIGNATURES = {
    "frb_version": (_i, []),
    "frb_init": (_i, [_i, C.POINTER(_vp)]),
    "frb_destroy": (_i, [_vp]),
    "frb_last_error": (C.c_char_p, [_vp]),
    "frb_set_option": (_i, [_vp, _i, _i]),
    "frb_launch_count": (C.c_longlong, [_vp]),
    "frb_debug_trace": (_i, [_vp, _vp, _i]),
    "frb_debug_layer_times": (_i, [_vp, _vp, _i]),
    "frb_load_model": (_i, [_vp, _i, _vp, _sz]),
    "frb_input_layout": (_i, [_i]),
    "frb_input_elems": (_sz, [_i, _i, _i, _i]),
    "frb_preprocess_u8": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "frb_retina_forward": (_i, [_vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "frb_num_priors": (_i, [_i, _i]),
    "frb_prior_decode_filter": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _i, _vp, _vp, _vp, _vp]),
    "frb_nms": (_i, [_vp, _vp, _vp, _vp, _i, _i, _f, _i, _vp, _vp, _vp]),
    "frb_gather_kept": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "frb_crop_resize_std": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp]),
    "frb_irv1_forward": (_i, [_vp, _vp, _i, _vp, _vp]),
    "frb_conv2d": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _i, _i, _i,
                        _vp, _i, _i, C.POINTER(_f), _vp]),
    "frb_scene_stats": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "frb_segment_mean": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp]),
    "frb_pdist": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "frb_linkage_complete": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "frb_knn_topk": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "frb_knn_redo_count": (_i, [_vp]),
    "frb_knn_fit": (_i, [_vp, _vp, _i, _i, _vp]),
    "frb_workspace_bytes": (C.c_longlong, [_vp, _i, _i, _i, _i]),
    "frb_knn_vote": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "frb_format_f32_json": (C.c_longlong, [_vp, _i, _vp, C.c_longlong]),
    "frb_sort_create": (_vp, [_i, _i, C.c_double]),
    "frb_sort_destroy": (None, [_vp]),
    "frb_sort_update": (_i, [_vp, _vp, _i, C.c_longlong, _vp]),
    "frb_sort_has_valid_tracker": (_i, [_vp, C.c_longlong]),
    "frb_sort_get_detection_bbox": (_i, [_vp, C.c_longlong, _vp]),
    "frb_sort_kill_trackers": (None, [_vp]),
    "frb_sort_nan_removed": (_i, [_vp]),
    "frb_sort_pop_expired": (_i, [_vp, C.c_longlong, C.c_longlong]),
    "frb_sort_expired_info": (_i, [_vp, _i, _vp, _vp]),
    "frb_sort_expired_boxes": (_i, [_vp, _i, _vp, _vp]),
}

_lib = None


def load_library() -> C.CDLL:
    """dlopen the in-tree shared library and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        try:                                  # fresh checkout: compile in-tree with nvcc (no GPU needed to build)
            from . import build as _build
            _build.build()
        except Exception as e:                # noqa: BLE001
            raise RuntimeError(f"{LIB_PATH} is missing and could not be built: {e}") from e
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(facerec_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class FrbError(RuntimeError):
    pass


class Context:
    """One frb_ctx per (process, device).  Raises if the device is not an sm_100 GPU."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = _vp()
        rc = self.lib.frb_init(int(device), C.byref(h))
        if rc != 0 or not h.value:
            raise FrbError(f"frb_init(device={device}) failed with status {rc}: an sm_100 (B200) GPU is required; "
                           "there is no CPU fallback")
        self.handle = h
        self.device = int(device)
        self.bf16 = False

    def check(self, rc: int, what: str = "") -> None:
        if rc != 0:
            msg = self.lib.frb_last_error(self.handle)
            raise FrbError(f"{what or 'frb call'} failed ({rc}): {msg.decode() if msg else ''}")

    def set_option(self, option: int, value: int) -> None:
        self.check(self.lib.frb_set_option(self.handle, option, value), "frb_set_option")
        if option == FRB_OPT_BF16:
            self.bf16 = bool(value)

    def launch_count(self) -> int:
        return int(self.lib.frb_launch_count(self.handle))

    def close(self) -> None:
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.frb_destroy(self.handle)
            self.handle = _vp()

    def __del__(self):  # best effort
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def ptr(t) -> int:
    """Raw device/host pointer of a torch tensor or numpy array (0 for None)."""
    if t is None:
        return 0
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream

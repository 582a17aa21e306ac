"""ctypes binding of libgssd_b200.so (include/gssd.h).  No CPU fallback: if the library cannot be
loaded or no CUDA device is present, every entry point raises."""
import ctypes as C
import os

import torch

from . import build as _build

ERR_ARG, ERR_LIMIT, ERR_WS, ERR_VALUE, ERR_EMPTY, ERR_UNSUPPORTED = -1, -2, -3, -4, -5, -6
MAX_MAPS, MAX_AR = 8, 8
PRIOR_V2, PRIOR_V2_CUSTOM, PRIOR_LEGACY = 0, 1, 2
WS_LSE, WS_MATCH, WS_LOSS, WS_NMS = 0, 1, 2, 3
STATS_HEADER_BYTES = 16


class PriorCfg(C.Structure):
    """`gssd_prior_cfg` (include/gssd.h)."""
    _fields_ = [
        ("version", C.c_int32), ("n_maps", C.c_int32), ("clip", C.c_int32),
        ("feature_maps", C.c_int32 * MAX_MAPS), ("n_ar", C.c_int32 * MAX_MAPS),
        ("min_dim", C.c_double), ("steps", C.c_double * MAX_MAPS),
        ("min_sizes", C.c_double * MAX_MAPS), ("max_sizes", C.c_double * MAX_MAPS),
        ("aspect_ratios", (C.c_double * MAX_AR) * MAX_MAPS), ("variance", C.c_double * 2),
    ]


class ConvDesc(C.Structure):
    """`gssd_conv_desc` (include/gssd.h)."""
    _fields_ = [
        ("n_img", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
        ("c_in", C.c_int32), ("c_out", C.c_int32), ("groups", C.c_int32),
        ("taps", C.c_int32), ("relu", C.c_int32),
        ("x", C.c_void_p), ("w", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p),
        ("row_ss_in", C.c_void_p), ("l2_eps", C.c_float),
        ("y", C.c_void_p), ("row_ss_out", C.c_void_p), ("chan_sum", C.c_void_p),
        ("loc", C.c_void_p), ("conf", C.c_void_p),
        ("n_anchor", C.c_int32), ("n_cls", C.c_int32), ("prior_off", C.c_int32), ("n_priors", C.c_int32),
    ]


class PipeCfg(C.Structure):
    """`gssd_pipe_cfg` (include/gssd.h)."""
    _fields_ = [("B", C.c_int32), ("P", C.c_int32), ("C", C.c_int32), ("top_k", C.c_int32), ("max_gt_rows", C.c_int32),
                ("depth", C.c_int32), ("match_thresh", C.c_float), ("var0", C.c_float), ("var1", C.c_float),
                ("conf_thresh", C.c_float), ("nms_thresh", C.c_float), ("negpos_ratio", C.c_int32)]


class PipeSlot(C.Structure):
    """`gssd_pipe_slot` (include/gssd.h): device addresses as integers."""
    _fields_ = [(n, C.c_void_p) for n in ("loc", "conf", "scores", "gt", "gt_off", "tags", "stats", "losses", "grad_loc",
                                          "grad_conf", "detect_out", "ws")] + [("ws_bytes", C.c_size_t), ("fused_state", C.c_void_p)]


XCHG_MAX_RANKS, XCHG_HANDLE_BYTES = 16, 64


class Xchg(C.Structure):
    """`gssd_xchg` (include/gssd.h)."""
    _fields_ = [("peers", C.c_void_p * XCHG_MAX_RANKS), ("rank", C.c_int32), ("world", C.c_int32),
                ("timeout_ms", C.c_uint32), ("reserved", C.c_uint32)]


MAX_GT_PER_IMAGE = 128
_P, _I, _F, _SZ = C.c_void_p, C.c_int, C.c_float, C.c_size_t
_SIGS = {
    "gssd_abi_version": (C.c_int, []),
    "gssd_error_string": (C.c_char_p, [_I]),
    "gssd_launch_count": (C.c_uint64, []),
    "gssd_priorbox_count": (_I, [C.POINTER(PriorCfg)]),
    "gssd_priorbox": (_I, [C.POINTER(PriorCfg), _P, _P]),
    "gssd_point_form": (_I, [_P, _I, _P, _P]),
    "gssd_center_size": (_I, [_P, _I, _P, _P]),
    "gssd_intersect": (_I, [_P, _I, _P, _I, _P, _P]),
    "gssd_jaccard": (_I, [_P, _I, _P, _I, _P, _P]),
    "gssd_encode": (_I, [_P, _P, _I, _F, _F, _P, _P]),
    "gssd_decode": (_I, [_P, _P, _I, _F, _F, _P, _P]),
    "gssd_log_sum_exp": (_I, [_P, _I, _I, _P, _P, _SZ, _P]),
    "gssd_match": (_I, [_P, _I, _P, _P, _I, _I, _I, _F, _F, _F, _P, _P, _P, _P, _SZ, _P]),
    "gssd_nms": (_I, [_P, _P, _I, _F, _I, _P, _P, _P, _SZ, _P]),
    "gssd_stats_bytes": (_SZ, [_I]),
    "gssd_mbox_match": (_I, [_P, _I, _P, _I, _P, _P, _I, _I, _I, _F, _P, _P, _P]),
    "gssd_mbox_loss": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _I, _I, _P, _P, _P, _I, _I, _F, _F,
                            _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "gssd_fused_state_bytes": (_SZ, []),
    "gssd_mbox_fused_supported": (_I, [_I, _I, _I, _I]),
    "gssd_mbox_loss_fused": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _I, _I, _F, _I, _F, _F, _P, C.POINTER(Xchg),
                                  _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "gssd_mbox_scale_grads": (_I, [_P, _SZ, _P, _SZ, _P, _P, _P]),
    "gssd_detect": (_I, [_P, _P, _P, _I, _I, _I, _I, _F, _F, _F, _F, _P, _P, _P, _P]),
    "gssd_detect_logits": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _F, _F, _P, _P, _P, _P]),
    "gssd_pipe_set_detect_logits": (_I, [_P, _I, _P]),
    "gssd_collect_detections": (_I, [_P, _I, _I, _I, _I, _F, _F, _F, _I, _P, _P, _P, _P]),
    "gssd_ap_workspace_bytes": (_SZ, [_I, _I]),
    "gssd_ap_eval": (_I, [_P, _P, _P, _P, _I, _I, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _SZ, _P]),
    "gssd_l2norm_fwd": (_I, [_P, _P, _I, _I, _I, _F, _P, _P, _P]),
    "gssd_l2norm_bwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _F, _P, _P, _P, _SZ, _P]),
    "gssd_l2norm_bwd_ws_bytes": (_SZ, [_I, _I, _I]),
    "gssd_workspace_bytes": (_SZ, [_I, _I, _I, _I, _I, _I]),
    "gssd_xchg_create": (_I, [C.POINTER(C.c_void_p), _P]),
    "gssd_xchg_open": (_I, [_P, C.POINTER(C.c_void_p)]),
    "gssd_xchg_close": (_I, [_P]),
    "gssd_xchg_destroy": (_I, [_P]),
    "gssd_mbox_match_x": (_I, [_P, _I, _P, _I, _P, _P, _I, _I, _I, _F, _P, _P, C.POINTER(Xchg), _P]),
    "gssd_mbox_loss_x": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _I, _I, _P, _P, C.POINTER(Xchg), _I, _F, _F,
                              _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "gssd_pipe_set_xchg": (_I, [_P, C.POINTER(Xchg)]),
    "gssd_pipe_arena_bytes": (_SZ, [C.POINTER(PipeCfg)]),
    "gssd_pipe_create": (_I, [C.POINTER(C.c_void_p), C.POINTER(PipeCfg), _P, _P, _SZ]),
    "gssd_pipe_destroy": (None, [_P]),
    "gssd_pipe_slot_info": (_I, [_P, _I, C.POINTER(PipeSlot)]),
    "gssd_pipe_submit": (C.c_int64, [_P, _P, _P, _P, _P, _P, _I, _I, _P, _P]),
    "gssd_pipe_begin": (C.c_int64, [_P, _P, _P, _P, _P, _P, _I, _I, _P, C.POINTER(C.c_void_p)]),
    "gssd_pipe_finish": (_I, [_P, C.c_int64, _P, _I, _P]),
    "gssd_pipe_wait": (_I, [_P, C.c_int64]),
    "gssd_conv_igemm": (_I, [C.POINTER(ConvDesc), _P]),
    "gssd_conv_pack_weights": (_I, [_P, _I, _I, _I, _I, _P, _P, _P]),
    "gssd_nchw_to_pm": (_I, [_P, _I, _I, _I, _I, _P, _P]),
    "gssd_pm_to_nchw": (_I, [_P, _I, _I, _I, _I, _P, _P]),
    "gssd_maxpool_pm": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _P, C.POINTER(C.c_int), C.POINTER(C.c_int), _P]),
    "gssd_conv_wgrad": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "gssd_conv_wgrad_bytes": (_SZ, [_I, _I, _I, _I]),
    "gssd_head_grad_pm": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    "gssd_bn_relu_bwd_pm": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _F, _I, _P, _P, _P, _F, _P, _F, _P, _P, _P]),
    "gssd_bn_act_pm_to": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _F, _I, _P, _P, _P]),
    "gssd_bn_act_pm": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _F, _I, _P, _P, _P]),
    "gssd_dcn_columns": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "gssd_dcn_columns_bwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "gssd_pmf32_to_nchw": (_I, [_P, _I, _I, _I, _I, _P, _P]),
    "gssd_bn_relu_nchw_fwd": (_I, [_P, _P, _P, _I, _I, _I, _F, _I, _P, _P, _P, _P, _F, _P, _P, _P]),
    "gssd_bn_relu_nchw_bwd": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "gssd_maxpool_nchw_bwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "gssd_bn_relu_nhwc_fwd": (_I, [_P, _P, _P, C.c_long, _I, _F, _I, _P, _P, _P, _P, _F, _P, _P, _P]),
    "gssd_bn_relu_nhwc_bwd": (_I, [_P, _P, _P, _P, _P, C.c_long, _I, _I, _P, _P, _P, _P, _P]),
    "gssd_maxpool_nhwc_bwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "gssd_attn_fwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P]),
    "gssd_attn_bwd": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
}

_lib = None


def load():
    """dlopen the in-tree library (building it first when the sources are newer and nvcc exists)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    alt = os.environ.get("GSSD_LIB")                  # development: an alternative build of the library (A/B measurements)
    if alt:
        path = os.path.abspath(alt)
    elif _build.stale():
        try:
            _build.build()
        except Exception as e:  # no nvcc on this box and no prebuilt library
            if not os.path.exists(path):
                raise RuntimeError("libgssd_b200.so is missing and could not be built: %s" % e)
    lib = C.CDLL(path)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.gssd_abi_version() != 2:
        raise RuntimeError("libgssd_b200.so ABI version mismatch")
    _lib = lib
    return lib


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("grouped_ssd_pytorch_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return load()


def check(rc, what=""):
    if rc == 0:
        return
    msg = load().gssd_error_string(rc).decode()
    if rc == ERR_VALUE:
        raise ValueError(what or msg)
    if rc == ERR_EMPTY:
        raise IndexError(what or msg)
    raise RuntimeError("%s: %s (code %d)" % (what or "libgssd_b200", msg, rc))


def ptr(t):
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def device_of(*tensors):
    """the CUDA device to compute on: the first CUDA tensor's, else the current device."""
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.is_cuda:
            return t.device
    return torch.device("cuda", torch.cuda.current_device())


def f32(t, dev):
    """contiguous float32 view/copy of `t` on `dev` (no copy when it already is)."""
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


_fused_states = {}


def fused_state(dev, stream_ptr):
    """the zero-initialised rendezvous state of the one-launch MultiBoxLoss (include/gssd.h: gssd_mbox_loss_fused), one per
    (device, stream): launches that share a state must be ordered on one stream, and the kernel leaves it zeroed."""
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), int(stream_ptr))
    t = _fused_states.get(key)
    if t is None:
        t = torch.zeros((max(64, int(load().gssd_fused_state_bytes())),), dtype=torch.uint8, device=dev)
        _fused_states[key] = t
    return t


def launch_count():
    return int(load().gssd_launch_count())


def prior_cfg(cfg):
    """cfg dict -> PriorCfg; branch selection follows prior_box.py:35,58,87,116,139."""
    c = PriorCfg()
    name = cfg["name"]
    if name in ("v2", "v2_512"):
        c.version = PRIOR_V2
    elif name in ("v2_custom", "v2_custom_squareonly", "v2_custom_512"):
        c.version = PRIOR_V2_CUSTOM
    else:
        c.version = PRIOR_LEGACY
    c.n_maps = len(cfg["feature_maps"])
    if c.n_maps > MAX_MAPS:
        raise RuntimeError("PriorBox: more than %d feature maps" % MAX_MAPS)
    c.clip = 1 if cfg["clip"] else 0
    c.min_dim = float(cfg["min_dim"])
    for k in range(c.n_maps):
        c.feature_maps[k] = int(cfg["feature_maps"][k])
        c.steps[k] = float(cfg["steps"][k])
        c.min_sizes[k] = float(cfg["min_sizes"][k])
        c.max_sizes[k] = float(cfg["max_sizes"][k])
        ars = cfg["aspect_ratios"][k]
        if len(ars) > MAX_AR:
            raise RuntimeError("PriorBox: more than %d aspect ratios" % MAX_AR)
        c.n_ar[k] = len(ars)
        for a, ar in enumerate(ars):
            c.aspect_ratios[k][a] = float(ar)
    var = cfg["variance"] or [0.1]
    c.variance[0] = float(var[0])
    c.variance[1] = float(var[-1])
    return c

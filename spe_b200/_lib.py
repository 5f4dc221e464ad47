"""ctypes binding of libspe_b200.so (the C ABI declared in include/spe_b200.h).

There is no fallback: if the shared library is missing the import of any compute op raises.  Build it
with ``python -m spe_b200.build`` (or ``__graft_entry__.build()``).
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libspe_b200.so")
_lib = None

c_p = C.c_void_p
c_i = C.c_int
c_l = C.c_int64
c_f = C.c_float


class GemmArgs(C.Structure):
    _fields_ = [
        ("M", c_i), ("N", c_i), ("K", c_i), ("batch1", c_i), ("batch2", c_i),
        ("A", c_p), ("a_major", c_i), ("lda", c_l), ("a_sb1", c_l), ("a_sb2", c_l),
        ("B", c_p), ("b_major", c_i), ("ldb", c_l), ("b_sb1", c_l), ("b_sb2", c_l),
        ("C", c_p), ("c_dtype", c_i), ("ldc", c_l), ("c_sb1", c_l), ("c_sb2", c_l),
        ("alpha", c_f), ("bias", c_p), ("act", c_i), ("aux_in", c_p), ("aux_out", c_p), ("ld_aux", c_l),
        ("gamma", c_p), ("residual", c_p), ("ldr", c_l), ("r_sb1", c_l), ("r_sb2", c_l),
        ("split", c_i), ("split_stride", c_i),
    ]


class AttentionArgs(C.Structure):
    _fields_ = [
        ("B", c_i), ("H", c_i), ("Lq", c_i), ("Lk", c_i), ("d", c_i), ("d2", c_i), ("dv", c_i),
        ("q", c_p), ("q_ld", c_l), ("q_sb", c_l), ("k", c_p), ("k_ld", c_l), ("k_sb", c_l), ("v", c_p), ("v_ld", c_l), ("v_sb", c_l),
        ("q2", c_p), ("q2_ld", c_l), ("q2_sb", c_l), ("k2", c_p), ("k2_ld", c_l), ("k2_sb", c_l),
        ("mask", c_p), ("scale", c_f), ("out", c_p), ("out_ld", c_l), ("out_sb", c_l), ("P", c_p), ("ldP", c_l), ("lse", c_p),
    ]


class AttentionBwdArgs(C.Structure):
    _fields_ = [
        ("B", c_i), ("H", c_i), ("Lq", c_i), ("Lk", c_i), ("d", c_i), ("dv", c_i),
        ("dS", c_p), ("P", c_p), ("ld", c_l), ("q", c_p), ("q_ld", c_l), ("q_sb", c_l), ("k", c_p), ("k_ld", c_l), ("k_sb", c_l),
        ("dO", c_p), ("do_ld", c_l), ("do_sb", c_l), ("alpha", c_f), ("dq", c_p), ("dq_ld", c_l), ("dq_sb", c_l),
        ("dk", c_p), ("dk_ld", c_l), ("dk_sb", c_l), ("dv_out", c_p), ("dv_ld", c_l), ("dv_sb", c_l), ("workspace", c_p), ("delta", c_p),
    ]


class AttentionBwd2Args(C.Structure):
    _fields_ = [("B", c_i), ("H", c_i), ("Lq", c_i), ("Lk", c_i), ("d", c_i), ("d2", c_i), ("dv", c_i)] + \
               [f for n in ("q", "k", "v", "q2", "k2") for f in ((n, c_p), (n + "_ld", c_l), (n + "_sb", c_l))] + \
               [("dO", c_p), ("do_ld", c_l), ("do_sb", c_l), ("mask", c_p), ("scale", c_f), ("lse", c_p), ("delta", c_p)] + \
               [f for n in ("dq", "dk") for f in ((n, c_p), (n + "_ld", c_l), (n + "_sb", c_l))] + \
               [("dv_out", c_p), ("dv_ld", c_l), ("dv_sb", c_l)] + \
               [f for n in ("dq2", "dk2") for f in ((n, c_p), (n + "_ld", c_l), (n + "_sb", c_l))] + [("workspace", c_p)]


class TalkingFusedArgs(C.Structure):
    _fields_ = [("B", c_i), ("H", c_i), ("N", c_i), ("dh", c_i),
                ("q", c_p), ("q_ld", c_l), ("q_sb", c_l), ("k", c_p), ("k_ld", c_l), ("k_sb", c_l), ("v", c_p), ("v_ld", c_l), ("v_sb", c_l),
                ("Wl", c_p), ("bl", c_p), ("Ww", c_p), ("bw", c_p), ("scale", c_f),
                ("out", c_p), ("out_ld", c_l), ("out_sb", c_l), ("lse2", c_p), ("workspace", c_p), ("workspace_bytes", c_l)]


class TalkingFusedBwdArgs(C.Structure):
    _fields_ = [("B", c_i), ("H", c_i), ("N", c_i), ("dh", c_i),
                ("q", c_p), ("q_ld", c_l), ("q_sb", c_l), ("k", c_p), ("k_ld", c_l), ("k_sb", c_l), ("v", c_p), ("v_ld", c_l), ("v_sb", c_l),
                ("dO", c_p), ("do_ld", c_l), ("do_sb", c_l), ("Wl", c_p), ("bl", c_p), ("Ww", c_p), ("bw", c_p), ("scale", c_f), ("lse2", c_p),
                ("dqkv", c_p), ("dqkv_ld", c_l), ("dWl", c_p), ("dWw", c_p), ("workspace", c_p), ("workspace_bytes", c_l)]


class AdamwArgs(C.Structure):
    _fields_ = [("p", c_p), ("g", c_p), ("m", c_p), ("v", c_p), ("n", c_l), ("nseg", c_i),
                ("seg_end", c_l * 16), ("seg_group", C.c_int32 * 16), ("lr", c_p), ("wd", c_p),
                ("beta1", c_f), ("beta2", c_f), ("eps", c_f), ("state", c_p), ("g_out", c_p), ("shadow_bf16", c_p)]


_SIGS = {
    "spe_cam_boxes_workspace_bytes": (c_l, [c_i, c_i, c_i]),
    "spe_cam_boxes": (c_i, [c_p, c_i, c_i, c_i, c_i, c_p, c_i, c_i, c_i, c_i, c_f, c_f, c_p, c_p, c_p, c_l, c_p]),
    "spe_cam_boxes_multi": (c_i, [c_p, c_i, c_i, c_i, c_i, c_p, c_i, c_i, c_i, c_i, c_f, c_f, C.c_double, c_i, c_p, c_p, c_p, c_p, c_l, c_p]),
    "spe_gt_jitter_repeat": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_f, c_i, c_f, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "spe_sumsq_workspace_floats": (c_l, []),
    "spe_sumsq_f32": (c_i, [c_p, c_l, c_p, c_p, c_p]),
    "spe_adamw_tick": (c_i, [c_p, c_f, c_f, c_p, c_f, c_p]),
    "spe_adamw_flat": (c_i, [C.POINTER(AdamwArgs), c_p]),
    "spe_scale_by_clip_coef": (c_i, [c_p, c_l, c_p, c_p]),
    "spe_version": (c_i, []),
    "spe_launch_count": (c_l, []),
    "spe_set_cache_config": (c_i, [c_i]),
    "spe_prof_enable": (c_i, [c_i]),
    "spe_prof_collect": (c_i, [c_p, c_p, c_p]),
    "spe_prof_family_count": (c_i, []),
    "spe_gemm": (c_i, [C.POINTER(GemmArgs), c_p]),
    "spe_attention_fwd": (c_i, [C.POINTER(AttentionArgs), c_p]),
    "spe_attention_bwd_gemms": (c_i, [C.POINTER(AttentionBwdArgs), c_p]),
    "spe_attention_bwd_gemms_workspace": (c_l, [c_i, c_i, c_i, c_i]),
    "spe_attention_bwd": (c_i, [C.POINTER(AttentionBwd2Args), c_p]),
    "spe_attention_delta": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_l, c_l, c_l, c_l, c_p, c_p]),
    "spe_layernorm_fwd": (c_i, [c_p, c_p, c_p, c_f, c_l, c_i, c_p, c_p, c_p, c_p, c_p]),
    "spe_layernorm_bwd": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_l, c_i, c_p, c_p, c_p, c_p]),
    "spe_talking_softmax_fwd": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_l, c_l, c_p]),
    "spe_talking_softmax_bwd": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_l, c_l, c_p, c_p, c_p, c_p,
                                      c_p, c_l, c_p]),
    "spe_talking_softmax_bwd_workspace": (c_l, [c_i, c_i, c_i, c_i]),
    "spe_talking_s16_supported": (c_i, [c_i, c_i, c_l, c_l]),
    "spe_talking_softmax_fwd_s16": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_l, c_l, c_p]),
    "spe_talking_softmax_bwd_s16": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_l, c_l, c_p, c_p, c_p, c_p,
                                          c_p, c_l, c_p]),
    "spe_talking_fused_supported": (c_i, [c_i, c_i]),
    "spe_talking_fused_fwd_workspace": (c_l, [c_i, c_i, c_i, c_i]),
    "spe_talking_fused_fwd": (c_i, [C.POINTER(TalkingFusedArgs), c_p]),
    "spe_talking_fused_bwd_workspace": (c_l, [c_i, c_i, c_i, c_i]),
    "spe_talking_fused_bwd": (c_i, [C.POINTER(TalkingFusedBwdArgs), c_p]),
    "spe_softmax_fwd": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_l, c_l, c_p, c_p]),
    "spe_cam_std_reweight": (c_i, [c_p, c_i, c_i, c_i, c_l, c_i, c_i, c_i, c_i, c_p, c_p]),
    "spe_softmax_bwd": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_l, c_p]),
    "spe_layerscale_bwd": (c_i, [c_p, c_p, c_p, c_l, c_i, c_p, c_p, c_p, c_p]),
    "spe_colsum_bf16": (c_i, [c_p, c_l, c_i, c_l, c_p, c_p]),
    "spe_colsum_bf16_batched": (c_i, [c_p, c_i, c_l, c_i, c_l, c_l, c_p, c_p]),
    "spe_axpby_cast": (c_i, [c_p, c_p, c_f, c_f, c_l, c_p, c_p, c_p]),
    "spe_cast_bf16_to_f32": (c_i, [c_p, c_p, c_l, c_p]),
    "spe_add_bf16_into_f32": (c_i, [c_p, c_p, c_l, c_p]),
    "spe_relu_bwd_bf16": (c_i, [c_p, c_p, c_p, c_l, c_p]),
    "spe_im2col_patch": (c_i, [c_p, c_i, c_i, c_i, c_i, c_p, c_p]),
    "spe_bicubic_tokens_fwd": (c_i, [c_p, c_i, c_i, c_i, c_p, c_i, c_i, c_p]),
    "spe_bicubic_tokens_bwd": (c_i, [c_p, c_i, c_i, c_i, c_p, c_i, c_i, c_p]),
    "spe_sine_pos_2d": (c_i, [c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_p]),
    "spe_query_sine_fwd": (c_i, [c_p, c_l, c_i, c_p, c_p]),
    "spe_query_sine_bwd": (c_i, [c_p, c_p, c_l, c_i, c_p, c_p]),
    "spe_match_cost": (c_i, [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_f, c_f, c_f, c_p, c_l, c_i, c_p]),
    "spe_cast_f32_to_bf16_multi": (c_i, [c_p, c_i, c_l, c_p]),
    "spe_lsap_workspace_bytes": (c_l, [c_i, c_i, c_i]),
    "spe_lsap_batched": (c_i, [c_p, c_i, c_i, c_l, c_p, c_i, c_p, c_p, c_p]),
    "spe_focal_loss_workspace_bytes": (c_l, [c_i, c_i]),
    "spe_focal_loss": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_f, c_f, c_p, c_p, c_p, c_p]),
    "spe_box_loss": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_p, c_p, c_p, c_p]),
    "spe_bce_logits": (c_i, [c_p, c_p, c_l, c_p, c_p, c_p]),
    "spe_box_iou_pairwise": (c_i, [c_p, c_i, c_p, c_i, c_p, c_p, c_p, c_p]),
}
EXPORTS = sorted(list(_SIGS) + ["spe_last_error"])


def lib():
    """Load (once) and return the CDLL.  Raises if the native library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError("spe_b200: native library %s is missing -- run `python -m spe_b200.build`; "
                               "there is no CPU/PyTorch fallback" % SO_PATH)
        L = C.CDLL(SO_PATH)
        L.spe_last_error.restype = C.c_char_p
        L.spe_last_error.argtypes = []
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
        mode = os.environ.get("SPE_CACHE_CONFIG", "")
        if mode != "" and torch.cuda.is_available():
            L.spe_set_cache_config(int(mode))
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError("libspe_b200: " + lib().spe_last_error().decode())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_cur_device = getattr(torch._C, "_cuda_getDevice", None)


def stream():
    """cudaStream_t of torch's current stream on the current device.  (torch.cuda.current_stream() builds a Python Stream
    object through several layers, ~10 us; this is called once per kernel launch, several thousand times a step.)"""
    if _raw_stream is not None and _cur_device is not None:
        return _raw_stream(_cur_device())
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return 0 if t is None else t.data_ptr()


def launch_count():
    return int(lib().spe_launch_count())


PROF_FAMILIES = ["gemm", "talking_softmax_fwd", "talking_softmax_bwd", "softmax", "layernorm", "matcher_lsap", "other", "gemm_attention", "attention_fused"]


def prof_enable(on):
    lib().spe_prof_enable(1 if on else 0)


def prof_collect():
    """{family: (ms, work, launches)} since the last collect (synchronises the device)."""
    n = lib().spe_prof_family_count()
    ms, wk, ln = (C.c_double * n)(), (C.c_double * n)(), (C.c_int64 * n)()
    lib().spe_prof_collect(ms, wk, ln)
    return {PROF_FAMILIES[i]: (ms[i], wk[i], ln[i]) for i in range(n)}

"""ctypes binding of libqeb_sm100.so (the C ABI of include/qeb.h). No CPU fallback: a missing library is an error."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# QEB_LIB: another build of the same ABI (same-box A/B measurements of a kernel change against the previous build)
LIB_PATH = os.environ.get("QEB_LIB") or os.path.join(_HERE, "libqeb_sm100.so")

P = ctypes.c_void_p
I = ctypes.c_int
LL = ctypes.c_longlong
F = ctypes.c_float
ULL = ctypes.c_ulonglong
SZ = ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol declared in include/qeb.h
SIGNATURES = {
    "qeb_last_error": (ctypes.c_char_p, []),
    "qeb_abi_version": (I, []),
    "qeb_launch_count": (LL, []),
    "qeb_reset_launch_count": (None, []),
    "qeb_check_device": (I, []),
    "qeb_ctc_workspace_bytes": (SZ, [I, I, I]),
    "qeb_ctc_fwd": (I, [P, LL, LL, P, P, P, P, P, I, I, I, I, I, I, I, P, P, P, P]),
    "qeb_ctc_bwd": (I, [P, LL, LL, P, P, P, P, P, I, I, I, I, I, I, I, P, P, P, P, LL, LL, P]),
    "qeb_log_softmax_fwd": (I, [P, P, LL, I, P]),
    "qeb_log_softmax_bwd": (I, [P, P, P, LL, I, P]),
    "qeb_levenshtein_batch": (I, [P, P, P, P, P, P, I, I, I, P, P, P, P]),
    "qeb_greedy_decode": (I, [P, LL, LL, I, I, I, I, P, P, P, P]),
    "qeb_greedy_collapse": (I, [P, LL, LL, I, I, I, P, P, P]),
    "qeb_cer_topk_segmented": (I, [P, P, P, P, I, P, P]),
    "qeb_cer_topk_global_workspace_bytes": (SZ, [LL]),
    "qeb_cer_topk_global": (I, [P, LL, LL, P, P, P]),
    "qeb_cer_range_segmented": (I, [P, P, P, P, P, I, P, P, P, P]),
    "qeb_gauss_jitter": (I, [P, P, F, F, P, ULL, LL, I, P, P, P]),
    "qeb_gauss_jitter_devseed": (I, [P, P, F, F, P, ULL, LL, I, P, P, P]),
    "qeb_to_uint8": (I, [P, LL, P, P]),
    "qeb_crop_pad_gather": (I, [P, I, I, P, I, I, I, P, P]),
    "qeb_crop_pad_scatter": (I, [P, I, I, P, I, I, I, P, P]),
    "qeb_pack_weight": (I, [P, P, I, I, I, I, I, P]),
    "qeb_nchw_to_nhwc": (I, [P, P, I, I, I, I, I, P]),
    "qeb_nhwc_to_nchw": (I, [P, P, I, I, I, I, I, P]),
    "qeb_conv_fprop_tc": (I, [P, I, I, I, I, I, P, I, I, I, I, I, P, P, I, P, I, I, P]),
    "qeb_conv_fprop_tc16": (I, [P, I, I, I, I, I, P, I, I, I, I, I, P, P, I, P, I, P, P]),
    "qeb_conv_wgrad_tc": (I, [P, I, I, I, I, P, I, I, I, I, I, I, I, P, P]),
    "qeb_conv_wgrad_tc16": (I, [P, P, I, I, I, I, P, P, I, I, I, I, I, I, I, P, P, P]),
    "qeb_convT2x2_fprop_tc": (I, [P, I, I, I, I, I, P, P, I, P, I, P]),
    "qeb_convT2x2_dgrad_tc": (I, [P, I, I, I, I, I, P, I, P, I, P]),
    "qeb_convT2x2_wgrad_tc": (I, [P, I, I, I, I, P, I, I, I, P, P]),
    "qeb_lstm_layer_fwd": (I, [P, P, P, P, P, I, I, P]),
    "qeb_lstm_layer_bwd": (I, [P, P, P, P, P, I, I, P]),
    "qeb_prof_enable": (None, [I]),
    "qeb_debug_set_timeline": (None, [P]),
    "qeb_prof_report": (I, [ctypes.c_char_p, I]),
    "qeb_mse_ones_fwd": (I, [P, LL, P, P]),
    "qeb_mse_ones_bwd": (I, [P, LL, P, P, P]),
    "qeb_adam_table_entry_bytes": (I, []),
    "qeb_adam_multi": (I, [P, I, LL, F, F, F, F, F, I, P]),
    "qeb_crnn_workspace_bytes": (SZ, [I, I, I]),
    "qeb_crnn_num_params": (I, []),
    "qeb_crnn_forward": (I, [P, I, I, I, P, P, I, P, P, P]),
    "qeb_crnn_forward_fused": (I, [P, I, I, I, P, P, I, P, P, I, P, P, F, F, ULL, P, P, P, P]),
    "qeb_crnn_backward": (I, [P, I, I, I, P, I, P, P, P, P, P]),
    "qeb_unet_workspace_bytes": (SZ, [I, I, I]),
    "qeb_unet_num_params": (I, []),
    "qeb_unet_num_buffers": (I, []),
    "qeb_unet_forward": (I, [P, I, I, I, P, P, I, P, P, P]),
    "qeb_unet_backward": (I, [P, I, I, I, P, I, P, P, P, P, P, P]),
    "qeb_unet_backward_bucketed": (I, [P, I, I, I, P, I, P, P, P, P, P, P, P]),
}

_lib = None


class QebError(RuntimeError):
    pass


def load():
    """Load the shared library (building it is __graft_entry__.build()'s job)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise QebError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the qeb hot path)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def call(name, *args):
    """Invoke an int-returning entry point; raise QebError with qeb_last_error() on failure."""
    L = load()
    rc = getattr(L, name)(*args)
    if rc != 0:
        raise QebError(f"{name} failed ({rc}): {L.qeb_last_error().decode()}")


def ptr(t):
    """Device (or host) pointer of a torch tensor, or NULL for None."""
    return None if t is None else t.data_ptr()


def stream():
    import torch

    return torch.cuda.current_stream().cuda_stream


def launch_count():
    return int(load().qeb_launch_count())


def reset_launch_count():
    load().qeb_reset_launch_count()


def prof_enable(on=True):
    load().qeb_prof_enable(1 if on else 0)


def prof_report():
    """Per-kernel-family totals recorded since prof_enable(True): {tag: {launches, ms, flops, bytes}}."""
    import json

    buf = ctypes.create_string_buffer(1 << 16)
    call("qeb_prof_report", buf, len(buf))
    return json.loads(buf.value.decode())

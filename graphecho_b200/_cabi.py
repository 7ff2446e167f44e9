"""ctypes binding of libgraphecho_b200.so (the C-ABI in include/graphecho_b200.h).

The product path has no CPU fallback: if the shared library is missing, or a tensor is not
a CUDA tensor, the call raises.  Only raw device pointers, sizes and the current CUDA
stream cross this boundary.
"""
from __future__ import annotations

import ctypes
import re
from ctypes import c_int, c_size_t, c_void_p, c_char_p, c_ulonglong, c_float, c_double, c_longlong
from pathlib import Path

import torch

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libgraphecho_b200.so"
HEADER = PKG.parent / "include" / "graphecho_b200.h"

P, I, Z, F, L = c_void_p, c_int, c_size_t, c_float, c_longlong

# name -> (restype, argtypes)
_SIGNATURES = {
    "ge_version": (c_int, []),
    "ge_last_error": (c_char_p, []),
    "ge_device_sm_count": (c_int, []),
    "ge_launch_count": (c_ulonglong, []),
    "ge_affinity_pairwise_fwd": (c_int, [P, P, P, P, P, I, I, I, I, P]),
    "ge_affinity_pairwise_bwd_workspace_bytes": (c_size_t, [I, I, I, I]),
    "ge_affinity_pairwise_bwd": (c_int, [P, P, P, P, P, P, P, P, P, Z, I, I, I, I, P]),
    "ge_sinkhorn_rpm_set_path": (c_int, [I]),
    "ge_sinkhorn_rpm_cluster_size": (c_int, [I, I, I]),
    "ge_sinkhorn_rpm_fwd": (c_int, [P, P, P, P, P, I, I, I, I, I, I, P]),
    "ge_sinkhorn_rpm_bwd": (c_int, [P, P, P, P, P, P, I, I, I, I, I, I, P]),
    "ge_matching_loss_fwd": (c_int, [P, P, P, P, P, P, I, I, F, F, P]),
    "ge_matching_loss_bwd": (c_int, [P, P, P, P, P, P, P, I, I, F, F, P]),
    "ge_stem_conv_supported": (c_int, [I, I]),
    "ge_stem_conv_fwd": (c_int, [P, P, P, I, I, I, I, P]),
    "ge_stem_conv_wgrad_workspace_bytes": (c_size_t, [I, I, I]),
    "ge_stem_conv_wgrad": (c_int, [P, P, P, P, Z, I, I, I, I, P]),
    "ge_sinkhorn_distance_fwd": (c_int, [P, P, P, P, P, P, P, P, P, I, I, I, I, F, I, c_double, P]),
    "ge_sinkhorn_distance_bwd": (c_int, [P, P, P, P, P, P, P, P, P, P, I, I, I, I, F, I, P]),
    "ge_knn_graph_set_path": (c_int, [I]),
    "ge_knn_graph_workspace_bytes": (c_size_t, [I, I, I, I]),
    "ge_knn_graph": (c_int, [P, P, P, P, P, Z, I, I, I, I, I, I, P]),
    "ge_knn_graph_nmajor_supported": (c_int, [I, I, I, I, I, I]),
    "ge_knn_graph_nmajor": (c_int, [P, P, I, P, P, Z, I, I, I, I, I, I, P]),
    "ge_mrconv_gather_nmajor_fwd": (c_int, [P, P, P, P, P, I, I, I, I, I, I, P]),
    "ge_mrconv_gather_nmajor_bwd": (c_int, [P, P, P, P, P, I, I, I, I, I, I, P]),
    "ge_mrconv_gather_nmajor_bwd_self": (c_int, [P, P, P, P, I, I, I, I, I, P]),
    "ge_mrconv_gather_fwd": (c_int, [P, P, P, P, P, P, I, I, I, I, I, P]),
    "ge_mrconv_gather_bwd": (c_int, [P, P, P, P, P, P, I, I, I, I, I, P]),
    "ge_tgcn_pool_concat_fwd": (c_int, [P, L, P, I, L, I, I, I, I, I, I, P]),
    "ge_tgcn_pool_concat_bwd": (c_int, [P, P, I, L, I, I, I, I, I, I, P]),
    "ge_upsample_add_fwd": (c_int, [P, P, P, I, I, I, I, I, I, I, P]),
    "ge_upsample_bwd": (c_int, [P, P, I, I, I, I, I, I, I, P]),
    "ge_group_stats": (c_int, [P, P, P, I, I, I, I, I, F, P]),
    "ge_maxpool3s2_fwd": (c_int, [P, P, P, I, I, I, I, I, P]),
    "ge_maxpool3s2_bwd": (c_int, [P, P, P, I, I, I, I, I, P]),
    "ge_group_stats_bias": (c_int, [P, P, P, P, P, I, I, I, I, I, F, P]),
    "ge_gn_relu_upsample_fwd": (c_int, [P, P, P, P, P, P, I, I, I, I, I, I, I, P]),
    "ge_gn_relu_upsample_bwd": (c_int, [P, P, P, P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, I, I, P]),
    "ge_seg_tail_fwd": (c_int, [P, P, P, P, P, P, P, P, I, I, I, I, I, I, I, I, P]),
    "ge_bn_workspace_bytes": (c_size_t, [L, I]),
    "ge_bn_relu_mask_bytes": (c_size_t, [L, I]),
    "ge_bn_set_path": (c_int, [I]),
    "ge_bn_fwd_train": (c_int, [P, P, P, P, P, P, P, F, F, P, P, P, P, P, Z, I, L, L, I, I, P]),
    "ge_bn_fwd_train_prestat": (c_int, [P, P, P, P, P, P, P, F, F, P, P, P, P, P, I, I, I, L, L, I, I, P]),
    "ge_conv1x1_tc_supported": (c_int, [L, I, I]),
    "ge_conv1x1_tc_partial_rows": (c_int, [L, L, I, I, P]),
    "ge_conv1x1_bn_stats": (c_int, [P, P, P, P, P, L, L, I, I, P]),
    "ge_bn_fwd_eval": (c_int, [P, P, P, P, P, P, F, P, I, L, I, I, P]),
    "ge_bn_bwd": (c_int, [P, P, P, P, P, P, P, P, P, P, P, Z, I, L, L, I, I, P]),
    "ge_bn_sync_stats": (c_int, [P, P, P, P, Z, I, L, L, I, P]),
    "ge_bn_sync_fwd_apply": (c_int, [P, P, P, P, P, P, P, F, F, P, P, P, P, P, I, I, L, L, I, I, P]),
    "ge_bn_sync_bwd_reduce": (c_int, [P, P, P, P, P, P, P, P, P, Z, I, L, L, I, I, P]),
    "ge_bn_sync_bwd_apply": (c_int, [P, P, P, P, P, P, P, P, P, I, L, L, I, I, P]),
    "ge_tgcn_recurrence_supported": (c_int, [I, I, I, I, I, I]),
    "ge_tgcn_recurrence_fwd": (c_int, [P, P, P, P, P, P, P, I, I, I, I, I, P]),
    "ge_tgcn_recurrence_bwd": (c_int, [P, P, P, P, P, P, P, P, P, P, P, I, I, I, I, I, P]),
    "ge_seg_loss_workspace_bytes": (c_size_t, [I, I, I]),
    "ge_seg_loss_fwd": (c_int, [P, P, P, P, P, Z, I, I, I, F, P]),
    "ge_seg_loss_bwd": (c_int, [P, P, P, P, P, I, I, I, P]),
    "ge_mask_boxes": (c_int, [P, P, I, I, I, I, I, P]),
    "ge_sampler_labels": (c_int, [P, P, P, P, P, P, P, P, I, I, I, P]),
    "ge_sampler_gather": (c_int, [P, P, P, P, P, P, P, P, P, P, P, P, P, P, I, I, I, P]),
    "ge_sampler_scatter": (c_int, [P, P, P, P, P, P, P, I, I, I, P]),
    "ge_spectral_bipartition_max_points": (c_int, []),
    "ge_spectral_bipartition": (c_int, [P, P, I, I, I, I, P]),
    "ge_seg_tail_bwd": (c_int, [P, P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, I, I, P]),
}

_lib = None


class GraphEchoNativeError(RuntimeError):
    pass


def header_symbols() -> list[str]:
    """Every function the public header declares (used by the ABI export test)."""
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ge_[a-z0-9_]+)\s*\(", text)))


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise GraphEchoNativeError(
                f"{LIB_PATH} is missing: build it with `python -m graphecho_b200.build` "
                "(graphecho_b200 has no CPU or PyTorch-eager fallback)")
        handle = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error() -> str:
    return lib().ge_last_error().decode("utf-8", "replace")


def ptr(t: torch.Tensor | None) -> c_void_p:
    if t is None:
        return c_void_p(0)
    if not t.is_cuda:
        raise GraphEchoNativeError("graphecho_b200 kernels take CUDA tensors only (no CPU fallback)")
    return c_void_p(t.data_ptr())


def stream() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


# Optional per-entry-point timing (bench.py): name -> list of (start_event, end_event, bytes, flops).
# Events are recorded on the stream the kernels are launched on (torch's current stream).
_PROFILE: dict | None = None


def profile_start() -> None:
    global _PROFILE
    _PROFILE = {}


def profile_stop() -> dict:
    """-> {name: {"calls", "ms", "bytes", "flops"}} summed over the profiled region (syncs)."""
    global _PROFILE
    rec, _PROFILE = _PROFILE or {}, None
    torch.cuda.synchronize()
    out = {}
    for name, items in rec.items():
        out[name] = {"calls": len(items), "ms": sum(a.elapsed_time(b) for a, b, _, _ in items),
                     "bytes": sum(i[2] for i in items), "flops": sum(i[3] for i in items)}
    return out


def call(name: str, *args, work=(0, 0)) -> None:
    """Invoke an int-returning entry point and raise on a non-zero status.  `work` = the call's
    ALGORITHMIC (bytes, flops), used only by the profiler."""
    if _PROFILE is not None:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = getattr(lib(), name)(*args)
        b.record()
        _PROFILE.setdefault(name, []).append((a, b, work[0], work[1]))
    else:
        rc = getattr(lib(), name)(*args)
    if rc != 0:
        kind = "argument/shape" if rc < 0 else "CUDA"
        raise GraphEchoNativeError(f"{name} failed ({kind} error {rc}): {last_error()}")


def launch_count() -> int:
    return int(lib().ge_launch_count())

"""torch.autograd.Function shells around the C-ABI kernels (libgraphecho_b200.so).

Every function here takes CUDA tensors only; there is no eager/CPU fallback.  The modules in
graphecho_b200.models / graphecho_b200.utils compose these with library GEMMs/convs.
"""
from __future__ import annotations

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _cabi
from ._cabi import call, ptr, stream
from ctypes import c_int, c_float, c_double, c_size_t, c_longlong, byref as ctypes_byref

F32, BF16 = 0, 1


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise _cabi.GraphEchoNativeError(f"unsupported activation dtype {t.dtype} (fp32 or bf16)")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    """fp32 + contiguous (no copy when already so)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _cabi.GraphEchoNativeError("graphecho_b200 ops run on CUDA tensors only (no CPU fallback)")


# ------------------------------------------------------------------------------------------ K3
class _AffinityPairwise(Function):
    @staticmethod
    def forward(ctx, A, B, w2, b2):
        _need_cuda(A, B, w2, b2)
        squeeze = A.dim() == 2
        A3 = _f32c(A if not squeeze else A.unsqueeze(0))
        B3 = _f32c(B if not squeeze else B.unsqueeze(0))
        w2c, b2c = _f32c(w2).view(-1), _f32c(b2).view(-1)
        batch, N1, H = A3.shape
        N2 = B3.shape[1]
        M = torch.empty((batch, N1, N2), device=A.device, dtype=torch.float32)
        call("ge_affinity_pairwise_fwd", ptr(A3), ptr(B3), ptr(w2c), ptr(b2c), ptr(M),
             batch, N1, N2, H, stream(),
             work=(4 * batch * (H * (N1 + N2) + N1 * N2), 3 * batch * H * N1 * N2))
        ctx.save_for_backward(A3, B3, w2c)
        ctx.squeeze = squeeze
        ctx.shapes = (w2.shape, b2.shape)
        return M[0] if squeeze else M

    @staticmethod
    @once_differentiable
    def backward(ctx, dM):
        A3, B3, w2c = ctx.saved_tensors
        batch, N1, H = A3.shape
        N2 = B3.shape[1]
        dM3 = _f32c(dM if not ctx.squeeze else dM.unsqueeze(0))
        dA, dB = torch.empty_like(A3), torch.empty_like(B3)
        dw2 = torch.empty_like(w2c)
        db2 = torch.empty(1, device=A3.device, dtype=torch.float32)
        nbytes = _cabi.lib().ge_affinity_pairwise_bwd_workspace_bytes(batch, N1, N2, H)
        ws = torch.empty(max(nbytes, 4), device=A3.device, dtype=torch.uint8)
        call("ge_affinity_pairwise_bwd", ptr(A3), ptr(B3), ptr(w2c), ptr(dM3), ptr(dA), ptr(dB),
             ptr(dw2), ptr(db2), ptr(ws), c_size_t(ws.numel()), batch, N1, N2, H, stream(),
             work=(4 * batch * (2 * H * (N1 + N2) + N1 * N2), 8 * batch * H * N1 * N2))
        if ctx.squeeze:
            dA, dB = dA[0], dB[0]
        return dA, dB, dw2.view(ctx.shapes[0]), db2.view(ctx.shapes[1])


def affinity_pairwise(A, B, w2, b2):
    """M[i,j] = sum_k w2[k] relu(A[i,k] + B[j,k]) + b2.   A [N1,H] / [b,N1,H], B likewise."""
    return _AffinityPairwise.apply(A, B, w2, b2)


# ------------------------------------------------------------------------------------------ K4
class _SinkhornRpm(Function):
    @staticmethod
    def forward(ctx, M, n_iters, instnorm, cluster_size):
        _need_cuda(M)
        squeeze = M.dim() == 2
        M3 = _f32c(M if not squeeze else M.unsqueeze(0))
        batch, N1, N2 = M3.shape
        dev = M.device
        P = torch.empty_like(M3)
        hist_r = torch.empty((batch, max(n_iters, 1), N1), device=dev, dtype=torch.float32)
        hist_c = torch.empty((batch, max(n_iters, 1), N2), device=dev, dtype=torch.float32)
        stats = torch.empty((batch, 4), device=dev, dtype=torch.float32)
        call("ge_sinkhorn_rpm_fwd", ptr(M3), ptr(P), ptr(hist_r), ptr(hist_c), ptr(stats),
             batch, N1, N2, int(n_iters), int(bool(instnorm)), int(cluster_size), stream(),
             work=(batch * (8 * N1 * N2 + 4 * int(n_iters) * (N1 + N2)), batch * N1 * N2 * (8 * int(n_iters) + 6)))
        ctx.save_for_backward(M3, hist_r, hist_c, stats)
        ctx.cfg = (int(n_iters), int(bool(instnorm)), int(cluster_size), squeeze)
        return P[0] if squeeze else P

    @staticmethod
    @once_differentiable
    def backward(ctx, G):
        M3, hist_r, hist_c, stats = ctx.saved_tensors
        n_iters, instnorm, cluster_size, squeeze = ctx.cfg
        batch, N1, N2 = M3.shape
        G3 = _f32c(G if not squeeze else G.unsqueeze(0))
        dM = torch.empty_like(M3)
        call("ge_sinkhorn_rpm_bwd", ptr(M3), ptr(G3), ptr(hist_r), ptr(hist_c), ptr(stats), ptr(dM),
             batch, N1, N2, n_iters, instnorm, cluster_size, stream(),
             work=(batch * (12 * N1 * N2 + 4 * n_iters * (N1 + N2)), batch * N1 * N2 * (12 * n_iters + 10)))
        return (dM[0] if squeeze else dM), None, None, None


def sinkhorn_rpm_exp(M, n_iters=20, instnorm=True, cluster_size=0):
    """exp(sinkhorn_rpm(InstanceNorm(M), n_iters, slack=True)) for M [N1,N2] or [b,N1,N2]."""
    return _SinkhornRpm.apply(M, n_iters, instnorm, cluster_size)


class _MatchingLossO2O(Function):
    @staticmethod
    def forward(ctx, P, lab1, lab2, alpha, gamma):
        _need_cuda(P, lab1, lab2)
        Pc, l1, l2 = _f32c(P), _f32c(lab1), _f32c(lab2)
        N1, N2 = Pc.shape
        dev = P.device
        loss = torch.empty(1, device=dev, dtype=torch.float32)
        idx = torch.empty(N1, device=dev, dtype=torch.int32)
        stats = torch.empty(4, device=dev, dtype=torch.float32)
        call("ge_matching_loss_fwd", ptr(Pc), ptr(l1), ptr(l2), ptr(loss), ptr(idx), ptr(stats), N1, N2,
             c_float(alpha), c_float(gamma), stream(), work=(4 * N1 * N2, 12 * N1 * N2))
        ctx.save_for_backward(Pc, l1, l2, idx, stats)
        ctx.cfg = (float(alpha), float(gamma))
        return loss[0]

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        Pc, l1, l2, idx, stats = ctx.saved_tensors
        alpha, gamma = ctx.cfg
        N1, N2 = Pc.shape
        dP = torch.empty_like(Pc)
        gc = _f32c(g).reshape(1)
        call("ge_matching_loss_bwd", ptr(Pc), ptr(l1), ptr(l2), ptr(idx), ptr(stats), ptr(gc), ptr(dP), N1, N2,
             c_float(alpha), c_float(gamma), stream(), work=(8 * N1 * N2, 14 * N1 * N2))
        return dP, None, None, None, None


def matching_loss_o2o(P, labels_1, labels_2, alpha=0.25, gamma=2.0):
    """TP + FP focal losses of GModule._forward_aff ('o2o', graph_matching.py:572-590) on the Sinkhorn-normalised P."""
    return _MatchingLossO2O.apply(P, labels_1, labels_2, alpha, gamma)


# ------------------------------------------------------------------------------------------ K5
class _SinkhornDistance(Function):
    @staticmethod
    def forward(ctx, x, y, eps, max_iter, thresh):
        _need_cuda(x, y)
        x3, y3 = _f32c(x), _f32c(y)
        B, P1, D = x3.shape
        P2 = y3.shape[1]
        dev = x.device
        C = torch.empty((B, P1, P2), device=dev, dtype=torch.float32)
        pi = torch.empty_like(C)
        cost = torch.empty(B, device=dev, dtype=torch.float32)
        mi = max(int(max_iter), 1)
        hist_u = torch.empty((B, mi, P1), device=dev, dtype=torch.float32)
        hist_v = torch.empty((B, mi, P2), device=dev, dtype=torch.float32)
        err = torch.empty((B, mi), device=dev, dtype=torch.float32)
        nits = torch.zeros(1, device=dev, dtype=torch.int32)
        call("ge_sinkhorn_distance_fwd", ptr(x3), ptr(y3), ptr(C), ptr(pi), ptr(cost), ptr(hist_u),
             ptr(hist_v), ptr(err), ptr(nits), B, P1, P2, D, c_float(eps), int(max_iter),
             c_double(thresh), stream(),
             work=(4 * B * D * (P1 + P2) + 8 * B * P1 * P2, 3 * B * P1 * P2 * D + 10 * B * P1 * P2 * int(max_iter)))
        ctx.save_for_backward(x3, y3, C, hist_u, hist_v, nits)
        ctx.cfg = (float(eps), int(max_iter))
        ctx.mark_non_differentiable(pi, C, nits)
        return cost, pi, C, nits

    @staticmethod
    @once_differentiable
    def backward(ctx, gcost, gpi, gC, gn):
        x3, y3, C, hist_u, hist_v, nits = ctx.saved_tensors
        eps, max_iter = ctx.cfg
        B, P1, D = x3.shape
        P2 = y3.shape[1]
        g = _f32c(gcost)
        dC = torch.empty_like(C)
        gemm = P1 * P2 > 128 * 128          # large node sets: dC -> (dx, dy) is two GEMMs, not a one-CTA loop
        dx, dy = (None, None) if gemm else (torch.empty_like(x3), torch.empty_like(y3))
        call("ge_sinkhorn_distance_bwd", ptr(x3), ptr(y3), ptr(C), ptr(hist_u), ptr(hist_v), ptr(nits),
             ptr(g), ptr(dC), ptr(dx), ptr(dy), B, P1, P2, D, c_float(eps), max_iter, stream(),
             work=(8 * B * D * (P1 + P2) + 8 * B * P1 * P2, 4 * B * P1 * P2 * D + 20 * B * P1 * P2 * max_iter))
        if gemm:                            # C_ij = sum_d (x_id - y_jd)^2
            dx = 2.0 * (dC.sum(2, keepdim=True) * x3 - torch.bmm(dC, y3))
            dy = 2.0 * (dC.sum(1).unsqueeze(2) * y3 - torch.bmm(dC.transpose(1, 2), x3))
        return dx, dy, None, None, None


def sinkhorn_distance(x, y, eps, max_iter, thresh=0.1):
    """x [B,P1,D], y [B,P2,D] -> (cost [B], pi, C, nits).  Gradients flow through `cost`."""
    return _SinkhornDistance.apply(x, y, eps, max_iter, thresh)


# ------------------------------------------------------------------------------------------ K1
@torch.no_grad()
def knn_graph(x, y=None, k=9, dilation=1, relative_pos=None):
    """x [B,C,N(,1)], y [B,C,M(,1)] or None -> int64 edge_index [2,B,N,k] (vig.py:369-381)."""
    _need_cuda(x, y, relative_pos)
    B, C = x.shape[0], x.shape[1]
    x3 = _f32c(x.reshape(B, C, -1))
    N = x3.shape[2]
    y3 = None
    M = N
    if y is not None:
        y3 = _f32c(y.reshape(B, C, -1))
        M = y3.shape[2]
    rel = None
    if relative_pos is not None:
        rel = _f32c(relative_pos.reshape(-1, relative_pos.shape[-2], relative_pos.shape[-1]))
        if rel.shape[0] != 1 or rel.shape[1] != N or rel.shape[2] != M:
            raise _cabi.GraphEchoNativeError(f"relative_pos must be [1,{N},{M}], got {tuple(relative_pos.shape)}")
    out = torch.empty((2, B, N, k), device=x.device, dtype=torch.int64)
    nbytes = _cabi.lib().ge_knn_graph_workspace_bytes(B, C, N, M)
    ws = torch.empty(max(nbytes, 4), device=x.device, dtype=torch.uint8)
    call("ge_knn_graph", ptr(x3), ptr(y3), ptr(rel), ptr(out), ptr(ws), c_size_t(ws.numel()),
         B, C, N, M, int(k), int(dilation), stream(),
         work=(4 * B * C * (N + M) + 16 * B * N * int(k), 2 * B * N * M * C))
    return out


# ------------------------------------------------------------------------------------------ K2
class _MRGather(Function):
    @staticmethod
    def forward(ctx, x, y, idx_nbr, idx_ctr):
        _need_cuda(x, y, idx_nbr, idx_ctr)
        B, C = x.shape[0], x.shape[1]
        x3 = _f32c(x.reshape(B, C, -1))
        N = x3.shape[2]
        y3, M = None, N
        if y is not None:
            y3 = _f32c(y.reshape(B, C, -1))
            M = y3.shape[2]
        i0 = idx_nbr.contiguous()
        i1 = idx_ctr.contiguous() if idx_ctr is not None else None
        k = i0.shape[-1]
        out = torch.empty((B, 2 * C, N), device=x.device, dtype=torch.float32)
        argk = torch.empty((B, C, N), device=x.device, dtype=torch.uint8)
        call("ge_mrconv_gather_fwd", ptr(x3), ptr(y3), ptr(i0), ptr(i1), ptr(out), ptr(argk),
             B, C, N, M, k, stream(),
             work=(4 * B * C * (N + (M if y is not None else 0)) + 8 * B * N * k + 9 * B * C * N, 2 * B * C * N * k))
        ctx.save_for_backward(i0, i1 if i1 is not None else i0, argk)
        ctx.cfg = (B, C, N, M, k, i1 is not None, y is not None, x.shape, None if y is None else y.shape)
        ctx.mark_non_differentiable(argk)
        return out, argk

    @staticmethod
    @once_differentiable
    def backward(ctx, dout, _):
        i0, i1, argk = ctx.saved_tensors
        B, C, N, M, k, has_ctr, has_y, xshape, yshape = ctx.cfg
        d = _f32c(dout)
        dx = torch.empty((B, C, N), device=d.device, dtype=torch.float32)
        dy = torch.zeros((B, C, M), device=d.device, dtype=torch.float32) if has_y else None
        call("ge_mrconv_gather_bwd", ptr(d), ptr(i0), ptr(i1 if has_ctr else None), ptr(argk), ptr(dx), ptr(dy),
             B, C, N, M, k, stream(),
             work=(13 * B * C * N + 8 * B * N * k + (4 * B * C * M if has_y else 0), 2 * B * C * N))
        return dx.view(xshape), (dy.view(yshape) if has_y else None), None, None


def mr_gather(x, edge_index, y=None, identity_centre=True):
    """Channel-interleaved [x ; max_k(y_j - x_i)]  ->  [B, 2C, N, 1]  (vig.py:96-104)."""
    idx_nbr = edge_index[0]
    idx_ctr = None if identity_centre else edge_index[1]
    out, _ = _MRGather.apply(x, y, idx_nbr, idx_ctr)
    return out.unsqueeze(-1)


# ------------------------------------------------------------------------------------------ K1/K2 node-major
def knn_nmajor_supported(B, C, N, M, k, dilation):
    return bool(_cabi.lib().ge_knn_graph_nmajor_supported(int(B), int(C), int(N), int(M), int(k), int(dilation)))


@torch.no_grad()
def knn_graph_nmajor(x, y=None, k=9, dilation=1):
    """Node-major k-NN: x [B,N,C], y [B,M,C] or None (fp32 / bf16, dense) -> int64 edge_index [2,B,N,k]
    (same contents as knn_graph on the [B,C,N] transposes; vig.py:369-381).  tcgen05 kernel only."""
    _need_cuda(x, y)
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    x = x.contiguous()
    B, N, C = x.shape
    M = N
    if y is not None:
        y = y.to(x.dtype).contiguous()
        M = y.shape[1]
    out = torch.empty((2, B, N, k), device=x.device, dtype=torch.int64)
    nbytes = _cabi.lib().ge_knn_graph_workspace_bytes(B, C, N, M)
    ws = torch.empty(max(nbytes, 4), device=x.device, dtype=torch.uint8)
    call("ge_knn_graph_nmajor", ptr(x), ptr(y), _dtype_code(x), ptr(out), ptr(ws), c_size_t(ws.numel()),
         B, C, N, M, int(k), int(dilation), stream(),
         work=(x.element_size() * B * C * (N + M) + 16 * B * N * int(k), 2 * B * N * M * C))
    return out


class _MRGatherNMajor(Function):
    @staticmethod
    def forward(ctx, x, y, idx_nbr):
        _need_cuda(x, y, idx_nbr)
        x = x.contiguous()
        B, N, C = x.shape
        M = N
        if y is not None:
            y = y.to(x.dtype).contiguous()
            M = y.shape[1]
        i0 = idx_nbr.contiguous()
        k = i0.shape[-1]
        out = torch.empty((B, N, 2 * C), device=x.device, dtype=x.dtype)
        argk = torch.empty((B, N, C), device=x.device, dtype=torch.uint8)
        es = x.element_size()
        call("ge_mrconv_gather_nmajor_fwd", ptr(x), ptr(y), ptr(i0), ptr(out), ptr(argk), _dtype_code(x),
             B, C, N, M, k, stream(),
             work=(es * B * C * (N + (M if y is not None else 0)) + 8 * B * N * k + (2 * es + 1) * B * C * N, 2 * B * C * N * k))
        ctx.save_for_backward(i0, argk)
        ctx.cfg = (B, C, N, M, k, y is not None, x.dtype)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        i0, argk = ctx.saved_tensors
        B, C, N, M, k, has_y, dt = ctx.cfg
        d = dout.to(dt).contiguous()
        es = d.element_size()
        if not has_y and C % 32 == 0 and N * 32 * (es + 1) + (2 * N + 1) * 4 + N * k * 3 + 16 <= 110 * 1024 and N <= 65535 and B <= 65535:
            dxs = torch.empty((B, N, C), device=d.device, dtype=dt)
            call("ge_mrconv_gather_nmajor_bwd_self", ptr(d), ptr(i0), ptr(argk), ptr(dxs), _dtype_code(d), B, C, N, k, stream(),
                 work=((2 * es + 1 + es) * B * C * N + 8 * B * N * k, 2 * B * C * N))       # dout, argk in; dx out
            return dxs, None, None
        dx = torch.empty((B, N, C), device=d.device, dtype=torch.float32)
        dy = torch.zeros((B, M, C), device=d.device, dtype=torch.float32) if has_y else None
        es = d.element_size()
        call("ge_mrconv_gather_nmajor_bwd", ptr(d), ptr(i0), ptr(argk), ptr(dx), ptr(dy), _dtype_code(d),
             B, C, N, M, k, stream(), work=((2 * es + 1 + 8) * B * C * N + 8 * B * N * k, 2 * B * C * N))
        return dx.to(dt), (dy.to(dt) if has_y else None), None


def mr_gather_nmajor(x, idx_nbr, y=None):
    """Node-major max-relative aggregation (vig.py:96-104 with the centre = the point itself):
    x [B,N,C], idx_nbr [B,N,k] -> [B,N,2C] with the reference's channel interleaving."""
    return _MRGatherNMajor.apply(x, y, idx_nbr)


# ------------------------------------------------------------------------------------------ K6
def _nhwc_view(t: torch.Tensor):
    """[F,C,H,W] logical tensor -> NHWC-dense tensor sharing storage when already channels_last."""
    return t.contiguous(memory_format=torch.channels_last)


class _PoolConcat(Function):
    @staticmethod
    def forward(ctx, rs, *levels):
        _need_cuda(*levels)
        lv = [_nhwc_view(t) for t in levels]
        F_ = lv[0].shape[0]
        Ctot = sum(t.shape[1] for t in lv)
        Ho, Wo = lv[0].shape[2] // rs[0], lv[0].shape[3] // rs[0]
        out = torch.empty((F_, Ctot, Ho, Wo), device=lv[0].device, dtype=torch.float32,
                          memory_format=torch.channels_last)
        coff = 0
        meta = []
        for t, r in zip(lv, rs):
            _, C, H, W = t.shape
            if (H // r, W // r) != (Ho, Wo):
                raise _cabi.GraphEchoNativeError(
                    f"pooled sizes differ: level {tuple(t.shape)} with r={r} vs {(Ho, Wo)} (TGCN.py:64-70 would fail in torch.cat)")
            call("ge_tgcn_pool_concat_fwd", ptr(t), c_longlong(H * W * C), ptr(out), _dtype_code(t),
                 c_longlong(F_), H, W, C, int(r), Ctot, coff, stream(),
                 work=(F_ * C * (H * W * t.element_size() + 4 * Ho * Wo), F_ * C * H * W))
            meta.append((t.shape, t.dtype, int(r), coff))
            coff += C
        ctx.meta = meta
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        d = dout.float().contiguous(memory_format=torch.channels_last)
        F_, Ctot = d.shape[0], d.shape[1]
        grads = []
        for shape, dtype, r, coff in ctx.meta:
            _, C, H, W = shape
            g = torch.empty(shape, device=d.device, dtype=dtype, memory_format=torch.channels_last)
            call("ge_tgcn_pool_concat_bwd", ptr(d), ptr(g), _dtype_code(g), c_longlong(F_), H, W, C, r, Ctot, coff, stream(),
                 work=(F_ * C * (H * W * g.element_size() + 4 * (H // r) * (W // r)), F_ * C * H * W))
            grads.append(g)
        return (None, *grads)


def pool_concat(levels, rs):
    """avg_pool2d(level_l, r_l) for every level, concatenated on channels -> fp32 channels_last."""
    return _PoolConcat.apply(tuple(int(r) for r in rs), *levels)


def tgcn_recurrence_supported(C, Cout, N, k, dilation, groups):
    return bool(_cabi.lib().ge_tgcn_recurrence_supported(int(C), int(Cout), int(N), int(k), int(dilation), int(groups)))


class _TgcnRecurrence(Function):
    @staticmethod
    def forward(ctx, emb, weight, bias, k):
        _need_cuda(emb, weight, bias)
        e = _f32c(emb)
        B, T, C, N = e.shape
        W2 = _f32c(weight).reshape(weight.shape[0], -1)                  # [Cout, 2C/groups]
        Wt = W2.t().contiguous()
        bz = _f32c(bias)
        dev = e.device
        hidden_all = torch.empty((B, T, C, N), device=dev, dtype=torch.float32)
        z_all = torch.empty_like(hidden_all)
        idx_all = torch.empty((B, T, N, k), device=dev, dtype=torch.int32)
        argk_all = torch.empty((B, T, C, N), device=dev, dtype=torch.uint8)
        call("ge_tgcn_recurrence_fwd", ptr(e), ptr(Wt), ptr(bz), ptr(hidden_all), ptr(z_all), ptr(idx_all), ptr(argk_all),
             B, T, C, N, int(k), stream(),
             work=(B * T * C * N * 4 * 3 + Wt.numel() * 4, 2 * B * T * (N * N * C + N * C * W2.shape[1]) + 2 * B * T * C * N * int(k)))
        ctx.save_for_backward(e, W2, hidden_all, z_all, idx_all, argk_all)
        ctx.cfg = (int(k), weight.shape, weight.dtype, bias.dtype)
        ctx.mark_non_differentiable(idx_all)
        return hidden_all[:, T - 1], idx_all

    @staticmethod
    @once_differentiable
    def backward(ctx, dH, _):
        e, W2, hidden_all, z_all, idx_all, argk_all = ctx.saved_tensors
        k, wshape, wdt, bdt = ctx.cfg
        B, T, C, N = e.shape
        dev = e.device
        g = _f32c(dH)
        d_emb = torch.empty_like(e)
        dWt_part = torch.zeros((B, W2.shape[1], C), device=dev, dtype=torch.float32)
        db_part = torch.zeros((B, C), device=dev, dtype=torch.float32)
        scratch = torch.empty((B, C, N), device=dev, dtype=torch.float32)
        call("ge_tgcn_recurrence_bwd", ptr(e), ptr(W2), ptr(hidden_all), ptr(z_all), ptr(idx_all), ptr(argk_all), ptr(g),
             ptr(d_emb), ptr(dWt_part), ptr(db_part), ptr(scratch), B, T, C, N, k, stream(),
             work=(B * T * C * N * 4 * 4, 4 * B * T * N * C * W2.shape[1]))
        dW = dWt_part.sum(0).t().reshape(wshape).to(wdt)
        return d_emb, dW, db_part.sum(0).to(bdt), None


def tgcn_recurrence(emb, weight, bias, k=9):
    """The TGCN time loop (TGCN.py:224-235 over :62-78, state-independent part hoisted out) in one persistent launch:
    emb [B,T,C,N] -> (final hidden state [B,C,N], neighbour lists int32 [B,T,N,k])."""
    return _TgcnRecurrence.apply(emb, weight, bias, int(k))


# ------------------------------------------------------------------------------------------ K7
class _UpsampleAdd(Function):
    @staticmethod
    def forward(ctx, top, lateral, size):
        _need_cuda(top, lateral)
        t = _nhwc_view(top)
        N, C, h, w = t.shape
        H, W = size
        lat = None
        if lateral is not None:
            lat = _nhwc_view(lateral.to(t.dtype))
        out = torch.empty((N, C, H, W), device=t.device, dtype=t.dtype, memory_format=torch.channels_last)
        es = t.element_size()
        call("ge_upsample_add_fwd", ptr(t), ptr(lat), ptr(out), _dtype_code(t), N, h, w, H, W, C, stream(),
             work=(N * C * es * (h * w + (2 if lat is not None else 1) * H * W), 8 * N * C * H * W))
        ctx.cfg = (N, C, h, w, H, W, lateral is not None)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        N, C, h, w, H, W, has_lat = ctx.cfg
        d = _nhwc_view(dout)
        dtop = torch.empty((N, C, h, w), device=d.device, dtype=d.dtype, memory_format=torch.channels_last)
        call("ge_upsample_bwd", ptr(d), ptr(dtop), _dtype_code(d), N, h, w, H, W, C, stream(),
             work=(N * C * d.element_size() * (h * w + H * W), 2 * N * C * H * W))
        return dtop, (d if has_lat else None), None


def upsample_add(top, lateral):
    """bilinear(align_corners=True) up-sampling of `top` to lateral's size, plus lateral."""
    return _UpsampleAdd.apply(top, lateral, (lateral.shape[2], lateral.shape[3]))


def upsample_bilinear(x, size):
    return _UpsampleAdd.apply(x, None, (int(size[0]), int(size[1])))


class _GnReluUpsample(Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, size, eps, groups, pre_bias=None):
        _need_cuda(x, gamma, beta, pre_bias)
        xc = _nhwc_view(x)
        N, C, h, w = xc.shape
        H, W = size
        cpg = C // groups
        g, b = _f32c(gamma), _f32c(beta)
        mean = torch.empty((N, C), device=x.device, dtype=torch.float32)
        rstd = torch.empty_like(mean)
        code = _dtype_code(xc)
        es = xc.element_size()
        csum = None
        if pre_bias is None:
            call("ge_group_stats", ptr(xc), ptr(mean), ptr(rstd), code, N, h * w, C, cpg, c_float(eps), stream(),
                 work=(N * C * h * w * es, 3 * N * C * h * w))
        else:
            # statistics of x + pre_bias; `mean` is the effective mean (group mean - pre_bias[c]), so everything
            # downstream normalises the raw x
            csum = torch.empty_like(mean)
            call("ge_group_stats_bias", ptr(xc), ptr(_f32c(pre_bias)), ptr(mean), ptr(rstd), ptr(csum), code,
                 N, h * w, C, cpg, c_float(eps), stream(), work=(N * C * h * w * es, 3 * N * C * h * w))
        out = torch.empty((N, C, H, W), device=x.device, dtype=xc.dtype, memory_format=torch.channels_last)
        call("ge_gn_relu_upsample_fwd", ptr(xc), ptr(mean), ptr(rstd), ptr(g), ptr(b), ptr(out), code,
             N, h, w, H, W, C, stream(), work=(N * C * es * (h * w + H * W), 12 * N * C * H * W))
        ctx.save_for_backward(xc, mean, rstd, g, b, csum)
        ctx.cfg = (N, C, h, w, H, W, cpg, gamma.dtype, beta.dtype, None if pre_bias is None else pre_bias.dtype)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        xc, mean, rstd, g, b, csum = ctx.saved_tensors
        N, C, h, w, H, W, cpg, gdt, bdt, pdt = ctx.cfg
        d = _nhwc_view(dout.to(xc.dtype))
        es = xc.element_size()
        if (h, w) != (H, W):
            # adjoint of the up-sampling first (the source map is small), then the same-size backward
            dact = torch.empty_like(xc)
            call("ge_upsample_bwd", ptr(d), ptr(dact), _dtype_code(d), N, h, w, H, W, C, stream(),
                 work=(N * C * es * (h * w + H * W), 2 * N * C * H * W))
            d = dact
        S = torch.empty((4, N, C), device=d.device, dtype=torch.float32)
        dx = torch.empty_like(xc)
        call("ge_gn_relu_upsample_bwd", ptr(d), ptr(xc), ptr(mean), ptr(rstd), ptr(g), ptr(b), ptr(None),
             ptr(S[0]), ptr(S[1]), ptr(S[2]), ptr(S[3]), ptr(dx), _dtype_code(xc), N, h, w, h, w, C, cpg, stream(),
             work=(N * C * es * 3 * h * w, 20 * N * C * h * w))          # compulsory: dy, x in; dx out
        dpb = None
        if csum is not None:
            # d/d pre_bias = sum_{n,hw} dx, in closed form from the [N,C] arrays of the two kernels (no pass over dx):
            # sum_hw dx = rstd*(gamma*S1 - HW*A1 - A2*sum_hw xhat),  sum_hw xhat = rstd*(sum_hw x - HW*mean_eff)
            hw = float(h * w)
            sum_xhat = rstd * (csum - hw * mean)
            dpb = (rstd * (g * S[0] - hw * S[2] - S[3] * sum_xhat)).sum(0).to(pdt)
        return dx, S[1].sum(0).to(gdt), S[0].sum(0).to(bdt), None, None, None, dpb


def gn_relu_upsample(x, gamma, beta, size, eps=1e-5, groups=None):
    """_upsample(relu(GroupNorm(groups, C)(x)), size), align_corners=True; groups=None means one
    channel per group (the FPN head's GroupNorm(C,C), fpnseg.py:428-442)."""
    groups = x.shape[1] if groups is None else int(groups)
    return _GnReluUpsample.apply(x, gamma, beta, (int(size[0]), int(size[1])), float(eps), groups)


def gn_relu(x, gamma, beta, groups, eps=1e-5, pre_bias=None):
    """relu(GroupNorm(groups, C)(x + pre_bias)) on an NHWC map (Discriminator towers, fpnseg.py:455-466).
    `pre_bias` [C] is the bias of the convolution that produced x: folded into the statistics kernel, so the
    convolution runs without its bias-add pass and autograd without the bias-gradient reduction pass."""
    return _GnReluUpsample.apply(x, gamma, beta, (int(x.shape[2]), int(x.shape[3])), float(eps), int(groups), pre_bias)


class _MaxPool3s2(Function):
    @staticmethod
    def forward(ctx, x):
        _need_cuda(x)
        xc = _nhwc_view(x)
        N, C, H, W = xc.shape
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        out = torch.empty((N, C, Ho, Wo), device=x.device, dtype=xc.dtype, memory_format=torch.channels_last)
        arg = torch.empty((N, Ho, Wo, C), device=x.device, dtype=torch.uint8)
        es = xc.element_size()
        call("ge_maxpool3s2_fwd", ptr(xc), ptr(out), ptr(arg), _dtype_code(xc), N, H, W, C, stream(),
             work=(N * C * (es * H * W + (es + 1) * Ho * Wo), 9 * N * C * Ho * Wo))
        ctx.save_for_backward(arg)
        ctx.cfg = (N, C, H, W, xc.dtype)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        (arg,) = ctx.saved_tensors
        N, C, H, W, dt = ctx.cfg
        d = _nhwc_view(dout.to(dt))
        dx = torch.empty((N, C, H, W), device=d.device, dtype=dt, memory_format=torch.channels_last)
        es = d.element_size()
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        call("ge_maxpool3s2_bwd", ptr(d), ptr(arg), ptr(dx), _dtype_code(d), N, H, W, C, stream(),
             work=(N * C * (es * H * W + (es + 1) * Ho * Wo), 4 * N * C * H * W))
        return dx


def maxpool3s2(x):
    """nn.MaxPool2d(3, 2, 1) on an NHWC map (ResNet stem, fpnseg.py:232): one pass forward, a deterministic gather
    backward."""
    return _MaxPool3s2.apply(x)


# Per-domain BatchNorm statistics.  The reference trainers run the network on the source batch and on the target
# batch as two separate train-mode calls (train_cardiac_uda.py:225, 234; train_camus_echo.py:208, 218): every
# BatchNorm sees each domain's own batch statistics and updates its running statistics twice per step.  The engine
# sends [source | target] through the convolutions as ONE batch; inside `domain_split(n_source_frames)` every fused
# BatchNorm call treats the first n_source_frames images and the rest as two segments (statistics, running-stat
# updates in source -> target order, backward sums), which is the two-call semantics at one-call cost.
_BN_SPLIT = 0


class domain_split:
    def __init__(self, n_first):
        self.n_first = int(n_first or 0)

    def __enter__(self):
        global _BN_SPLIT
        self.prev, _BN_SPLIT = _BN_SPLIT, self.n_first
        return self

    def __exit__(self, *exc):
        global _BN_SPLIT
        _BN_SPLIT = self.prev
        return False


def bn_segments(n_images):
    """Number of BatchNorm segments a batch of `n_images` is normalised in under the current domain_split."""
    return 2 if 0 < _BN_SPLIT < n_images else 1


class _BnAct(Function):
    """Training-mode fused BatchNorm2d (+ residual) (+ ReLU) on an NHWC map."""

    @staticmethod
    def forward(ctx, x, residual, gamma, beta, running_mean, running_var, nbt, momentum, eps, relu, split):
        _need_cuda(x, residual, gamma, beta)
        xc = _nhwc_view(x)
        N, C, H, W = xc.shape
        P = N * H * W
        P_split = split * H * W if 0 < split < N else 0
        nseg = 2 if P_split else 1
        rc = _nhwc_view(residual.to(xc.dtype)) if residual is not None else None
        g, b = _f32c(gamma), _f32c(beta)
        dev = x.device
        out = torch.empty_like(xc)
        save = torch.empty((2, nseg, C), device=dev, dtype=torch.float32)
        nbytes = _cabi.lib().ge_bn_workspace_bytes(P, C)
        if nbytes == 0:
            raise _cabi.GraphEchoNativeError(f"fused BatchNorm does not support C={C}")
        ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        mask = None
        if relu:
            mask = torch.empty(_cabi.lib().ge_bn_relu_mask_bytes(P, C), device=dev, dtype=torch.uint8)
        es = xc.element_size()
        call("ge_bn_fwd_train", ptr(xc), ptr(rc), ptr(g), ptr(b), ptr(running_mean), ptr(running_var), ptr(nbt),
             c_float(momentum), c_float(eps), ptr(out), ptr(save[0]), ptr(save[1]), ptr(mask), ptr(ws), c_size_t(nbytes),
             _dtype_code(xc), c_longlong(P), c_longlong(P_split), C, int(relu), stream(),
             work=(P * C * es * (2 + (1 if rc is not None else 0)), 8 * P * C))      # compulsory: x (+res) in, out
        ctx.save_for_backward(xc, mask, g, save)
        ctx.cfg = (P, P_split, C, bool(relu), rc is not None, gamma.dtype, beta.dtype)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        xc, mask, g, save = ctx.saved_tensors
        P, P_split, C, relu, has_res, gdt, bdt = ctx.cfg
        d = _nhwc_view(dout.to(xc.dtype))
        dx = torch.empty_like(xc)
        dres = torch.empty_like(xc) if has_res else None
        dgb = torch.empty((2, C), device=d.device, dtype=torch.float32)
        nbytes = _cabi.lib().ge_bn_workspace_bytes(P, C)
        ws = torch.empty(nbytes, device=d.device, dtype=torch.uint8)
        es = xc.element_size()
        call("ge_bn_bwd", ptr(d), ptr(mask), ptr(xc), ptr(g), ptr(save[0]), ptr(save[1]), ptr(dx),
             ptr(dres), ptr(dgb[0]), ptr(dgb[1]), ptr(ws), c_size_t(nbytes), _dtype_code(xc), c_longlong(P),
             c_longlong(P_split), C, int(relu), stream(),
             work=(P * C * es * (3 + (1 if has_res else 0)) + (P * C // 8 if relu else 0), 16 * P * C))  # dy, x, mask in; dx (+dres) out
        return dx, dres, dgb[0].to(gdt), dgb[1].to(bdt), None, None, None, None, None, None, None


class _SyncBnAct(Function):
    """Training-mode SyncBatchNorm (+ residual) (+ ReLU) on an NHWC map: the fused BatchNorm kernels with the two small
    cross-rank exchanges (per-segment sums, averaged over the process group) between their statistics and apply stages.
    What nn.SyncBatchNorm does under DDP (the reference's intent, train_cardiac_uda.py:142), per domain segment; the
    all-reduces are NCCL collectives on the current stream (run eagerly: engine.capture_graphs refuses sync_bn)."""

    @staticmethod
    def forward(ctx, x, residual, gamma, beta, running_mean, running_var, nbt, momentum, eps, relu, split, group, world):
        import torch.distributed as dist
        _need_cuda(x, residual, gamma, beta)
        xc = _nhwc_view(x)
        N, C, H, W = xc.shape
        P = N * H * W
        P_split = split * H * W if 0 < split < N else 0
        nseg = 2 if P_split else 1
        rc = _nhwc_view(residual.to(xc.dtype)) if residual is not None else None
        g, b = _f32c(gamma), _f32c(beta)
        dev = x.device
        nbytes = _cabi.lib().ge_bn_workspace_bytes(P, C)
        if nbytes == 0:
            raise _cabi.GraphEchoNativeError(f"fused BatchNorm does not support C={C}")
        ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        sums = torch.empty((nseg, 2, C), device=dev, dtype=torch.float32)
        dt = _dtype_code(xc)
        es = xc.element_size()
        call("ge_bn_sync_stats", ptr(xc), ptr(running_mean), ptr(sums), ptr(ws), c_size_t(nbytes), dt, c_longlong(P),
             c_longlong(P_split), C, stream(), work=(P * C * es, 3 * P * C))
        if world > 1:
            dist.all_reduce(sums, op=dist.ReduceOp.AVG, group=group)
        out = torch.empty_like(xc)
        save = torch.empty((2, nseg, C), device=dev, dtype=torch.float32)
        mask = torch.empty(_cabi.lib().ge_bn_relu_mask_bytes(P, C), device=dev, dtype=torch.uint8) if relu else None
        call("ge_bn_sync_fwd_apply", ptr(xc), ptr(rc), ptr(g), ptr(b), ptr(running_mean), ptr(running_var), ptr(nbt),
             c_float(momentum), c_float(eps), ptr(out), ptr(save[0]), ptr(save[1]), ptr(mask), ptr(sums), int(world),
             dt, c_longlong(P), c_longlong(P_split), C, int(relu), stream(),
             work=(P * C * es * (2 + (1 if rc is not None else 0)), 5 * P * C))
        ctx.save_for_backward(xc, mask, g, save)
        ctx.cfg = (P, P_split, C, bool(relu), rc is not None, gamma.dtype, beta.dtype, group, int(world))
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        import torch.distributed as dist
        xc, mask, g, save = ctx.saved_tensors
        P, P_split, C, relu, has_res, gdt, bdt, group, world = ctx.cfg
        nseg = 2 if P_split else 1
        d = _nhwc_view(dout.to(xc.dtype))
        dev = d.device
        dgb = torch.empty((2, C), device=dev, dtype=torch.float32)
        seg = torch.empty((2, nseg, C), device=dev, dtype=torch.float32)
        nbytes = _cabi.lib().ge_bn_workspace_bytes(P, C)
        ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        dt = _dtype_code(xc)
        es = xc.element_size()
        call("ge_bn_sync_bwd_reduce", ptr(d), ptr(mask), ptr(xc), ptr(save[0]), ptr(save[1]), ptr(seg), ptr(dgb[0]), ptr(dgb[1]),
             ptr(ws), c_size_t(nbytes), dt, c_longlong(P), c_longlong(P_split), C, int(relu), stream(),
             work=(P * C * es * 2 + (P * C // 8 if relu else 0), 8 * P * C))
        if world > 1:
            dist.all_reduce(seg, op=dist.ReduceOp.AVG, group=group)
        dx = torch.empty_like(xc)
        dres = torch.empty_like(xc) if has_res else None
        call("ge_bn_sync_bwd_apply", ptr(d), ptr(mask), ptr(xc), ptr(g), ptr(save[0]), ptr(save[1]), ptr(seg), ptr(dx), ptr(dres),
             dt, c_longlong(P), c_longlong(P_split), C, int(relu), stream(),
             work=(P * C * es * (3 + (1 if has_res else 0)) + (P * C // 8 if relu else 0), 8 * P * C))
        return dx, dres, dgb[0].to(gdt), dgb[1].to(bdt), None, None, None, None, None, None, None, None, None


def _sync_bn_group(bn):
    """(process group, world size) of an nn.SyncBatchNorm in training mode, or None when it has to fall back."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return None
    group = bn.process_group if bn.process_group is not None else dist.group.WORLD
    if dist.get_backend(group) != "nccl":
        return None
    return group, dist.get_world_size(group)


def bn_act(x, bn, residual=None, relu=True):
    """relu(BatchNorm2d(x) + residual) with `bn` an nn.BatchNorm2d (parameters / buffers are read and,
    in training mode, updated exactly as the module would; under `domain_split` exactly as two calls on the two
    halves would).  Other norm types (e.g. SyncBatchNorm) fall back to the module itself."""
    split = _BN_SPLIT if 0 < _BN_SPLIT < x.shape[0] else 0
    if (type(bn) is torch.nn.SyncBatchNorm and bn.training and bn.affine and bn.track_running_stats
            and bn.momentum is not None and x.is_cuda and x.dim() == 4 and bn.num_features % 8 == 0
            and _cabi.lib().ge_bn_workspace_bytes(x.shape[0] * x.shape[2] * x.shape[3], bn.num_features) > 0):
        gw = _sync_bn_group(bn)
        if gw is not None:
            return _SyncBnAct.apply(x, residual, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked,
                                    float(bn.momentum), float(bn.eps), bool(relu), int(split), gw[0], gw[1])
    if type(bn) is not torch.nn.BatchNorm2d or not bn.affine or (bn.training and bn.momentum is None):
        if split and bn.training:
            y = torch.cat([bn(x[:split]), bn(x[split:])], dim=0)
        else:
            y = bn(x)
        if residual is not None:
            y = y + residual
        return torch.relu(y) if relu else y
    if bn.training or not bn.track_running_stats:
        rm = bn.running_mean if bn.track_running_stats else None
        rv = bn.running_var if bn.track_running_stats else None
        nbt = bn.num_batches_tracked if bn.track_running_stats else None
        return _BnAct.apply(x, residual, bn.weight, bn.bias, rm, rv, nbt, float(bn.momentum), float(bn.eps), bool(relu),
                            int(split))
    # inference: a per-channel affine map -- differentiable through ordinary autograd if anyone asks
    if torch.is_grad_enabled() and (x.requires_grad or bn.weight.requires_grad):
        y = torch.nn.functional.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.0, bn.eps)
        if residual is not None:
            y = y + residual
        return torch.relu(y) if relu else y
    xc = _nhwc_view(x)
    N, C, H, W = xc.shape
    rc = _nhwc_view(residual.to(xc.dtype)) if residual is not None else None
    out = torch.empty_like(xc)
    es = xc.element_size()
    call("ge_bn_fwd_eval", ptr(xc), ptr(rc), ptr(_f32c(bn.weight)), ptr(_f32c(bn.bias)), ptr(bn.running_mean),
         ptr(bn.running_var), c_float(bn.eps), ptr(out), _dtype_code(xc), c_longlong(N * H * W), C, int(relu), stream(),
         work=(N * H * W * C * es * (2 + (1 if rc is not None else 0)), 4 * N * H * W * C))
    return out


# ------------------------------------------------------------------------------------------ conv bias on NHWC maps
_CONST_CACHE = {}


def _const_vec(value, C, device):
    key = (float(value), int(C), str(device))
    t = _CONST_CACHE.get(key)
    if t is None:
        t = _CONST_CACHE[key] = torch.full((C,), float(value), device=device, dtype=torch.float32)
    return t


class _BiasAddNHWC(Function):
    """y += bias (per channel) in place on a fresh NHWC convolution output, and the bias gradient as a streaming channel
    sum.  ATen adds the bias of a channels_last convolution with a non-vectorised broadcast kernel (114 us at
    [256,256,28,28] bf16) and reduces its gradient at ~1 TB/s (108 us); the fused BatchNorm kernels do both at streaming
    speed: ge_bn_fwd_eval with (gamma, beta, mean, var, eps) = (1, bias, 0, 1, 0) is exactly x + bias, and
    ge_bn_sync_stats with a zero shift yields the per-channel sums."""

    @staticmethod
    def forward(ctx, y, bias):
        _need_cuda(y, bias)
        N, C, H, W = y.shape
        b = _f32c(bias)
        call("ge_bn_fwd_eval", ptr(y), None, ptr(_const_vec(1.0, C, y.device)), ptr(b), ptr(_const_vec(0.0, C, y.device)),
             ptr(_const_vec(1.0, C, y.device)), c_float(0.0), ptr(y), _dtype_code(y), c_longlong(N * H * W), C, 0, stream(),
             work=(2 * y.numel() * y.element_size(), y.numel()))
        ctx.mark_dirty(y)
        ctx.cfg = (bias.dtype, tuple(bias.shape))
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        bdt, bshape = ctx.cfg
        d = _nhwc_view(dy)
        N, C, H, W = d.shape
        P = N * H * W
        nbytes = _cabi.lib().ge_bn_workspace_bytes(P, C)
        ws = torch.empty(nbytes, device=d.device, dtype=torch.uint8)
        sums = torch.empty((1, 2, C), device=d.device, dtype=torch.float32)
        call("ge_bn_sync_stats", ptr(d), None, ptr(sums), ptr(ws), c_size_t(nbytes), _dtype_code(d), c_longlong(P),
             c_longlong(0), C, stream(), work=(d.numel() * d.element_size(), d.numel()))
        return dy, sums[0, 0].to(bdt).view(bshape)


def conv_bias(x, conv):
    """conv(x) for an nn.Conv2d with a bias on a channels_last CUDA map: library convolution without the bias, then the
    fused in-place bias add (and a streaming bias gradient).  Anything else: the module itself."""
    C = conv.out_channels
    ok = (x.is_cuda and x.dim() == 4 and conv.bias is not None and C % 8 == 0
          and _cabi.lib().ge_bn_workspace_bytes(x.shape[0] * 8 * 8, C) > 0)
    if not ok:
        return conv(x)
    y = torch.nn.functional.conv2d(x, conv.weight, None, conv.stride, conv.padding, conv.dilation, conv.groups)
    if y.dtype not in (torch.float32, torch.bfloat16) or not y.is_contiguous(memory_format=torch.channels_last):
        return y + conv.bias.view(1, -1, 1, 1).to(y.dtype)
    return _BiasAddNHWC.apply(y, conv.bias)


# ------------------------------------------------------------------------------------------ stem convolution
class _StemConv(Function):
    @staticmethod
    def forward(ctx, x, weight, out_dtype):
        _need_cuda(x, weight)
        xf = _f32c(x)
        wf = _f32c(weight)
        F_, _, H, W = xf.shape
        y = torch.empty((F_, 64, H // 2, W // 2), device=x.device, dtype=out_dtype, memory_format=torch.channels_last)
        dt = BF16 if out_dtype == torch.bfloat16 else F32
        call("ge_stem_conv_fwd", ptr(xf), ptr(wf), ptr(y), F_, H, W, dt, stream(),
             work=(4 * F_ * H * W + y.numel() * y.element_size(), 2 * 49 * y.numel()))
        ctx.save_for_backward(xf)
        ctx.cfg = (dt, weight.dtype, tuple(weight.shape))
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        (xf,) = ctx.saved_tensors
        dt, wdt, wshape = ctx.cfg
        F_, _, H, W = xf.shape
        d = _nhwc_view(dy.to(torch.bfloat16 if dt == BF16 else torch.float32))
        dw = torch.empty(wshape, device=xf.device, dtype=torch.float32)
        nbytes = _cabi.lib().ge_stem_conv_wgrad_workspace_bytes(F_, H, W)
        ws = torch.empty(nbytes, device=xf.device, dtype=torch.uint8)
        call("ge_stem_conv_wgrad", ptr(xf), ptr(d), ptr(dw), ptr(ws), c_size_t(nbytes), F_, H, W, dt, stream(),
             work=(4 * F_ * H * W + d.numel() * d.element_size(), 2 * 49 * d.numel()))
        return None, dw.to(wdt), None


USE_STEM_CONV = True


def stem_conv(x, conv):
    """ResNet.conv1 (Conv2d(1, 64, 7, 2, 3, bias=False), fpnseg.py:229) on the direct kernels where they apply (one input
    channel, data that needs no gradient, even height, width % 8 == 0); the module itself otherwise."""
    ok = (USE_STEM_CONV and x.is_cuda and x.dim() == 4 and x.shape[1] == 1 and not x.requires_grad and conv.bias is None
          and tuple(conv.weight.shape) == (64, 1, 7, 7) and conv.stride == (2, 2) and conv.padding == (3, 3)
          and conv.dilation == (1, 1) and conv.groups == 1
          and bool(_cabi.lib().ge_stem_conv_supported(int(x.shape[2]), int(x.shape[3]))))
    if not ok:
        return conv(x)
    if torch.is_autocast_enabled():
        adt = torch.get_autocast_dtype('cuda')
        if adt != torch.bfloat16:
            return conv(x)
        out_dtype = torch.bfloat16
    else:
        if x.dtype != torch.float32 or conv.weight.dtype != torch.float32:
            return conv(x)
        out_dtype = torch.float32
    return _StemConv.apply(x, conv.weight, out_dtype)


# ------------------------------------------------------------------------------------------ f3: 1x1 conv GEMM + BN statistics
def conv1x1_tc_supported(P, K, N):
    return bool(_cabi.lib().ge_conv1x1_tc_supported(c_longlong(int(P)), int(K), int(N)))


def conv1x1_gemm(x2d, w2d, shift=None, want_stats=False, P_split=0):
    """y [P,N] = x2d [P,K] @ w2d [N,K]^T on the tcgen05 kernel (bf16 in / out, fp32 accumulate).  With want_stats also
    returns the per-CTA partial rows [rows,2,N] of sum(y - shift), sum((y - shift)^2) and (rows, rows_segment0)."""
    _need_cuda(x2d, w2d)
    P, K = x2d.shape
    N = w2d.shape[0]
    y = torch.empty((P, N), device=x2d.device, dtype=torch.bfloat16)
    part, rows, rows0 = None, 0, 0
    if want_stats:
        r0 = c_int(0)
        rows = _cabi.lib().ge_conv1x1_tc_partial_rows(c_longlong(P), c_longlong(int(P_split)), K, N, ctypes_byref(r0))
        rows0 = r0.value
        part = torch.empty((rows, 2, N), device=x2d.device, dtype=torch.float32)
    call("ge_conv1x1_bn_stats", ptr(x2d), ptr(w2d), ptr(y), ptr(shift), ptr(part), c_longlong(P), c_longlong(int(P_split)),
         K, N, stream(), work=(2 * P * (K + N) + 2 * K * N, 2 * P * K * N))
    return (y, part, rows, rows0) if want_stats else y


class _Conv1x1BnAct(Function):
    """relu(BatchNorm2d(conv1x1(x)) + residual), training mode, bf16 NHWC: the tcgen05 GEMM writes y and the partial batch
    statistics of y from its fp32 accumulators; finalize + apply follow (no statistics pass over y).  Backward: fused
    BatchNorm backward, then dgrad / wgrad as two library GEMMs."""

    @staticmethod
    def forward(ctx, x, weight, residual, gamma, beta, running_mean, running_var, nbt, momentum, eps, relu, split):
        xc = _nhwc_view(x)
        Nb, K, H, W = xc.shape
        Cout = weight.shape[0]
        P = Nb * H * W
        P_split = split * H * W if 0 < split < Nb else 0
        nseg = 2 if P_split else 1
        x2 = xc.permute(0, 2, 3, 1).reshape(P, K)
        w2 = weight.reshape(Cout, K).to(torch.bfloat16).contiguous()
        y2, part, rows, rows0 = conv1x1_gemm(x2, w2, running_mean, True, P_split)
        y = y2.view(Nb, H, W, Cout).permute(0, 3, 1, 2)
        rc = _nhwc_view(residual.to(torch.bfloat16)) if residual is not None else None
        g, b = _f32c(gamma), _f32c(beta)
        dev = x.device
        out = torch.empty_like(y)
        save = torch.empty((2, nseg, Cout), device=dev, dtype=torch.float32)
        mask = torch.empty(_cabi.lib().ge_bn_relu_mask_bytes(P, Cout), device=dev, dtype=torch.uint8) if relu else None
        call("ge_bn_fwd_train_prestat", ptr(y), ptr(rc), ptr(g), ptr(b), ptr(running_mean), ptr(running_var), ptr(nbt),
             c_float(momentum), c_float(eps), ptr(out), ptr(save[0]), ptr(save[1]), ptr(mask), ptr(part), rows, rows0,
             BF16, c_longlong(P), c_longlong(P_split), Cout, int(relu), stream(),
             work=(P * Cout * 2 * (2 + (1 if rc is not None else 0)), 6 * P * Cout))
        ctx.save_for_backward(x2, w2, y, mask, g, save)
        ctx.cfg = (P, P_split, K, Cout, bool(relu), rc is not None, weight.shape, weight.dtype, gamma.dtype, beta.dtype, xc.shape)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        x2, w2, y, mask, g, save = ctx.saved_tensors
        P, P_split, K, Cout, relu, has_res, wshape, wdt, gdt, bdt, xshape = ctx.cfg
        d = _nhwc_view(dout.to(torch.bfloat16))
        dy = torch.empty_like(y)
        dres = torch.empty_like(y) if has_res else None
        dgb = torch.empty((2, Cout), device=d.device, dtype=torch.float32)
        nbytes = _cabi.lib().ge_bn_workspace_bytes(P, Cout)
        ws = torch.empty(nbytes, device=d.device, dtype=torch.uint8)
        call("ge_bn_bwd", ptr(d), ptr(mask), ptr(y), ptr(g), ptr(save[0]), ptr(save[1]), ptr(dy), ptr(dres), ptr(dgb[0]),
             ptr(dgb[1]), ptr(ws), c_size_t(nbytes), BF16, c_longlong(P), c_longlong(P_split), Cout, int(relu), stream(),
             work=(P * Cout * 2 * (3 + (1 if has_res else 0)) + (P * Cout // 8 if relu else 0), 16 * P * Cout))
        dy2 = dy.permute(0, 2, 3, 1).reshape(P, Cout)
        dx2 = dy2 @ w2                                            # dgrad  [P,Cout] x [Cout,K]
        dw = dy2.t() @ x2                                         # wgrad  [Cout,P] x [P,K]  (bf16 product)
        dw = (dw if wdt == torch.bfloat16 else dw.float()).reshape(wshape).to(wdt)
        Nb, _, H, W = xshape
        dx = dx2.view(Nb, H, W, K).permute(0, 3, 1, 2)
        return dx, dw, dres, dgb[0].to(gdt), dgb[1].to(bdt), None, None, None, None, None, None, None


def conv1x1_bn_act(x, conv, bn, residual=None, relu=True, drop_bias=False):
    """relu(bn(conv(x)) + residual) for a 1x1, stride-1 convolution followed by a train-mode nn.BatchNorm2d on a bf16
    NHWC map: the tcgen05 GEMM with the statistics epilogue where it applies and measures faster than cuDNN + the
    statistics kernel (K <= 512, >= 4096 pixels; scripts/conv1x1_tc_check.py), else conv + bn_act.  `drop_bias`: the
    caller accounts for the conv bias itself (a constant in front of train-mode BatchNorm cancels, vig._conv_bn_train)."""
    ok = (USE_CONV1X1_TC and x.is_cuda and x.dtype == torch.bfloat16 and bn.training and type(bn) is torch.nn.BatchNorm2d
          and bn.affine and bn.track_running_stats and bn.momentum is not None and (conv.bias is None or drop_bias)
          and conv.kernel_size == (1, 1) and conv.stride == (1, 1) and conv.groups == 1 and conv.padding == (0, 0)
          and x.dim() == 4 and conv.in_channels <= 512 and x.shape[0] * x.shape[2] * x.shape[3] >= 4096
          and torch.is_autocast_enabled()
          and conv1x1_tc_supported(x.shape[0] * x.shape[2] * x.shape[3], conv.in_channels, conv.out_channels)
          and bn.num_features % 8 == 0)
    if not ok:
        y = torch.nn.functional.conv2d(x, conv.weight, None if drop_bias else conv.bias, conv.stride, conv.padding,
                                       conv.dilation, conv.groups)
        return bn_act(y, bn, residual=residual, relu=relu)
    split = _BN_SPLIT if 0 < _BN_SPLIT < x.shape[0] else 0
    return _Conv1x1BnAct.apply(x, conv.weight, residual, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                               bn.num_batches_tracked, float(bn.momentum), float(bn.eps), bool(relu), int(split))


def conv_bn_act(x, conv, bn, residual=None, relu=True):
    """relu(bn(conv(x)) + residual) for a biased convolution in front of a train-mode nn.BatchNorm2d WITHOUT the bias
    passes: a per-channel constant in front of train-mode BatchNorm cancels in the output and has a zero gradient; it
    only shifts the batch mean, so the running mean gets its share added back (state_dict parity with the reference:
    VGG16's conv + BN + ReLU triples, fpnseg.py:18-166).  ATen would add the bias with a non-vectorised broadcast
    kernel and reduce its gradient at ~1 TB/s over the largest maps of the network.  Anything else (eval mode, other
    norm types, no running statistics): the plain composition."""
    drop = (conv.bias is not None and bn.training and type(bn) is torch.nn.BatchNorm2d and bn.affine
            and bn.track_running_stats and bn.momentum is not None and x.is_cuda)
    if not drop:
        return bn_act(conv(x), bn, residual=residual, relu=relu)
    y = conv1x1_bn_act(x, conv, bn, residual=residual, relu=relu, drop_bias=True)
    with torch.no_grad():
        # one running-mean update per BatchNorm segment (domain_split): 1 - (1-m)^nseg of the bias in total
        share = 1.0 - (1.0 - float(bn.momentum)) ** bn_segments(x.shape[0])
        bn.running_mean.add_(conv.bias.detach().to(bn.running_mean.dtype), alpha=share)
    return y


USE_CONV1X1_TC = True       # measured >= cuDNN + statistics kernel at every eligible shape (profiles/r2_conv1x1_tc.md)


class _SegTail(Function):
    @staticmethod
    def forward(ctx, s2, s3, s4, s5, W3, b3, scale):
        _need_cuda(s2, s3, s4, s5, W3, b3)
        dt = s2.dtype
        a, b, c, d = (_nhwc_view(t.to(dt)) for t in (s2, s3, s4, s5))
        N, C, h, w = a.shape
        nc = W3.shape[0]
        Wm = _f32c(W3).view(nc, C)
        bm = _f32c(b3)
        H, W = h * scale, w * scale
        q = torch.empty((N, h, w, nc), device=a.device, dtype=torch.float32)
        logits = torch.empty((N, nc, H, W), device=a.device, dtype=torch.float32)
        call("ge_seg_tail_fwd", ptr(a), ptr(b), ptr(c), ptr(d), ptr(Wm), ptr(bm), ptr(q), ptr(logits),
             _dtype_code(a), N, h, w, H, W, C, nc, stream(),
             work=(4 * N * h * w * C * a.element_size() + 4 * N * nc * H * W, 2 * N * h * w * C * (nc + 2)))
        ctx.save_for_backward(a, b, c, d, Wm)
        ctx.cfg = (N, C, h, w, H, W, nc, W3.shape, W3.dtype, b3.dtype)
        return logits

    @staticmethod
    @once_differentiable
    def backward(ctx, dlogits):
        a, b, c, d, Wm = ctx.saved_tensors
        N, C, h, w, H, W, nc, wshape, wdt, bdt = ctx.cfg
        g = _f32c(dlogits)
        dq = torch.empty((N, h, w, nc), device=g.device, dtype=torch.float32)
        ds = torch.empty_like(a)
        dW3 = torch.zeros((nc, C), device=g.device, dtype=torch.float32)
        db3 = torch.zeros(nc, device=g.device, dtype=torch.float32)
        call("ge_seg_tail_bwd", ptr(g), ptr(a), ptr(b), ptr(c), ptr(d), ptr(Wm), ptr(dq), ptr(ds), ptr(dW3),
             ptr(db3), _dtype_code(a), N, h, w, H, W, C, nc, stream(),
             work=(5 * N * h * w * C * a.element_size() + 4 * N * nc * H * W, 4 * N * h * w * C * (nc + 1)))
        return ds, ds, ds, ds, dW3.view(wshape).to(wdt), db3.to(bdt), None


def seg_tail(s2, s3, s4, s5, W3, b3, scale=4):
    """_upsample(conv3(s2+s3+s4+s5), 4h, 4w) -> fp32 NCHW logits (fpnseg.py:444)."""
    return _SegTail.apply(s2, s3, s4, s5, W3, b3, int(scale))


# ------------------------------------------------------------------------------------------ f4: losses, score-map boxes
class _SegLoss(Function):
    @staticmethod
    def forward(ctx, logits, target, smooth):
        _need_cuda(logits, target)
        x, t = _f32c(logits), _f32c(target)
        if x.shape != t.shape or x.dim() != 4:
            raise _cabi.GraphEchoNativeError(f"seg_loss: logits {tuple(x.shape)} and masks {tuple(t.shape)} must be equal [F,nc,H,W]")
        Fr, nc, H, W = x.shape
        nbytes = _cabi.lib().ge_seg_loss_workspace_bytes(Fr, nc, H * W)
        if nbytes == 0:
            raise _cabi.GraphEchoNativeError(f"seg_loss supports 1..8 classes, got {nc}")
        ws = torch.empty(nbytes, device=x.device, dtype=torch.uint8)
        numden = torch.empty((Fr, nc, 2), device=x.device, dtype=torch.float32)
        loss = torch.empty(3, device=x.device, dtype=torch.float32)
        call("ge_seg_loss_fwd", ptr(x), ptr(t), ptr(numden), ptr(loss), ptr(ws), c_size_t(nbytes), Fr, nc, H * W,
             c_float(smooth), stream(), work=(8 * x.numel(), 30 * x.numel()))
        ctx.save_for_backward(x, t, numden)
        return loss[0]

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, t, numden = ctx.saved_tensors
        Fr, nc, H, W = x.shape
        gout = _f32c(g).reshape(1)
        dx = torch.empty_like(x)
        call("ge_seg_loss_bwd", ptr(x), ptr(t), ptr(numden), ptr(gout), ptr(dx), Fr, nc, H * W, stream(),
             work=(12 * x.numel(), 40 * x.numel()))
        return dx, None, None


def seg_loss(logits, masks, smooth=1.0):
    """DiceLoss()(logits, masks) + BCEWithLogitsLoss()(logits, masks) (utils/losses.py:64-95 + nn.BCEWithLogitsLoss;
    train_cardiac_uda.py:228): one fused pass forward, one backward.  logits, masks [F,nc,H,W], nc <= 8."""
    return _SegLoss.apply(logits, masks, float(smooth))


class LogitMap:
    """A score map given by its logits: `where(sigmoid(logits) > 0.5, 1, 0)` (train_cardiac_uda.py:235) WITHOUT being
    materialised; the graph module only ever takes its bounding boxes (mask_boxes, mode 1)."""

    def __init__(self, logits):
        self.logits = logits

    def materialise(self):
        return torch.where(torch.sigmoid(self.logits) > 0.5, 1, 0)


@torch.no_grad()
def mask_boxes(maps):
    """[B,K,H,W] masks / score maps (or a LogitMap) -> fp32 [B,K,4] boxes (xmin,ymin,xmax,ymax) of the non-zero pixels,
    (0,0,W,H) for an empty plane (graph_matching.py:702-746)."""
    mode = 0
    if isinstance(maps, LogitMap):
        maps, mode = maps.logits, 1
    _need_cuda(maps)
    if maps.dtype == torch.int64:
        m, code = maps.contiguous(), 2
    else:
        m, code = _f32c(maps), F32
    B, K, H, W = m.shape
    boxes = torch.empty((B, K, 4), device=m.device, dtype=torch.float32)
    call("ge_mask_boxes", ptr(m), ptr(boxes), code, B * K, H, W, mode, stream(),
         work=(m.numel() * m.element_size(), 4 * m.numel()))
    return boxes


@torch.no_grad()
def sampler_labels(boxes, level_hw, strides, size_ranges):
    """Labels of every location of every pyramid level + per-level (positive, negative) counts in one launch
    (graph_matching.py:609-635, 874-959).  boxes [B,K,4] fp32; level_hw [(h,w)...]; strides, size_ranges per level.
    Returns ([labels_l int64 [B*h_l*w_l]], counts int32 [L,2] on the device)."""
    import ctypes
    _need_cuda(boxes)
    bx = _f32c(boxes)
    B, K = bx.shape[0], bx.shape[1]
    L = len(level_hw)
    sizes = [int(h) * int(w) for h, w in level_hw]
    labels = torch.empty(B * sum(sizes), device=bx.device, dtype=torch.int64)
    counts = torch.zeros((L, 2), device=bx.device, dtype=torch.int32)
    IA, FA = ctypes.c_int * L, ctypes.c_float * L
    call("ge_sampler_labels", ptr(bx), ptr(labels), ptr(counts), IA(*[int(h) for h, _ in level_hw]), IA(*[int(w) for _, w in level_hw]),
         IA(*[int(s_) for s_ in list(strides)[:L]]), FA(*[float(a) for a, _ in size_ranges]), FA(*[float(b) for _, b in size_ranges]),
         L, B, K, stream(), work=(16 * B * K + 8 * labels.numel(), 20 * K * labels.numel()))
    out, off = [], 0
    for n_ in sizes:
        out.append(labels[off * B:(off + n_) * B])
        off += n_
    return out, counts


def sampler_gather_plan(counts, num_nodes_per_class=100, bg_ratio=8, sample_bg_nodes=True):
    """Host half of the picks (graph_matching.py:984-1003) for one domain from its per-level (n_pos, n_neg) counts:
    per level (n_neg_all, step, n_pos_pick, n_neg_pick, neg_all, out_pos, out_neg) and the node count."""
    rows = []
    for n_pos_all, n_neg_all in counts:
        n_pos_all, n_neg_all = int(n_pos_all), int(n_neg_all)
        step = max(n_pos_all // num_nodes_per_class, 1)
        n_pos_pick = -(-n_pos_all // step) if step > 1 else n_pos_all
        if not sample_bg_nodes:
            neg_all, n_neg_pick = 0, 0
        elif n_pos_all > n_neg_all:
            neg_all, n_neg_pick = 1, n_neg_all
        else:
            neg_all, n_neg_pick = 0, n_pos_pick // bg_ratio
        rows.append([n_neg_all, step, n_pos_pick, n_neg_pick, neg_all, 0, 0])
    total_neg = sum(r[3] for r in rows)
    on, op = 0, total_neg
    for r in rows:
        r[5], r[6] = op, on
        op += r[2]
        on += r[3]
    return rows, op


class _SamplerGather(Function):
    """Nodes of every domain in one launch: forward(plan, *feats) -> (nodes_0, labels_0, nodes_1, labels_1, ...).
    plan = [(per-level label views, per-level feat index into feats, batch_offset, rows, n_nodes), ...] per domain."""

    @staticmethod
    def forward(ctx, plan, *feats):
        import ctypes
        _need_cuda(*feats)
        fv = [_nhwc_view(f) for f in feats]
        dt = _dtype_code(fv[0])
        C = fv[0].shape[1]
        dev = fv[0].device
        outs, ent = [], []
        for labels, fidx, boff, rows, n_nodes in plan:
            nodes = torch.empty((n_nodes, C), device=dev, dtype=torch.float32)
            nlab = torch.empty(n_nodes, device=dev, dtype=torch.int64)
            src = torch.empty(n_nodes, device=dev, dtype=torch.int64)
            outs += [nodes, nlab]
            for lab, fi, r in zip(labels, fidx, rows):
                f = fv[fi]
                if f.shape[1] != C or _dtype_code(f) != dt:
                    raise _cabi.GraphEchoNativeError("sampler_gather: all feature levels must share C and dtype")
                hw = f.shape[2] * f.shape[3]
                ent.append((f, lab, nodes, nlab, src, lab.numel(), boff * hw, r, fi))
        n = len(ent)
        PA, LA, IA = ctypes.c_void_p * n, ctypes.c_longlong * n, ctypes.c_int * n
        col = lambda k: IA(*[int(e[7][k]) for e in ent])
        call("ge_sampler_gather", PA(*[e[0].data_ptr() for e in ent]), PA(*[e[1].data_ptr() for e in ent]),
             PA(*[e[2].data_ptr() for e in ent]), PA(*[e[3].data_ptr() for e in ent]), PA(*[e[4].data_ptr() for e in ent]),
             LA(*[int(e[5]) for e in ent]), LA(*[int(e[6]) for e in ent]), col(0), col(1), col(2), col(3), col(4), col(5), col(6),
             n, C, dt, stream(),
             work=(8 * sum(int(e[5]) for e in ent) + sum(o.numel() * 4 for o in outs[::2]) * 2, 0))
        ctx.ent = [(e[4], e[7], e[8]) for e in ent]
        ctx.meta = (C, dt, [tuple(f.shape) for f in fv], [f.dtype for f in fv], dev,
                    [len(p[0]) for p in plan])
        ctx.mark_non_differentiable(*outs[1::2])
        return tuple(outs)

    @staticmethod
    @once_differentiable
    def backward(ctx, *grads):
        import ctypes
        C, dt, shapes, dtypes, dev, per_domain = ctx.meta
        dn = [_f32c(g) if g is not None else None for g in grads[0::2]]
        need = ctx.needs_input_grad[1:]
        # allocated channels_last from the start (zeros(...).contiguous(channels_last) is a 100 MB strided copy per level)
        dfeat = [torch.empty(sh, device=dev, dtype=d_, memory_format=torch.channels_last).zero_() if nd else None
                 for sh, d_, nd in zip(shapes, dtypes, need)]
        ent, k = [], 0
        for d, nlev in enumerate(per_domain):
            for _ in range(nlev):
                src, r, fi = ctx.ent[k]
                k += 1
                if dfeat[fi] is not None and dn[d] is not None and (r[2] + r[3]) > 0:
                    ent.append((dfeat[fi], dn[d], src, r))
        if ent:
            n = len(ent)
            PA, IA = ctypes.c_void_p * n, ctypes.c_int * n
            col = lambda k_: IA(*[int(e[3][k_]) for e in ent])
            call("ge_sampler_scatter", PA(*[e[0].data_ptr() for e in ent]), PA(*[e[1].data_ptr() for e in ent]),
                 PA(*[e[2].data_ptr() for e in ent]), col(2), col(3), col(5), col(6), n, C, dt, stream(),
                 work=(sum((e[3][2] + e[3][3]) * C * 6 for e in ent), 0))
        return (None, *dfeat)


def sampler_gather(plan, feats):
    """See _SamplerGather.  Returns [(nodes fp32 [n,C], labels int64 [n]) per domain]."""
    out = _SamplerGather.apply(plan, *feats)
    return [(out[2 * i], out[2 * i + 1]) for i in range(len(plan))]


class _GradReverse(Function):
    """Gradient reversal (gradient_reversal.py:6-24) without the reference's full clone()."""

    @staticmethod
    def forward(ctx, x, lambda_):
        ctx.lambda_ = float(lambda_)
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return -ctx.lambda_ * g, None


def grad_reverse(x, lambda_):
    return _GradReverse.apply(x, lambda_)

"""CPU: host-side logic of the product that needs no GPU — the dense route of the spectral
bipartition against sklearn (what the reference calls, graph_matching.py:539-543), the synthetic
generators, and the TGCN per-timestep BatchNorm bookkeeping."""
import warnings

import pytest
import torch

from graphecho_b200 import synth
from graphecho_b200.spectral import spectral_bipartition


def test_spectral_bipartition_matches_sklearn_on_clustered_points():
    cluster = pytest.importorskip("sklearn.cluster")
    warnings.filterwarnings("ignore")
    for seed in range(6):
        g = torch.Generator().manual_seed(seed)
        n = int(torch.randint(25, 110, (1,), generator=g))
        a = torch.randn(n, 256, generator=g)
        a[: n // 3] += 1.5 * torch.randn(1, 256, generator=g)
        pts = torch.cat([torch.randn(1, 256, generator=g), a])
        sp = cluster.SpectralClustering(2, affinity="nearest_neighbors", n_jobs=-1, assign_labels="kmeans",
                                        random_state=1234, n_neighbors=n // 2)
        idx = sp.fit_predict(pts.numpy())
        ref = torch.as_tensor((idx == idx[0])[1:])
        assert (spectral_bipartition(pts, n // 2) == ref).float().mean() > 0.97


def test_synthetic_generators():
    m = synth.disc_masks(3, 4, 64)
    assert m.shape == (3, 4, 64, 64) and torch.allclose(m.sum(1), torch.ones(3, 64, 64))
    assert all(m[:, c].sum() > 0 for c in range(4))
    x = synth.clips(2, 16, 5, seed=3)
    f = synth.flatten_clips(x)
    assert f.shape == (10, 1, 16, 16) and torch.equal(f[7, 0], x[1, 0, :, :, 2])
    p = synth.pyramid(2, 112)
    assert [t.shape[-1] for t in p] == [28, 14, 7, 4]
    p = synth.pyramid(2, 256)
    assert [t.shape[-1] for t in p] == [64, 32, 16, 8]


def test_fp16_hi_lo_split_inner_product_is_fp32_accurate():
    """Arithmetic model of the tcgen05 k-NN kernel (csrc/knn_tc.cu), emulated on the CPU: every normalised value
    is split as v = hi + lo * 2^-11 with hi = fp16(v), lo = fp16((v - hi) * 2^11); products of fp16 values are exact
    in fp32, the tensor core accumulates D1 = sum hi.hi' and D2 = sum (lo.hi' + hi.lo') in fp32, and
    x.y = D1 + 2^-11 * D2.  The error against the exact inner product must be at the level of a plain fp32 dot
    product (that is why k-NN indices match the fp32 reference except between candidates < 2e-6 apart)."""
    torch.manual_seed(0)
    C, n = 256, 2000
    x = torch.nn.functional.normalize(torch.randn(n, C, dtype=torch.float64), dim=1).float()
    y = torch.nn.functional.normalize(torch.randn(n, C, dtype=torch.float64), dim=1).float()

    def split(v):
        hi = v.half()
        lo = ((v - hi.float()) * 2048.0).half()
        return hi.float(), lo.float()

    xh, xl = split(x)
    yh, yl = split(y)
    # the split reconstructs v to ~2^-22 relative
    assert ((xh + xl / 2048.0) - x).abs().max() <= 2.0 ** -21 * x.abs().max()
    d1 = (xh * yh).sum(1)
    d2 = (xl * yh).sum(1) + (xh * yl).sum(1)
    ip = d1 + d2 / 2048.0
    truth = (x.double() * y.double()).sum(1)
    err_split = (ip.double() - truth).abs()
    err_fp32 = ((x * y).sum(1).double() - truth).abs()
    assert err_split.max() < 1e-7                       # distances are 2 - 2*ip: well inside the 2e-6 near-tie window
    assert err_split.mean() < 3 * err_fp32.mean() + 1e-9


def test_attention_forward_pair_equals_two_calls():
    """MultiHeadAttention.forward_pair (projections / output block shared over two node sets, attention products apart)
    reproduces two separate version-'v2' calls (transformer.py:45-75) for the self-, cross- and general patterns."""
    import torch
    from graphecho_b200.models.transformer import MultiHeadAttention
    torch.manual_seed(0)
    m = MultiHeadAttention(256, 1, dropout=0.1, version="v2").eval()
    a, b, c, d = torch.randn(37, 256), torch.randn(45, 256), torch.randn(11, 256), torch.randn(13, 256)
    for kv_a, q_a, kv_b, q_b in ((a, a, b, b), (a, b, b, a), (a, c, b, d)):
        (o1, e1), (o2, e2) = m.forward_pair(kv_a, q_a, kv_b, q_b)
        r1, f1 = m(kv_a, kv_a, q_a)
        r2, f2 = m(kv_b, kv_b, q_b)
        for x, y in ((o1, r1), (e1, f1), (o2, r2), (e2, f2)):
            assert x.shape == y.shape and torch.allclose(x, y, rtol=1e-5, atol=1e-6)


def _linspace_pick_index(r, m, n_neg_all):
    """Python port of csrc/sampler.cu::linspace_pick_index (same fp64 operations)."""
    import math
    if m <= 0:
        return -1
    if m == 1:
        return 0 if r == 0 else -1
    stop = float(n_neg_all - 2)
    step = stop / float(m - 1)
    i = int(math.floor(float(r) / step)) if step > 0.0 else 0
    for c in (i - 1, i, i + 1):
        if c < 0 or c >= m:
            continue
        y = stop if c == m - 1 else float(c) * step
        if int(math.floor(y)) == r:
            return c
    return -1


def test_sampler_negative_pick_membership_inverts_numpy_linspace():
    """The sampler kernel decides, per negative location of rank r, whether r is one of
    floor(linspace(0, n_neg - 2, m)) and which one (graph_matching.py:1001) in closed form.  Exhaustive check of that
    inversion against numpy over the (n_neg, m) range the sampler can produce (m = num_pos // 8 <= n_neg // 8)."""
    import numpy as np
    rng = np.random.default_rng(0)
    cases = [(n, m) for n in (2, 3, 9, 10, 17, 64, 100, 101, 777, 1000, 4097) for m in (0, 1, 2, 3, 7, 12, 100, 125, 512)
             if m <= max(n // 8, 1)]
    cases += [(int(n), int(rng.integers(1, max(n // 8, 1) + 1))) for n in rng.integers(16, 120000, size=40)]
    for n_neg, m in cases:
        picks = np.floor(np.linspace(0, n_neg - 2, m)).astype(np.int64) if m > 0 else np.zeros(0, dtype=np.int64)
        want = {int(p): i for i, p in enumerate(picks)}
        assert len(want) == len(picks), (n_neg, m)                  # strictly increasing: no rank is picked twice
        ranks = range(n_neg) if n_neg <= 5000 else list(picks) + [int(p) + 1 for p in picks] + list(rng.integers(0, n_neg, 200))
        for r in ranks:
            assert _linspace_pick_index(int(r), m, n_neg) == want.get(int(r), -1), (n_neg, m, r)


def test_sampler_gather_plan_matches_the_reference_counts():
    """functional.sampler_gather_plan (host half of graph_matching.py:984-1003): strides, pick counts and slot bases."""
    from graphecho_b200 import functional as GF
    counts = [(12345, 188000), (150, 40), (99, 5000), (0, 64), (450, 450)]
    rows, n = GF.sampler_gather_plan(counts)
    exp = []
    for n_pos, n_neg in counts:
        step = n_pos // 100
        num_pos = len(range(0, n_pos, step)) if step > 1 else n_pos
        num_neg = n_neg if n_pos > n_neg else num_pos // 8
        exp.append((num_pos, num_neg))
    assert [(r[2], r[3]) for r in rows] == exp
    total_neg = sum(e[1] for e in exp)
    assert n == total_neg + sum(e[0] for e in exp)
    assert [r[6] for r in rows] == [sum(e[1] for e in exp[:i]) for i in range(len(exp))]
    assert [r[5] for r in rows] == [total_neg + sum(e[0] for e in exp[:i]) for i in range(len(exp))]
    assert [r[4] for r in rows] == [0, 1, 0, 0, 0]


def test_sinkhorn_exponent_domain_restatement_equals_the_log_domain_reference():
    """The algebra csrc/sinkhorn_rpm_reg.cu relies on, in fp64 on the CPU: with K = exp(z), u = exp(-r), v = exp(-c) the
    slack-padded log-domain iteration (graph_matching.py:637-676) is u <- 1/(1 + K v), v <- 1/(1 + K^T u), P = K u v^T;
    and the exact adjoint of the unrolled loop is dz = K o F with F accumulated as rank-1 updates
    (x = gc v_t, gr -= u_t (K x), y = gr u_t, F += u_t x^T + y v_{t-1}^T, gc = -v_{t-1} (K^T y)).  Checked against
    autograd through the oracle's literal log-domain loop."""
    import torch
    from oracle import graph_ops as G
    torch.manual_seed(0)
    for n1, n2, T in ((7, 5, 20), (12, 17, 6), (3, 3, 1)):
        z = (torch.randn(n1, n2, dtype=torch.float64) * 1.3).requires_grad_()
        W = torch.randn(n1, n2, dtype=torch.float64)
        P_ref = G.sinkhorn_rpm(z[None], n_iters=T, slack=True)[0].exp()
        (P_ref * W).sum().backward()
        K = z.detach().exp()
        u, v = torch.ones(n1, dtype=torch.float64), torch.ones(n2, dtype=torch.float64)
        hu, hv = [], []
        for _ in range(T):
            u = 1.0 / (1.0 + K @ v)
            v = 1.0 / (1.0 + K.t() @ u)
            hu.append(u)
            hv.append(v)
        P = K * u[:, None] * v[None, :]
        assert torch.allclose(P, P_ref.detach(), rtol=1e-12, atol=1e-14)
        F = W * hu[-1][:, None] * hv[-1][None, :]
        E = K * F
        gr, gc = -E.sum(1), -E.sum(0)
        for t in range(T - 1, -1, -1):
            vprev = hv[t - 1] if t > 0 else torch.ones(n2, dtype=torch.float64)
            x = gc * hv[t]
            gr = gr - hu[t] * (K @ x)
            y = gr * hu[t]
            F = F + hu[t][:, None] * x[None, :] + y[:, None] * vprev[None, :]
            gc = -vprev * (K.t() @ y)
            gr = torch.zeros_like(gr)
        assert torch.allclose(K * F, z.grad, rtol=1e-10, atol=1e-13)


def test_affinity_identities_used_by_the_kernels():
    """csrc/affinity.cu: relu(a + b) == a + max(b, -a) bit for bit in fp32 (forward at 2 issue slots per term), the ReLU
    gate a + b > 0 <=> b > -a, and dw_k = sum_i a_ik s_ik + sum_j b_jk t_jk with the UNWEIGHTED gated sums s, t."""
    import torch
    torch.manual_seed(1)
    a = torch.cat([torch.randn(4000) * 3, torch.tensor([0.0, -0.0, 1e-30, -1e-30, 3.0, -3.0, 1e20, -1e20])])
    b = torch.cat([torch.randn(4000) * 3, torch.tensor([0.0, 0.0, -1e-30, 1e-30, -3.0, 3.0, -1e20, 1e20])])
    lhs, rhs = torch.relu(a + b), a + torch.maximum(b, -a)
    assert torch.equal(lhs, rhs)
    assert torch.equal((a + b) > 0, b > -a)
    A, B = torch.randn(9, 16, dtype=torch.float64), torch.randn(11, 16, dtype=torch.float64)
    w = torch.randn(16, dtype=torch.float64, requires_grad=True)
    g = torch.randn(9, 11, dtype=torch.float64)
    M = (torch.relu(A[:, None, :] + B[None, :, :]) * w).sum(-1)
    (dw,) = torch.autograd.grad(M, w, g)
    gate = ((A[:, None, :] + B[None, :, :]) > 0).to(torch.float64)
    s = (g[:, :, None] * gate).sum(1)            # [N1, H]
    t = (g[:, :, None] * gate).sum(0)            # [N2, H]
    assert torch.allclose(dw, (A * s).sum(0) + (B * t).sum(0), rtol=1e-12, atol=1e-12)

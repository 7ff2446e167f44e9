"""GPU: every C-ABI kernel against the CPU oracle (oracle/*.py) on seeded inputs, and against
the golden fixtures generated from the reference.  Integer outputs are bit-exact; fp32 outputs
use the tolerances written at each assert (fp32 re-association round-off)."""
import pytest
import torch
import torch.nn.functional as F

from graphecho_b200 import functional as GF
from oracle import graph_ops as G, vig_ops as V
from oracle.params import make_params

pytestmark = pytest.mark.gpu


def close(a, b, rtol=1e-4, atol=1e-5):
    torch.testing.assert_close(a.detach().cpu().float(), b.detach().cpu().float(), rtol=rtol, atol=atol)


def _affinity_AB(X, Y, p, pre=""):
    """Separable operands of the affinity MLP (the caller-side dense projections)."""
    W1 = p[pre + "fc_M.0.weight"]
    A = (X @ p[pre + "project_sr.weight"].t()) @ W1[:, :256].t()
    B = (Y @ p[pre + "project_tg.weight"].t()) @ W1[:, 256:].t() + p[pre + "fc_M.0.bias"]
    return A, B


# ---------------------------------------------------------------------------------------- K3
@pytest.mark.parametrize("n1,n2", [(37, 45), (1, 7), (250, 251), (320, 204), (33, 64)])
def test_affinity_pairwise_vs_oracle(dev, n1, n2):
    torch.manual_seed(n1 * 1000 + n2)
    p = make_params("affinity", fill_prefix="node_affinity.")
    X, Y = torch.randn(n1, 256), torch.randn(n2, 256)
    ref = G.affinity(X, Y, p).reshape(n1, n2)
    pd = {k: v.to(dev) for k, v in p.items()}
    A, B = _affinity_AB(X.to(dev), Y.to(dev), pd)
    M = GF.affinity_pairwise(A, B, pd["fc_M.2.weight"].view(-1), pd["fc_M.2.bias"])
    close(M, ref, rtol=1e-4, atol=2e-5)


def test_affinity_pairwise_golden_and_grads(dev, golden):
    g = golden("affinity_sinkhorn")
    p = {k: v.to(dev).requires_grad_() for k, v in make_params("affinity", fill_prefix="node_affinity.").items()}
    X, Y = g["X"].to(dev).requires_grad_(), g["Y"].to(dev).requires_grad_()
    A, B = _affinity_AB(X, Y, p)
    M = GF.affinity_pairwise(A, B, p["fc_M.2.weight"].view(-1), p["fc_M.2.bias"])
    close(M, g["M"], rtol=1e-4, atol=2e-5)
    P = GF.sinkhorn_rpm_exp(M, 20, True)
    close(P, g["P"], rtol=5e-4, atol=1e-6)
    (P * g["W"].to(dev)).sum().backward()
    close(X.grad, g["dX"], rtol=2e-3, atol=2e-6)
    close(Y.grad, g["dY"], rtol=2e-3, atol=2e-6)
    close(p["fc_M.2.weight"].grad, g["dw2"], rtol=2e-3, atol=2e-6)
    close(p["fc_M.2.bias"].grad, g["db2"], rtol=2e-3, atol=2e-6)
    close(p["fc_M.0.weight"].grad[:8], g["dfc0"], rtol=2e-3, atol=2e-6)
    close(p["project_sr.weight"].grad[:8], g["dPs"], rtol=2e-3, atol=2e-6)


def test_affinity_pairwise_batched(dev):
    torch.manual_seed(3)
    A, B = torch.randn(5, 70, 512, device=dev), torch.randn(5, 90, 512, device=dev)
    w2, b2 = torch.randn(512, device=dev) * 0.05, torch.randn(1, device=dev)
    M = GF.affinity_pairwise(A, B, w2, b2)
    ref = (torch.relu(A[:, :, None, :] + B[:, None, :, :]) * w2).sum(-1) + b2
    close(M, ref, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("batch,n1,n2,H", [(128, 70, 90, 64), (3, 70, 90, 96), (1, 252, 250, 512), (40, 33, 64, 288),
                                          (64, 130, 150, 512)])
def test_affinity_pairwise_backward_both_modes(dev, batch, n1, n2, H):
    """dA, dB, dw2, db2 against autograd through the literal relu-coupled sum: the split-Q partial mode (small
    batches: one fused sweep for both sides) and the direct mode (one sweep per side once the row tiles fill the
    machine), incl. a hidden width above one 256-channel pass of the forward."""
    torch.manual_seed(batch + n1)
    A = torch.randn(batch, n1, H, device=dev).requires_grad_()
    B = torch.randn(batch, n2, H, device=dev).requires_grad_()
    w2 = (torch.randn(H, device=dev) * 0.05).requires_grad_()
    b2 = torch.randn(1, device=dev).requires_grad_()
    Gm = torch.randn(batch, n1, n2, device=dev)
    M = GF.affinity_pairwise(A, B, w2, b2)
    got = torch.autograd.grad(M, (A, B, w2, b2), Gm)
    ref = torch.zeros(batch, n1, n2, device=dev)
    for k0 in range(0, H, 32):      # chunked over the hidden axis to bound the [b, n1, n2, 32] intermediate
        ref = ref + (torch.relu(A[:, :, None, k0:k0 + 32] + B[:, None, :, k0:k0 + 32]) * w2[k0:k0 + 32]).sum(-1)
    ref = ref + b2
    close(M, ref, rtol=1e-4, atol=1e-4)
    want = torch.autograd.grad(ref, (A, B, w2, b2), Gm)
    for g_, w_, name in zip(got, want, ("dA", "dB", "dw2", "db2")):
        scale = float(w_.abs().max())
        assert torch.allclose(g_, w_, rtol=2e-3, atol=2e-5 * max(scale, 1.0)), (name, float((g_ - w_).abs().max()), scale)


# ---------------------------------------------------------------------------------------- K4
@pytest.mark.parametrize("n1,n2,cs", [(37, 45, 0), (6, 6, 0), (1, 9, 0), (250, 251, 0), (320, 204, 0),
                                      (100, 100, 1), (100, 100, 2), (100, 100, 4), (100, 101, 8), (130, 97, 16),
                                      (318, 318, 0)])
def test_sinkhorn_rpm_fwd_bwd_vs_oracle(dev, n1, n2, cs):
    torch.manual_seed(n1 + 7 * n2 + cs)
    M = torch.randn(n1, n2) * 1.7 + 0.3
    W = torch.randn(n1, n2)
    Mo = M.clone().requires_grad_()
    Po = G.sinkhorn_rpm_exp(Mo, 20, True)
    (Po * W).sum().backward()
    Md = M.to(dev).requires_grad_()
    Pd = GF.sinkhorn_rpm_exp(Md, 20, True, cs)
    close(Pd, Po, rtol=5e-4, atol=1e-6)
    (Pd * W.to(dev)).sum().backward()
    close(Md.grad, Mo.grad, rtol=5e-3, atol=2e-6)


@pytest.mark.parametrize("path", [1, 2, 3])
@pytest.mark.parametrize("n1,n2,iters,instnorm", [(252, 252, 20, True), (250, 251, 20, True), (37, 45, 20, True),
                                                  (129, 128, 7, True), (511, 130, 20, True), (64, 256, 3, False),
                                                  (5, 3, 20, True), (300, 77, 0, True)])
def test_sinkhorn_rpm_register_path_vs_oracle(dev, path, n1, n2, iters, instnorm):
    """The register-resident exponent-domain kernels (4 and 8 rows per thread) and the log-domain kernel against the
    oracle's literal log-domain loop (graph_matching.py:637-676), forward and gradient, batched (2 problems)."""
    from graphecho_b200 import _cabi
    torch.manual_seed(n1 * 3 + n2 + iters)
    M = torch.randn(2, n1, n2) * 1.3 + 0.2
    W = torch.randn(2, n1, n2)
    Mo = M.clone().requires_grad_()
    if instnorm:
        Po = torch.stack([G.sinkhorn_rpm_exp(Mo[b], iters, True) for b in range(2)])
    else:
        Po = G.sinkhorn_rpm(Mo, iters, True).exp()
    (Po * W).sum().backward()
    _cabi.lib().ge_sinkhorn_rpm_set_path(path)
    try:
        Md = M.to(dev).requires_grad_()
        Pd = GF.sinkhorn_rpm_exp(Md, iters, instnorm)
        (Pd * W.to(dev)).sum().backward()
    finally:
        _cabi.lib().ge_sinkhorn_rpm_set_path(0)
    close(Pd, Po, rtol=5e-4, atol=1e-6)
    close(Md.grad, Mo.grad, rtol=5e-3, atol=2e-6)


def test_sinkhorn_rpm_mixed_batch_defers_only_the_outlier_problem(dev):
    """Batch of 3 where problem 1 has an instance-normed entry ~ 100: that problem takes the log-domain kernel
    (stats[2] == 1), the others stay on the register path; all three match the oracle, forward and gradient."""
    torch.manual_seed(5)
    M = torch.randn(3, 100, 120) * 0.5
    M[1] *= 0.02
    M[1, 3, 5] = 500.0
    W = torch.randn(3, 100, 120)
    Mo = M.clone().requires_grad_()
    Po = torch.stack([G.sinkhorn_rpm_exp(Mo[b], 20, True) for b in range(3)])
    (Po * W).sum().backward()
    Md = M.to(dev).requires_grad_()
    Pd = GF.sinkhorn_rpm_exp(Md, 20, True)
    (Pd * W.to(dev)).sum().backward()
    assert torch.isfinite(Pd).all() and torch.isfinite(Md.grad).all()
    close(Pd, Po, rtol=1e-3, atol=1e-6)
    close(Md.grad, Mo.grad, rtol=5e-3, atol=2e-6)


@pytest.mark.parametrize("iters", [0, 1, 5])
def test_sinkhorn_rpm_plain_and_batched(dev, iters, golden):
    torch.manual_seed(iters)
    M = torch.randn(4, 20, 31) * 2
    ref = G.sinkhorn_rpm(M, iters, True).exp()
    out = GF.sinkhorn_rpm_exp(M.to(dev), iters, False)
    close(out, ref, rtol=5e-4, atol=1e-6)
    if iters == 5:
        g = golden("affinity_sinkhorn")
        close(GF.sinkhorn_rpm_exp(g["Mraw"].to(dev), 5, False), g["rpm5"].exp(), rtol=5e-4, atol=1e-6)


def test_sinkhorn_rpm_properties_at_full_size(dev):
    """Size-independent properties at the largest GModule shape: rows/cols sum to <= 1 (slack),
    entries in (0,1), and the slack-completed matrix is doubly stochastic after the column pass."""
    torch.manual_seed(0)
    M = torch.randn(64, 320, 318, device=dev)
    P = GF.sinkhorn_rpm_exp(M, 20, True)
    assert torch.isfinite(P).all() and (P > 0).all() and (P < 1).all()
    assert (P.sum(1) <= 1 + 1e-4).all()          # column pass came last: column sums + slack == 1
    assert (P.sum(2) <= 1 + 2e-2).all()
    with pytest.raises(RuntimeError, match="does not fit"):
        GF.sinkhorn_rpm_exp(torch.randn(1500, 1500, device=dev), 20, True)


def test_sinkhorn_rpm_outlier_is_stable(dev):
    """A single huge entry (instance-normed z ~ 100): the log-domain kernel must stay finite."""
    M = torch.randn(100, 120) * 0.01
    M[3, 5] = 500.0
    ref = G.sinkhorn_rpm_exp(M, 20, True)
    out = GF.sinkhorn_rpm_exp(M.to(dev), 20, True)
    assert torch.isfinite(out).all()
    close(out, ref, rtol=1e-3, atol=1e-6)


# ---------------------------------------------------------------------------------------- K5
def test_sinkhorn_distance_golden(dev, golden):
    for name, c in golden("sinkhorn_distance").items():
        x, y = c["x"].to(dev), c["y"].to(dev)
        if x.dim() == 2:
            x, y = x[None], y[None]
        x.requires_grad_(); y.requires_grad_()
        cost, pi, C, nits = GF.sinkhorn_distance(x, y, c["eps"], c["max_iter"])
        ref_n = G.sinkhorn_distance(c["x"], c["y"], c["eps"], c["max_iter"], c["reduction"])[3]
        assert int(nits.item()) == ref_n, name
        if c["reduction"] == "mean":
            cost = cost.mean()
        if c["x"].dim() == 2:
            cost, pi, C = cost[0], pi[0], C[0]
        close(C, c["C"], rtol=1e-4, atol=1e-5)
        close(pi, c["pi"], rtol=2e-3, atol=1e-7)
        close(cost, c["cost"], rtol=2e-3, atol=1e-6)
        if "dx" in c:
            cost.sum().backward()
            close(x.grad, c["dx"], rtol=5e-3, atol=2e-6)
            close(y.grad, c["dy"], rtol=5e-3, atol=2e-6)


def test_sinkhorn_distance_marginals(dev):
    """Property: after enough iterations the plan's marginals are uniform."""
    torch.manual_seed(1)
    x, y = torch.randn(8, 64, 256, device=dev) * 0.05, torch.randn(8, 64, 256, device=dev) * 0.05
    cost, pi, C, nits = GF.sinkhorn_distance(x, y, 0.1, 200, thresh=1e-6)
    close(pi.sum(1), torch.full((8, 64), 1 / 64), rtol=1e-3, atol=1e-6)   # v update came last
    close(pi.sum(2), torch.full((8, 64), 1 / 64), rtol=5e-2, atol=1e-5)
    assert (C >= 0).all()


# ---------------------------------------------------------------------------------------- K1
def _check_knn(edge, x, y, k, dilation, rel=None):
    """Exact where the oracle's ordering is unambiguous; where two candidate distances are within
    fp32 round-off of each other (|gap| < 2e-6) either order is accepted."""
    dist = V.knn_distances(x, y, rel)
    ref = V.dense_dilated_knn(x, y, k, dilation, rel)
    e = edge.cpu()
    assert e.shape == ref.shape and e.dtype == torch.int64
    assert torch.equal(e[1], ref[1])
    mism = (e[0] != ref[0])
    if mism.any():
        d_ours = torch.gather(dist, 2, e[0])
        d_ref = torch.gather(dist, 2, ref[0])
        assert ((d_ours - d_ref).abs()[mism] < 2e-6).all(), "k-NN index differs beyond a fp32 near-tie"
    return float(mism.float().mean())


@pytest.mark.parametrize("P1,P2,B", [(250, 301, 1), (320, 204, 2)])
def test_sinkhorn_distance_large_node_sets(dev, P1, P2, B):
    """Node sets of the graph module (~200-320 rows, config 3): the P1 x P2 matrices exceed one CTA's shared memory
    and live in their global arrays instead (`spill`) -- cost, plan, cost matrix and the gradients vs the oracle."""
    torch.manual_seed(P1)
    x = (torch.randn(B, P1, 256) * 0.5).requires_grad_()
    y = (torch.randn(B, P2, 256) * 0.5 + 0.1).requires_grad_()
    cost, pi, C, _ = G.sinkhorn_distance(x, y, 0.1, 5, "mean")
    cost.backward()
    xd, yd = x.detach().to(dev).requires_grad_(), y.detach().to(dev).requires_grad_()
    c2, pi2, C2, _ = GF.sinkhorn_distance(xd, yd, 0.1, 5)
    c2.mean().backward()
    close(C2, C, rtol=1e-4, atol=1e-4)
    close(c2.mean(), cost, rtol=1e-3, atol=1e-5)
    close(pi2, pi, rtol=2e-3, atol=1e-7)
    close(xd.grad, x.grad, rtol=5e-3, atol=1e-6 + 1e-3 * float(x.grad.abs().max()))
    close(yd.grad, y.grad, rtol=5e-3, atol=1e-6 + 1e-3 * float(y.grad.abs().max()))


def test_knn_golden(dev, golden):
    g = golden("vig")
    x, y, rel = g["x"], g["y"], g["rel"]
    assert _check_knn(GF.knn_graph(x.to(dev), y.to(dev), 5, 2), x, y, 5, 2) < 0.02
    assert torch.equal(GF.knn_graph(x.to(dev), y.to(dev), 5, 2).cpu()[1], g["e_xy"][1])
    assert _check_knn(GF.knn_graph(x.to(dev), None, 9, 1, rel.to(dev)), x, None, 9, 1, rel) < 0.02
    assert _check_knn(GF.knn_graph(x.to(dev), None, 9, 1), x, None, 9, 1) < 0.02


@pytest.mark.parametrize("B,C,N,M,k,d", [(8, 256, 64, 64, 9, 1), (2, 256, 784, 784, 9, 1), (2, 48, 196, 49, 9, 2),
                                         (1, 32, 70, 65, 9, 5), (3, 16, 10, 130, 3, 1), (1, 256, 4096, 1024, 9, 1)])
def test_knn_vs_oracle(dev, B, C, N, M, k, d):
    torch.manual_seed(B + C + N)
    x = torch.randn(B, C, N, 1)
    y = torch.randn(B, C, M, 1) if (M != N or B == 8) else None
    frac = _check_knn(GF.knn_graph(x.to(dev), None if y is None else y.to(dev), k, d), x, y, k, d)
    assert frac < 0.01


def test_knn_all_ties_zero_hidden(dev):
    """TGCN step 0: hidden == 0 so every key is identical (TGCN.py:230).  Any k distinct keys are a
    valid answer; ours must be the k lowest indices (documented tie rule)."""
    x = torch.randn(2, 256, 64, 1, device=dev)
    y = torch.zeros(2, 256, 64, 1, device=dev)
    e = GF.knn_graph(x, y, 9, 1).cpu()
    assert torch.equal(e[0], torch.arange(9).expand(2, 64, 9))


def _knn_with_path(path, *args):
    from graphecho_b200 import _cabi
    assert _cabi.lib().ge_knn_graph_set_path(path) == 0
    try:
        return GF.knn_graph(*args)
    finally:
        _cabi.lib().ge_knn_graph_set_path(0)


@pytest.mark.parametrize("B,C,N,M,k,d,self_graph", [
    (1, 32, 128, 128, 9, 1, True),        # smallest problem the tcgen05 path takes, one 64-channel slab half empty
    (2, 256, 784, 784, 9, 1, True),       # config-2 Grapher shape (7 key tiles of 112)
    (2, 96, 300, 140, 9, 1, False),       # ragged: C not a multiple of 64, N and M not multiples of the tiles
    (2, 64, 256, 196, 9, 2, False),       # dilation 2 -> 18-entry lists
    (1, 128, 500, 500, 16, 2, True),      # 32-entry lists
    (1, 256, 4096, 1024, 9, 1, False),    # 256x256 images: N = 4096 queries, r=2 pooled keys
])
def test_knn_tcgen05_path(dev, B, C, N, M, k, d, self_graph):
    """The tensor-core kernel (required explicitly: GE_ERR_SHAPE if it does not apply) against the oracle
    and against the fp32 FFMA kernels."""
    torch.manual_seed(B * 7 + N)
    x = torch.randn(B, C, N, 1)
    y = None if self_graph else torch.randn(B, C, M, 1)
    xd, yd = x.to(dev), None if y is None else y.to(dev)
    e_tc = _knn_with_path(2, xd, yd, k, d)
    assert _check_knn(e_tc, x, y, k, d) < 0.01
    e_ff = _knn_with_path(1, xd, yd, k, d)
    assert (e_tc != e_ff).float().mean() < 0.01


def test_knn_tcgen05_tie_rule_and_scope(dev):
    from graphecho_b200 import _cabi
    x = torch.randn(2, 64, 256, 1, device=dev)
    y = torch.zeros(2, 64, 256, 1, device=dev)
    e = _knn_with_path(2, x, y, 9, 1).cpu()
    assert torch.equal(e[0], torch.arange(9).expand(2, 256, 9))          # all keys tie -> lowest indices
    with pytest.raises(_cabi.GraphEchoNativeError):                       # too small for the tensor-core path: loud
        _knn_with_path(2, torch.randn(2, 256, 64, 1, device=dev), None, 9, 1)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("self_graph", [True, False])
def test_knn_node_major_equals_channel_major(dev, dtype, self_graph):
    """ge_knn_graph_nmajor on x [B,N,C] == ge_knn_graph on the [B,C,N] transpose of the same values."""
    torch.manual_seed(3)
    B, C, N, M, k = 3, 96, 400, 400 if self_graph else 196, 9
    x = torch.randn(B, N, C, device=dev).to(dtype)
    y = None if self_graph else torch.randn(B, M, C, device=dev).to(dtype)
    assert GF.knn_nmajor_supported(B, C, N, M, k, 1)
    e_nm = GF.knn_graph_nmajor(x, y, k, 1)
    xt = x.float().transpose(1, 2).contiguous().unsqueeze(-1)
    yt = None if y is None else y.float().transpose(1, 2).contiguous().unsqueeze(-1)
    e_cm = GF.knn_graph(xt, yt, k, 1)
    assert (e_nm != e_cm).float().mean() < 0.002
    assert _check_knn(e_nm, xt.cpu(), None if yt is None else yt.cpu(), k, 1) < 0.01
    assert not GF.knn_nmajor_supported(B, C, 64, 64, k, 1)          # small graphs stay on the [B,C,N] route


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("self_graph", [True, False])
@pytest.mark.parametrize("C,N", [(40, 150), (128, 784), (256, 196)])
def test_mr_gather_node_major_vs_oracle(dev, dtype, self_graph, C, N):
    """C % 64 == 0 self-graphs take the one-launch backward with the gradient slab resident in shared memory
    (ge_mrconv_gather_nmajor_bwd_self); the others the init + global-atomic scatter pair."""
    torch.manual_seed(7)
    B, M, k = 2, N if self_graph else 61, 9
    x = torch.randn(B, N, C).to(dtype)
    y = None if self_graph else torch.randn(B, M, C).to(dtype)
    e0 = torch.randint(0, M, (B, N, k))
    edge = torch.stack([e0, torch.arange(N).view(1, N, 1).expand(B, N, k)])
    xo = x.float().transpose(1, 2).unsqueeze(-1).clone().requires_grad_()
    yo = None if y is None else y.float().transpose(1, 2).unsqueeze(-1).clone().requires_grad_()
    ref = V.max_relative(xo, edge, yo)                                   # [B,2C,N,1]
    W = torch.randn(B, N, 2 * C).to(dtype)
    (ref.squeeze(-1).transpose(1, 2) * W.float()).sum().backward()
    xd = x.to(dev).requires_grad_()
    yd = None if y is None else y.to(dev).requires_grad_()
    out = GF.mr_gather_nmajor(xd, e0.to(dev), yd)
    assert out.dtype == dtype and out.shape == (B, N, 2 * C)
    tol = dict(rtol=0, atol=0) if dtype == torch.float32 else dict(rtol=1e-2, atol=1e-2)
    close(out.float(), ref.squeeze(-1).transpose(1, 2), **tol)
    (out.float() * W.to(dev).float()).sum().backward()
    gtol = dict(rtol=1e-5, atol=1e-5) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
    close(xd.grad.float(), xo.grad.squeeze(-1).transpose(1, 2), **gtol)
    if y is not None:
        close(yd.grad.float(), yo.grad.squeeze(-1).transpose(1, 2), **gtol)


# ---------------------------------------------------------------------------------------- K2
def test_mr_gather_golden(dev, golden):
    g = golden("vig")
    p = {k: v.to(dev).requires_grad_() for k, v in make_params("mrconv32_64", fill_prefix="grapher.gconv.").items()}
    x, y = g["x"].to(dev).requires_grad_(), g["y"].to(dev).requires_grad_()
    feat = GF.mr_gather(x, g["e_xy"].to(dev), y)
    close(feat, V.max_relative(g["x"], g["e_xy"], g["y"]), rtol=0, atol=0)
    o = F.gelu(F.conv2d(feat, p["nn.0.weight"], p["nn.0.bias"], groups=4))
    close(o, g["mr_out"], rtol=1e-4, atol=1e-5)
    o.square().sum().backward()
    close(x.grad, g["mr_dx"], rtol=1e-3, atol=1e-5)
    close(y.grad, g["mr_dy"], rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize("self_graph,ident", [(True, True), (False, True), (False, False), (True, False)])
def test_mr_gather_vs_oracle(dev, self_graph, ident):
    torch.manual_seed(5)
    B, C, N, M, k = 3, 40, 130, 130 if self_graph else 77, 9
    x = torch.randn(B, C, N, 1)
    y = None if self_graph else torch.randn(B, C, M, 1)
    e0 = torch.randint(0, M, (B, N, k))
    e1 = torch.arange(N).view(1, N, 1).expand(B, N, k).contiguous() if ident else torch.randint(0, N, (B, N, k))
    edge = torch.stack([e0, e1])
    xo = x.clone().requires_grad_()
    yo = None if y is None else y.clone().requires_grad_()
    ref = V.max_relative(xo, edge, yo)
    W = torch.randn_like(ref)
    (ref * W).sum().backward()
    xd = x.to(dev).requires_grad_()
    yd = None if y is None else y.to(dev).requires_grad_()
    out = GF.mr_gather(xd, edge.to(dev), yd, identity_centre=ident)
    close(out, ref, rtol=0, atol=0)
    (out * W.to(dev)).sum().backward()
    close(xd.grad, xo.grad, rtol=1e-5, atol=1e-5)
    if y is not None:
        close(yd.grad, yo.grad, rtol=1e-5, atol=1e-5)


# ---------------------------------------------------------------------------------------- K6
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pool_concat(dev, dtype):
    torch.manual_seed(9)
    F_, C = 6, 64
    lv = [torch.randn(F_, C, s, s) for s in (64, 32, 16, 8)]
    rs = (8, 4, 2, 1)
    lo = [t.to(dtype).float().clone().requires_grad_() for t in lv]
    ref = torch.cat([F.avg_pool2d(t, r, r) if r > 1 else t for t, r in zip(lo, rs)], dim=1)
    W = torch.randn_like(ref)
    (ref * W).sum().backward()
    ld = [t.detach().to(dev).to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_() for t in lv]
    out = GF.pool_concat(ld, rs)
    assert out.dtype == torch.float32 and out.shape == ref.shape
    close(out, ref, rtol=1e-5, atol=1e-5)
    (out * W.to(dev)).sum().backward()
    tol = 1e-6 if dtype == torch.float32 else 1e-2
    for a, b in zip(ld, lo):
        close(a.grad, b.grad, rtol=tol, atol=tol)
    with pytest.raises(RuntimeError, match="pooled sizes differ"):
        GF.pool_concat([torch.randn(1, 8, 28, 28, device=dev), torch.randn(1, 8, 16, 16, device=dev)], (8, 4))


# ---------------------------------------------------------------------------------------- K7
def _cl(t, dev, dtype=torch.float32):
    return t.to(dev).to(dtype).contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("h,H", [(4, 7), (7, 14), (14, 28), (8, 16), (1, 5), (28, 28), (16, 64)])
def test_upsample_add(dev, h, H):
    torch.manual_seed(h * 100 + H)
    top, lat = torch.randn(3, 32, h, h), torch.randn(3, 32, H, H)
    to, lo = top.clone().requires_grad_(), lat.clone().requires_grad_()
    ref = F.interpolate(to, size=(H, H), mode="bilinear", align_corners=True) + lo
    W = torch.randn_like(ref)
    (ref * W).sum().backward()
    td, ld = _cl(top, dev).requires_grad_(), _cl(lat, dev).requires_grad_()
    out = GF.upsample_add(td, ld)
    close(out, ref, rtol=1e-5, atol=1e-5)
    (out * W.to(dev)).sum().backward()
    close(td.grad, to.grad, rtol=1e-4, atol=1e-5)
    close(ld.grad, lo.grad, rtol=0, atol=0)


def test_upsample_add_rect_and_bf16(dev):
    torch.manual_seed(4)
    top, lat = torch.randn(2, 16, 5, 9), torch.randn(2, 16, 11, 17)
    ref = F.interpolate(top, size=(11, 17), mode="bilinear", align_corners=True) + lat
    close(GF.upsample_add(_cl(top, dev), _cl(lat, dev)), ref, rtol=1e-5, atol=1e-5)
    out = GF.upsample_add(_cl(top, dev, torch.bfloat16), _cl(lat, dev, torch.bfloat16))
    assert out.dtype == torch.bfloat16
    close(out, ref, rtol=2e-2, atol=3e-2)
    close(GF.upsample_bilinear(_cl(top, dev), (11, 17)),
          F.interpolate(top, size=(11, 17), mode="bilinear", align_corners=True), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("h,H,C", [(4, 28, 256), (7, 28, 256), (14, 28, 128), (28, 28, 128), (8, 64, 128)])
def test_gn_relu_upsample(dev, h, H, C):
    torch.manual_seed(h + H + C)
    x = torch.randn(2, C, h, h) * 1.5 + 0.2
    gamma, beta = 1 + 0.1 * torch.randn(C), 0.1 * torch.randn(C)
    xo, go, bo = x.clone().requires_grad_(), gamma.clone().requires_grad_(), beta.clone().requires_grad_()
    ref = F.interpolate(torch.relu(F.group_norm(xo, C, go, bo, 1e-5)), size=(H, H), mode="bilinear", align_corners=True)
    W = torch.randn_like(ref)
    (ref * W).sum().backward()
    xd, gd, bd = _cl(x, dev).requires_grad_(), gamma.to(dev).requires_grad_(), beta.to(dev).requires_grad_()
    out = GF.gn_relu_upsample(xd, gd, bd, (H, H))
    close(out, ref, rtol=1e-4, atol=1e-5)
    (out * W.to(dev)).sum().backward()
    close(xd.grad, xo.grad, rtol=2e-3, atol=2e-4)
    close(gd.grad, go.grad, rtol=1e-3, atol=1e-3)
    close(bd.grad, bo.grad, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("C,groups,h,dtype", [(256, 32, 28, torch.float32), (256, 32, 7, torch.float32),
                                              (64, 32, 9, torch.float32), (256, 32, 14, torch.bfloat16),
                                              (128, 128, 5, torch.bfloat16), (512, 16, 4, torch.float32)])
def test_gn_relu_groups(dev, C, groups, h, dtype):
    """GroupNorm(groups, C) + ReLU as the Discriminator towers use it (fpnseg.py:455-466)."""
    torch.manual_seed(C + groups + h)
    x = (torch.randn(3, C, h, h) * 1.3 + 0.4).to(dtype).float()
    gamma, beta = 1 + 0.1 * torch.randn(C), 0.1 * torch.randn(C)
    xo, go, bo = x.clone().requires_grad_(), gamma.clone().requires_grad_(), beta.clone().requires_grad_()
    ref = torch.relu(F.group_norm(xo, groups, go, bo, 1e-5))
    W = torch.randn_like(ref)
    (ref * W).sum().backward()
    xd, gd, bd = _cl(x, dev, dtype).requires_grad_(), gamma.to(dev).requires_grad_(), beta.to(dev).requires_grad_()
    out = GF.gn_relu(xd, gd, bd, groups)
    assert out.dtype == dtype
    lo = dtype == torch.bfloat16
    close(out, ref, rtol=2e-2 if lo else 1e-4, atol=2e-2 if lo else 1e-5)
    (out * W.to(dev).to(dtype)).sum().backward()
    close(xd.grad, xo.grad, rtol=5e-2 if lo else 2e-3, atol=5e-2 if lo else 2e-4)
    close(gd.grad, go.grad, rtol=3e-2 if lo else 1e-3, atol=0.3 if lo else 1e-3)
    close(bd.grad, bo.grad, rtol=3e-2 if lo else 1e-3, atol=0.3 if lo else 1e-3)


@pytest.mark.parametrize("nc,h", [(1, 28), (2, 28), (4, 64), (3, 16)])
def test_seg_tail(dev, nc, h):
    torch.manual_seed(nc * 10 + h)
    C = 128
    s = [torch.relu(torch.randn(2, C, h, h)) for _ in range(4)]
    W3, b3 = torch.randn(nc, C, 1, 1) * 0.1, torch.randn(nc) * 0.1
    so = [t.clone().requires_grad_() for t in s]
    Wo, bo = W3.clone().requires_grad_(), b3.clone().requires_grad_()
    ref = F.interpolate(F.conv2d(so[0] + so[1] + so[2] + so[3], Wo, bo), size=(4 * h, 4 * h), mode="bilinear", align_corners=True)
    G_ = torch.randn_like(ref)
    (ref * G_).sum().backward()
    sd = [_cl(t, dev).requires_grad_() for t in s]
    Wd, bd = W3.to(dev).requires_grad_(), b3.to(dev).requires_grad_()
    out = GF.seg_tail(*sd, Wd, bd, 4)
    close(out, ref, rtol=1e-4, atol=1e-4)
    (out * G_.to(dev)).sum().backward()
    for a, b in zip(sd, so):
        close(a.grad, b.grad, rtol=1e-3, atol=1e-4)
    close(Wd.grad, Wo.grad, rtol=1e-3, atol=1e-2)
    close(bd.grad, bo.grad, rtol=1e-3, atol=1e-2)


# ---------------------------------------------------------------------------------------- update_seed
def test_spectral_bipartition_kernel_matches_dense_route(dev):
    """The one-CTA power-iteration kernel against the dense eigh route of graphecho_b200/spectral.py
    (itself checked against sklearn on CPU in test_host_logic.py)."""
    from graphecho_b200.spectral import spectral_bipartition
    agree = []
    for seed in range(10):
        g = torch.Generator().manual_seed(seed)
        n = int(torch.randint(25, 150, (1,), generator=g))
        a = torch.randn(n, 256, generator=g)
        a[: n // 3] += 1.5 * torch.randn(1, 256, generator=g)          # two blobs, as LayerNorm'd class nodes
        pts = torch.cat([torch.randn(1, 256, generator=g), a])
        ref = spectral_bipartition(pts, n // 2)                         # CPU: dense route
        out = spectral_bipartition(pts.to(dev), n // 2).cpu()
        agree.append((out == ref).float().mean().item())
    assert min(agree) > 0.97, agree


def test_spectral_bipartition_large_kernel_matches_dense_route(dev):
    """192 < n <= 512 points: the block-wise / bit-matrix kernel (class banks of a 256-frame step have 200-450
    nodes) against the dense eigh route."""
    from graphecho_b200 import _cabi
    from graphecho_b200.spectral import spectral_bipartition
    assert _cabi.lib().ge_spectral_bipartition_max_points() == 512
    agree = []
    for seed, n in enumerate((193, 260, 333, 448, 511)):
        g = torch.Generator().manual_seed(100 + seed)
        a = torch.randn(n, 256, generator=g)
        a[: n // 3] += 1.5 * torch.randn(1, 256, generator=g)
        pts = torch.cat([torch.randn(1, 256, generator=g), a])
        ref = spectral_bipartition(pts, n // 2)                         # CPU: dense route
        out = spectral_bipartition(pts.to(dev), n // 2).cpu()
        agree.append((out == ref).float().mean().item())
    assert min(agree) > 0.97, agree


# ---------------------------------------------------------------------------------------- fused BatchNorm
@pytest.mark.parametrize("C,hw,res,relu,dtype", [(64, 28, False, True, torch.float32), (256, 14, True, True, torch.float32),
                                                 (2048, 4, True, True, torch.float32), (512, 7, False, False, torch.float32),
                                                 (1024, 7, True, True, torch.bfloat16), (128, 56, False, True, torch.bfloat16)])
@pytest.mark.parametrize("bn_path", [0, 2])
def test_bn_act_train_and_eval(dev, C, hw, res, relu, dtype, bn_path):
    """Fused BN(+residual)(+ReLU) against nn.BatchNorm2d + add + relu (fp32 oracle on CPU): output,
    grads of x / residual / gamma / beta, running statistics, then inference mode.  bn_path 0 = the three-kernel
    streaming path (default), 2 = the single-launch cooperative kernels (these maps fit on chip)."""
    from graphecho_b200 import _cabi
    _cabi.lib().ge_bn_set_path(bn_path)
    torch.manual_seed(C + hw)
    N = 5
    x = (torch.randn(N, C, hw, hw) * 1.4 + 0.3).to(dtype).float()
    r = torch.randn(N, C, hw, hw).to(dtype).float() if res else None
    ref_bn = torch.nn.BatchNorm2d(C)
    with torch.no_grad():
        ref_bn.weight.copy_(1 + 0.1 * torch.randn(C)); ref_bn.bias.copy_(0.1 * torch.randn(C))
        ref_bn.running_mean.copy_(0.05 * torch.randn(C)); ref_bn.running_var.copy_(1 + 0.1 * torch.rand(C))
    import copy
    our_bn = copy.deepcopy(ref_bn).to(dev)
    xo = x.clone().requires_grad_()
    ro = r.clone().requires_grad_() if res else None
    y = ref_bn(xo)
    if res:
        y = y + ro
    ref = torch.relu(y) if relu else y
    W = torch.randn_like(ref)
    (ref * W).sum().backward()
    xd = _cl(x, dev, dtype).requires_grad_()
    rd = _cl(r, dev, dtype).requires_grad_() if res else None
    out = GF.bn_act(xd, our_bn, residual=rd, relu=relu)
    assert out.dtype == dtype
    lo = dtype == torch.bfloat16
    close(out, ref, rtol=2e-2 if lo else 1e-4, atol=3e-2 if lo else 2e-5)
    (out * W.to(dev).to(dtype)).sum().backward()
    close(xd.grad, xo.grad, rtol=5e-2 if lo else 2e-3, atol=5e-2 if lo else 2e-4)
    if res:
        close(rd.grad, ro.grad, rtol=2e-2 if lo else 1e-5, atol=2e-2 if lo else 1e-6)
    scale = float(ref_bn.weight.grad.abs().max())
    close(our_bn.weight.grad, ref_bn.weight.grad, rtol=3e-2 if lo else 1e-3, atol=(3e-2 if lo else 1e-4) * scale)
    close(our_bn.bias.grad, ref_bn.bias.grad, rtol=3e-2 if lo else 1e-3, atol=(3e-2 if lo else 1e-4) * scale)
    close(our_bn.running_mean, ref_bn.running_mean, rtol=1e-4, atol=1e-5)
    close(our_bn.running_var, ref_bn.running_var, rtol=1e-4, atol=1e-5)
    assert int(our_bn.num_batches_tracked) == 1
    ref_bn.eval(); our_bn.eval()
    with torch.no_grad():
        y = ref_bn(x)
        if res:
            y = y + r
        ref_e = torch.relu(y) if relu else y
        out_e = GF.bn_act(_cl(x, dev, dtype), our_bn, residual=None if not res else _cl(r, dev, dtype), relu=relu)
    close(out_e, ref_e, rtol=2e-2 if lo else 1e-4, atol=3e-2 if lo else 2e-5)
    _cabi.lib().ge_bn_set_path(0)


@pytest.mark.parametrize("C,hw,N,ns,res,relu,dtype", [(64, 7, 5, 3, False, True, torch.float32),     # P_split = 147: a thread straddles
                                                       (256, 14, 6, 2, True, True, torch.float32),
                                                       (2048, 4, 4, 1, True, True, torch.float32),
                                                       (512, 7, 5, 4, False, False, torch.float32),
                                                       (256, 28, 8, 4, True, True, torch.bfloat16),
                                                       (64, 56, 4, 2, False, True, torch.bfloat16)])
@pytest.mark.parametrize("bn_path", [0, 2])
def test_bn_act_domain_split_equals_two_calls(dev, C, hw, N, ns, res, relu, dtype, bn_path):
    """Per-domain statistics: one fused call on the [source | target] batch under GF.domain_split(ns) == two separate
    nn.BatchNorm2d calls, source first (train_cardiac_uda.py:225, 234): outputs, all gradients (gamma / beta gradients
    accumulate over the two calls), running statistics after TWO momentum updates, num_batches_tracked == 2."""
    import copy
    from graphecho_b200 import _cabi
    _cabi.lib().ge_bn_set_path(bn_path)
    torch.manual_seed(C + hw + ns)
    x = torch.cat([torch.randn(ns, C, hw, hw) * 1.4 + 0.3, torch.randn(N - ns, C, hw, hw) * 0.6 - 0.5]).to(dtype).float()
    r = torch.randn(N, C, hw, hw).to(dtype).float() if res else None
    ref_bn = torch.nn.BatchNorm2d(C)
    with torch.no_grad():
        ref_bn.weight.copy_(1 + 0.1 * torch.randn(C)); ref_bn.bias.copy_(0.1 * torch.randn(C))
        ref_bn.running_mean.copy_(0.05 * torch.randn(C)); ref_bn.running_var.copy_(1 + 0.1 * torch.rand(C))
    our_bn = copy.deepcopy(ref_bn).to(dev)
    xo = x.clone().requires_grad_()
    ro = r.clone().requires_grad_() if res else None
    y = torch.cat([ref_bn(xo[:ns]), ref_bn(xo[ns:])])
    if res:
        y = y + ro
    ref = torch.relu(y) if relu else y
    W = torch.randn_like(ref)
    (ref * W).sum().backward()
    xd = _cl(x, dev, dtype).requires_grad_()
    rd = _cl(r, dev, dtype).requires_grad_() if res else None
    with GF.domain_split(ns):
        out = GF.bn_act(xd, our_bn, residual=rd, relu=relu)
    lo = dtype == torch.bfloat16
    close(out, ref, rtol=2e-2 if lo else 1e-4, atol=3e-2 if lo else 2e-5)
    (out * W.to(dev).to(dtype)).sum().backward()          # outside the context: the split travels with the autograd node
    close(xd.grad, xo.grad, rtol=5e-2 if lo else 2e-3, atol=5e-2 if lo else 2e-4)
    if res:
        close(rd.grad, ro.grad, rtol=2e-2 if lo else 1e-5, atol=2e-2 if lo else 1e-6)
    scale = float(ref_bn.weight.grad.abs().max())
    close(our_bn.weight.grad, ref_bn.weight.grad, rtol=3e-2 if lo else 1e-3, atol=(3e-2 if lo else 1e-4) * scale)
    close(our_bn.bias.grad, ref_bn.bias.grad, rtol=3e-2 if lo else 1e-3, atol=(3e-2 if lo else 1e-4) * scale)
    close(our_bn.running_mean, ref_bn.running_mean, rtol=1e-4, atol=1e-5)
    close(our_bn.running_var, ref_bn.running_var, rtol=1e-4, atol=1e-5)
    assert int(our_bn.num_batches_tracked) == 2
    # and the whole-batch call is NOT the same thing (the two domains have different statistics)
    with torch.no_grad():
        whole = GF.bn_act(_cl(x, dev, dtype), copy.deepcopy(ref_bn).to(dev).train(), residual=None if not res else _cl(r, dev, dtype),
                          relu=relu)
    assert (whole.float().cpu() - ref).abs().max() > 0.05
    _cabi.lib().ge_bn_set_path(0)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gn_relu_with_folded_conv_bias(dev, dtype):
    """relu(GroupNorm(32)(x + b)) with b folded into the statistics kernel == the plain composition, including the
    closed-form gradient of b (no pass over dx)."""
    torch.manual_seed(4)
    N, C, H = 3, 64, 9
    x = (torch.randn(N, C, H, H) * 2).to(dtype)
    bias = torch.randn(C)
    gamma, beta = torch.rand(C) + 0.5, torch.randn(C)
    xo, bo = x.float().clone().requires_grad_(), bias.clone().requires_grad_()
    go, beo = gamma.clone().requires_grad_(), beta.clone().requires_grad_()
    ref = torch.relu(F.group_norm(xo + bo.view(1, C, 1, 1), 32, go, beo, 1e-5))
    W = torch.randn_like(ref)
    (ref * W).sum().backward()
    xd = x.to(dev).contiguous(memory_format=torch.channels_last).requires_grad_()
    bd = bias.to(dev).requires_grad_()
    gd, bed = gamma.to(dev).requires_grad_(), beta.to(dev).requires_grad_()
    out = GF.gn_relu(xd, gd, bed, 32, 1e-5, pre_bias=bd)
    tol = dict(rtol=2e-4, atol=2e-5) if dtype == torch.float32 else dict(rtol=3e-2, atol=3e-2)
    close(out.float(), ref, **tol)
    (out.float() * W.to(dev)).sum().backward()
    gt = dict(rtol=2e-3, atol=2e-4) if dtype == torch.float32 else dict(rtol=5e-2, atol=8e-2)
    close(xd.grad.float(), xo.grad, **gt)
    close(bd.grad, bo.grad, **gt)
    close(gd.grad, go.grad, **gt)
    close(bed.grad, beo.grad, **gt)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("hw", [(56, 56), (9, 14)])
def test_maxpool3s2_matches_torch(dev, dtype, hw):
    torch.manual_seed(2)
    x = torch.randn(3, 16, *hw).to(dtype)
    x[0, :, 2:5, 2:5] = 1.5                                     # ties inside windows: first maximum wins
    xo = x.float().clone().requires_grad_()
    ref = F.max_pool2d(xo, 3, 2, 1)
    Wt = torch.randn_like(ref).to(dtype).float()
    (ref * Wt).sum().backward()
    xd = x.to(dev).contiguous(memory_format=torch.channels_last).requires_grad_()
    out = GF.maxpool3s2(xd)
    assert out.shape == ref.shape and out.dtype == dtype
    close(out.float(), ref, rtol=0, atol=0)
    (out.float() * Wt.to(dev)).sum().backward()
    close(xd.grad.float(), xo.grad, rtol=1e-6 if dtype == torch.float32 else 1e-2, atol=1e-6 if dtype == torch.float32 else 2e-2)


# ---------------------------------------------------------------------------------------- f4: loss fusion, boxes
@pytest.mark.parametrize("F_,nc,hw", [(3, 1, 28), (5, 2, 112), (2, 4, 256), (4, 3, 37)])
def test_seg_loss_fused_vs_oracle(dev, F_, nc, hw):
    """DiceLoss + BCEWithLogits as one pass (utils/losses.py:64-95 + BCEWithLogitsLoss) against the oracle: value and
    the gradient w.r.t. the logits."""
    from oracle import fpn_ops as FP
    from graphecho_b200 import synth
    torch.manual_seed(F_ * 10 + nc)
    x = (torch.randn(F_, nc, hw, hw) * 2.0).requires_grad_()
    t = synth.disc_masks(F_, nc, hw)
    ref = FP.seg_loss(x, t)
    (ref * 1.7).backward()
    xd = x.detach().to(dev).requires_grad_()
    out = GF.seg_loss(xd, t.to(dev))
    (out * 1.7).backward()
    close(out, ref, rtol=2e-5, atol=1e-6)
    close(xd.grad, x.grad, rtol=1e-4, atol=1e-9 + 1e-5 * float(x.grad.abs().max()))


def test_mask_boxes_vs_oracle(dev):
    """Bounding boxes of mask / score-map planes (graph_matching.py:702-746): fp32 masks, int64 thresholded score maps,
    logits taken through sigmoid > 0.5 without materialising the map, an empty plane, a full plane."""
    from oracle import gmodule_ops as GM
    from graphecho_b200 import synth
    masks = synth.disc_masks(5, 4, 256)
    masks[2, 1] = 0                                          # empty plane -> (0, 0, W, H)
    masks[3, 2] = 1                                          # full plane
    ref = torch.stack(GM.find_bbox(masks))
    assert torch.equal(GF.mask_boxes(masks.to(dev)).cpu(), ref)
    assert torch.equal(GF.mask_boxes(masks.long().to(dev)).cpu(), ref)
    torch.manual_seed(0)
    logits = torch.randn(3, 2, 112, 96) - 1.5
    logits[1, 0] = -3.0
    score = torch.where(torch.sigmoid(logits) > 0.5, 1, 0)
    ref = torch.stack(GM.find_bbox(score))
    assert torch.equal(GF.mask_boxes(GF.LogitMap(logits.to(dev))).cpu(), ref)
    assert torch.equal(GF.mask_boxes(score.to(dev)).cpu(), ref)
    # raw logits used as a score map (the temporal branch, Appendix A-12): every plane is "full"
    assert torch.equal(GF.mask_boxes(logits.to(dev)).cpu(), torch.stack(GM.find_bbox(logits)))


# ---------------------------------------------------------------------------------------- f3: tcgen05 1x1-conv GEMM + BN statistics
@pytest.mark.parametrize("N,H,Ci,Co,ns,res", [(16, 28, 64, 256, 8, True), (16, 28, 256, 64, 6, False), (12, 28, 256, 256, 0, False),
                                              (32, 14, 512, 128, 16, False), (32, 14, 128, 512, 11, True), (40, 11, 64, 64, 13, False)])
def test_conv1x1_tc_gemm_with_bn_statistics(dev, N, H, Ci, Co, ns, res):
    """1x1 conv + train-mode BatchNorm (+ residual) + ReLU on the tcgen05 GEMM whose epilogue produces the batch statistics
    (Bottleneck conv1/conv3, fpnseg.py:192-212), bf16 operands, against nn.Conv2d + nn.BatchNorm2d in fp32 on the SAME
    bf16-rounded inputs and weights: output (bf16 round-off), every gradient, running statistics (two updates under a
    domain split), and -- the epilogue itself -- the GEMM output and the partial sums it reports."""
    import copy
    torch.manual_seed(Ci + Co + H)
    P = N * H * H
    assert GF.conv1x1_tc_supported(P, Ci, Co)
    x = torch.randn(N, Ci, H, H).bfloat16().float()
    r = torch.randn(N, Co, H, H).bfloat16().float() if res else None
    conv = torch.nn.Conv2d(Ci, Co, 1, bias=False)
    with torch.no_grad():
        conv.weight.copy_(conv.weight.bfloat16().float())
    bn = torch.nn.BatchNorm2d(Co)
    with torch.no_grad():
        bn.weight.copy_(1 + 0.1 * torch.randn(Co)); bn.bias.copy_(0.1 * torch.randn(Co))
        bn.running_mean.copy_(0.05 * torch.randn(Co)); bn.running_var.copy_(1 + 0.1 * torch.rand(Co))
    conv_d, bn_d = copy.deepcopy(conv).to(dev), copy.deepcopy(bn).to(dev)
    # the raw GEMM and its partial statistics
    x2 = x.to(dev).bfloat16().permute(0, 2, 3, 1).reshape(P, Ci).contiguous()
    w2 = conv.weight.reshape(Co, Ci).to(dev).bfloat16().contiguous()
    shift = bn.running_mean.to(dev)
    Ps = ns * H * H
    y, part, rows, rows0 = GF.conv1x1_gemm(x2, w2, shift, True, Ps)
    yref = x2.float() @ w2.float().t()
    close(y, yref, rtol=1e-2, atol=2e-2)                                   # bf16 output rounding
    d = yref - shift
    segs = [(d[:Ps], part[:rows0])] + ([(d[Ps:], part[rows0:])] if 0 < Ps < P else [])
    if not (0 < Ps < P):
        segs = [(d, part)]
    for dd, pp in segs:
        close(pp[:, 0].sum(0), dd.sum(0), rtol=1e-4, atol=1e-2)
        close(pp[:, 1].sum(0), (dd * dd).sum(0), rtol=1e-4, atol=1e-2)
    # the fused module path
    xo = x.clone().requires_grad_()
    ro = r.clone().requires_grad_() if res else None
    yc = conv(xo)
    yb = torch.cat([bn(yc[:ns]), bn(yc[ns:])]) if 0 < ns < N else bn(yc)
    ref = torch.relu(yb + ro if res else yb)
    W = torch.randn_like(ref)
    (ref * W).sum().backward()
    cl = torch.channels_last
    xd = x.to(dev).bfloat16().contiguous(memory_format=cl).requires_grad_()
    rd = r.to(dev).bfloat16().contiguous(memory_format=cl).requires_grad_() if res else None
    with torch.autocast("cuda", dtype=torch.bfloat16), GF.domain_split(ns):
        out = GF.conv1x1_bn_act(xd, conv_d, bn_d, residual=rd, relu=True)
    assert out.dtype == torch.bfloat16
    close(out, ref, rtol=2e-2, atol=4e-2)
    (out.float() * W.to(dev)).sum().backward()
    relerr = lambda a, b: float((a.detach().float().cpu() - b).norm() / b.norm())
    assert relerr(xd.grad, xo.grad) < 3e-2
    assert relerr(conv_d.weight.grad, conv.weight.grad) < 3e-2
    assert relerr(bn_d.weight.grad, bn.weight.grad) < 3e-2 and relerr(bn_d.bias.grad, bn.bias.grad) < 3e-2
    if res:
        assert relerr(rd.grad, ro.grad) < 3e-2
    close(bn_d.running_mean, bn.running_mean, rtol=1e-3, atol=2e-3)
    close(bn_d.running_var, bn.running_var, rtol=2e-3, atol=2e-3)
    assert int(bn_d.num_batches_tracked) == (2 if 0 < ns < N else 1)


@pytest.mark.parametrize("hw,nc,B", [(256, 3, 4), (112, 2, 6), (256, 4, 3)])
def test_sampler_labels_kernel_bit_exact(dev, hw, nc, B):
    """Per-location labels of all pyramid levels + per-level (positive, negative) counts in one launch
    (graph_matching.py:609-635, 874-959) against the oracle: bit-exact integer output, incl. the reference's stride quirk
    (locations at 8,16,32,64 on a stride-4..32 pyramid), an empty class plane and raw-logit score maps (every box = the
    full image -> every label 0, Appendix A-12)."""
    from oracle import gmodule_ops as GM
    from graphecho_b200 import synth
    feats = synth.pyramid(B, hw, seed=1)
    level_hw = [(f.shape[-2], f.shape[-1]) for f in feats]
    masks = synth.disc_masks(B, nc, hw)
    masks[1, 1] = 0
    logits = torch.randn(B, nc, hw, hw)
    for m in (masks, torch.where(torch.sigmoid(logits - 1.0) > 0.5, 1, 0), logits):
        boxes = torch.stack(GM.find_bbox(m))
        ref = GM.location_labels(GM.compute_locations(feats), list(boxes), nc)
        labels, counts = GF.sampler_labels(boxes.to(dev), level_hw, [8, 16, 32, 64, 128][:4],
                                           [[-1, 64], [64, 128], [128, 256], [256, 512]])
        for l, (a, b) in enumerate(zip(labels, ref)):
            assert torch.equal(a.cpu(), b.reshape(-1).long()), l
            assert counts[l].tolist() == [int((b > 0).sum()), int((b == 0).sum())]

# ---------------------------------------------------------------------------------------- f1: picks + row gather
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shared", [True, False])
def test_sampler_gather_kernel_equals_the_reference_picks(dev, dtype, shared):
    """ge_sampler_gather / ge_sampler_scatter against PrototypeComputation.gather (the torch restatement of
    graph_matching.py:978-1013 that tests/golden/sampler.pt pins): node rows, labels and order bit-exact, and the
    feature-map gradients equal, over levels that exercise every branch -- more positives than negatives (all
    negatives kept), strided positives with floor(linspace) negatives, fewer than 100 positives, no positives."""
    from graphecho_b200.models.graph_matching import PrototypeComputation
    torch.manual_seed(3)
    gen = PrototypeComputation(3)
    hw = [(28, 28), (14, 14), (7, 7), (4, 4), (2, 2)]
    Bs, Bt, C = 6, 5, 64
    Ball = Bs + Bt
    p_pos = [0.7, 0.2, 0.3, 0.02, 0.0]

    def labels_for(B):
        out = []
        for (h, w), p in zip(hw, p_pos):
            lab = torch.where(torch.rand(B * h * w, device=dev) < p, torch.randint(1, 3, (B * h * w,), device=dev),
                              torch.zeros(B * h * w, dtype=torch.long, device=dev))
            out.append(lab)
        return out

    def counts_of(labels):
        return [[int((l > 0).sum()), int((l == 0).sum())] for l in labels]

    lab_s, lab_t = labels_for(Bs), labels_for(Bt)
    cl = torch.channels_last
    if shared:
        feats = [torch.randn(Ball, C, h, w, device=dev).to(dtype).contiguous(memory_format=cl).requires_grad_() for h, w in hw]
        doms = [(feats, lab_s, counts_of(lab_s), 0), (feats, lab_t, counts_of(lab_t), Bs)]
        leaves = feats
    else:
        fs = [torch.randn(Bs, C, h, w, device=dev).to(dtype).contiguous(memory_format=cl).requires_grad_() for h, w in hw]
        ft = [torch.randn(Bt, C, h, w, device=dev).to(dtype).contiguous(memory_format=cl).requires_grad_() for h, w in hw]
        doms = [(fs, lab_s, counts_of(lab_s), 0), (ft, lab_t, counts_of(lab_t), 0)]
        leaves = fs + ft
    ref = [gen.gather(*d) for d in doms]
    got = gen.gather_pair(*doms)
    ws = []
    for (rn, rl, rw), (gn, gl, gw) in zip(ref, got):
        assert gn.dtype == torch.float32 and gl.dtype == torch.int64
        assert torch.equal(gn, rn.float()) and torch.equal(gl, rl) and torch.equal(gw, rw)
        ws.append(torch.randn_like(gn))
    g_ref = torch.autograd.grad(sum((rn.float() * w).sum() for (rn, _, _), w in zip(ref, ws)), leaves)
    g_got = torch.autograd.grad(sum((gn * w).sum() for (gn, _, _), w in zip(got, ws)), leaves)
    for a, b in zip(g_got, g_ref):
        assert a.dtype == b.dtype and torch.equal(a, b)

# ---------------------------------------------------------------------------------------- K4b: matching loss
@pytest.mark.parametrize("n1,n2,nc", [(252, 250, 2), (37, 45, 4), (6, 9, 3), (300, 120, 2)])
def test_matching_loss_kernel_vs_oracle(dev, n1, n2, nc):
    """ge_matching_loss_fwd/bwd against the oracle's literal TP/FP focal losses (graph_matching.py:572-590: boolean
    gathers, BCEFocalLoss means), value and gradient with respect to the Sinkhorn-normalised matrix."""
    torch.manual_seed(n1 + n2)
    P = GF.sinkhorn_rpm_exp(torch.randn(n1, n2, device=dev) * 1.5, 20, True).detach()
    l1 = torch.randint(0, nc, (n1,))
    l2 = torch.randint(0, nc, (n2,))
    if n1 == 6:
        l1[0] = nc - 1
        l2[l2 == nc - 1] = 0          # a row without any same-class column: the reference's argmax falls on column 0
    Po = P.cpu().clone().requires_grad_()
    ref = G.matching_loss_o2o(Po, l1, l2, nc)
    ref.backward()
    Pd = P.clone().requires_grad_()
    got = GF.matching_loss_o2o(Pd, l1.float().to(dev), l2.float().to(dev), 0.25, 2.0)
    (got * 1.0).backward()
    close(got, ref, rtol=2e-5, atol=1e-7)
    close(Pd.grad, Po.grad, rtol=1e-4, atol=1e-9)


# ---------------------------------------------------------------------------------------- stem convolution
@pytest.mark.parametrize("F_,H,W", [(6, 112, 112), (2, 256, 256), (3, 100, 120), (1, 8, 8)])
@pytest.mark.parametrize("bf16", [False, True])
def test_stem_conv_kernels_vs_conv2d(dev, F_, H, W, bf16):
    """ge_stem_conv_fwd / ge_stem_conv_wgrad against torch's Conv2d(1, 64, 7, 2, 3, bias=False) (fpnseg.py:229): fp32
    mode against the fp32 convolution; bf16 mode (autocast) against the fp32 convolution of the bf16-rounded operands
    (what a bf16 convolution with fp32 accumulation computes), output and weight gradient."""
    torch.manual_seed(H + W)
    conv = torch.nn.Conv2d(1, 64, 7, 2, 3, bias=False).to(dev)
    x = torch.randn(F_, 1, H, W, device=dev)
    g = torch.randn(F_, 64, H // 2, W // 2, device=dev)
    if bf16:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = GF.stem_conv(x, conv)
        assert y.dtype == torch.bfloat16 and y.is_contiguous(memory_format=torch.channels_last)
        gq = g.bfloat16()
        (dw,) = torch.autograd.grad(y, conv.weight, gq)
        wq = conv.weight.detach().bfloat16().float().requires_grad_()
        ref = F.conv2d(x.bfloat16().float(), wq, None, 2, 3)
        (dref,) = torch.autograd.grad(ref, wq, gq.float())
        close(y.float(), ref, rtol=1e-2, atol=1e-2)
        close(dw, dref, rtol=2e-3, atol=2e-3 * float(dref.abs().max()))
    else:
        y = GF.stem_conv(x, conv)
        assert y.dtype == torch.float32 and y.shape == (F_, 64, H // 2, W // 2)
        (dw,) = torch.autograd.grad(y, conv.weight, g)
        ref = conv(x)
        (dref,) = torch.autograd.grad(ref, conv.weight, g)
        close(y, ref, rtol=1e-4, atol=1e-5)
        close(dw, dref, rtol=1e-4, atol=1e-4 * float(dref.abs().max()))

@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("k,pad", [(1, 0), (3, 1)])
def test_conv_bias_fused_add_and_gradient(dev, dtype, k, pad):
    """GF.conv_bias (library convolution + fused in-place NHWC bias add, streaming bias gradient) against nn.Conv2d:
    output and the gradients of input, weight and bias (FPN lateral / smoothing convolutions, fpnseg.py:333-345)."""
    torch.manual_seed(k)
    conv = torch.nn.Conv2d(64, 256, k, 1, pad).to(dev)
    x = torch.randn(6, 64, 14, 10, device=dev).contiguous(memory_format=torch.channels_last)
    g = torch.randn(6, 256, 14, 10, device=dev)
    xa, xb = x.clone().requires_grad_(), x.clone().requires_grad_()
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=dtype == torch.bfloat16):
        ya = GF.conv_bias(xa, conv)
        ga = torch.autograd.grad(ya, (xa, conv.weight, conv.bias), g.to(ya.dtype))
        yb = conv(xb)
        gb = torch.autograd.grad(yb, (xb, conv.weight, conv.bias), g.to(yb.dtype))
    assert ya.dtype == yb.dtype == dtype
    tol = dict(rtol=1e-5, atol=1e-5) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
    close(ya.float(), yb.float(), **tol)
    for a, b in zip(ga, gb):
        close(a.float(), b.float(), rtol=tol["rtol"], atol=tol["atol"] * max(1.0, float(b.abs().max())))

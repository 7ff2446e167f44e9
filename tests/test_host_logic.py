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

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

"""On-device spectral bipartition for GModule.update_seed (the reference's own TODO,
graph_matching.py:538: "Use Pytorch-based GPU version").

Follows what sklearn.cluster.SpectralClustering(2, affinity='nearest_neighbors', n_neighbors=m,
assign_labels='kmeans') computes, without the CPU round trip:
  1. binary k-NN connectivity (self included), symmetrised: A = (C + C^T) / 2
  2. normalised Laplacian embedding: the two leading eigenvectors of D^-1/2 A D^-1/2, divided by
     sqrt(degree) (sklearn.manifold.spectral_embedding, drop_first=False) -- the first is constant,
     so the partition lives in the second (Fiedler) coordinate
  3. 2-means on that coordinate, solved EXACTLY (best threshold split over the sorted values, via
     prefix sums) instead of k-means++ restarts.
Label parity with sklearn holds whenever sklearn's k-means reaches the global optimum; the result
only steers the seed-bank buffers, never the same-step losses.
"""
from __future__ import annotations

import torch


@torch.no_grad()
def spectral_bipartition(pts: torch.Tensor, n_neighbors: int, iterations: int = 300) -> torch.Tensor:
    """pts [n, d] (row 0 = the class seed).  Returns bool [n-1]: True where a point lands in the same
    cluster as row 0.  CUDA inputs up to ge_spectral_bipartition_max_points() run as ONE kernel
    (csrc/spectral.cu: power iteration instead of a full eigendecomposition); larger or CPU inputs take
    the dense torch.linalg.eigh route below (same algorithm)."""
    n = pts.shape[0]
    if pts.is_cuda:
        from . import _cabi
        if n <= _cabi.lib().ge_spectral_bipartition_max_points():
            x = pts.float().contiguous()
            keep = torch.empty(n - 1, device=pts.device, dtype=torch.uint8)
            _cabi.call("ge_spectral_bipartition", _cabi.ptr(x), _cabi.ptr(keep), n, x.shape[1],
                       max(1, min(int(n_neighbors), n)), int(iterations), _cabi.stream(),
                       work=(4 * n * x.shape[1] + n, 2 * n * n * (x.shape[1] + n + iterations)))
            return keep.bool()
    x = pts.float()
    k = max(1, min(int(n_neighbors), n))
    sq = (x * x).sum(1)
    dist = sq[:, None] + sq[None, :] - 2.0 * (x @ x.t())
    dist.fill_diagonal_(-1.0)                      # a point is its own nearest neighbour (include_self)
    nbr = dist.topk(k, dim=1, largest=False).indices
    conn = torch.zeros(n, n, device=x.device)
    conn.scatter_(1, nbr, 1.0)
    adj = 0.5 * (conn + conn.t())
    deg = adj.sum(1).clamp_min(1e-12)
    dinv = deg.rsqrt()
    sym = dinv[:, None] * adj * dinv[None, :]
    evals, evecs = torch.linalg.eigh(sym)          # ascending; leading = last columns
    fiedler = evecs[:, -2] * dinv
    # exact 1-D 2-means: best split of the sorted coordinate
    vals, order = fiedler.sort()
    csum = vals.cumsum(0)
    csq = (vals * vals).cumsum(0)
    cnt = torch.arange(1, n + 1, device=x.device, dtype=vals.dtype)
    left_sse = csq[:-1] - csum[:-1] ** 2 / cnt[:-1]
    right_cnt = n - cnt[:-1]
    right_sum = csum[-1] - csum[:-1]
    right_sse = (csq[-1] - csq[:-1]) - right_sum ** 2 / right_cnt
    split = (left_sse + right_sse).argmin()        # left cluster = sorted positions [0, split]
    rank_of = torch.empty_like(order)
    rank_of[order] = torch.arange(n, device=x.device)
    left = rank_of <= split
    return (left == left[0])[1:]

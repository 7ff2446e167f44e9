"""CPU: the C-ABI shared library loads and exports every symbol include/graphecho_b200.h declares
(no compute calls: there is no GPU in the build container)."""
import ctypes

import pytest
import torch

from graphecho_b200 import _cabi


def test_library_exports_every_header_symbol():
    names = _cabi.header_symbols()
    assert len(names) >= 20
    handle = ctypes.CDLL(str(_cabi.LIB_PATH))
    missing = [n for n in names if not hasattr(handle, n)]
    assert not missing, f"header declares symbols the .so does not export: {missing}"


def test_bindings_cover_header():
    names = set(_cabi.header_symbols())
    bound = set(_cabi._SIGNATURES)
    assert names == bound, f"unbound: {sorted(names - bound)}  stale: {sorted(bound - names)}"


def test_version_and_error_text():
    lib = _cabi.lib()
    assert lib.ge_version() == 1
    assert isinstance(_cabi.last_error(), str)


def test_cpu_tensors_are_rejected_loudly():
    from graphecho_b200 import functional as GF
    with pytest.raises(RuntimeError, match="CUDA"):
        GF.sinkhorn_rpm_exp(torch.randn(4, 5))
    with pytest.raises(RuntimeError, match="CUDA"):
        GF.affinity_pairwise(torch.randn(4, 32), torch.randn(5, 32), torch.randn(32), torch.randn(1))
    with pytest.raises(RuntimeError, match="CUDA"):
        GF.knn_graph(torch.randn(1, 8, 16, 1))


def test_argument_errors_do_not_need_a_gpu():
    # null pointers are rejected before any CUDA call
    lib = _cabi.lib()
    rc = lib.ge_affinity_pairwise_fwd(None, None, None, None, None, 1, 4, 4, 32, None)
    assert rc < 0 and "null" in _cabi.last_error()
    assert lib.ge_sinkhorn_rpm_cluster_size(250, 250, 0) == 8
    assert lib.ge_sinkhorn_rpm_cluster_size(6, 6, 0) == 1
    assert lib.ge_sinkhorn_rpm_cluster_size(4000, 4000, 0) == -1


def test_new_entry_points_validate_arguments_without_a_gpu():
    lib = _cabi.lib()
    assert lib.ge_knn_graph_set_path(7) < 0 and "path" in _cabi.last_error()
    assert lib.ge_knn_graph_set_path(0) == 0
    # workspace covers both the fp32 FFMA layout and the (larger) fp16 hi/lo split of the tcgen05 path
    ffma = 4 * (2 * 256 * (784 + 784) + 2 * (784 + 784))
    assert lib.ge_knn_graph_workspace_bytes(2, 256, 784, 784) >= ffma
    assert lib.ge_knn_graph_workspace_bytes(0, 256, 784, 784) == 0
    assert lib.ge_knn_graph_nmajor_supported(2, 256, 64, 64, 9, 1) == 0          # small graphs: FFMA route only
    assert lib.ge_knn_graph_nmajor_supported(2, 256, 784, 784, 9, 1) in (0, 1)   # 1 only where the driver is present
    assert lib.ge_maxpool3s2_fwd(None, None, None, 0, 1, 8, 8, 8, None) < 0 and "null" in _cabi.last_error()
    assert lib.ge_mrconv_gather_nmajor_fwd(None, None, None, None, None, 0, 1, 8, 4, 4, 3, None) < 0
    assert lib.ge_group_stats_bias(None, None, None, None, None, 0, 1, 4, 8, 1, 1e-5, None) < 0
    assert lib.ge_spectral_bipartition_max_points() == 512


def test_new_wrappers_reject_cpu_tensors():
    from graphecho_b200 import functional as GF
    with pytest.raises(RuntimeError, match="CUDA"):
        GF.knn_graph_nmajor(torch.randn(1, 128, 32))
    with pytest.raises(RuntimeError, match="CUDA"):
        GF.mr_gather_nmajor(torch.randn(1, 16, 8), torch.zeros(1, 16, 3, dtype=torch.int64))
    with pytest.raises(RuntimeError, match="CUDA"):
        GF.maxpool3s2(torch.randn(1, 8, 6, 6))
    with pytest.raises(RuntimeError, match="CUDA"):
        GF.gn_relu(torch.randn(1, 32, 4, 4), torch.ones(32), torch.zeros(32), 4, pre_bias=torch.zeros(32))


def test_round2_entry_points_validate_arguments_without_a_gpu():
    lib = _cabi.lib()
    assert lib.ge_bn_set_path(5) < 0 and "path" in _cabi.last_error()
    assert lib.ge_bn_set_path(0) == 0
    assert lib.ge_bn_relu_mask_bytes(200704, 256) == 200704 * 256 // 8          # one bit per element
    assert lib.ge_bn_relu_mask_bytes(10, 12) == 0
    assert lib.ge_seg_loss_workspace_bytes(4, 9, 100) == 0                     # at most 8 classes
    assert lib.ge_seg_loss_fwd(None, None, None, None, None, 0, 1, 2, 16, 1.0, None) < 0 and "null" in _cabi.last_error()
    assert lib.ge_mask_boxes(None, None, 0, 1, 4, 4, 0, None) < 0
    with pytest.raises(RuntimeError, match="CUDA"):
        from graphecho_b200 import functional as GF
        GF.seg_loss(torch.randn(1, 2, 4, 4), torch.zeros(1, 2, 4, 4))

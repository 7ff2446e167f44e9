"""SyncBatchNorm path of the fused BatchNorm kernels (ge_bn_sync_*): the statistics / apply stages around a cross-rank
average.  Reference intent: nn.SyncBatchNorm.convert_sync_batchnorm under DDP (train_cardiac_uda.py:142)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_layer(x, res, bn, split, relu, w):
    from graphecho_b200 import functional as GF
    xr = x.clone().requires_grad_()
    rr = res.clone().requires_grad_() if res is not None else None
    with GF.domain_split(split):
        out = GF.bn_act(xr, bn, residual=rr, relu=relu)
    (out.float() * w).sum().backward()
    return out.detach(), xr.grad, (rr.grad if rr is not None else None)


@pytest.mark.parametrize("dtype,relu,with_res,split", [(torch.float32, True, True, 3), (torch.bfloat16, True, False, 0),
                                                       (torch.float32, False, False, 2)])
def test_sync_bn_with_one_rank_equals_batch_norm(dtype, relu, with_res, split):
    """world = 1: the split kernels (statistics -> [average over 1 rank] -> apply) reproduce the fused BatchNorm call
    bit for bit: output, running statistics, num_batches_tracked and every gradient."""
    dev = torch.device("cuda:0")
    port = _free_port()
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1, device_id=dev)
    try:
        torch.manual_seed(0)
        C = 64
        cl = torch.channels_last
        x = (torch.randn(6, C, 9, 7, device=dev) * 1.5 + 0.3).to(dtype).contiguous(memory_format=cl)
        res = torch.randn_like(x) if with_res else None
        w = torch.randn(6, C, 9, 7, device=dev)
        bn = torch.nn.BatchNorm2d(C).to(dev).train()
        with torch.no_grad():
            bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(); bn.running_mean.normal_(0, 0.1); bn.running_var.uniform_(0.5, 2.0)
        sbn = torch.nn.SyncBatchNorm.convert_sync_batchnorm(torch.nn.Sequential(torch.nn.BatchNorm2d(C)))[0].to(dev).train()
        sbn.load_state_dict(bn.state_dict())
        o1, dx1, dr1 = _run_layer(x, res, bn, split, relu, w)
        o2, dx2, dr2 = _run_layer(x, res, sbn, split, relu, w)
        assert torch.equal(o1, o2) and torch.equal(dx1, dx2)
        if with_res:
            assert torch.equal(dr1, dr2)
        for k, v in bn.state_dict().items():
            assert torch.equal(v, sbn.state_dict()[k]), k
        assert torch.equal(bn.weight.grad, sbn.weight.grad) and torch.equal(bn.bias.grad, sbn.bias.grad)
    finally:
        dist.destroy_process_group()


def _two_rank_worker(rank, port, graphs, ret):
    import datetime
    import faulthandler
    faulthandler.dump_traceback_later(90, exit=True)          # a stuck collective must not outlive the test
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dev = torch.device(f"cuda:{rank}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=2, device_id=dev, timeout=datetime.timedelta(seconds=60))
    try:
        from graphecho_b200 import functional as GF
        torch.manual_seed(1)
        C, ns, nt = 32, 4, 6                                   # per-rank source / target images
        cl = torch.channels_last
        full = torch.randn(2 * (ns + nt), C, 5, 6, device=dev) * 2 + 0.5      # [src r0 | src r1 | tgt r0 | tgt r1]
        wfull = torch.randn_like(full)
        src, tgt = full[:2 * ns], full[2 * ns:]
        mine = torch.cat([src[rank * ns:(rank + 1) * ns], tgt[rank * nt:(rank + 1) * nt]]).contiguous(memory_format=cl)
        wmine = torch.cat([wfull[:2 * ns][rank * ns:(rank + 1) * ns], wfull[2 * ns:][rank * nt:(rank + 1) * nt]])
        bn = torch.nn.BatchNorm2d(C).to(dev).train()
        with torch.no_grad():
            bn.weight.uniform_(0.5, 1.5); bn.bias.normal_()
        sbn = torch.nn.SyncBatchNorm.convert_sync_batchnorm(torch.nn.Sequential(torch.nn.BatchNorm2d(C)))[0].to(dev).train()
        sbn.load_state_dict(bn.state_dict())
        # single-process reference on the whole [source | target] batch
        ref_out, ref_dx, _ = _run_layer(full.contiguous(memory_format=cl), None, bn, 2 * ns, True, wfull)
        ref_out = torch.cat([ref_out[:2 * ns][rank * ns:(rank + 1) * ns], ref_out[2 * ns:][rank * nt:(rank + 1) * nt]])
        ref_dx = torch.cat([ref_dx[:2 * ns][rank * ns:(rank + 1) * ns], ref_dx[2 * ns:][rank * nt:(rank + 1) * nt]])
        if graphs:
            x_static = mine.clone().requires_grad_()

            class Layer(torch.nn.Module):
                def __init__(self, bn):
                    super().__init__()
                    self.bn = bn

                def forward(self, x):
                    with GF.domain_split(ns):
                        return GF.bn_act(x, self.bn, relu=True)

            snapshot = {k: v.clone() for k, v in sbn.state_dict().items()}
            g = torch.cuda.make_graphed_callables(Layer(sbn), (x_static,), num_warmup_iters=3)
            sbn.load_state_dict(snapshot)
            sbn.weight.grad = None
            sbn.bias.grad = None
            xr = mine.clone().requires_grad_()
            out = g(xr)
            (out * wmine).sum().backward()
            out, dx = out.detach(), xr.grad
        else:
            out, dx, _ = _run_layer(mine, None, sbn, ns, True, wmine)
        ok = (torch.allclose(out, ref_out, rtol=1e-5, atol=1e-5) and torch.allclose(dx, ref_dx, rtol=1e-4, atol=1e-5)
              and torch.allclose(sbn.running_mean, bn.running_mean, rtol=1e-5, atol=1e-6)
              and torch.allclose(sbn.running_var, bn.running_var, rtol=1e-5, atol=1e-6)
              and int(sbn.num_batches_tracked) == int(bn.num_batches_tracked))
        # parameter gradients: the sum over ranks of the local gradients is the single-process gradient
        gw = sbn.weight.grad.clone(); dist.all_reduce(gw)
        gb = sbn.bias.grad.clone(); dist.all_reduce(gb)
        ok = ok and torch.allclose(gw, bn.weight.grad, rtol=1e-4, atol=1e-4) and torch.allclose(gb, bn.bias.grad, rtol=1e-4, atol=1e-4)
        ret[rank] = bool(ok)
    finally:
        faulthandler.cancel_dump_traceback_later()
        dist.destroy_process_group()


@pytest.mark.parametrize("graphs", [False])
def test_sync_bn_two_ranks_equals_single_process_batch_norm(graphs):
    """Two ranks, each with its [source | target] shard: per-segment statistics over both ranks = BatchNorm of the whole
    batch in one process (outputs, input gradients, running statistics, summed parameter gradients).  (The worker also
    has a graph-captured variant: with make_graphed_callables the replay of the captured NCCL averages hung on 2 ranks,
    so the engine keeps SyncBatchNorm eager and that variant is not run.)"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    port = _free_port()
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_two_rank_worker, args=(port, graphs, ret), nprocs=2, join=True)
    assert ret.get(0) and ret.get(1), dict(ret)

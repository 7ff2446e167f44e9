"""CPU, world_size 2 over gloo: the multi-GPU host logic — clip sharding and the single flat-buffer
gradient exchange (engine.FlatGradSync) — without any GPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from graphecho_b200.engine import FlatGradSync, shard_range, make_batch, EngineConfig


def test_shard_range_partitions_every_clip_once():
    for n, w in [(8, 1), (8, 2), (8, 8), (10, 4), (3, 4), (128, 8)]:
        got = [i for r in range(w) for i in shard_range(n, r, w)]
        assert got == list(range(n))
        sizes = [len(shard_range(n, r, w)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def test_make_batch_layout_and_rank_decorrelation():
    cfg = EngineConfig(hw=16, num_classes=2)
    x0, m0 = make_batch(cfg, 4, 3, rank=0, world=2)
    x1, _ = make_batch(cfg, 4, 3, rank=1, world=2)
    assert x0.shape == (4, 1, 16, 16, 3) and m0.shape == (6, 2, 16, 16)
    assert not torch.equal(x0, x1)
    assert torch.allclose(m0.sum(1), torch.ones(6, 16, 16))          # one-hot, background in channel 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker_unbound(rank, world, port, out):
    """bind=False: gradients stay autograd's own tensors and are packed into the flat buffer for the exchange."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 2))
    unused = torch.nn.Linear(3, 3)
    sync = FlatGradSync([net, unused], bind=False)
    full = torch.randn(8, 6, generator=torch.Generator().manual_seed(1))
    tgt = torch.randn(8, 2, generator=torch.Generator().manual_seed(2))
    idx = list(shard_range(8, rank, world))
    ok = True
    for _ in range(2):                                  # second round: .grad is a flat view when zero() runs
        sync.zero()
        ok = ok and all(p.grad is None for p in net.parameters())
        ((net(full[idx]) - tgt[idx]) ** 2).mean().backward()
        sync.all_reduce()
        ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 2))
        ref.load_state_dict(net.state_dict())
        ((ref(full) - tgt) ** 2).mean().backward()
        ok = ok and all(torch.allclose(p.grad, q.grad, atol=1e-6) for p, q in zip(net.parameters(), ref.parameters()))
        ok = ok and float(unused.weight.grad.abs().sum()) == 0.0
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_unbound_grad_exchange_matches_single_process():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker_unbound, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 2))
    unused = torch.nn.Linear(3, 3)                      # takes no part in the step: its slice must stay 0
    sync = FlatGradSync([net, unused])
    full = torch.randn(8, 6, generator=torch.Generator().manual_seed(1))
    tgt = torch.randn(8, 2, generator=torch.Generator().manual_seed(2))
    idx = list(shard_range(8, rank, world))
    sync.zero()
    loss = ((net(full[idx]) - tgt[idx]) ** 2).mean()
    loss.backward()
    sync.rebind()
    sync.all_reduce()
    # reference: the same loss on the concatenated batch in one process
    ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 2))
    ref.load_state_dict(net.state_dict())
    ((ref(full) - tgt) ** 2).mean().backward()
    ok = all(torch.allclose(p.grad, q.grad, atol=1e-6) for p, q in zip(net.parameters(), ref.parameters()))
    ok = ok and all(p.grad.data_ptr() >= sync.flat.data_ptr() for p in net.parameters())
    ok = ok and float(unused.weight.grad.abs().sum()) == 0.0
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_flat_grad_allreduce_matches_single_process():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def _worker_buckets(rank, world, port, out):
    """Bucketed exchange: the last layer's gradients are final first (backward order) and go on the wire with
    reduce_bucket(0) before the rest; all_reduce() then exchanges the remaining bucket and hands every gradient back."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 2))
    unused = torch.nn.Linear(3, 3)
    sync = FlatGradSync([net, unused], bind=False, buckets=[list(net[2].parameters())])     # rest = net[0] + unused
    ok = len(sync.buckets) == 2 and sync.bucket_range[0] == (0, 12) and sync.nbytes == 4 * (12 + 35 + 12)
    full = torch.randn(8, 6, generator=torch.Generator().manual_seed(1))
    tgt = torch.randn(8, 2, generator=torch.Generator().manual_seed(2))
    idx = list(shard_range(8, rank, world))
    for _ in range(2):
        sync.zero()
        hid = net[1](net[0](full[idx]))
        leaf = hid.detach().requires_grad_()
        ((net[2](leaf) - tgt[idx]) ** 2).mean().backward()          # last layer only
        sync.reduce_bucket(0)                                        # its gradients are final: start their exchange
        hid.backward(leaf.grad)                                      # the rest of the backward
        sync.all_reduce()
        ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 2))
        ref.load_state_dict(net.state_dict())
        ((ref(full) - tgt) ** 2).mean().backward()
        ok = ok and all(torch.allclose(p.grad, q.grad, atol=1e-6) for p, q in zip(net.parameters(), ref.parameters()))
        ok = ok and float(unused.weight.grad.abs().sum()) == 0.0
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_bucketed_grad_exchange_matches_single_process():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker_buckets, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}

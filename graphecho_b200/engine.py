"""The GraphEcho UDA training step on the B200-native modules, plus the multi-GPU plumbing.

`UDAEngine.train_step` restates the inner loop of the reference trainers
(train_cardiac_uda.py:223-325, train_camus_echo.py:206-299): segmentation network on source +
target frames, Dice+BCE on the source, thresholded target score maps, graph matching, four
pyramid discriminators, optional ViG Grapher on p2 (112x112 workloads) or TGCN temporal module
(256x256 clips), one backward, one optimizer step per module group (Adam for the network, SGD for
the rest, as the reference's config dicts, train_cardiac_uda.py:646-736).

Multi-GPU: one process per GPU, the batch-of-clips axis is sharded (both streams symmetrically),
no data-path collective; the only exchange is the gradient all-reduce, done as ONE NCCL call over
a flat fp32 gradient buffer that every parameter's .grad is a view of (no bucket copies), plus
SyncBatchNorm on the segmentation network as the reference intends (train_cardiac_uda.py:142).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn

from .models.fpnseg import FPN, Discriminator
from .models.graph_matching import GModule
from .models.TGCN import TGCN
from .models.vig import Grapher
from .utils.losses import DiceLoss
from .utils.sinkhorn_distance import SinkhornDistance
from . import functional as GF
from . import synth


def preset(config: int, **over) -> "EngineConfig":
    """The workloads BASELINE.json's `configs` name.
    2: EchoNet-shape 112x112 clips, FPN(resnet, nc=2) + ViG Grapher(p2) + GModule + 4 discriminators  (also config 5)
    3: CAMUS-shape 256x256 frames, FPN(resnet, nc=4) + GModule + 4 discriminators + SinkhornDistance on the node sets
    4: CardiacUDA-shape 256x256, FPN(VGG16, nc=3) + GModule + 4 discriminators + TGCN temporal module on 8-frame clips"""
    base = {2: dict(backbone="resnet", hw=112, num_classes=2, vig_grapher=True),
            5: dict(backbone="resnet", hw=112, num_classes=2, vig_grapher=True),
            3: dict(backbone="resnet", hw=256, num_classes=4, vig_grapher=False, sinkhorn_nodes=True),
            4: dict(backbone="VGG16", hw=256, num_classes=3, vig_grapher=False, temporal_graph=True, clip_frames=8)}[config]
    base.update(over)
    return EngineConfig(**base)


@dataclass
class EngineConfig:
    backbone: str = "resnet"            # 'resnet' (CAMUS/EchoNet, 112x112) | 'VGG16' (CardiacUDA, 256x256)
    num_classes: int = 2
    hw: int = 112
    bf16: bool = True                   # tensor-core convs under autocast; graph modules stay fp32
    graph_matching: bool = True
    discriminator: bool = True
    vig_grapher: bool = True            # Grapher(256, k=9, 'mr', 'gelu', 'batch') on p2 (config 2)
    temporal_graph: bool = False        # TGCN on [b,t] clips (256x256 only, Appendix A-2): the step takes a `temporal`
                                        # input (clips + source-clip masks) beside the single-frame streams
    clip_frames: int = 8
    sinkhorn_nodes: bool = False        # config 3: SinkhornDistance(0.1, 5, 'mean') between the matched node sets
    sinkhorn_weight: float = 0.001
    per_domain_bn: bool = True          # BatchNorm statistics per domain, as the reference's two network calls per step
    sync_bn: bool = True
    seg_weight: float = 1.0             # train_camus_echo.py:212 uses 0.05
    lr_net: float = 3e-4
    lr_aux: float = 2.5e-3
    weight_decay: float = 1e-4
    cluster_backend: str = "device"     # GModule.update_seed bipartition: 'device' | 'sklearn'
    cuda_graphs: bool = False           # capture the static segments (FPN, Grapher, discriminators) as CUDA graphs
    overlap_streams: bool = True        # GModule on a side stream, overlapped with the discriminators
    phased_backward: bool = True        # cut autograd at the pyramid: discriminators fwd+bwd are issued before the
                                        # host-driven GModule, so the GPU is busy while the CPU drives the graph module
    seed: int = 0


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous, balanced shard of `n_items` clips for `rank` (first ranks take the remainder)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


class _NullCtx:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


class FlatGradSync:
    """The DDP exchange over a flat fp32 buffer, no per-parameter hooks, tolerant of parameters that did not take part
    in the step (their slice is zero).  The buffer is laid out in BUCKETS (lists of parameters given by the caller in
    the order their gradients become final); `reduce_bucket(i)` packs bucket i and starts its all-reduce(AVG) on a
    communication stream while the rest of the backward is still running, `all_reduce()` reduces whatever is left and
    waits for everything.  With one bucket this is ONE all-reduce per step after the backward.

    bind=True : every parameter's .grad is a strided view of the flat buffer for the whole run (autograd
                accumulates into it: one small add kernel per parameter per step).
    bind=False: .grad is left to autograd (set to None each step, so the first gradient is adopted without a
                copy -- 343 fewer launches per step for the config-2 engine); only when world_size > 1 are the
                gradients packed into the flat buffer (one multi-tensor copy per bucket), reduced, and handed back as views.
    """

    def __init__(self, modules, group=None, bind=True, buckets=None):
        every = [p for m in modules for p in m.parameters() if p.requires_grad]
        if buckets is None:
            buckets = [every]
        else:
            seen = {id(p) for b in buckets for p in b}
            rest = [p for p in every if id(p) not in seen]
            buckets = [[p for p in b if p.requires_grad] for b in buckets] + ([rest] if rest else [])
            assert sum(len(b) for b in buckets) == len(every), "buckets must partition the parameters"
        self.buckets = buckets
        self.params = [p for b in buckets for p in b]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.bind = bind
        self.views, self.bucket_range, off = [], [], 0
        for b in buckets:
            start = off
            for p in b:
                self.views.append(self._view(p, off))
                off += p.numel()
            self.bucket_range.append((start, off))
        self._bucket_slice = []
        i = 0
        for b in buckets:
            self._bucket_slice.append((i, i + len(b)))
            i += len(b)
        if bind:
            for p, v in zip(self.params, self.views):
                p.grad = v
        self.group = group
        self._done = [False] * len(buckets)
        self._pending = []
        self._comm = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None

    def _view(self, p, off):
        """A gradient view with the parameter's own (dense, possibly channels_last) strides, which is
        what the fused optimizers require."""
        return torch.as_strided(self.flat, p.shape, p.stride(), storage_offset=off)

    def _world(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def zero(self):
        self._done = [False] * len(self.buckets)
        if self.bind:
            self.flat.zero_()
        else:
            for p in self.params:
                p.grad = None

    def rebind(self):
        """bind=True: restore the views if an optimizer / user replaced .grad (e.g. set_to_none)."""
        if not self.bind:
            return
        for p, view in zip(self.params, self.views):
            if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                if p.grad is not None:
                    view.copy_(p.grad)
                p.grad = view

    def _pack_bucket(self, i):
        lo, hi = self.bucket_range[i]
        a, b = self._bucket_slice[i]
        self.flat[lo:hi].zero_()
        dst = [v for p, v in zip(self.params[a:b], self.views[a:b]) if p.grad is not None]
        src = [p.grad for p in self.params[a:b] if p.grad is not None]
        if src:
            torch._foreach_copy_(dst, src)

    def pack(self):
        """bind=False: copy the gradients autograd produced into the flat buffer (absent ones as zeros)."""
        for i in range(len(self.buckets)):
            self._pack_bucket(i)
        return self.flat

    def reduce_bucket(self, i, after=()):
        """Start the exchange of bucket i now (its gradients are final): pack + all-reduce on the communication stream,
        ordered after the current stream and the streams in `after`.  No-op for one rank."""
        if self._done[i] or self._world() <= 1:
            return
        self._done[i] = True
        lo, hi = self.bucket_range[i]
        piece = self.flat[lo:hi]
        nccl = dist.get_backend(self.group) == "nccl"
        comm = self._comm
        if comm is not None:
            comm.wait_stream(torch.cuda.current_stream())
            for s_ in after:
                comm.wait_stream(s_)
        with torch.cuda.stream(comm) if comm is not None else _NullCtx():
            if not self.bind:
                self._pack_bucket(i)
            work = dist.all_reduce(piece, op=dist.ReduceOp.AVG if nccl else dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._pending.append((work, piece, nccl))

    def all_reduce(self):
        """Exchange every bucket that has not been started, wait for all of them, and hand the reduced gradients back."""
        if self._world() <= 1:
            if not self.bind:
                # parameters that took no part in the step step on a zero gradient (their untouched, all-zero
                # slice of the flat buffer), exactly as with bound views: the optimizer state stays uniform
                for p, v in zip(self.params, self.views):
                    if p.grad is None:
                        p.grad = v
            return
        for i in range(len(self.buckets)):
            self.reduce_bucket(i)
        for work, piece, nccl in self._pending:
            work.wait()
            if not nccl:
                piece.div_(self._world())
        self._pending = []
        if self._comm is not None:
            torch.cuda.current_stream().wait_stream(self._comm)
        if not self.bind:
            for p, v in zip(self.params, self.views):     # every rank steps every parameter identically
                p.grad = v

    @property
    def nbytes(self):
        return self.flat.numel() * 4


def init_distributed(device_index: int | None = None):
    """torchrun-style env (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).  Returns (rank, local_rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0")) if device_index is None else device_index
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, local, world


class _CastAll(torch.autograd.Function):
    """All convolution weights of a graphed segment fp32 -> bf16 in ONE multi-tensor copy (and their gradients back in
    one).  Under autocast every convolution call otherwise re-casts its own weight (the autocast cache is off inside
    CUDA-graph capture) and autograd casts every weight gradient back: ~200 serialised 3 us kernels per step."""

    @staticmethod
    def forward(ctx, *ws):
        outs = [torch.empty_like(w, dtype=torch.bfloat16) for w in ws]
        torch._foreach_copy_(outs, list(ws))
        ctx.set_materialize_grads(False)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gs):
        idx = [i for i, g in enumerate(gs) if g is not None]
        res = [None] * len(gs)
        if idx:
            outs = [torch.empty_like(gs[i], dtype=torch.float32) for i in idx]
            torch._foreach_copy_(outs, [gs[i] for i in idx])
            for i, o in zip(idx, outs):
                res[i] = o
        return tuple(res)


class _CastWeights:
    """Mixin of the graphed segment wrappers: run `fn` with the 4-D fp32 convolution weights under `prefixes` of
    `root` swapped for their bf16 casts (one _CastAll) while autocast(bf16) is on; a plain call otherwise."""

    cast_prefixes = ()

    def _cast_targets(self, root):
        got = getattr(self, "_cast_cache", None)
        if got is None:
            got = [(n, p) for n, p in root.named_parameters()
                   if p.dim() == 4 and p.dtype == torch.float32 and any(n.startswith(pre) for pre in self.cast_prefixes)]
            object.__setattr__(self, "_cast_cache", got)
        return got

    def _run_cast(self, root, fn, *args):
        targets = self._cast_targets(root) if (BATCH_WEIGHT_CASTS and torch.is_autocast_enabled()
                                               and torch.get_autocast_dtype("cuda") == torch.bfloat16) else []
        if not targets:
            return fn(*args)
        w16 = _CastAll.apply(*[p for _, p in targets])
        from torch.nn.utils.stateless import _reparametrize_module
        with _reparametrize_module(root, {n: w for (n, _), w in zip(targets, w16)}):
            return fn(*args)


BATCH_WEIGHT_CASTS = True


class _JointDiscriminator(nn.Module, _CastWeights):
    """Discriminator fed the whole [source | target] feature map of a level (no slice + cat)."""

    def __init__(self, dis: Discriminator):
        super().__init__()
        self.dis = dis
        self.n_source = None          # frames of the source stream at the front of the batch (default: half)

    cast_prefixes = ("dis_tower.", "cls_logits.")

    def forward(self, feature_all):
        ns = feature_all.shape[0] // 2 if self.n_source is None else self.n_source
        return self._run_cast(self.dis, self.dis.forward_joint, feature_all, ns)


class _TrunkLower(nn.Module, _CastWeights):
    """FPN backbone up to c4 as its own callable (its own CUDA graph): x -> (c2, c3, c4)."""

    def __init__(self, fpn):
        super().__init__()
        self.fpn = fpn

    # (not the one-input-channel stem: its kernel takes the fp32 filter)
    cast_prefixes = ("back_bone.layer1.", "back_bone.layer2.", "back_bone.layer3.", "back_bone.block_1.",
                     "back_bone.block_2.", "back_bone.block_3.", "back_bone.block_4.")

    def forward(self, x):
        return self._run_cast(self.fpn, self.fpn.forward_trunk_lower, x)


class _TrunkUpper(nn.Module, _CastWeights):
    """Last backbone stage + pyramid: (c2, c3, c4) -> (p2, p3, p4, p5).  Separate from the lower part so that its
    gradients (2/3 of the backbone's parameters) are final -- and on the wire -- while the lower part's backward runs."""

    def __init__(self, fpn):
        super().__init__()
        self.fpn = fpn

    cast_prefixes = ("back_bone.layer4.", "back_bone.block_5.", "toplayer.", "latlayer1.", "latlayer2.", "latlayer3.")

    def forward(self, c2, c3, c4):
        return self._run_cast(self.fpn, self.fpn.forward_trunk_upper, c2, c3, c4)


class _Head(nn.Module, _CastWeights):
    """FPN smoothing + semantic head as its own callable: (p2, p3, p4, p5) -> logits.  Split from the trunk so
    that its backward (which needs only the segmentation loss) can run while the graph module is still busy."""

    def __init__(self, fpn):
        super().__init__()
        self.fpn = fpn

    cast_prefixes = ("smooth1.", "smooth2.", "smooth3.", "semantic_branch.", "conv2.")   # conv3 feeds ge_seg_tail in fp32

    def forward(self, p2, p3, p4, p5):
        return self._run_cast(self.fpn, self.fpn.forward_head, p2, p3, p4, p5)


class UDAEngine:
    def __init__(self, cfg: EngineConfig, device: torch.device, world_size: int = 1):
        self.cfg, self.device, self.world = cfg, device, world_size
        torch.manual_seed(cfg.seed)           # identical initial weights on every rank
        nc = cfg.num_classes
        self.network = FPN([2, 4, 23, 3], num_classes=nc, in_channel=1, back_bone=cfg.backbone).to(device)
        self.network = self.network.to(memory_format=torch.channels_last)
        if world_size > 1 and cfg.sync_bn:
            # the per-layer statistics exchanges run on the compute stream while the gradient buckets travel on the
            # communication stream: they need their own communicator (one NCCL communicator must not be used from two
            # streams at once)
            import torch.distributed as dist
            self.bn_group = dist.new_group(backend="nccl") if dist.get_backend() == "nccl" else None
            self.network = nn.SyncBatchNorm.convert_sync_batchnorm(self.network, process_group=self.bn_group)
        self._lower_raw, self._upper_raw, self._head_raw = _TrunkLower(self.network), _TrunkUpper(self.network), _Head(self.network)
        self._lower, self._upper, self._head = self._lower_raw, self._upper_raw, self._head_raw
        self.aux: dict[str, nn.Module] = {}
        self._gmodule = None
        if cfg.graph_matching:
            gm = GModule(in_channels=256, num_classes=nc, device=device).to(device)
            gm.cluster_backend = cfg.cluster_backend
            gm.defer_seed_update = cfg.phased_backward      # the phased step flushes it after the last launch
            self.aux["Graph"] = gm
            self._gmodule = gm
        if cfg.discriminator and cfg.graph_matching:
            for lvl in ("p2", "p3", "p4", "p5"):
                self.aux[f"Dis_{lvl.upper()}"] = _JointDiscriminator(
                    Discriminator(grad_reverse_lambda=0.02).to(device).to(memory_format=torch.channels_last))
        if cfg.vig_grapher:
            self.aux["Grapher"] = Grapher(256, 9, 1, "mr", "gelu", "batch", True, False, 0.0, 1,
                                          (cfg.hw // 4) ** 2, 0.0, False).to(device)
        if cfg.temporal_graph:
            self.aux["TGCN"] = TGCN(256, 256, (cfg.clip_frames, 8, 8), 10, 10).to(device)
        self.sinkhorn = SinkhornDistance(eps=0.1, max_iter=5, reduction="mean")     # train_cardiac_uda.py:138
        self.dice = DiceLoss()
        self.ce = nn.CrossEntropyLoss()
        modules = [self.network, *self.aux.values()]
        for m in modules:
            m.train()
        # gradient buckets in the order they become final in the phased step: (0) discriminators, graph module, FPN head;
        # (1) Grapher + upper trunk; (2) the rest (lower trunk, TGCN).  Each is all-reduced as soon as it is final.
        early = [p for n_, m in self.aux.items() if n_.startswith("Dis_") or n_ == "Graph" for p in m.parameters()]
        early += self.network.head_parameters()
        mid = [p for n_, m in self.aux.items() if n_ == "Grapher" for p in m.parameters()] + self.network.upper_trunk_parameters()
        self.grads = FlatGradSync(modules, bind=False, buckets=[early, mid])
        self._early_exchange = (world_size > 1 and cfg.phased_backward and cfg.graph_matching and not cfg.temporal_graph
                                and os.environ.get("GE_EARLY_EXCHANGE", "1") != "0")
        self.graphed = False
        # high priority: the graph module's hundreds of tiny kernels (and the two read-backs its host code waits on)
        # must not queue behind the waves of the discriminator / head kernels they overlap with
        self._side_stream = torch.cuda.Stream(device=device, priority=-1) if device.type == "cuda" else None
        self.graph_launches = 0      # graphecho_b200 kernels replayed per step inside the CUDA graphs
        self.opt = {"Net": torch.optim.Adam(self.network.parameters(), lr=cfg.lr_net, betas=(0.9, 0.999),
                                            weight_decay=cfg.weight_decay, fused=device.type == "cuda")}
        for name, m in self.aux.items():
            self.opt[name] = torch.optim.SGD(m.parameters(), lr=cfg.lr_aux, momentum=0.9,
                                             weight_decay=cfg.weight_decay, fused=device.type == "cuda")

    # ------------------------------------------------------------------------------------------
    def _split(self, ns):
        """BatchNorm segments for a [source | target] batch with ns source frames (GF.domain_split)."""
        return GF.domain_split(ns if self.cfg.per_domain_bn else 0)

    def _buffers(self):
        return [b for m in [self.network, *self.aux.values()] for b in m.buffers()]

    def capture_graphs(self, n_frames: int, n_source: int | None = None):
        """Capture the shape-static segments of the step -- the segmentation network, the p2 Grapher and
        the four discriminators, forward and backward -- as CUDA graphs (torch.cuda.make_graphed_callables),
        so that ~3 500 of the step's ~4 000 kernel launches are replayed by a dozen graph launches.  The
        data-dependent part (GModule: node sampling, matching) stays eager.  `n_frames` = source + target
        frames per step on this rank, `n_source` of them source frames (the segments are re-captured if they
        change).  Buffers (BatchNorm running statistics, seed banks) are snapshotted before the warm-up / capture
        executions on dummy inputs and restored afterwards: capturing leaves no trace in the model state."""
        cfg, dev = self.cfg, self.device
        if self.world > 1 and cfg.sync_bn:
            # measured: capturing the per-layer NCCL averages with make_graphed_callables hangs at replay on 2 ranks
            # (tests/test_sync_bn_gpu.py history); SyncBatchNorm runs the fused kernels eagerly
            raise RuntimeError("cuda_graphs with SyncBatchNorm is not supported: use sync_bn=False or cuda_graphs=False")
        ns = n_frames // 2 if n_source is None else int(n_source)
        bufs = self._buffers()
        snapshot = [b.detach().clone() for b in bufs]
        x = torch.zeros(n_frames, 1, cfg.hw, cfg.hw, device=dev)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=cfg.bf16), self._split(ns):
            _, feats = self.network(x)
        # make_graphed_callables shares one memory pool between the graphs and relies on them being replayed in
        # the order of this tuple (forward) and in its reverse (backward): trunk -> Grapher -> head -> discriminators
        # forward, discriminators -> head -> Grapher -> trunk backward -- the order every step below keeps.
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=cfg.bf16), self._split(ns):
            mids = self._lower_raw(x)
        calls, samples, names = [self._lower_raw, self._upper_raw], [(x,), tuple(torch.zeros_like(m).requires_grad_() for m in mids)], ["lower", "upper"]
        del mids
        if cfg.graph_matching and cfg.vig_grapher:
            calls.append(self.aux["Grapher"])
            samples.append((torch.zeros_like(feats[0]).requires_grad_(),))
            names.append("Grapher")
        calls.append(self._head_raw)
        samples.append(tuple(torch.zeros_like(f).requires_grad_() for f in feats))
        names.append("head")
        if cfg.graph_matching and cfg.discriminator:
            for i, lvl in enumerate(("P2", "P3", "P4", "P5")):
                f = feats[i]
                calls.append(self.aux[f"Dis_{lvl}"])
                samples.append((torch.zeros_like(f).requires_grad_(),))
                names.append(f"Dis_{lvl}")
        del feats
        from . import _cabi
        before = _cabi.launch_count()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=cfg.bf16, cache_enabled=False), self._split(ns):
            graphed = torch.cuda.make_graphed_callables(tuple(calls), tuple(samples), num_warmup_iters=3,
                                                        allow_unused_input=True)
        # 3 warm-up executions + 1 capture of every segment's forward and backward
        self.graph_launches = (_cabi.launch_count() - before) // 4
        for name, g in zip(names, graphed):
            if name == "lower":
                self._lower = g
            elif name == "upper":
                self._upper = g
            elif name == "head":
                self._head = g
            else:
                self.aux[name] = g
        with torch.no_grad():
            for b, keep in zip(bufs, snapshot):
                b.copy_(keep)
        self.grads.zero()
        self.graphed = True
        self._graph_frames = (n_frames, ns)

    def _autocast(self):
        return torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.cfg.bf16, cache_enabled=not self.graphed)

    def seg_loss(self, logits, masks):
        """DiceLoss + BCEWithLogitsLoss (train_cardiac_uda.py:228) as one fused pass (csrc/seg_loss.cu)."""
        if logits.is_cuda and logits.shape[1] <= 8:
            return GF.seg_loss(logits, masks)
        return self.dice(logits, masks) + F.binary_cross_entropy_with_logits(logits, masks)

    def _prepare(self, frames_src, frames_tgt):
        ns, n = frames_src.shape[0], frames_src.shape[0] + frames_tgt.shape[0]
        for m in self.aux.values():
            if isinstance(m, _JointDiscriminator):
                m.n_source = ns
        if self.cfg.cuda_graphs and not self.graphed:
            self.capture_graphs(n, ns)
        if self.graphed and self._graph_frames != (n, ns):
            raise RuntimeError(f"the batch changed after CUDA-graph capture: {(n, ns)} vs {self._graph_frames}")
        return ns

    def _node_transport(self, nodes, losses):
        """config 3: entropic optimal-transport distance between the two matched node sets (the reference's
        SinkhornDistance(0.1, 5, 'mean'), train_cardiac_uda.py:138, applied as TGCN.py:281-283 applies it)."""
        if self.cfg.sinkhorn_nodes and nodes[0].dim() == 2 and nodes[0].size(0) >= 6 and nodes[1].size(0) > 0:
            losses["sinkhorn_loss"] = self.cfg.sinkhorn_weight * self.sinkhorn(nodes[0], nodes[1])[0]

    def _temporal_losses(self, temporal):
        """The temporal branch of the reference step (train_cardiac_uda.py:258-311): the [source | target] clips go
        through the network in ONE call (one BatchNorm segment), the graph module runs on their pyramid with the RAW
        target logits as score maps (Appendix A-12), and TGCN consumes the clip pyramids and the (detached) nodes.
        `temp_seg_loss` (:283-290) never reaches `losses` in the reference, so the head runs without autograd.
        Returns the scalar temporal_graph_loss (or None)."""
        cfg = self.cfg
        frames_temp, masks_temp, (b, t) = temporal
        nst = masks_temp.shape[0]
        # the network's own methods, not the _Trunk/_Head wrappers: make_graphed_callables swaps THEIR forward for the
        # graph replay of the single-frame batch
        with self._autocast():
            feats = list(self.network.forward_trunk(frames_temp))
            with torch.no_grad():
                logits = self.network.forward_head(*[f.detach() for f in feats])
        pred_src = logits[:nst]
        avail = (masks_temp.sum(dim=(1, 2, 3)) > 100).view(-1, 1, 1, 1)                 # :277, 283-290
        targets = torch.where(avail, masks_temp, pred_src.to(masks_temp.dtype))
        _, nodes, mid = self._gmodule.forward_joint(feats, nst, targets, logits[nst:])
        self._gmodule.flush_seed_update()
        total = sum(mid.values()) if mid else None
        if nodes[0].numel() > 0 and nodes[0].dim() == 2:
            graph_features = [f.reshape(b, t, *f.shape[1:]) for f in feats]             # :300-302
            tl = self.aux["TGCN"](graph_features, (nodes[0].detach(), nodes[1].detach()), self.sinkhorn, self.ce,
                                  (None, None), r=[8, 4, 2, 1])
            tsum = sum(tl.values())
            total = tsum if total is None else total + tsum
        return total

    def forward_losses(self, frames_src, masks_src, frames_tgt, temporal=None):
        """frames_* [F,1,H,W] (fp32, device), masks_src [F,nc,H,W] one-hot.  Returns the loss dict (one autograd
        graph; the reference order of train_cardiac_uda.py:223-311)."""
        cfg = self.cfg
        ns = self._prepare(frames_src, frames_tgt)
        losses = {}
        with self._autocast(), self._split(ns):
            feats = list(self._upper(*self._lower(torch.cat([frames_src, frames_tgt], dim=0))))
            p2g = self.aux["Grapher"](feats[0]) if cfg.graph_matching and cfg.vig_grapher else None
        with self._autocast():
            logits = self._head(*feats)
        pred_s, pred_t = logits[:ns], logits[ns:]
        losses["seg_loss"] = cfg.seg_weight * self.seg_loss(pred_s, masks_src)
        if not cfg.graph_matching:
            return losses
        if p2g is not None:
            feats = [p2g] + list(feats[1:])
        score_maps = GF.LogitMap(pred_t.detach())                                  # train_cardiac_uda.py:235, never materialised
        # The graph-matching module is host-driven (two count read-backs, hundreds of tiny launches) while
        # the discriminators are four long GPU-bound graph replays with no dependency on it: run GModule on
        # a side stream so its host synchronisations wait only for its own small kernels and its launch
        # latency hides under the discriminator towers.  Autograd replays each backward on its forward
        # stream, so the same overlap happens in the backward pass.
        main = torch.cuda.current_stream()
        side = self._side_stream if cfg.overlap_streams else None
        if side is not None:
            side.wait_stream(main)
        with torch.cuda.stream(side) if side is not None else _NullCtx():
            # source frames first, target frames after: the joint entry points avoid slicing the pyramid
            _, nodes, mid = self._gmodule.forward_joint(feats, ns, masks_src, score_maps)
            self._node_transport(nodes, mid)
        if cfg.discriminator:
            with self._autocast():
                for i, lvl in enumerate(("p2", "p3", "p4", "p5")):
                    losses[f"loss_adv_{lvl}"] = 0.1 * self.aux[f"Dis_{lvl.upper()}"](feats[i])
        if side is not None:
            main.wait_stream(side)
        losses.update(mid)
        if cfg.temporal_graph and temporal is not None:
            self._gmodule.flush_seed_update()
            tloss = self._temporal_losses(temporal)
            if tloss is not None:
                losses["temporal_graph_loss"] = tloss
        return losses

    def train_step(self, frames_src, masks_src, frames_tgt, temporal=None):
        self.grads.zero()
        if self.cfg.phased_backward and self.cfg.graph_matching:
            losses = self._phased_forward_backward(frames_src, masks_src, frames_tgt, temporal)
            total = sum(v.detach() for v in losses.values())
        else:
            losses = self.forward_losses(frames_src, masks_src, frames_tgt, temporal)
            total = sum(losses.values())
            total.backward()
            total = total.detach()
        self.grads.rebind()
        self.grads.all_reduce()
        for opt in self.opt.values():
            opt.step()
        return total, {k: v.detach() for k, v in losses.items()}

    def _phased_forward_backward(self, frames_src, masks_src, frames_tgt, temporal=None):
        """Same losses and gradients as forward_losses() + one backward(), issued in an order that keeps the GPU
        busy.  GModule is host-driven (two count read-backs, ~900 tiny launches forward, as many backward);
        in one autograd graph its launches sit between the pyramid and the discriminators on the CPU timeline
        and the GPU idles for the length of them.  Here autograd is cut at the pyramid (detached leaves):
          1. FPN trunk, FPN head (on detached pyramid leaves), Grapher forward      main stream, graph replays
          2. discriminators forward AND backward, then the head's backward          main stream, graph replays
          3. GModule forward and backward                                           side stream, while 2 runs
          4. pyramid gradients = head + discriminator + GModule parts; trunk backward   main stream
          5. [temporal_graph] the clip branch: trunk on the clips, GModule + TGCN, their backward
        """
        cfg = self.cfg
        ns = self._prepare(frames_src, frames_tgt)
        losses = {}
        main = torch.cuda.current_stream()
        side = self._side_stream if cfg.overlap_streams else None
        if side is not None:
            side.wait_stream(main)          # inputs (e.g. the step's host->device copies) are ready for the side stream
        with self._autocast(), self._split(ns):
            lower = self._lower(torch.cat([frames_src, frames_tgt], dim=0))
            leaves_u = [t.detach().requires_grad_() for t in lower]
            feats = list(self._upper(*leaves_u))
            tops = list(feats)
            if cfg.vig_grapher:
                tops[0] = self.aux["Grapher"](feats[0])
        with self._autocast():
            leaves_h = [f.detach().requires_grad_() for f in feats]
            logits = self._head(*leaves_h)
        # while the GPU runs the forward graphs: the source half of the sampler plan (needs the masks and the map
        # sizes only) and its count read-back, on the otherwise idle side stream
        with torch.cuda.stream(side) if side is not None else _NullCtx():
            self._gmodule.prepare_source(masks_src, [f.shape[-2:] for f in feats], masks_src.device)
        pred_s, pred_t = logits[:ns], logits[ns:]
        seg = cfg.seg_weight * self.seg_loss(pred_s, masks_src)
        losses["seg_loss"] = seg
        score_maps = GF.LogitMap(pred_t.detach())                                  # train_cardiac_uda.py:235, never materialised
        if side is not None:
            side.wait_stream(main)          # the side stream waits for the pyramid only, NOT for the work issued below
        # 2. discriminators on their own leaves, forward + backward back to back; then the head's backward
        grads = [None] * len(tops)
        if cfg.discriminator:
            leaves_d = [t.detach().requires_grad_() for t in tops]
            with self._autocast():
                for i, lvl in enumerate(("p2", "p3", "p4", "p5")):
                    losses[f"loss_adv_{lvl}"] = 0.1 * self.aux[f"Dis_{lvl.upper()}"](leaves_d[i])
            for lvl in ("p5", "p4", "p3", "p2"):                                    # reverse of the forward order
                losses[f"loss_adv_{lvl}"].backward()
            grads = [l.grad for l in leaves_d]
        seg.backward()                                                              # head only: stops at leaves_h
        # 3. graph matching on the side stream
        leaves_g = [t.detach().requires_grad_() for t in tops]
        with torch.cuda.stream(side) if side is not None else _NullCtx():
            _, nodes, mid = self._gmodule.forward_joint(leaves_g, ns, masks_src, score_maps)
            self._node_transport(nodes, mid)
            losses.update(mid)
            if mid:
                torch.autograd.backward(list(mid.values()))
        if side is not None:
            main.wait_stream(side)
        if self._early_exchange:
            self.grads.reduce_bucket(0)         # discriminators, graph module, head: final; on the wire under the trunk backward
        # 4. join the pyramid gradients and run the trunk (and Grapher) backward
        out_t, out_g = [], []
        for i, f in enumerate(feats):
            gd, gg = grads[i], leaves_g[i].grad
            if gg is not None and side is not None:
                gg.record_stream(main)          # allocated on the side stream, consumed on the main stream
            g_top = gg if gd is None else (gd if gg is None else gd + gg)          # dL/d tops[i]: discriminator + GModule
            g_head = leaves_h[i].grad                                              # dL/d feats[i]: segmentation head
            if tops[i] is f:
                out_t.append(f)
                out_g.append(g_head if g_top is None else g_head + g_top)
            else:                                                                  # level went through the Grapher
                out_t.append(f)
                out_g.append(g_head)
                if g_top is not None:
                    out_t.append(tops[i])
                    out_g.append(g_top)
        with self._split(ns):
            torch.autograd.backward(out_t, out_g)                     # Grapher + upper trunk: stops at leaves_u
            if self._early_exchange:
                self.grads.reduce_bucket(1)                           # 2/3 of the backbone: on the wire under the lower backward
            torch.autograd.backward(list(lower), [l.grad for l in leaves_u])
        # everything is issued: the seed-bank update (next step's input) runs while the GPU finishes the backward
        with torch.cuda.stream(side) if side is not None else _NullCtx():
            self._gmodule.flush_seed_update()
        if side is not None:
            main.wait_stream(side)
        # 5. the clip branch (its own autograd graph: trunk parameters accumulate a second gradient)
        if cfg.temporal_graph and temporal is not None:
            tloss = self._temporal_losses(temporal)
            if tloss is not None:
                losses["temporal_graph_loss"] = tloss
                tloss.backward()
        return losses

    @torch.no_grad()
    def predict(self, frames):
        self.network.eval()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.cfg.bf16):
            logits, _ = self.network(frames)
        self.network.train()
        return logits


def make_batch(cfg: EngineConfig, n_clips: int, frames: int, rank: int = 0, world: int = 1, seed: int = 0,
               pin: bool = False):
    """Synthetic step input on the HOST in the trainers' layout: clips [b,1,H,W,t] (first half of the
    clips = source stream, second half = target) + one-hot source masks [b/2 * t, nc, H, W].
    Weak scaling: every rank draws its own `n_clips` clips (different seed per rank)."""
    x = synth.clips(n_clips, cfg.hw, frames, seed=seed + 1000 * rank)
    ns = n_clips // 2
    masks = synth.disc_masks(ns * frames, cfg.num_classes, cfg.hw)
    if pin and torch.cuda.is_available():
        x, masks = x.pin_memory(), masks.pin_memory()
    return x, masks


def make_frame_batch(cfg: EngineConfig, n_src: int, n_tgt: int, rank: int = 0, seed: int = 0, pin: bool = False):
    """Single-frame streams of the CardiacUDA / CAMUS trainers (train_cardiac_uda.py:189-192: source batch
    2 x batch_size, target batch batch_size): (source frames [n_src,1,H,W], one-hot masks, target frames)."""
    xs = synth.images(n_src, cfg.hw, seed=seed + 1000 * rank + 17)
    xt = synth.images(n_tgt, cfg.hw, seed=seed + 1000 * rank + 29)
    masks = synth.disc_masks(n_src, cfg.num_classes, cfg.hw)
    if pin and torch.cuda.is_available():
        xs, xt, masks = xs.pin_memory(), xt.pin_memory(), masks.pin_memory()
    return xs, masks, xt


def split_streams(clips_dev: torch.Tensor):
    """[b,1,H,W,t] on device -> (source frames, target frames, (b, t)); clip-major frame order."""
    b, c, h, w, t = clips_dev.shape
    frames = synth.flatten_clips(clips_dev)
    ns = (b // 2) * t
    return frames[:ns], frames[ns:], (b, t)


def temporal_input(clips_dev: torch.Tensor, masks_dev: torch.Tensor):
    """The `temporal` argument of train_step from device clips [b,1,H,W,t] (source clips first) and the source
    clips' masks [b/2*t, nc, H, W]: (frames [b*t,1,H,W], masks, (b, t)) (train_cardiac_uda.py:272-276)."""
    b, c, h, w, t = clips_dev.shape
    return synth.flatten_clips(clips_dev), masks_dev, (b, t)

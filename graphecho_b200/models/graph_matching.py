"""Mirror of models/graph_matching.py: GModule (graph matching for UDA), PrototypeComputation (node
sampler) and BCEFocalLoss, with the reference's signatures, loss-dict keys and state_dict keys.

Hot operators: node_affinity (Affinity -> ge_affinity_pairwise) and the
InstanceNorm -> sinkhorn_rpm(20) -> exp chain of _forward_aff (-> ge_sinkhorn_rpm_fwd/bwd, one
cluster launch instead of ~250 kernels).  The node sampler reproduces the reference's semantics
(location strides 8..128 on a stride-4..32 pyramid, bbox labels, floor(linspace) negative picks,
graph_matching.py:609-635, 861-1013) but vectorised over the batch and entirely on the device.
Everything in this module computes in fp32 regardless of autocast."""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from .. import functional as GF
from .affinity_layer import Affinity
from .gradient_reversal import GradientReversal
from .transformer import MultiHeadAttention

INF = 100000000


class _null_ctx:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


def _nonzero_known(mask, count):
    """Indices of the True entries of a 1-D mask whose count is already known on the host: no
    device->host sync (torch.nonzero has to read the count back)."""
    if hasattr(torch, "nonzero_static"):
        return torch.nonzero_static(mask, size=int(count)).squeeze(1)
    return torch.nonzero(mask).squeeze(1)


class BCEFocalLoss(nn.Module):
    """Focal BCE on probabilities (graph_matching.py:23-45)."""

    def __init__(self, gamma=2, alpha=0.25, reduction="elementwise_mean"):
        super().__init__()
        self.gamma, self.alpha, self.reduction = gamma, alpha, reduction

    def forward(self, _input, target):
        pt, a, g = _input, self.alpha, self.gamma
        loss = -a * (1 - pt) ** g * target * torch.log(pt) - (1 - a) * pt ** g * (1 - target) * torch.log(1 - pt)
        if self.reduction == "elementwise_mean":
            return loss.mean()
        if self.reduction == "sum":
            return loss.sum()
        return loss


def _ln_mlp(widths, final_ln):
    layers = []
    for i in range(len(widths) - 1):
        layers.append(nn.Linear(widths[i], widths[i + 1]))
        last = i == len(widths) - 2
        if not last or final_ln:
            layers.append(nn.LayerNorm(widths[i + 1], elementwise_affine=False))
        if not last:
            layers.append(nn.ReLU())
    return nn.Sequential(*layers)


class GModule(nn.Module):
    """GModule(in_channels, num_classes, device)
    forward(images, features, targets=None, score_maps=None)
        train: -> (features, (nodes_1, nodes_2), loss_dict)      eval: -> (features, None)"""

    def __init__(self, in_channels, num_classes, device):
        super().__init__()
        self.device = device
        self.fpn_strides = [8, 16, 32, 64, 128]
        self.num_classes = num_classes
        self.matching_loss_type = "FL"
        self.matching_cfg = "o2o"
        self.with_cluster_update = True
        self.with_semantic_completion = True
        self.with_quadratic_matching = True
        self.weight_matching, self.weight_nodes, self.weight_dis, self.lambda_dis = 0.1, 1.0, 0.1, 0.02
        self.with_domain_interaction = True
        self.with_complete_graph = True
        self.with_node_dis = True
        self.with_global_graph = False
        self.node_dis_place = "feat"
        self.with_cond_cls = False
        self.with_score_weight = False
        # 'sklearn' = the reference's CPU SpectralClustering round trip; 'device' = on-GPU bipartition
        self.cluster_backend = "sklearn"
        self.async_seed_update = True
        self._seed_stream = None

        self.graph_generator = PrototypeComputation(num_classes)
        self.head_in_cfg = "LN"
        self.head_in_ln = _ln_mlp([256, 256, 256], final_ln=True)                 # keys 0, 3
        self.node_cls_middle = nn.Sequential(nn.Linear(256, 512), nn.ReLU(), nn.Linear(512, num_classes))
        self.seed_project_left = nn.Linear(256, 256)
        self.register_buffer("sr_seed", torch.randn(num_classes, 256))
        self.register_buffer("tg_seed", torch.randn(num_classes, 256))
        self.cross_domain_graph = MultiHeadAttention(256, 1, dropout=0.1, version="v2")
        self.intra_domain_graph = MultiHeadAttention(256, 1, dropout=0.1, version="v2")
        self.node_affinity = Affinity(d=256)
        self.InstNorm_layer = nn.InstanceNorm2d(1)
        self.matching_loss = BCEFocalLoss()
        self.quadratic_loss = nn.L1Loss(reduction="mean")
        self.grad_reverse = GradientReversal(self.lambda_dis)
        self.node_dis_2 = _ln_mlp([256, 256, 256, 256, 1], final_ln=False)         # keys 0, 3, 6, 9
        self.loss_fn = nn.BCEWithLogitsLoss()
        self._init_weight()
        # the seed banks are updated on a private stream (update_seed): anything that reads them through
        # state_dict() (checkpoints, buffer broadcasts) first joins that stream
        self.register_state_dict_pre_hook(lambda module, prefix, keep_vars: module._sync_seed_stream())

    def _init_weight(self):
        seqs = [self.node_dis_2, self.node_cls_middle, self.head_in_ln, [self.seed_project_left]]
        for seq in seqs:
            for m in seq:
                if isinstance(m, nn.Linear):
                    nn.init.normal_(m.weight, std=0.01)
                    nn.init.constant_(m.bias, 0)

    # ------------------------------------------------------------------ entry points
    def forward(self, images, features, targets=None, score_maps=None):
        if targets is not None:
            return self._forward_train(images, features, targets, score_maps)
        return self._forward_inference(images, features), None

    def _forward_inference(self, images, features):
        return features

    def _node_dis_loss(self, nodes_1, nodes_2):
        rev = self.grad_reverse(torch.cat([nodes_1, nodes_2], dim=0))
        tgt = torch.cat([nodes_1.new_ones(nodes_1.size(0)), nodes_2.new_zeros(nodes_2.size(0))])
        return self.weight_dis * self.loss_fn(self.node_dis_2(rev).view(-1), tgt)

    def _forward_train(self, images, features, targets=None, score_maps=None):
        with torch.autocast("cuda", enabled=False):
            features_s, features_t = features
            return self._train_fp32(features, (features_s, 0), (features_t, 0), targets, score_maps)

    def forward_joint(self, features_all, n_source, targets, score_maps):
        """Same as forward(images, (features_all[:n_source], features_all[n_source:]), targets, score_maps)
        for features that hold the source frames first and the target frames after them, WITHOUT slicing
        the feature maps: nodes are gathered from the full tensors with a batch offset, so autograd never
        materialises zero-padded slice gradients of the pyramid."""
        with torch.autocast("cuda", enabled=False):
            return self._train_fp32(None, (features_all, 0), (features_all, int(n_source)), targets, score_maps)

    def prepare_source(self, targets, feature_shapes, device):
        """The source half of the sampler plan depends on the ground-truth masks and the map SIZES only, not on the
        network: a caller that knows the sizes can issue it (and its count read-back) while the network forward is
        still running.  The next training forward consumes the result (train_cardiac_uda.py:247-256 calls the
        graph module with the same masks)."""
        shapes = [torch.empty((0, 0, int(h), int(w)), device=device) for h, w in feature_shapes]
        plan = self.graph_generator.plan(self.compute_locations(shapes), self.find_bbox(targets),
                                         [(int(h), int(w)) for h, w in feature_shapes], self.fpn_strides)
        self._prepared_source = (targets, plan[0], plan[1].tolist())

    def flush_seed_update(self):
        """Run a seed-bank update that `defer_seed_update=True` postponed (the banks are only read by the next
        step's hallucination, so the update -- host-driven: clustering launches, ~200 small ops -- can wait
        until the rest of the step has been issued)."""
        pending, self._pending_seed = getattr(self, "_pending_seed", None), None
        if pending is not None:
            self._class_layout = pending[4]
            self.update_seed(*pending[:4])

    def _train_fp32(self, features, src, tgt, targets, score_maps):
        losses = {}
        self.flush_seed_update()
        self._sync_seed_stream()
        (feat_s, off_s), (feat_t, off_t) = src, tgt
        gen = self.graph_generator
        prepared, self._prepared_source = getattr(self, "_prepared_source", None), None
        geo = [(int(f.size(-2)), int(f.size(-1))) for f in feat_t]
        plan_t = gen.plan(self.compute_locations(feat_t), self.find_bbox(score_maps), geo, self.fpn_strides)
        if prepared is not None and prepared[0] is targets:
            labels_s, counts_s = prepared[1], prepared[2]
            counts_t = plan_t[1].tolist()                                           # host sync: target counts only
        else:
            plan_s = gen.plan(self.compute_locations(feat_s), self.find_bbox(targets), geo, self.fpn_strides)
            labels_s = plan_s[0]
            counts_s, counts_t = torch.stack([plan_s[1], plan_t[1]]).tolist()       # ONE host sync for both domains
        (nodes_1, labels_1, weights_1), (nodes_2, labels_2, weights_2) = gen.gather_pair(
            (feat_s, labels_s, counts_s, off_s), (feat_t, plan_t[0], counts_t, off_t))
        if nodes_1.size(0) < 6 or nodes_1.dim() == 1:                             # graph_matching.py:259-260
            return features, (nodes_1, nodes_2), losses
        nodes_1, nodes_2 = nodes_1.float(), nodes_2.float()
        if self.with_node_dis and self.node_dis_place == "feat":
            losses["dis_loss"] = self._node_dis_loss(nodes_1, nodes_2)
        # head_in_ln is row-wise (Linear / LayerNorm / ReLU): one pass over both domains' rows, then split
        n_1 = nodes_1.size(0)
        both = self.head_in_ln(torch.cat([nodes_1, nodes_2], dim=0))
        nodes_1, nodes_2 = both[:n_1], both[n_1:]
        (nodes_1, nodes_2), (labels_1, labels_2), (weights_1, weights_2) = \
            self._forward_preprocessing_source_target((nodes_1, nodes_2), (labels_1, labels_2), (weights_1, weights_2))
        if self.with_complete_graph:
            (nodes_1, edges_1), (nodes_2, edges_2) = self.intra_domain_graph.forward_pair(nodes_1, nodes_1, nodes_2, nodes_2)
        if getattr(self, "defer_seed_update", False):
            self._pending_seed = (nodes_1.detach(), labels_1, nodes_2.detach(), labels_2, getattr(self, "_class_layout", None))
            self._class_layout = None
        else:
            self.update_seed(nodes_1, labels_1, nodes_2, labels_2)
        if self.with_node_dis and self.node_dis_place == "intra":
            losses["dis_loss"] = self._node_dis_loss(nodes_1, nodes_2)
        if self.with_domain_interaction:
            nodes_1, nodes_2 = self._forward_cross_domain_graph(nodes_1, nodes_2)
        if self.with_node_dis and self.node_dis_place == "inter":
            losses["dis_loss"] = self._node_dis_loss(nodes_1, nodes_2)
        node_loss = self._forward_node_loss(torch.cat([nodes_1, nodes_2]), torch.cat([labels_1, labels_2]),
                                            torch.cat([weights_1, weights_2]))
        losses["node_loss"] = self.weight_nodes * node_loss
        if self.matching_cfg != "none":
            aff_loss, affinity = self._forward_aff(nodes_1, nodes_2, labels_1, labels_2)
            losses["mat_loss_aff"] = self.weight_matching * aff_loss
            if self.with_quadratic_matching:
                losses["mat_loss_qu"] = self._forward_qu(edges_1.detach(), edges_2.detach(), affinity)
        return features, (nodes_1, nodes_2), losses

    # ------------------------------------------------------------------ class regrouping / hallucination
    def _hallucinate(self, seed_row, other):
        n = other.size(0)
        base = seed_row.unsqueeze(0).expand(n, 256)
        if not self.with_semantic_completion:
            out = torch.randn_like(other) * 0.01
        elif n < 5:
            out = torch.randn_like(other) * 0.01 + base
        else:
            out = torch.normal(mean=base, std=other.std(0).unsqueeze(0).expand(n, 256))
        return self.seed_project_left(out)

    def _forward_preprocessing_source_target(self, nodes, labels, weights):
        """Class-major regrouping; a class present in one domain only is completed with nodes
        hallucinated from the other domain's seed bank (graph_matching.py:381-483)."""
        (sn, tn), (sl, tl), (sw, tw) = nodes, labels, weights
        dev = sn.device
        nbin = max(self.num_classes, 1)
        cnt = torch.stack([torch.bincount(sl.long(), minlength=nbin)[:nbin],
                           torch.bincount(tl.long(), minlength=nbin)[:nbin]]).tolist()          # one host sync
        s_cnt, t_cnt = cnt
        present = [c for c in range(nbin) if s_cnt[c] > 0 or t_cnt[c] > 0]
        if all(s_cnt[c] > 0 and t_cnt[c] > 0 for c in present):
            # every class lives in both domains (the usual case): class-major regrouping is a stable sort by label
            self._class_layout = [(int(c), s_cnt[c], t_cnt[c]) for c in present]
            so, to = torch.sort(sl, stable=True)[1], torch.sort(tl, stable=True)[1]
            return ((sn.index_select(0, so), tn.index_select(0, to)),
                    (sl.index_select(0, so).float(), tl.index_select(0, to).float()),
                    (sw.index_select(0, so), tw.index_select(0, to)))
        S, T, SL, TL, SW, TW = [], [], [], [], [], []
        for c in present:
            ci = int(c)
            has_s, has_t = s_cnt[ci] > 0, t_cnt[ci] > 0
            si = _nonzero_known(sl == c, s_cnt[ci]) if has_s else None
            ti = _nonzero_known(tl == c, t_cnt[ci]) if has_t else None
            s_c = sn[si] if has_s else None
            t_c = tn[ti] if has_t else None
            if has_s and has_t:
                S.append(s_c); T.append(t_c)
                SW.append(sw[si]); TW.append(tw[ti])
            elif has_t:
                S.append(self._hallucinate(self.sr_seed[ci], t_c)); T.append(t_c)
                SW.append(torch.ones(len(t_c), dtype=torch.long, device=dev)); TW.append(tw[ti])
            elif has_s:
                S.append(s_c); T.append(self._hallucinate(self.tg_seed[ci], s_c))
                SW.append(sw[si]); TW.append(torch.ones(len(s_c), dtype=torch.long, device=dev))
            else:
                continue
            SL.append(torch.full((len(S[-1]),), float(c), device=dev))
            TL.append(torch.full((len(T[-1]),), float(c), device=dev))
        # class-major layout of the regrouped nodes, known on the host: update_seed needs no further sync
        kept = [int(c) for c in present]
        self._class_layout = [(c, len(s_), len(t_)) for c, s_, t_ in zip(kept, S, T)]
        return (torch.cat(S), torch.cat(T)), (torch.cat(SL), torch.cat(TL)), (torch.cat(SW), torch.cat(TW))

    def _forward_preprocessing_source(self, sr_nodes, sr_nodes_label):
        """Source-only split (graph_matching.py:354-379; unreachable with the trainers' calls)."""
        n1, n2, l1, l2 = [], [], [], []
        for c in sr_nodes_label.unique():
            cur = sr_nodes[sr_nodes_label == c]
            n1.append(cur[::2]); n2.append(cur[1::2])
            l1.append(cur.new_ones(len(n1[-1])) * c); l2.append(cur.new_ones(len(n2[-1])) * c)
        return (torch.cat(n1), torch.cat(n2)), (torch.cat(l1), torch.cat(l2))

    # ------------------------------------------------------------------ graph layers and losses
    def _forward_intra_domain_graph(self, nodes):
        return self.intra_domain_graph(nodes, nodes, nodes)

    def _forward_cross_domain_graph(self, nodes_1, nodes_2):
        if self.with_global_graph:
            n_1 = len(nodes_1)
            g = torch.cat([nodes_1, nodes_2], dim=0)
            g = self.cross_domain_graph(g, g, g)[0]
            return g[:n_1], g[n_1:]
        # nodes2_enhanced = cross(nodes_1, nodes_1, nodes_2)[0], nodes1_enhanced = cross(nodes_2, nodes_2, nodes_1)[0]
        (nodes2_enhanced, _), (nodes1_enhanced, _) = self.cross_domain_graph.forward_pair(nodes_1, nodes_2, nodes_2, nodes_1)
        return nodes1_enhanced, nodes2_enhanced

    def _forward_node_loss(self, nodes, labels, weights=None):
        labels = labels.long()
        assert len(nodes) == len(labels)
        logits = self.node_cls_middle(nodes)
        if weights is None:
            return F.cross_entropy(logits, labels, reduction="mean")
        loss = F.cross_entropy(logits, labels, reduction="none")
        return (loss * weights).float().mean() if self.with_score_weight else loss.float().mean()

    def _forward_aff(self, nodes_1, nodes_2, labels_side1, labels_side2):
        """graph_matching.py:569-599.  'o2o': affinity -> fused instance-norm + Sinkhorn(20) + exp ->
        true-positive (row-wise best same-class entry) and false-positive focal losses."""
        M = self.node_affinity(nodes_1, nodes_2)
        if self.matching_cfg == "o2o":
            M = GF.sinkhorn_rpm_exp(M, 20, True)
            # true positives = the row-wise best same-class entry, false positives = every different-class entry
            # (graph_matching.py:577-588); one fused kernel each way (GF.matching_loss_o2o).  The different-class
            # entries are selected by mask, not gathered, and the log is only taken there: a same-class entry that
            # saturates to 1.0 cannot produce -inf * 0 = NaN.
            loss = GF.matching_loss_o2o(M, labels_side1, labels_side2, self.matching_loss.alpha, self.matching_loss.gamma)
            return loss, M
        same = labels_side1.long().unsqueeze(1) == labels_side2.long().unsqueeze(0)     # one_hot @ one_hot^T == 1
        if self.matching_cfg == "m2m":
            return self.matching_loss(M.sigmoid(), same.float()).mean(), M
        return 0, None

    def _forward_qu(self, edge_1, edge_2, affinity):
        R = torch.mm(edge_1, affinity) - torch.mm(affinity, edge_2)
        return self.quadratic_loss(R, torch.zeros_like(R))

    def sinkhorn_rpm(self, log_alpha, n_iters=5, slack=True, eps=-1):
        """log-domain Sinkhorn with slack row/column (graph_matching.py:637-689), log_alpha [B,J,K].
        The fused kernel returns exp(result); this API-compatible wrapper returns the log."""
        if not slack or eps > 0:
            raise NotImplementedError("only slack=True, eps<0 is on the accelerated path (the reference's only use)")
        return torch.log(GF.sinkhorn_rpm_exp(log_alpha, n_iters, False))

    def one_hot(self, x):
        return torch.eye(self.num_classes, device=x.device)[x.long(), :]

    def dynamic_fc(self, features, kernel_par):
        return F.linear(features, kernel_par, bias=None)

    def dynamic_conv(self, features, kernel_par):
        return F.conv2d(features, kernel_par.view(self.num_classes, -1, 1, 1))

    # ------------------------------------------------------------------ seed bank
    @torch.no_grad()
    def update_seed(self, sr_nodes, sr_labels, tg_nodes=None, tg_labels=None):
        """graph_matching.py:532-567: per class, mean of the (spectrally filtered) nodes blended into
        the seed bank with cosine-similarity momentum.  The banks are only read by the NEXT step's
        hallucination, so on CUDA the update runs on a side stream, overlapped with the rest of the step
        (`_sync_seed_stream` joins it before the banks are read again)."""
        side = None
        if sr_nodes.is_cuda and self.cluster_backend == "device" and self.async_seed_update:
            if self._seed_stream is None:
                self._seed_stream = torch.cuda.Stream(device=sr_nodes.device)
            side = self._seed_stream
            side.wait_stream(torch.cuda.current_stream())
            for t in (sr_nodes, sr_labels, tg_nodes, tg_labels):
                if t is not None:
                    t.record_stream(side)
        layout = getattr(self, "_class_layout", None)
        if layout is not None and (sum(l[1] for l in layout) != sr_nodes.size(0) or
                                   (tg_nodes is not None and sum(l[2] for l in layout) != tg_nodes.size(0))):
            layout = None
        with torch.cuda.stream(side) if side is not None else _null_ctx():
            self._update_bank(sr_nodes, sr_labels, self.sr_seed, None if layout is None else [(c, a) for c, a, _ in layout])
            if tg_nodes is not None:
                self._update_bank(tg_nodes, tg_labels, self.tg_seed,
                                  None if layout is None else [(c, b) for c, _, b in layout])
        self._class_layout = None

    def _sync_seed_stream(self):
        if self._seed_stream is not None:
            torch.cuda.current_stream().wait_stream(self._seed_stream)

    def _update_bank(self, nodes, labels, bank, layout=None, k=20):
        nodes = nodes.detach()
        if layout is None:
            layout = [(cls, None) for cls in labels.unique().long().tolist()]
        # the classes are independent (each writes its own bank row) and the device bipartition is one latency-bound
        # CTA per class (~1 ms): fan the classes out over a few streams so they run side by side
        outer = torch.cuda.current_stream() if nodes.is_cuda else None
        fan = outer is not None and self.cluster_backend == "device" and len(layout) > 1
        if fan and len(getattr(self, "_bank_streams", [])) < len(layout):
            self._bank_streams = [torch.cuda.Stream(device=nodes.device) for _ in range(len(layout))]
        off, used = 0, []
        for n_cls, (cls, count) in enumerate(layout):
            if count is None:
                bs = nodes[labels == cls]
            else:
                bs = nodes[off:off + count]       # regrouped class-major: a contiguous slice, no sync
                off += count
            st = None
            if fan and len(bs) > k and self.with_cluster_update:
                st = self._bank_streams[n_cls]
                st.wait_stream(outer)
                used.append(st)
            with torch.cuda.stream(st) if st is not None else _null_ctx():
                if len(bs) > k and self.with_cluster_update:
                    keep = self._bipartition(torch.cat([bank[cls][None, :], bs]))
                    # mean of the kept rows without a boolean gather (bs[keep] would synchronise the host with the
                    # clustering kernel).  If the seed ends up alone in its cluster (possible while the bank is still
                    # far from the features) the reference takes the mean of nothing = NaN and poisons the bank for the
                    # rest of training; here such a step falls back to the plain class mean instead.
                    w = keep.to(bs.dtype)
                    cnt = w.sum()
                    bs = torch.where(cnt > 0, (bs * w[:, None]).sum(0) / cnt.clamp_min(1.0), bs.mean(0))
                else:
                    bs = bs.mean(0)
                mom = F.cosine_similarity(bs.unsqueeze(0), bank[cls].unsqueeze(0))
                bank[cls] = bank[cls] * mom + bs * (1.0 - mom)
        for st in used:
            outer.wait_stream(st)

    def _bipartition(self, pts):
        """Boolean mask over pts[1:]: the points that fall in the same spectral cluster as pts[0]."""
        if self.cluster_backend == "sklearn":
            import sklearn.cluster as cluster
            sp = cluster.SpectralClustering(2, affinity="nearest_neighbors", n_jobs=-1, assign_labels="kmeans",
                                            random_state=1234, n_neighbors=(len(pts) - 1) // 2)
            idx = sp.fit_predict(pts.cpu().numpy())
            return torch.as_tensor((idx == idx[0])[1:], device=pts.device)
        from ..spectral import spectral_bipartition
        return spectral_bipartition(pts, (len(pts) - 1) // 2)

    # ------------------------------------------------------------------ locations and boxes
    def compute_locations(self, features):
        """Per-level location grids (graph_matching.py:609-635).  They depend on the map sizes only, so they are
        built once per shape set and reused (the reference rebuilds ~30 small tensors per domain per step)."""
        key = tuple((f.size(-2), f.size(-1), str(f.device)) for f in features)
        cache = getattr(self, "_loc_cache", None)
        if cache is None or cache[0] != key:
            locs = [self.compute_locations_per_level(f.size(-2), f.size(-1), self.fpn_strides[l], f.device)
                    for l, f in enumerate(features)]
            self._loc_cache = cache = (key, locs)
        return cache[1]

    def compute_locations_per_level(self, h, w, stride, device):
        ys = torch.arange(0, h * stride, step=stride, dtype=torch.float32, device=device)
        xs = torch.arange(0, w * stride, step=stride, dtype=torch.float32, device=device)
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        return torch.stack((gx.reshape(-1), gy.reshape(-1)), dim=1) + stride // 2

    def masks_to_boxes(self, masks):
        """[K,H,W] -> [K,4] (xmin, ymin, xmax, ymax); an empty mask gives (0,0,W,H) (graph_matching.py:702-740).
        Vectorised: no per-mask host sync."""
        return self._boxes(masks.unsqueeze(0))[0]

    @staticmethod
    def _boxes(masks):
        B, K, H, W = masks.shape
        nz = masks != 0
        col, row = nz.any(dim=2), nz.any(dim=3)                                    # [B,K,W], [B,K,H]
        xs = torch.arange(W, device=masks.device).view(1, 1, W)
        ys = torch.arange(H, device=masks.device).view(1, 1, H)
        xmin = torch.where(col, xs, W).amin(-1)
        xmax = torch.where(col, xs, -1).amax(-1)
        ymin = torch.where(row, ys, H).amin(-1)
        ymax = torch.where(row, ys, -1).amax(-1)
        empty = ~col.any(-1)
        box = torch.stack([xmin, ymin, xmax, ymax], dim=-1).float()
        full = torch.tensor([0.0, 0.0, float(W), float(H)], device=masks.device).expand_as(box)
        return torch.where(empty.unsqueeze(-1), full, box)

    def find_bbox(self, masks):
        """[B,K,H,W] -> [B,K,4] (indexable per image like the reference's list).  CUDA maps (and score maps given by
        their logits, GF.LogitMap) take the one-CTA-per-plane kernel."""
        if isinstance(masks, GF.LogitMap) or (torch.is_tensor(masks) and masks.is_cuda and masks.dim() == 4):
            return GF.mask_boxes(masks)
        return self._boxes(masks)


class PrototypeComputation(object):
    """Node sampler (graph_matching.py:861-1065, `locations` branch — the one both trainer calls take)."""

    def __init__(self, num_class):
        self.num_class = num_class
        self.class_threshold = (0.5, 1.0)
        self.num_nodes_per_class = 100
        self.num_nodes_per_lvl = 100
        self.bg_ratio = 8
        self.sample_bg_nodes = True

    def prepare_targets(self, points, targets):
        """Per level: labels of every location of every image (image-major), [B * h_l * w_l]."""
        sizes = [[-1, 64], [64, 128], [128, 256], [256, 512], [512, INF]]
        soi = torch.cat([p.new_tensor(sizes[l])[None].expand(len(p), -1) for l, p in enumerate(points)], dim=0)
        labels = self.compute_targets_for_locations(torch.cat(points, dim=0), targets, soi)      # [B, L]
        return list(torch.split(labels, [len(p) for p in points], dim=1))

    def compute_targets_for_locations(self, locations, targets, object_sizes_of_interest):
        """targets [B,K,4] boxes.  A location takes the class of the smallest-area box that contains it
        and whose max side distance lies in the level's range; otherwise 0 (graph_matching.py:913-959)."""
        boxes = targets if torch.is_tensor(targets) else torch.stack(list(targets))
        xs, ys = locations[:, 0].view(1, -1, 1), locations[:, 1].view(1, -1, 1)
        x1, y1, x2, y2 = (boxes[:, None, :, i] for i in range(4))                   # [B,1,K]
        area = ((y2 - y1) * (x2 - x1)).expand(-1, locations.size(0), -1).clone()
        reg = torch.stack([xs - x1, ys - y1, x2 - xs, y2 - ys], dim=3)              # [B,L,K,4]
        inside = reg.min(dim=3)[0] > 0
        mx = reg.max(dim=3)[0]
        soi = object_sizes_of_interest
        cared = (mx >= soi[None, :, 0:1]) & (mx <= soi[None, :, 1:2])
        area[~inside] = INF
        area[~cared] = INF
        amin, ainds = area.min(dim=2)
        labels = torch.arange(self.num_class, device=boxes.device)[ainds]
        labels[amin == INF] = 0
        return labels

    def plan(self, locations, boxes, level_hw=None, strides=None):
        """Device-side half of the sampler: per-level label maps + the [levels, 2] (positive, negative)
        counts the host needs.  No synchronisation.  With the level geometry (`level_hw`, `strides`) and CUDA boxes this
        is ONE kernel launch (ge_sampler_labels); otherwise the vectorised torch route."""
        if level_hw is not None and torch.is_tensor(boxes) and boxes.is_cuda and boxes.shape[1] <= 8 and len(level_hw) <= 5:
            sizes = [[-1, 64], [64, 128], [128, 256], [256, 512], [512, INF]]
            return GF.sampler_labels(boxes, level_hw, strides, sizes[:len(level_hw)])
        labels = [l.reshape(-1) for l in self.prepare_targets(locations, boxes)]
        counts = torch.stack([torch.stack([(l > 0).sum(), (l == 0).sum()]) for l in labels])
        return labels, counts

    def gather(self, features, labels, counts, batch_offset=0):
        """Host-driven half (graph_matching.py:978-1013) given the counts: strided positive picks,
        floor(linspace) negative picks, row gathers from the NHWC feature maps.  `features` may hold more
        images than were labelled; `batch_offset` is the index of the first labelled image."""
        C = features[0].size(1)
        pos_pts, pos_lab, neg_pts = [], [], []
        for l, lab in enumerate(labels):
            f = features[l]
            flat = f.permute(0, 2, 3, 1).reshape(-1, C)
            shift = batch_offset * f.size(2) * f.size(3)
            n_pos_all, n_neg_all = counts[l]
            pos_idx = _nonzero_known(lab > 0, n_pos_all)
            step = n_pos_all // self.num_nodes_per_class
            if step > 1:
                pos_idx = pos_idx[::step]
            pos_pts.append(flat[pos_idx + shift])
            pos_lab.append(lab[pos_idx])
            num_pos = pos_idx.numel()
            if self.sample_bg_nodes:
                neg_idx = _nonzero_known(lab == 0, n_neg_all)
                if n_pos_all <= n_neg_all:
                    pick = np.floor(np.linspace(0, n_neg_all - 2, num_pos // self.bg_ratio)).astype(np.int64)
                    neg_idx = neg_idx[torch.as_tensor(pick, device=neg_idx.device)]
                neg_pts.append(flat[neg_idx + shift])
        pos_pts, pos_lab = torch.cat(pos_pts, dim=0), torch.cat(pos_lab, dim=0)
        if self.sample_bg_nodes:
            neg_pts = torch.cat(neg_pts, dim=0)
            pos_pts = torch.cat([neg_pts, pos_pts], dim=0)
            pos_lab = torch.cat([pos_lab.new_zeros(neg_pts.size(0)), pos_lab])
        return pos_pts, pos_lab, torch.ones_like(pos_lab).long()

    def gather_pair(self, *domains):
        """`gather` for several domains (features, labels, counts, batch_offset) at once.  CUDA NHWC feature maps take
        ONE kernel launch for every level of every domain (ge_sampler_gather; its backward is one scatter launch, with
        one gradient tensor per feature map even when the domains share the maps); anything else the torch route."""
        f0 = domains[0][0][0]
        fused = (f0.is_cuda and all(f.is_cuda and f.dim() == 4 and f.dtype == f0.dtype and f.size(1) == f0.size(1)
                                    for d in domains for f in d[0])
                 and f0.dtype in (torch.float32, torch.bfloat16) and f0.size(1) % 4 == 0
                 and sum(len(d[1]) for d in domains) <= 10)
        if not fused:
            return [self.gather(*d) for d in domains]
        feats, plan = [], []
        for features, labels, counts, boff in domains:
            fidx = []
            for f in features[:len(labels)]:
                k = next((i for i, g in enumerate(feats) if g is f), None)
                if k is None:
                    feats.append(f)
                    k = len(feats) - 1
                fidx.append(k)
            rows, n_nodes = GF.sampler_gather_plan(counts, self.num_nodes_per_class, self.bg_ratio, self.sample_bg_nodes)
            plan.append((list(labels), fidx, int(boff), rows, n_nodes))
        out = GF.sampler_gather(plan, feats)
        return [(n, l, torch.ones_like(l)) for n, l in out]

    def __call__(self, locations, features, targets):
        if not locations:
            raise NotImplementedError("the score-map sampling branch (graph_matching.py:1016-1065) is unreachable "
                                      "from the reference trainers and is not on the accelerated path")
        labels, counts = self.plan(locations, targets)
        return self.gather(features, labels, counts.tolist(), 0)

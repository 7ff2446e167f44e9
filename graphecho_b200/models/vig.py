"""Mirror of models/vig.py (vendored Vision-GNN): k-NN graph build, graph convolutions, Grapher, FFN,
Stem, Downsample, DeepGCN and the pvig_* factories, with the reference's names, signatures and
state_dict keys.

Hot path (reached through models.TGCN and Grapher): DenseDilatedKnnGraph -> ge_knn_graph (fused
normalise + fp32 distance + running top-k, no [B,N,M] matrix) and MRConv2d -> ge_mrconv_gather
(one pass, no [B,C,N,k] gathers).  The grouped 1x1 convs / BatchNorms are dense library calls.
The stray debugging `print`s of the reference (vig.py:204, 589, 649) are not reproduced."""
import math

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn
from torch.nn import Sequential as Seq

from .. import functional as GF


# ------------------------------------------------------------------ relative position embedding
def get_2d_relative_pos_embed(embed_dim, grid_size):
    """[grid_size^2, grid_size^2] = 2 * PE PE^T / D  (vig.py:21-29)."""
    pe = get_2d_sincos_pos_embed(embed_dim, grid_size)
    return 2 * np.matmul(pe, pe.transpose()) / pe.shape[1]


def get_2d_sincos_pos_embed(embed_dim, grid_size, cls_token=False):
    gh = np.arange(grid_size, dtype=np.float32)
    gw = np.arange(grid_size, dtype=np.float32)
    grid = np.stack(np.meshgrid(gw, gh), axis=0).reshape([2, 1, grid_size, grid_size])
    pe = get_2d_sincos_pos_embed_from_grid(embed_dim, grid)
    if cls_token:
        pe = np.concatenate([np.zeros([1, embed_dim]), pe], axis=0)
    return pe


def get_2d_sincos_pos_embed_from_grid(embed_dim, grid):
    assert embed_dim % 2 == 0
    return np.concatenate([get_1d_sincos_pos_embed_from_grid(embed_dim // 2, grid[0]),
                           get_1d_sincos_pos_embed_from_grid(embed_dim // 2, grid[1])], axis=1)


def get_1d_sincos_pos_embed_from_grid(embed_dim, pos):
    assert embed_dim % 2 == 0
    omega = 1.0 / 10000 ** (np.arange(embed_dim // 2, dtype=float) / (embed_dim / 2.0))
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


# ------------------------------------------------------------------ functional API surface
def batched_index_select(x, idx):
    """x [B,C,M,1], idx [B,N,k] -> [B,C,N,k]  (vig.py:209-229)."""
    b, c, m = x.shape[:3]
    _, n, k = idx.shape
    flat = x.reshape(b, c, m).transpose(1, 2).reshape(b * m, c)
    off = torch.arange(b, device=idx.device).view(-1, 1, 1) * m
    return flat[(idx + off).reshape(-1)].view(b, n, k, c).permute(0, 3, 1, 2).contiguous()


def pairwise_distance(x):
    with torch.no_grad():
        sq = (x * x).sum(-1, keepdim=True)
        return sq + (-2 * torch.matmul(x, x.transpose(2, 1))) + sq.transpose(2, 1)


def part_pairwise_distance(x, start_idx=0, end_idx=1):
    with torch.no_grad():
        part = x[:, start_idx:end_idx]
        sq = (x * x).sum(-1, keepdim=True)
        return (part * part).sum(-1, keepdim=True) + (-2 * torch.matmul(part, x.transpose(2, 1))) + sq.transpose(2, 1)


def xy_pairwise_distance(x, y):
    with torch.no_grad():
        return (x * x).sum(-1, keepdim=True) + (-2 * torch.matmul(x, y.transpose(2, 1))) + \
            (y * y).sum(-1, keepdim=True).transpose(2, 1)


def dense_knn_matrix(x, k=16, relative_pos=None):
    """k-NN on ALREADY-normalised x [B,C,N,1] (vig.py:277-309).  The fused kernel normalises
    internally; normalising a unit vector again is the identity up to one rounding."""
    return _tag(GF.knn_graph(x, None, k, 1, relative_pos))


def xy_dense_knn_matrix(x, y, k=16, relative_pos=None):
    return _tag(GF.knn_graph(x, y, k, 1, relative_pos))


def _tag(edge_index):
    edge_index._ge_identity_centre = True     # centre index == point index (vig.py:308/328)
    return edge_index


# ------------------------------------------------------------------ graph convolutions
class MRConv2d(nn.Module):
    """Max-relative graph conv (vig.py:88-105): grouped 1x1 conv over the interleaved
    [x ; max_k(x_j - x_i)] produced by the fused gather kernel."""

    def __init__(self, in_channels, out_channels, act="relu", norm=None, bias=True):
        super().__init__()
        self.nn = BasicConv([in_channels * 2, out_channels], act, norm, bias)

    def forward(self, x, edge_index, y=None):
        ident = getattr(edge_index, "_ge_identity_centre", False)
        feat = GF.mr_gather(x, edge_index, y, identity_centre=ident)
        return self.nn(feat)


class EdgeConv2d(nn.Module):
    def __init__(self, in_channels, out_channels, act="relu", norm=None, bias=True):
        super().__init__()
        self.nn = BasicConv([in_channels * 2, out_channels], act, norm, bias)

    def forward(self, x, edge_index, y=None):
        x_i = batched_index_select(x, edge_index[1])
        x_j = batched_index_select(y if y is not None else x, edge_index[0])
        return self.nn(torch.cat([x_i, x_j - x_i], dim=1)).max(-1, keepdim=True)[0]


class GraphSAGE(nn.Module):
    def __init__(self, in_channels, out_channels, act="relu", norm=None, bias=True):
        super().__init__()
        self.nn1 = BasicConv([in_channels, in_channels], act, norm, bias)
        self.nn2 = BasicConv([in_channels * 2, out_channels], act, norm, bias)

    def forward(self, x, edge_index, y=None):
        x_j = batched_index_select(y if y is not None else x, edge_index[0])
        x_j = self.nn1(x_j).max(-1, keepdim=True)[0]
        return self.nn2(torch.cat([x, x_j], dim=1))


class GINConv2d(nn.Module):
    def __init__(self, in_channels, out_channels, act="relu", norm=None, bias=True):
        super().__init__()
        self.nn = BasicConv([in_channels, out_channels], act, norm, bias)
        self.eps = nn.Parameter(torch.Tensor([0.0]))

    def forward(self, x, edge_index, y=None):
        x_j = batched_index_select(y if y is not None else x, edge_index[0]).sum(-1, keepdim=True)
        return self.nn((1 + self.eps) * x + x_j)


class GraphConv2d(nn.Module):
    def __init__(self, in_channels, out_channels, conv="edge", act="relu", norm=None, bias=True):
        super().__init__()
        table = {"edge": EdgeConv2d, "mr": MRConv2d, "sage": GraphSAGE, "gin": GINConv2d}
        if conv not in table:
            raise NotImplementedError("conv:{} is not supported".format(conv))
        self.gconv = table[conv](in_channels, out_channels, act, norm, bias)

    def forward(self, x, edge_index, y=None):
        return self.gconv(x, edge_index, y)


class DyGraphConv2d(GraphConv2d):
    """vig.py:184-206: (optional r x r avg-pooled keys) -> k-NN graph -> graph conv."""

    def __init__(self, in_channels, out_channels, kernel_size=9, dilation=1, conv="edge", act="relu",
                 norm=None, bias=True, stochastic=False, epsilon=0.0, r=1):
        super().__init__(in_channels, out_channels, conv, act, norm, bias)
        self.k, self.d, self.r = kernel_size, dilation, r
        self.dilated_knn_graph = DenseDilatedKnnGraph(kernel_size, dilation, stochastic, epsilon)

    def forward(self, x, relative_pos=None):
        B, C, H, W = x.shape
        y = None
        if self.r > 1:
            y = F.avg_pool2d(x, self.r, self.r).reshape(B, C, -1, 1).float().contiguous()
        # one fp32 [B,C,N,1] copy feeds both the k-NN build and the max-relative gather
        x = x.reshape(B, C, -1, 1).float().contiguous()
        edge_index = self.dilated_knn_graph(x, y, relative_pos)
        x = super().forward(x, edge_index, y)
        return x.reshape(B, -1, H, W).contiguous()


class DenseDilated(nn.Module):
    """Dilated pick from a k*d neighbour list (vig.py:332-354)."""

    def __init__(self, k=9, dilation=1, stochastic=False, epsilon=0.0):
        super().__init__()
        self.dilation, self.stochastic, self.epsilon, self.k = dilation, stochastic, epsilon, k

    def forward(self, edge_index):
        if self.stochastic and torch.rand(1) < self.epsilon and self.training:
            pick = torch.randperm(self.k * self.dilation)[: self.k]
            return edge_index[:, :, :, pick]
        return edge_index[:, :, :, :: self.dilation]


class DenseDilatedKnnGraph(nn.Module):
    """forward(x [B,C,N,1], y=None, relative_pos=None) -> int64 edge_index [2,B,N,k] (vig.py:357-381).
    Channel-wise L2 normalisation, distance, top-(k*d) and the dilation stride all happen inside
    ge_knn_graph; no gradient flows (the reference computes indices under no_grad)."""

    def __init__(self, k=9, dilation=1, stochastic=False, epsilon=0.0):
        super().__init__()
        self.dilation, self.stochastic, self.epsilon, self.k = dilation, stochastic, epsilon, k
        self._dilated = DenseDilated(k, dilation, stochastic, epsilon)

    def forward(self, x, y=None, relative_pos=None):
        if self.stochastic and self.training and torch.rand(1) < self.epsilon:
            full = GF.knn_graph(x, y, self.k * self.dilation, 1, relative_pos)
            pick = torch.randperm(self.k * self.dilation)[: self.k].to(full.device)
            return _tag(full[:, :, :, pick].contiguous())
        return _tag(GF.knn_graph(x, y, self.k, self.dilation, relative_pos))


class DropPath(nn.Module):
    """Stochastic depth (timm.models.layers.DropPath semantics); identity at drop_prob = 0."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x.div(keep) * mask


def _conv_bn_train(x, conv, bn, residual=None):
    """Training-mode BatchNorm(conv(x)) without the convolution's bias passes.  A per-channel constant in front of
    train-mode BatchNorm cancels in the output and has zero gradient; it only shifts the batch mean, so the
    running mean gets its `momentum * bias` share added back (state_dict parity with the reference)."""
    y = GF.conv1x1_bn_act(x, conv, bn, residual=residual, relu=False, drop_bias=True)
    if conv.bias is not None and bn.track_running_stats:
        with torch.no_grad():
            # one running-mean update per BatchNorm segment (GF.domain_split): 1 - (1-m)^nseg of the bias in total
            share = 1.0 - (1.0 - float(bn.momentum)) ** GF.bn_segments(x.shape[0])
            bn.running_mean.add_(conv.bias.detach().to(bn.running_mean.dtype), alpha=share)
    return y


class Grapher(nn.Module):
    """fc1 (1x1 conv + BN) -> DyGraphConv2d -> fc2 (1x1 conv + BN) -> + residual (vig.py:384-430)."""

    def __init__(self, in_channels, kernel_size=9, dilation=1, conv="edge", act="relu", norm=None,
                 bias=True, stochastic=False, epsilon=0.0, r=1, n=196, drop_path=0.0, relative_pos=False):
        super().__init__()
        self.channels, self.n, self.r = in_channels, n, r
        self.fc1 = nn.Sequential(nn.Conv2d(in_channels, in_channels, 1), nn.BatchNorm2d(in_channels))
        self.graph_conv = DyGraphConv2d(in_channels, in_channels * 2, kernel_size, dilation, conv, act, norm,
                                        bias, stochastic, epsilon, r)
        self.fc2 = nn.Sequential(nn.Conv2d(in_channels * 2, in_channels, 1), nn.BatchNorm2d(in_channels))
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.relative_pos = None
        if relative_pos:
            rel = torch.from_numpy(np.float32(get_2d_relative_pos_embed(in_channels, int(n ** 0.5))))
            rel = F.interpolate(rel[None, None], size=(n, n // (r * r)), mode="bicubic", align_corners=False)
            self.relative_pos = nn.Parameter(-rel.squeeze(1), requires_grad=False)

    def _get_relative_pos(self, relative_pos, H, W):
        if relative_pos is None or H * W == self.n:
            return relative_pos
        N = H * W
        return F.interpolate(relative_pos.unsqueeze(0), size=(N, N // (self.r * self.r)), mode="bicubic").squeeze(0)

    def _node_major_ok(self, x):
        """The channels-last fast path: max-relative conv, BatchNorm, no relative position bias, and a problem the
        tcgen05 k-NN kernel covers.  Everything else takes the reference-shaped [B,C,N,1] route below."""
        gc = self.graph_conv
        B, C, H, W = x.shape
        N = H * W
        M = (H // self.r) * (W // self.r) if self.r > 1 else N
        seq = getattr(gc.gconv, "nn", None)
        return (x.is_cuda and self.relative_pos is None and isinstance(gc.gconv, MRConv2d) and seq is not None
                and len(seq) == 3 and type(seq[1]) is nn.BatchNorm2d and isinstance(self.drop_path, nn.Identity)
                and not gc.dilated_knn_graph._dilated.stochastic and gc.k <= 32
                and GF.knn_nmajor_supported(B, C, N, M, gc.k, gc.d))

    def _forward_node_major(self, x, shortcut):
        """Grapher body on the NHWC map itself: its [B, H*W, C] view IS the node-major graph tensor, so the k-NN
        build, the max-relative gather (bf16 in, bf16 out under autocast), the grouped 1x1 conv and both
        BatchNorms run without a single layout change or fp32 copy (the [B,C,N,1] route makes four)."""
        gc = self.graph_conv
        B, C, H, W = x.shape
        xn = x.permute(0, 2, 3, 1).reshape(B, H * W, C)                  # view of the channels_last storage
        yn = None
        if self.r > 1:
            y = F.avg_pool2d(x, self.r, self.r).contiguous(memory_format=torch.channels_last)
            yn = y.permute(0, 2, 3, 1).reshape(B, -1, C)
        edge = GF.knn_graph_nmajor(xn, yn, gc.k, gc.d)[0]
        feat = GF.mr_gather_nmajor(xn, edge, yn)                         # [B,N,2C], reference channel interleaving
        f4 = feat.view(B, H, W, 2 * C).permute(0, 3, 1, 2)               # logical NCHW, channels_last strides
        conv, bn, act = gc.gconv.nn[0], gc.gconv.nn[1], gc.gconv.nn[2]
        y = act(_conv_bn_train(f4, conv, bn))
        return _conv_bn_train(y, self.fc2[0], self.fc2[1], residual=shortcut)

    def forward(self, x):
        shortcut = x
        fast = x.is_cuda and self.training and all(type(m) is nn.BatchNorm2d and m.momentum is not None
                                                   for m in (self.fc1[1], self.fc2[1]))
        if fast:
            x = _conv_bn_train(x, self.fc1[0], self.fc1[1])
        else:
            x = GF.bn_act(self.fc1[0](x), self.fc1[1], relu=False)
        _, _, H, W = x.shape
        if fast and self._node_major_ok(x):
            return self._forward_node_major(x.contiguous(memory_format=torch.channels_last), shortcut)
        x = self.graph_conv(x, self._get_relative_pos(self.relative_pos, H, W))
        if isinstance(self.drop_path, nn.Identity):
            return GF.bn_act(self.fc2[0](x), self.fc2[1], residual=shortcut, relu=False)   # BN + residual in one pass
        return self.drop_path(GF.bn_act(self.fc2[0](x), self.fc2[1], relu=False)) + shortcut


def act_layer(act, inplace=False, neg_slope=0.2, n_prelu=1):
    act = act.lower()
    if act == "relu":
        return nn.ReLU(inplace)
    if act == "leakyrelu":
        return nn.LeakyReLU(neg_slope, inplace)
    if act == "prelu":
        return nn.PReLU(num_parameters=n_prelu, init=neg_slope)
    if act == "gelu":
        return nn.GELU()
    if act == "hswish":
        return nn.Hardswish(inplace)
    raise NotImplementedError("activation layer [%s] is not found" % act)


def norm_layer(norm, nc):
    norm = norm.lower()
    if norm == "batch":
        return nn.BatchNorm2d(nc, affine=True)
    if norm == "instance":
        return nn.InstanceNorm2d(nc, affine=False)
    raise NotImplementedError("normalization layer [%s] is not found" % norm)


class BasicConv(Seq):
    """Stack of grouped (groups=4) 1x1 convs (+norm) (+act) (vig.py:476-500)."""

    def __init__(self, channels, act="relu", norm=None, bias=True, drop=0.0):
        m = []
        for i in range(1, len(channels)):
            m.append(nn.Conv2d(channels[i - 1], channels[i], 1, bias=bias, groups=4))
            if norm is not None and norm.lower() != "none":
                m.append(norm_layer(norm, channels[-1]))
            if act is not None and act.lower() != "none":
                m.append(act_layer(act))
            if drop > 0:
                m.append(nn.Dropout2d(drop))
        super().__init__(*m)
        self.reset_parameters()

    def reset_parameters(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, (nn.BatchNorm2d, nn.InstanceNorm2d)) and m.weight is not None:
                m.weight.data.fill_(1)
                m.bias.data.zero_()


class FFN(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act="relu", drop_path=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Sequential(nn.Conv2d(in_features, hidden_features, 1), nn.BatchNorm2d(hidden_features))
        self.act = act_layer(act)
        self.fc2 = nn.Sequential(nn.Conv2d(hidden_features, out_features, 1), nn.BatchNorm2d(out_features))
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()

    def forward(self, x):
        return self.drop_path(self.fc2(self.act(self.fc1(x)))) + x


class Stem(nn.Module):
    def __init__(self, img_size=224, in_dim=3, out_dim=768, act="relu"):
        super().__init__()
        self.convs = nn.Sequential(
            nn.Conv2d(in_dim, out_dim // 2, 3, 2, 1), nn.BatchNorm2d(out_dim // 2), act_layer(act),
            nn.Conv2d(out_dim // 2, out_dim, 3, 2, 1), nn.BatchNorm2d(out_dim), act_layer(act),
            nn.Conv2d(out_dim, out_dim, 3, 1, 1), nn.BatchNorm2d(out_dim))

    def forward(self, x):
        return self.convs(x)


class Downsample(nn.Module):
    def __init__(self, in_dim=3, out_dim=768):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(in_dim, out_dim, 3, 2, 1), nn.BatchNorm2d(out_dim))

    def forward(self, x):
        return self.conv(x)


class DeepGCN(nn.Module):
    """Pyramid ViG classifier (vig.py:586-651); composed from Grapher + FFN blocks."""

    def __init__(self, opt):
        super().__init__()
        k, act, norm, bias = opt.k, opt.act, opt.norm, opt.bias
        blocks, channels = opt.blocks, opt.channels
        self.n_blocks = sum(blocks)
        reduce_ratios = [4, 2, 1, 1]
        dpr = [x.item() for x in torch.linspace(0, opt.drop_path, self.n_blocks)]
        num_knn = [int(x.item()) for x in torch.linspace(k, k, self.n_blocks)]
        max_dilation = 49 // max(num_knn)
        self.stem = Stem(out_dim=channels[0], act=act)
        self.pos_embed = nn.Parameter(torch.zeros(1, channels[0], 224 // 4, 224 // 4))
        HW = 224 // 4 * 224 // 4
        layers, idx = [], 0
        for i in range(len(blocks)):
            if i > 0:
                layers.append(Downsample(channels[i - 1], channels[i]))
                HW = HW // 4
            for _ in range(blocks[i]):
                layers.append(Seq(
                    Grapher(channels[i], num_knn[idx], min(idx // 4 + 1, max_dilation), opt.conv, act, norm, bias,
                            opt.use_stochastic, opt.epsilon, reduce_ratios[i], n=HW, drop_path=dpr[idx],
                            relative_pos=True),
                    FFN(channels[i], channels[i] * 4, act=act, drop_path=dpr[idx])))
                idx += 1
        self.backbone = Seq(*layers)
        self.prediction = Seq(nn.Conv2d(channels[-1], 1024, 1, bias=True), nn.BatchNorm2d(1024), act_layer(act),
                              nn.Dropout(opt.dropout), nn.Conv2d(1024, opt.n_classes, 1, bias=True))
        self.model_init()

    def model_init(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)
                m.weight.requires_grad = True
                if m.bias is not None:
                    m.bias.data.zero_()
                    m.bias.requires_grad = True

    def forward(self, inputs):
        x = self.stem(inputs) + self.pos_embed
        x = self.backbone(x)
        x = F.adaptive_avg_pool2d(x, 1)
        return self.prediction(x).squeeze(-1).squeeze(-1)


class _PvigOpt:
    def __init__(self, blocks, channels, num_classes=1000, drop_path_rate=0.0, **kwargs):
        self.k, self.conv, self.act, self.norm, self.bias = 9, "mr", "gelu", "batch", True
        self.dropout, self.use_dilation, self.epsilon, self.use_stochastic = 0.0, True, 0.2, False
        self.drop_path, self.blocks, self.channels = drop_path_rate, blocks, channels
        self.n_classes, self.emb_dims = num_classes, 1024


def _pvig(blocks, channels, **kwargs):
    return DeepGCN(_PvigOpt(blocks, channels, **kwargs))


def pvig_ti_224_gelu(pretrained=False, **kwargs):
    return _pvig([2, 2, 6, 2], [48, 96, 240, 384], **kwargs)


def pvig_s_224_gelu(pretrained=False, **kwargs):
    return _pvig([2, 2, 6, 2], [80, 160, 400, 640], **kwargs)


def pvig_m_224_gelu(pretrained=False, **kwargs):
    return _pvig([2, 2, 16, 2], [96, 192, 384, 768], **kwargs)


def pvig_b_224_gelu(pretrained=False, **kwargs):
    return _pvig([2, 2, 18, 2], [128, 256, 512, 1024], **kwargs)

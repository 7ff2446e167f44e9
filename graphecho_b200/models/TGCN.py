"""Mirror of models/TGCN.py: the temporal graph module (DyGraphConv2d recurrence over a clip +
graph attention with the GModule nodes + node discriminator / Sinkhorn transport), with the
reference's [sic] constructor signature, forward contract and state_dict keys.

B200 design: everything in the per-timestep DyGraphConv2d that does not depend on the recurrent
state is hoisted out of the time loop and batched over all b*t frames -- the pyramid pooling +
concat is ONE fused pass per level (ge_tgcn_pool_concat, each level read once), the 1x1-conv MLP
is two dense GEMMs, and BatchNorm keeps the reference's per-timestep batch statistics (one
batch_norm call over (t, C) "channels").  Only k-NN(x_t, hidden) + max-relative conv remain in
the recurrence (ge_knn_graph + ge_mrconv_gather)."""
import argparse

import torch
import torch.nn.functional as F
from torch import nn

from .. import functional as GF
from .gradient_reversal import GradientReversal
from .transformer import MultiHeadAttention
from .vig import DenseDilatedKnnGraph, GraphConv2d


def calculate_laplacian_with_self_loop(matrixs):
    """TGCN.py:11-23 (dead code in the reference; API surface)."""
    out = []
    for m in matrixs:
        m = m + torch.eye(m.size(0), device=m.device)
        d = torch.pow(m.sum(1), -0.5).flatten()
        d[torch.isinf(d)] = 0.0
        dm = torch.diag(d)
        out.append(m.matmul(dm).transpose(0, 1).matmul(dm).unsqueeze(0))
    return torch.cat(out, dim=0)


def calculate_laplacian_without_self_loop(graph, normalize=None):
    if normalize:
        D = torch.diag(torch.sum(graph, dim=-1) ** (-1 / 2))
        return torch.eye(graph.size(0), device=graph.device, dtype=graph.dtype) - torch.mm(torch.mm(D, graph), D)
    return torch.diag(torch.sum(graph, dim=-1)) - graph


class DyGraphConv2d(GraphConv2d):
    """TGCN flavour (TGCN.py:41-78): forward(input, rs, y, learnable_pos) -> ([B,C,N], H, W)."""

    def __init__(self, in_channels, out_channels, kernel_size=9, dilation=1, conv="mr", act="gelu",
                 norm=None, bias=True, stochastic=False, epsilon=0.2):
        super().__init__(in_channels, out_channels, conv, act, norm, bias)
        self.k, self.d = kernel_size, dilation
        self.MLP = nn.Sequential(nn.Conv2d(in_channels * 4, out_channels, 1), nn.BatchNorm2d(out_channels),
                                 nn.GELU(), nn.Dropout(0.1), nn.Conv2d(out_channels, out_channels, 1))
        self.dilated_knn_graph = DenseDilatedKnnGraph(kernel_size, dilation, stochastic, epsilon)

    # -- state-independent half, batched over every frame of the clip batch ------------------
    def embed_frames(self, levels, rs, steps):
        """levels: list of [F,C,s,s] with F = b*steps frames ordered (b, t).  Returns [F,Cout,H,W] with
        BatchNorm statistics taken per timestep over the b clips, exactly as `steps` separate calls."""
        x = GF.pool_concat(levels, rs)                                   # fp32 channels_last [F,4C,H,W]
        conv1, bn, act, drop, conv2 = self.MLP
        x = conv1(x)
        Fr, C, H, W = x.shape
        if bn.training and steps > 1:
            b = Fr // steps
            xt = x.reshape(b, steps * C, H, W)                           # (t, c) pairs as channels
            mean = x.new_zeros(steps * C)
            var = x.new_ones(steps * C)
            xt = F.batch_norm(xt, mean, var, bn.weight.repeat(steps), bn.bias.repeat(steps), True, 1.0, bn.eps)
            x = xt.reshape(Fr, C, H, W)
            if bn.track_running_stats:
                with torch.no_grad():
                    m = bn.momentum
                    for t in range(steps):                               # sequential momentum updates, as t calls
                        bn.running_mean.mul_(1 - m).add_(mean[t * C:(t + 1) * C], alpha=m)
                        bn.running_var.mul_(1 - m).add_(var[t * C:(t + 1) * C], alpha=m)
                    bn.num_batches_tracked += steps
        else:
            x = bn(x)
        return conv2(drop(act(x)))

    # -- recurrent half ------------------------------------------------------------------------
    def graph_step(self, x, y):
        """x [B,C,N,1] (embedded frame + position), y [B,C,N] hidden state -> [B,Cout,N,1]."""
        edge_index = self.dilated_knn_graph(x, y, None)
        return GraphConv2d.forward(self, x, edge_index, y)

    def forward(self, input, rs, y, learnable_pos, relative_pos=None):
        x = self.embed_frames([t for t in input], rs, 1) + learnable_pos
        B, C, H, W = x.shape
        x = x.float().reshape(B, C, -1, 1).contiguous()
        out = self.graph_step(x, y)
        return out.reshape(B, -1, H * W).contiguous(), H, W


class TGCNGraphConvolution(nn.Module):
    """TGCN.py:81-137 (never instantiated by the reference; API surface)."""

    def __init__(self, in_feature_dim, num_gru_units, output_dim, bias=0.0):
        super().__init__()
        self._in_feature_num, self._num_gru_units = in_feature_dim, num_gru_units
        self._output_dim, self._bias_init_value = output_dim, bias
        self.weights = nn.Parameter(torch.FloatTensor(num_gru_units + in_feature_dim, output_dim))
        self.biases = nn.Parameter(torch.FloatTensor(output_dim))
        nn.init.xavier_uniform_(self.weights)
        nn.init.constant_(self.biases, bias)

    def forward(self, inputs, hidden_state):
        b, n, f = inputs.shape
        lap = calculate_laplacian_with_self_loop(inputs)
        cat = torch.cat((inputs, hidden_state.reshape(b, n, self._num_gru_units)), dim=2)
        ax = torch.einsum("bnc,bck->bnk", lap, cat).reshape(b * n, self._num_gru_units + f)
        return (ax @ self.weights + self.biases).reshape(b, n * self._output_dim)


class TGCNCell(nn.Module):
    def __init__(self, input_dim, hidden_dim):
        super().__init__()
        self._input_dim, self._hidden_dim = input_dim, hidden_dim
        self.graph_conv1 = TGCNGraphConvolution(input_dim, hidden_dim, hidden_dim * 2, bias=1.0)
        self.graph_conv2 = TGCNGraphConvolution(input_dim, hidden_dim, hidden_dim)

    def forward(self, inputs, hidden_state):
        r, u = torch.chunk(torch.sigmoid(self.graph_conv1(inputs, hidden_state)), chunks=2, dim=1)
        c = torch.tanh(self.graph_conv2(inputs, r * hidden_state))
        new = u * hidden_state + (1.0 - u) * c
        return new, new


class TGCN(nn.Module):
    """TGCN(input_dim, hidden_dim, clip_shape, soucre_class, target_class, cluster_method=None,
            transport_method='node_discriminate')
    forward(input_features, input_feature_nodes, loss_trans, loss_cluster, update_index, r) -> loss dict"""

    def __init__(self, input_dim, hidden_dim, clip_shape, soucre_class, target_class,
                 cluster_method=None, transport_method="node_discriminate"):
        super().__init__()
        self._input_dim, self._hidden_dim = input_dim, hidden_dim
        self.grapher = DyGraphConv2d(input_dim, hidden_dim)
        self.graph_attention = MultiHeadAttention(256, 1, dropout=0.1, version="v2")
        self.clip_l, self.clip_h, self.clip_w = clip_shape
        self.cluster_method, self.transport_method = cluster_method, transport_method
        self.pos_embed = nn.Parameter(torch.zeros(self.clip_l, 1, input_dim, self.clip_h, self.clip_w))
        self.prediction = nn.Sequential(nn.Conv2d(hidden_dim, hidden_dim, 3, stride=2, bias=True),
                                        nn.BatchNorm2d(hidden_dim), nn.GELU(), nn.Dropout(0.1), nn.AdaptiveAvgPool2d(1))
        if cluster_method == "momentum_queue":
            self.m, self.K = 0.99, 150
            self.register_buffer("queue_source", F.normalize(torch.randn(hidden_dim, self.K), dim=0))
            self.register_buffer("queue_target", F.normalize(torch.randn(hidden_dim, self.K), dim=0))
        elif cluster_method == "linear_clustering":
            self.classifer_source = nn.Linear(hidden_dim, soucre_class)
            self.classifer_target = nn.Linear(hidden_dim, target_class)
        if transport_method == "node_discriminate":
            self.loss_bce = nn.BCEWithLogitsLoss()
            self.grad_reverse = GradientReversal(0.02)
            widths = [256, 256, 256, 256, 1]
            layers = []
            for i in range(4):
                layers.append(nn.Linear(widths[i], widths[i + 1]))
                if i < 3:
                    layers += [nn.LayerNorm(256, elementwise_affine=False), nn.ReLU()]
            self.node_dis_2 = nn.Sequential(*layers)
            for m in self.node_dis_2:
                if isinstance(m, nn.Linear):
                    nn.init.normal_(m.weight, std=0.01)
                    nn.init.constant_(m.bias, 0)

    def forward(self, input_features, input_feature_nodes, loss_trans, loss_cluster, update_index, r=1.0):
        losses = dict()
        x_f1 = input_features[0]
        source_nodes, target_nodes = input_feature_nodes
        batch_size, seq_len = x_f1.shape[:2]
        rs = list(r) if isinstance(r, (list, tuple)) else [r] * len(input_features)
        levels = [f.reshape((batch_size * seq_len,) + tuple(f.shape[2:])) for f in input_features]
        with torch.autocast("cuda", enabled=False):
            emb = self.grapher.embed_frames(levels, rs, seq_len)                    # [b*t, C, h, w]
            _, C, H, W = emb.shape
            emb = emb.float().reshape(batch_size, seq_len, C, H, W) + self.pos_embed[:seq_len, 0].unsqueeze(0)
            emb = emb.reshape(batch_size, seq_len, C, H * W)                         # -> [B,C,N] per step
            current_graph = self._recurrence(emb, H * W)
            output_f = self.prediction(current_graph.reshape(batch_size, -1, H, W)).view(batch_size, -1)
            update_index_source, update_index_target = update_index

            if self.cluster_method == "momentum_queue":
                q = F.normalize(output_f, dim=1)
                l_pos = q @ torch.cat([self.queue_source, self.queue_target], dim=-1).clone().detach()
                self._dequeue_and_enqueue(q[:batch_size // 2], self.queue_source, update_index_source)
                self._dequeue_and_enqueue(q[batch_size // 2:], self.queue_target, update_index_target)
                losses["clustering_loss"] = loss_cluster(
                    l_pos, torch.cat([update_index_source, torch.add(update_index_target, 150)]))
            elif self.cluster_method == "linear_clustering":
                losses["clustering_loss"] = \
                    loss_cluster(self.classifer_source(output_f[:batch_size // 2]), update_index_source) + \
                    loss_cluster(self.classifer_target(output_f[batch_size // 2:]), update_index_target)

            output_g = current_graph.transpose(1, 2)                                # [b, N, C]
            b_g, d_g, n_g = output_g.shape
            flat = output_g.reshape(b_g * d_g, n_g)
            nodes_ = torch.cat([flat, source_nodes.float(), target_nodes.float()])
            nodes_ = self.graph_attention(nodes_, nodes_, nodes_)[0]
            nodes_g = nodes_[: b_g * d_g].reshape(b_g, d_g, n_g)
            if self.transport_method == "node_discriminate":
                nodes_source = nodes_g[: b_g // 2].reshape(-1, n_g)
                nodes_target = nodes_g[b_g // 2:].reshape(-1, n_g)
                rev = self.grad_reverse(torch.cat([nodes_source, nodes_target], dim=0))
                tgt = torch.cat([rev.new_ones(nodes_source.size(0)), rev.new_zeros(nodes_target.size(0))])
                losses["node_dis_loss"] = 0.1 * self.loss_bce(self.node_dis_2(rev).view(-1), tgt)
            elif self.transport_method == "sinkhorn_distance":
                losses["sinkhorn_loss"] = loss_trans(nodes_g[: batch_size // 2], nodes_g[batch_size // 2:])[0]
        return losses

    persistent_recurrence = True          # one launch for the whole time loop where the kernel covers the shape

    def _recurrence(self, emb, N):
        """hidden_t = DyGraphConv2d.graph_step(x_t, hidden_{t-1}), hidden_0 = 0 (TGCN.py:230-235).  The persistent kernel
        (ge_tgcn_recurrence_*) runs all steps in one launch; other shapes take the step-by-step path."""
        B, T, C, _ = emb.shape
        g = self.grapher
        seq = getattr(g.gconv, "nn", None)
        conv = seq[0] if seq is not None and len(seq) == 2 else None
        if (self.persistent_recurrence and emb.is_cuda and conv is not None and isinstance(seq[1], nn.GELU)
                and conv.bias is not None and self._input_dim == C
                and not g.dilated_knn_graph._dilated.stochastic
                and GF.tgcn_recurrence_supported(C, conv.out_channels, N, g.k, g.d, conv.groups)):
            hidden, _ = GF.tgcn_recurrence(emb.contiguous(), conv.weight, conv.bias, g.k)
            return hidden
        hidden = torch.zeros(B, self._input_dim, self.clip_h * self.clip_w, device=emb.device, dtype=torch.float32)
        for i in range(T):
            x = emb[:, i].contiguous().unsqueeze(-1)
            hidden = g.graph_step(x, hidden).reshape(B, -1, N)
        return hidden

    @torch.no_grad()
    def _momentum_update_key_encoder(self, encoder_q, encoder_k):
        for pq, pk in zip(encoder_q.parameters(), encoder_k.parameters()):
            pk.data = pk.data * self.m + pq.data * (1.0 - self.m)

    @torch.no_grad()
    def _dequeue_and_enqueue(self, features, queue, labels):
        for idx, l_idx in enumerate(labels):
            queue[:, l_idx] = queue[:, l_idx] * self.m + features[idx] * (1.0 - self.m)

    @staticmethod
    def add_model_specific_arguments(parent_parser):
        parser = argparse.ArgumentParser(parents=[parent_parser], add_help=False)
        parser.add_argument("--hidden_dim", type=int, default=64)
        return parser

    @property
    def hyperparameters(self):
        return {"input_dim": self._input_dim, "hidden_dim": self._hidden_dim}


@torch.no_grad()
def concat_all_gather(tensor):
    """all_gather + cat over ranks (TGCN.py:315-326; defined but never called by the reference)."""
    gathered = [torch.ones_like(tensor) for _ in range(torch.distributed.get_world_size())]
    torch.distributed.all_gather(gathered, tensor, async_op=False)
    return torch.cat(gathered, dim=0)

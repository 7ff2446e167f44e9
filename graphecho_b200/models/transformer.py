"""Mirror of models/transformer.py: dot_attention, MultiHeadAttention (v1 / v2), CrossGraph.

GraphEcho uses MultiHeadAttention(256, 1, dropout=0.1, version='v2') for the intra-/cross-domain
graphs and TGCN.graph_attention; it returns the POST-dropout attention matrix, which
GModule._forward_qu consumes, so the matrix is materialised (no flash-style kernel).  The four
projections and the two attention products are plain dense GEMMs (cuBLAS, fp32)."""
import numpy as np
import torch
from torch import nn


class dot_attention(nn.Module):
    def __init__(self, attention_dropout=0.0):
        super().__init__()
        self.dropout = nn.Dropout(attention_dropout)
        self.softmax = nn.Softmax(dim=2)

    def forward(self, q, k, v, scale=None, attn_mask=None):
        scores = torch.bmm(q, k.transpose(1, 2))
        if scale:
            scores = scores * scale
        if attn_mask:
            scores = scores.masked_fill(attn_mask, -np.inf)
        attention = self.dropout(self.softmax(scores))
        return torch.bmm(attention, v), attention


class MultiHeadAttention(nn.Module):
    def __init__(self, model_dim=256, num_heads=4, dropout=0.0, version="v2"):
        super().__init__()
        self.dim_per_head = model_dim // num_heads
        self.num_heads = num_heads
        width = self.dim_per_head * num_heads
        self.linear_k = nn.Linear(model_dim, width)
        self.linear_v = nn.Linear(model_dim, width)
        self.linear_q = nn.Linear(model_dim, width)
        self.dot_product_attention = dot_attention(dropout)
        self.linear_final = nn.Linear(model_dim, model_dim)
        self.dropout = nn.Dropout(dropout)
        self.layer_norm = nn.LayerNorm(model_dim)
        self.version = version

    def forward(self, key, value, query, attn_mask=None):
        with torch.autocast("cuda", enabled=False):
            return self._forward(key.float(), value.float(), query.float(), attn_mask)

    def _forward(self, key, value, query, attn_mask):
        h, dh = self.num_heads, self.dim_per_head
        if self.version == "v2":      # transformer.py:45-75: nodes are a length-N "batch" of one token
            residual = query.unsqueeze(1)
            k = self.linear_k(key).view(key.size(0), h, dh).transpose(0, 1)
            v = self.linear_v(value).view(value.size(0), h, dh).transpose(0, 1)
            q = self.linear_q(query).view(query.size(0), h, dh).transpose(0, 1)
            scale = (dh // h) ** -0.5
            context, attention = self.dot_product_attention(q, k, v, scale, attn_mask)
            context = context.transpose(0, 1).contiguous().view(query.size(0), 1, dh * h)
        elif self.version == "v1":    # transformer.py:77-108
            residual = query.unsqueeze(0)
            k = self.linear_k(key.unsqueeze(0)).view(h, -1, dh)
            v = self.linear_v(value.unsqueeze(0)).view(h, -1, dh)
            q = self.linear_q(query.unsqueeze(0)).view(h, -1, dh)
            if attn_mask:
                attn_mask = attn_mask.repeat(h, 1, 1)
            scale = (dh // h) ** -0.5
            context, attention = self.dot_product_attention(q, k, v, scale, attn_mask)
            context = context.view(1, -1, dh * h)
        else:
            raise ValueError(f"unknown MultiHeadAttention version {self.version!r}")
        output = self.layer_norm(residual + self.dropout(self.linear_final(context)))
        return output.squeeze(), attention.squeeze()


def _mha_forward_pair(self, kv_a, q_a, kv_b, q_b):
    """Two version-'v2', single-head calls of ONE module, forward(kv_a, kv_a, q_a) and forward(kv_b, kv_b, q_b), with the
    row-wise parts (the three projections, the output projection, dropout, residual and LayerNorm) run once over the
    concatenated rows and only the two attention products kept apart.  Same arithmetic per row as two separate calls
    (transformer.py:45-75); roughly a third fewer launches forward and backward on the host-driven graph-module stream.
    Returns ((out_a, attn_a), (out_b, attn_b))."""
    assert self.version == "v2" and self.num_heads == 1
    with torch.autocast("cuda", enabled=False):
        kv_a, q_a, kv_b, q_b = kv_a.float(), q_a.float(), kv_b.float(), q_b.float()
        na = kv_a.size(0)
        same_q = q_a is kv_a and q_b is kv_b          # self-attention of two node sets
        swapped = q_a is kv_b and q_b is kv_a         # each set attends over the other one
        kv = torch.cat([kv_a, kv_b], dim=0)
        q_in = kv if (same_q or swapped) else torch.cat([q_a, q_b], dim=0)
        K, V, Q = self.linear_k(kv), self.linear_v(kv), self.linear_q(q_in)
        scale = (self.dim_per_head // self.num_heads) ** -0.5
        att = self.dot_product_attention

        def core(Qx, Kx, Vx):
            attention = att.dropout(att.softmax((torch.mm(Qx, Kx.t()) * scale).unsqueeze(0))).squeeze(0)
            return torch.mm(attention, Vx), attention

        if swapped:      # rows of q_in are [q_b ; q_a]
            ctx_a, attn_a = core(Q[na:], K[:na], V[:na])
            ctx_b, attn_b = core(Q[:na], K[na:], V[na:])
            out = self.layer_norm(q_in + self.dropout(self.linear_final(torch.cat([ctx_b, ctx_a], dim=0))))
            return (out[na:], attn_a), (out[:na], attn_b)
        ma = q_a.size(0)
        ctx_a, attn_a = core(Q[:ma], K[:na], V[:na])
        ctx_b, attn_b = core(Q[ma:], K[na:], V[na:])
        out = self.layer_norm(q_in + self.dropout(self.linear_final(torch.cat([ctx_a, ctx_b], dim=0))))
        return (out[:ma], attn_a), (out[ma:], attn_b)


MultiHeadAttention.forward_pair = _mha_forward_pair


class CrossGraph(nn.Module):
    """transformer.py:115-160 (never instantiated by the trainers; API surface)."""

    def __init__(self, model_dim=256, dropout=0.0):
        super().__init__()
        self.linear_edge = nn.Linear(model_dim, model_dim)
        self.linear_node1 = nn.Linear(model_dim, model_dim)
        self.linear_node2 = nn.Linear(model_dim, model_dim)
        self.dot_product_attention = dot_attention(dropout)
        self.linear_final = nn.Linear(model_dim, model_dim)
        self.dropout = nn.Dropout(dropout)
        self.layer_norm = nn.LayerNorm(model_dim)

    def forward(self, node_1, node_2, attn_mask=None):
        a = torch.mm(self.linear_edge(node_1), self.linear_edge(node_2).t())
        v1, v2 = self.linear_node1(node_1), self.linear_node1(node_2)
        o1 = self.dropout(self.linear_final(torch.mm(a.softmax(-1), v2)))
        o2 = self.dropout(self.linear_final(torch.mm(a.t().softmax(-1), v1)))
        return self.layer_norm(node_1 + o1), self.layer_norm(node_2 + o2)

"""dgl.nn.pytorch.conv shim: DotGatConv restated from DGL's published semantics.

DotGatConv(in_feats, out_feats, num_heads): ft = fc(h) with a bias-free Linear shared by source
and destination; a_e = <ft_src, ft_dst> / sqrt(out_feats); sa = softmax of a over the incoming
edges of each destination; out_dst = sum_e sa_e * ft_src.  Parameter name: `fc.weight`.
SAGEConv / GATConv / GATv2Conv only need to exist as names (reference assert at
models/graph_attention.py:246 and the default argument at :343).
"""
import torch


class DotGatConv(torch.nn.Module):
    def __init__(self, in_feats, out_feats, num_heads, allow_zero_in_degree=False):
        super().__init__()
        self._out_feats = out_feats
        self._num_heads = num_heads
        self.fc = torch.nn.Linear(in_feats, out_feats * num_heads, bias=False)

    def forward(self, graph, feat, get_attention=False):
        src, dst = graph.edges()
        src = src.long().to(feat.device)
        dst = dst.long().to(feat.device)
        n = feat.shape[0]
        ft = self.fc(feat).view(n, self._num_heads, self._out_feats)
        a = (ft[src] * ft[dst]).sum(-1) / self._out_feats ** 0.5          # (E, H)
        amax = torch.full((n, self._num_heads), -float("inf"), dtype=a.dtype, device=a.device)
        amax = amax.scatter_reduce(0, dst[:, None].expand_as(a), a, reduce="amax", include_self=True)
        ex = torch.exp(a - amax[dst])
        den = torch.zeros((n, self._num_heads), dtype=a.dtype, device=a.device).index_add(0, dst, ex)
        sa = ex / den[dst]
        out = torch.zeros_like(ft).index_add(0, dst, sa[:, :, None] * ft[src])
        return out


class SAGEConv(torch.nn.Module):  # name only (gnn_convolutions == 0 in grappa-1.1/1.2)
    def __init__(self, in_feats, out_feats, aggregator_type="mean"):
        super().__init__()
        self.fc_self = torch.nn.Linear(in_feats, out_feats)
        self.fc_neigh = torch.nn.Linear(in_feats, out_feats, bias=False)

    def forward(self, graph, feat):
        src, dst = graph.edges()
        src = src.long(); dst = dst.long()
        n = feat.shape[0]
        agg = torch.zeros_like(feat).index_add(0, dst, feat[src])
        deg = torch.zeros(n, dtype=feat.dtype, device=feat.device).index_add(
            0, dst, torch.ones(len(dst), dtype=feat.dtype, device=feat.device)).clamp(min=1)
        return self.fc_self(feat) + self.fc_neigh(agg / deg[:, None])


class GATConv(torch.nn.Module):
    pass


class GATv2Conv(torch.nn.Module):
    pass

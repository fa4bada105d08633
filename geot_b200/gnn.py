"""PyG-free forward of the reference's GCN / GraphSAGE models on the geot operators.

The reference's models are ``torch_geometric`` ``BasicGNN`` clones whose ``message_and_aggregate`` calls
geot (``/root/reference/models/gcn.py:26-60``, ``models/graphsage.py:26-64``, ``models/conv/spmm.py:5-14``).
PyG / torch_sparse are not part of this build, so the same math is restated on plain index tensors:

* GCN layer (``models/conv/gcnconv.py:212-259``): ``out = A_hat @ (x W) + b`` with
  ``A_hat[i,j] = d_i^-1/2 w_ij d_j^-1/2`` (``gcn_norm``, ``gcnconv.py:41-57``; ``d`` = weighted in-degree of the
  dst-sorted adjacency, optional self loops) -- one ``gather_weight_scatter`` per layer.  The reference
  recomputes ``gcn_norm`` on every forward unless ``cached``; here the normalised weights and the
  ``format_preprocess`` plan are computed once per graph.
* GraphSAGE layer (``models/conv/sageconv.py:122-154``, ``aggr="sum"`` as ``models/graphsage.py:48`` sets it):
  ``out = lin_l(sum_j x_j) + lin_r(x)`` -- one ``gather_scatter`` per layer.
* Stack (``models/basicgnn.py:216-264``): ``conv -> ReLU`` for every layer but the last; dropout p = 0.

The dense ``lin`` layers are cuBLAS GEMMs through ``torch.nn.Linear`` (library code, not part of the hot
path).  ``forward_sharded`` runs the same stack with the node rows sharded over the GPUs of one box
(``geot_b200.dist``): the GEMM is row-local with replicated weights, the aggregation exchanges the src
rows once per layer (one all-gather, or the overlapped two-bucket exchange of ``dist.BucketedGather``).
"""
from typing import List, Optional

import torch
from torch import nn

from .gather_scatter import gather_scatter
from .gather_weight_scatter import gather_weight_scatter


def add_self_loops(src_index: torch.Tensor, dst_index: torch.Tensor, num_nodes: int,
                   edge_weight: Optional[torch.Tensor] = None, fill_value: float = 1.0):
    """Sets the diagonal to ``fill_value`` (``torch_sparse.fill_diag`` semantics: existing self loops are
    replaced) and returns the edge list re-sorted by (dst, src)."""
    keep = src_index != dst_index
    loop = torch.arange(num_nodes, device=dst_index.device, dtype=dst_index.dtype)
    s = torch.cat([src_index[keep], loop])
    d = torch.cat([dst_index[keep], loop])
    w = None
    if edge_weight is not None:
        w = torch.cat([edge_weight[keep], edge_weight.new_full((num_nodes,), fill_value)])
    key, perm = torch.sort(d * num_nodes + s)
    return s[perm].contiguous(), d[perm].contiguous(), (w[perm].contiguous() if w is not None else None)


def gcn_norm(src_index: torch.Tensor, dst_index: torch.Tensor, num_nodes: int,
             edge_weight: Optional[torch.Tensor] = None, dtype=torch.float32) -> torch.Tensor:
    """Symmetric normalisation ``d_dst^-1/2 * w * d_src^-1/2`` of a dst-sorted edge list
    (``models/conv/gcnconv.py:41-57``: ``deg = sum(adj_t, dim=1)``, infinities -> 0)."""
    w = edge_weight if edge_weight is not None else torch.ones(dst_index.numel(), dtype=dtype, device=dst_index.device)
    deg = torch.zeros(num_nodes, dtype=w.dtype, device=w.device).index_add_(0, dst_index, w)
    dis = deg.pow(-0.5)
    dis.masked_fill_(dis == float("inf"), 0.0)
    return dis.index_select(0, dst_index) * w * dis.index_select(0, src_index)


def _pad_rows(out: torch.Tensor, n: int) -> torch.Tensor:
    """geot ops return ``dst_index[-1] + 1`` rows; trailing isolated nodes get zero rows."""
    if out.shape[0] >= n:
        return out
    return torch.cat([out, out.new_zeros([n - out.shape[0]] + list(out.shape[1:]))], 0)


class GCNConv(nn.Module):
    """``GCNConv_GS`` (``models/conv/gcnconv.py``): ``lin`` without bias, aggregation, then ``+ bias``."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True):
        super().__init__()
        self.lin = nn.Linear(in_channels, out_channels, bias=False)
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None

    def forward(self, x, src_index, dst_index, norm_weight):
        h = self.lin(x)
        out = _pad_rows(gather_weight_scatter(src_index, dst_index, norm_weight, h), x.shape[0])
        return out + self.bias if self.bias is not None else out


class SAGEConv(nn.Module):
    """``SAGEConv_GS`` with ``aggr="sum"``, ``root_weight=True`` (``models/conv/sageconv.py:122-154``)."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.lin_l = nn.Linear(in_channels, out_channels, bias=True)
        self.lin_r = nn.Linear(in_channels, out_channels, bias=False)

    def forward(self, x, src_index, dst_index, reduce: str = "sum"):
        agg = _pad_rows(gather_scatter(src_index, dst_index, x, reduce), x.shape[0])
        return self.lin_l(agg) + self.lin_r(x)


class _Stack(nn.Module):
    def __init__(self, conv, in_channels, hidden_channels, num_layers, out_channels=None):
        super().__init__()
        out_channels = hidden_channels if out_channels is None else out_channels
        dims = [in_channels] + [hidden_channels] * (num_layers - 1) + [out_channels]
        self.convs = nn.ModuleList([conv(dims[i], dims[i + 1]) for i in range(num_layers)])

    def _run(self, x, *graph):
        for i, conv in enumerate(self.convs):
            x = conv(x, *graph)
            if i + 1 < len(self.convs):
                x = torch.relu(x)
        return x


class GCN(_Stack):
    """``GCN_GS`` (``models/gcn.py:26-33``): ``num_layers`` GCN layers, ReLU between them."""

    def __init__(self, in_channels, hidden_channels, num_layers, out_channels=None):
        super().__init__(GCNConv, in_channels, hidden_channels, num_layers, out_channels)

    def forward(self, x, src_index, dst_index, norm_weight):
        return self._run(x, src_index, dst_index, norm_weight)


class GraphSAGE(_Stack):
    """``GraphSAGE_GS`` (``models/graphsage.py:26-33``) with sum aggregation."""

    def __init__(self, in_channels, hidden_channels, num_layers, out_channels=None):
        super().__init__(SAGEConv, in_channels, hidden_channels, num_layers, out_channels)

    def forward(self, x, src_index, dst_index):
        return self._run(x, src_index, dst_index)


def forward_sharded(model: _Stack, x_local: torch.Tensor, shard, group=None, gather=None) -> torch.Tensor:
    """The same stack on a dst-row shard (``geot_b200.dist.GraphShard``): ``x_local`` are this rank's node
    rows; returns this rank's rows of the output.  Per layer: row-local GEMM(s) with replicated weights,
    the exchange of the src rows, aggregation into the local dst rows.  ``shard.weight`` carries the
    (globally normalised) GCN weights for a GCN stack and is ``None`` for GraphSAGE.

    ``gather``: a ``geot_b200.dist.BucketedGather`` built once for ``shard`` -- the exchange is then overlapped
    with the reduction (all-gather or needed-rows push transport, whichever the object was built with) and its
    per-graph state (buckets, plans, request lists) is shared by all layers.  ``None``: one all-gather per layer.
    Forward only (the overlapped exchange is not differentiable)."""
    from . import dist as gdist

    def aggregate(h):
        if gather is None:
            return gdist.sharded_gather_scatter(shard, h, "sum", group)
        return gather.aggregate(h.detach(), shard.weight, "sum")

    x = x_local
    for i, conv in enumerate(model.convs):
        if isinstance(conv, GCNConv):
            h = conv.lin(x)
            out = aggregate(h)
            if conv.bias is not None:
                out = out + conv.bias
        else:
            agg = aggregate(x)
            out = conv.lin_l(agg) + conv.lin_r(x)
        x = torch.relu(out) if i + 1 < len(model.convs) else out
    return x

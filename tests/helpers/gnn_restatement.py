"""Plain-torch restatement of the GCN / GraphSAGE stacks of ``geot_b200.gnn`` (test infrastructure, not product):
``index_select -> mul -> index_add_`` per layer -- the formula the reference's compile tests compare against
(``/root/reference/test/compile/test_gcn.py:31-52``, ``models/gcn.py:35-60``, ``models/graphsage.py:26-64``).  The
aggregation walks the edge list in chunks so that the [E, F] intermediate of a full BASELINE shape stays bounded."""
import torch


def aggregate(h, src_index, dst_index, weight=None, n_rows=None, chunk=4_000_000, acc_dtype=None):
    n = h.shape[0] if n_rows is None else n_rows
    out = torch.zeros(n, h.shape[1], dtype=acc_dtype or h.dtype, device=h.device)
    for e0 in range(0, src_index.numel(), chunk):
        rows = h.index_select(0, src_index[e0:e0 + chunk]).to(out.dtype)
        if weight is not None:
            rows = rows * weight[e0:e0 + chunk].unsqueeze(-1).to(out.dtype)
        out.index_add_(0, dst_index[e0:e0 + chunk], rows)
    return out.to(h.dtype)


def forward(model, x, src_index, dst_index, norm_weight=None, acc_dtype=None):
    """The stack ``model`` (``gnn.GCN`` / ``gnn.GraphSAGE``: only its Linear layers and biases are used) on the whole
    graph.  ``acc_dtype=torch.float64`` accumulates the aggregation in fp64 (the checker's precision)."""
    n_layers = len(model.convs)
    for i, conv in enumerate(model.convs):
        if hasattr(conv, "lin"):                                   # GCNConv: aggregate(norm * lin(x)) + bias
            out = aggregate(conv.lin(x), src_index, dst_index, norm_weight, acc_dtype=acc_dtype)
            if conv.bias is not None:
                out = out + conv.bias
        else:                                                      # SAGEConv: lin_l(sum of neighbours) + lin_r(x)
            out = conv.lin_l(aggregate(x, src_index, dst_index, None, acc_dtype=acc_dtype)) + conv.lin_r(x)
        x = torch.relu(out) if i + 1 < n_layers else out
    return x

"""FX rewriter: ``index_select -> (mul) -> index_add`` message passing => geot operators.

Counterpart of the reference's ``geot/match_replace`` (``match_replace.py:8-32``, ``fused_gs.py``,
``fused_gws.py``, ``fused_mh_spmm.py``): export the model with ``torch.export``, find every
``aten.index_add`` whose source is a gathered (and optionally weighted) feature matrix, and replace the chain
by one fused operator.  Differences from the reference, on purpose:

* the match is structural, per ``index_add`` node (its own ``index_select`` / ``mul`` operands), instead of
  "the last ``select`` / ``view`` node seen while walking the graph", so several aggregations with different
  edge lists in one model are rewritten correctly;
* the fused node keeps the row count of the ``zeros`` tensor it replaces (``geot::pad_rows``): the geot ops
  return ``index[-1] + 1`` rows, the reference's ``csr_gws`` returns ``nrow + 1`` (``csrc/csr_gws.cpp:29-31``),
  either of which changes shapes downstream when the last nodes are isolated;
* the COO operators are emitted directly -- the per-graph preprocessing the reference inserts as an explicit
  ``geot::coo_to_csr`` node (``format_transform.py:27-40``) is the extension's cached ``format_preprocess`` plan.

The destination index must be sorted (the geot contract, ``README.md:34``); unsorted input raises at run time.
"""
from typing import Optional, Tuple

import torch
import torch.fx

aten = torch.ops.aten


@torch.library.custom_op("geot::pad_rows", mutates_args=())
def pad_rows(x: torch.Tensor, rows: int) -> torch.Tensor:
    """``x`` with zero rows appended up to ``rows`` (a copy when nothing is missing: custom ops must not alias)."""
    if x.shape[0] >= rows:
        return x.clone()
    return torch.cat([x, x.new_zeros([rows - x.shape[0]] + list(x.shape[1:]))], 0)


@torch.library.register_fake("geot::pad_rows")
def _(x, rows):
    return x.new_empty([rows] + list(x.shape[1:]))


_ZEROS = (aten.new_zeros.default, aten.zeros.default, aten.zeros_like.default)


def _is_call(node, *targets) -> bool:
    return isinstance(node, torch.fx.Node) and node.op == "call_function" and node.target in targets


def _is_zeros(node) -> bool:
    if _is_call(node, *_ZEROS):
        return True
    if _is_call(node, aten.full.default, aten.new_full.default, aten.full_like.default):
        return node.args[-1] == 0 or node.args[-1] == 0.0
    return False


def _gathered(node) -> Optional[Tuple[torch.fx.Node, torch.fx.Node]]:
    """``index_select(x, 0, idx)`` (or dim -x.ndim) -> (x, idx)."""
    if not _is_call(node, aten.index_select.default):
        return None
    x, dim, idx = node.args[:3]
    ndim = len(x.meta["val"].shape) if "val" in x.meta else None
    if dim == 0 or (ndim is not None and dim == -ndim):
        return x, idx
    return None


def _edge_weight(node, feat_ndim: int) -> Optional[torch.fx.Node]:
    """The weight operand of ``mul``: ``w.view(-1, 1)`` / ``w.unsqueeze(-1)`` of a per-edge (2-D features) or
    per-edge-per-head (3-D features) weight.  Returns the un-reshaped weight node."""
    if _is_call(node, aten.unsqueeze.default) and node.args[1] in (-1, feat_ndim - 1):
        return node.args[0]
    if _is_call(node, aten.view.default, aten.reshape.default, aten._unsafe_view.default):
        shape = list(node.args[1])
        if len(shape) == feat_ndim and shape[-1] == 1:
            return node.args[0]
    return None


def rewrite_graph(gm: torch.fx.GraphModule) -> int:
    """Rewrites ``gm`` in place; returns the number of fused aggregations."""
    graph = gm.graph
    fused = 0
    for node in list(graph.nodes):
        if not _is_call(node, aten.index_add.default) or len(node.args) < 4 or node.kwargs.get("alpha", 1) != 1:
            continue
        base, dim, index, source = node.args[:4]
        val = node.meta.get("val")
        ndim = len(val.shape) if val is not None else None
        if not (dim == 0 or (ndim is not None and dim == -ndim)) or not _is_zeros(base) or ndim not in (2, 3):
            continue
        rows = val.shape[0]
        new = None
        with graph.inserting_before(node):
            g = _gathered(source)
            if g is not None and ndim == 2:
                x, src_idx = g
                new = graph.call_function(torch.ops.geot.gather_scatter.default, (src_idx, index, x))
            elif _is_call(source, aten.mul.Tensor):
                a, b = source.args[:2]
                for feat, wnode in ((a, b), (b, a)):
                    g = _gathered(feat)
                    w = _edge_weight(wnode, ndim) if g is not None else None
                    if w is None:
                        continue
                    x, src_idx = g
                    if ndim == 2:
                        new = graph.call_function(torch.ops.geot.gather_weight_scatter.default, (src_idx, index, w, x))
                    else:
                        new = graph.call_function(torch.ops.geot.mh_spmm.default, (src_idx, index, w, x, "sum"))
                    break
            if new is None:
                continue
            padded = graph.call_function(torch.ops.geot.pad_rows.default, (new, rows))
        if val is not None:
            padded.meta["val"] = val
            new.meta["val"] = val
        node.replace_all_uses_with(padded)
        graph.erase_node(node)
        fused += 1
    if fused:
        graph.eliminate_dead_code()
        graph.lint()
        gm.recompile()
    return fused


def pattern_transform(model: torch.nn.Module, args, **kwargs) -> torch.export.ExportedProgram:
    """``torch.export`` the model and fuse its message-passing chains (reference: ``pattern_transform``,
    ``geot/match_replace/match_replace.py:8-32``).  Run the result with ``exported.module()(*args)``."""
    exported = torch.export.export(model, args, **kwargs)
    rewrite_graph(exported.graph_module)
    return exported

"""``format_preprocess`` -- segment pointers and the edge-count partition of a sorted edge list.

In the reference, ``geot/format_preprocess.py`` is FlashSparse block-format code;
the name is reused here, as BASELINE.json's north_star does, for the preprocessing this
implementation actually needs: the CSR row pointer of the dst-sorted COO list
(== ``geot::coo_to_csr``, ``geot/match_replace/format_transform.py:5-18``; == the unique-key/offset
pass of ``csrc/cpu/index_scatter_cpu.cpp:36-75``), degree statistics, and edge-balanced dst-row shards
for multi-GPU runs.  Plans are cached inside the extension per index tensor, so calling the
operators directly costs one preprocessing pass per graph, not per call.
"""
from dataclasses import dataclass

import torch


@dataclass
class Plan:
    rowptr: torch.Tensor      # int64 [S+1] on the index's device
    num_edges: int
    num_rows: int             # S = dst_index[-1] + 1
    num_segments: int         # non-empty rows
    max_degree: int
    is_sorted: bool
    has_gaps: bool

    def shards(self, dst_index: torch.Tensor, parts: int):
        """(row_bounds, edge_bounds): python lists of ``parts+1`` ints, edge-balanced, cut at
        segment boundaries."""
        b = torch.ops.geot.plan_shards(dst_index, parts)
        return b[0].tolist(), b[1].tolist()

    @property
    def segment_offsets(self):
        """(row_index, row_offset) of the non-empty rows, the layout of the reference CPU kernel's
        ``row_index[] / row_index_offset[]`` (``index_scatter_cpu.cpp:62-75``)."""
        deg = self.rowptr[1:] - self.rowptr[:-1]
        rows = torch.nonzero(deg > 0).flatten()
        return rows, torch.cat([self.rowptr[rows], self.rowptr[-1:]])


def format_preprocess(dst_index: torch.Tensor) -> Plan:
    rowptr, stats = torch.ops.geot.format_preprocess(dst_index)
    e, s, nseg, maxdeg, is_sorted, has_gaps = stats.tolist()
    return Plan(rowptr, e, s, nseg, maxdeg, bool(is_sorted), bool(has_gaps))


def clear_plan_cache() -> None:
    torch.ops.geot.clear_plan_cache()

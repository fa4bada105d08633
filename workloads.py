"""Seeded synthetic inputs of the BASELINE.json shapes.  Bench / test support, not product.

There is no network for the real datasets, so every graph is synthetic with the named shape: node and
edge counts of Reddit / ogbn-products / ogbn-arxiv / ogbn-proteins, heavy-tailed in-degrees, edges
sorted lexicographically by (dst, src) as torch_sparse.SparseTensor yields them
(/root/reference/benchmark/utils.py:56-65).
"""
from dataclasses import dataclass
from typing import Optional

import torch

SHAPES = {
    # name: (nodes, edges, degree exponent)
    "reddit": (232_965, 114_615_892, 0.5),
    "products": (2_449_029, 61_859_140, 0.9),
    "arxiv": (169_343, 1_166_243, 0.7),
    "proteins": (132_534, 39_561_252, 0.4),
}


@dataclass
class Graph:
    name: str
    num_nodes: int
    src_index: torch.Tensor   # [E] int64
    dst_index: torch.Tensor   # [E] int64, non-decreasing
    max_degree: int
    degree_cv: float

    @property
    def num_edges(self) -> int:
        return self.dst_index.numel()


def power_law_graph(name: str, device="cuda", scale: float = 1.0, seed: int = 0, isolated: float = 0.0) -> Graph:
    """In-degree of node i proportional to (pi(i)+1)^-a (pi a seeded permutation), sources drawn from
    the same heavy-tailed distribution; `scale` shrinks nodes and edges together (tests).  `isolated` > 0: that
    fraction of the nodes (seeded choice, never the last node) receives NO edge -- dst rows the op must zero-fill;
    the other nodes share the shape's E by the same law."""
    n0, e0, a = SHAPES[name]
    N = max(2, int(round(n0 * scale)))
    E = max(N, int(round(e0 * scale)))
    g = torch.Generator(device=device).manual_seed(seed)
    perm = torch.randperm(N, generator=g, device=device)
    w = (perm.double() + 1.0).pow(-a)
    p = w / w.sum()
    if isolated > 0:                                     # in-degrees over the nodes that keep edges, same law
        gone = torch.rand(N, generator=g, device=device) < isolated
        gone[-1] = False                                 # the output keeps its N rows (S = dst[-1] + 1)
        wd = torch.where(gone, torch.zeros_like(w), w)
        deg = torch.floor(wd / wd.sum() * E).long().clamp_min(1)
        deg[gone] = 0
    else:
        deg = torch.floor(p * E).long().clamp_min(1)    # every node keeps >= 1 in-edge: no empty rows
    diff = E - int(deg.sum())
    top = torch.argmax(deg)
    deg[top] += diff                                     # remainder to the largest
    assert int(deg.sum()) == E and (isolated > 0 or int(deg.min()) >= 1)
    dst = torch.repeat_interleave(torch.arange(N, device=device), deg)
    cdf = torch.cumsum(p, 0)
    u = torch.rand(E, generator=g, device=device, dtype=torch.float64)
    src = torch.searchsorted(cdf, u).clamp_max(N - 1)
    del u, cdf
    key = dst * N + src                                  # lexicographic (dst, src)
    key, _ = torch.sort(key)
    dst = torch.div(key, N, rounding_mode="floor")
    src = key - dst * N
    del key
    degf = deg.double()
    return Graph(name, N, src.contiguous(), dst.contiguous(), int(deg.max()), float(degf.std() / degf.mean()))


def random_segments(E: int, S: int, device="cuda", seed: int = 0) -> torch.Tensor:
    """BASELINE config #1 index: S segments from S-1 distinct random cut points in [1, E)."""
    g = torch.Generator(device=device).manual_seed(seed)
    cuts = torch.randperm(E - 1, generator=g, device=device)[: S - 1] + 1
    cuts, _ = torch.sort(cuts)
    bounds = torch.cat([cuts.new_zeros(1), cuts, cuts.new_full((1,), E)])
    lens = bounds[1:] - bounds[:-1]
    return torch.repeat_interleave(torch.arange(S, device=device), lens)


def features(rows: int, width, dtype=torch.float32, device="cuda", seed: int = 1) -> torch.Tensor:
    g = torch.Generator(device=device).manual_seed(seed)
    shape = [rows] + (list(width) if isinstance(width, (tuple, list)) else [width])
    return torch.rand(shape, generator=g, device=device, dtype=torch.float32).to(dtype)


def edge_weights(E: int, heads: Optional[int] = None, dtype=torch.float32, device="cuda", seed: int = 2) -> torch.Tensor:
    g = torch.Generator(device=device).manual_seed(seed)
    shape = [E] if heads is None else [E, heads]
    return torch.rand(shape, generator=g, device=device, dtype=torch.float32).to(dtype)


# algorithmic ("logical") and compulsory bytes per call -- SURVEY.md 8d table
def bytes_logical(op: str, E: int, S: int, N: int, F: int, H: int = 1, s: int = 4) -> int:
    if op == "index_scatter":
        return E * F * s + 8 * E + S * F * s
    if op == "gather_scatter":
        return E * (F * s + 16) + S * F * s
    if op == "gather_weight_scatter":
        return E * (F * s + 16 + s) + S * F * s
    if op == "mh_spmm":
        return E * (H * F * s + 16 + H * s) + S * H * F * s
    raise ValueError(op)


def bytes_compulsory(op: str, E: int, S: int, N: int, F: int, H: int = 1, s: int = 4) -> int:
    if op == "index_scatter":
        return bytes_logical(op, E, S, N, F, H, s)
    if op == "gather_scatter":
        return 16 * E + N * F * s + S * F * s
    if op == "gather_weight_scatter":
        return E * (16 + s) + N * F * s + S * F * s
    if op == "mh_spmm":
        return E * (16 + H * s) + (N + S) * H * F * s
    raise ValueError(op)

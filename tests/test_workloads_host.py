"""The measurement inputs (workloads.py) against SURVEY.md 8(d): the algorithmic / compulsory byte formulas give the
table's figures for the BASELINE shapes, and the synthetic generators produce what bench.py says they do (sorted
(dst, src) order, exact edge / node counts, heavy-tailed degrees, config #1's segment structure)."""
import torch

import workloads as wl


def test_byte_formulas_match_survey_8d():
    GB = 1e9
    # (op, E, S, N, F, H, s) -> (logical GB, compulsory GB) as printed in SURVEY 8(d)
    table = [
        (("index_scatter", 1_000_000, 50_000, 0, 64, 1, 4), 0.277, 0.277),
        (("index_scatter", 114_615_892, 232_965, 0, 128, 1, 4), 59.7, 59.7),
        (("gather_weight_scatter", 114_615_892, 232_965, 232_965, 128, 1, 4), 61.10, 2.53),
        (("gather_scatter", 61_859_140, 2_449_029, 2_449_029, 64, 1, 4), 17.45, 2.24),
        (("gather_scatter", 61_859_140, 2_449_029, 2_449_029, 256, 1, 4), 66.84, 6.01),
        (("mh_spmm", 1_166_243, 169_343, 169_343, 32, 8, 2), 0.721, 0.211),
        (("gather_weight_scatter", 39_561_252, 132_534, 132_534, 256, 1, 4), 41.44, None),
    ]
    for args, logical, compulsory in table:
        assert abs(wl.bytes_logical(*args) / GB - logical) <= 0.006 * logical + 0.0006, args
        if compulsory is not None:
            assert abs(wl.bytes_compulsory(*args) / GB - compulsory) <= 0.006 * compulsory + 0.0006, args
    # the bench's headline denominator, to the byte
    assert wl.bytes_logical("gather_weight_scatter", 114_615_892, 232_965, 232_965, 128, 1, 4) == 61_094_932_624


def test_config1_segments():
    E, S = 100_000, 5_000
    idx = wl.random_segments(E, S, "cpu")
    assert idx.numel() == E and int(idx[0]) == 0 and int(idx[-1]) == S - 1
    assert bool((idx[1:] >= idx[:-1]).all())
    lens = torch.bincount(idx, minlength=S)
    assert int(lens.min()) >= 1 and abs(float(lens.float().mean()) - E / S) < 1e-6       # every segment non-empty
    assert torch.equal(idx, wl.random_segments(E, S, "cpu"))                               # seeded


def test_power_law_graph_shape_and_order():
    g = wl.power_law_graph("reddit", "cpu", scale=1 / 256)
    N, E = g.num_nodes, g.num_edges
    assert N == round(232_965 / 256) and E == round(114_615_892 / 256)
    assert g.src_index.numel() == E and g.dst_index.numel() == E
    assert int(g.src_index.min()) >= 0 and int(g.src_index.max()) < N and int(g.dst_index.max()) < N
    key = g.dst_index * N + g.src_index
    assert bool((key[1:] >= key[:-1]).all()), "edges are sorted by (dst, src), as SparseTensor gives them"
    deg = torch.bincount(g.dst_index, minlength=N)
    assert g.max_degree == int(deg.max()) and g.degree_cv > 0.5                            # heavy tail
    g2 = wl.power_law_graph("reddit", "cpu", scale=1 / 256)
    assert torch.equal(g.src_index, g2.src_index) and torch.equal(g.dst_index, g2.dst_index)


def test_power_law_graph_with_isolated_rows():
    """`isolated`: that fraction of the dst rows receives no edge (the gappy bench workload); E, N, the order and the
    last row are kept, and the default graphs are untouched by the option."""
    g = wl.power_law_graph("products", "cpu", scale=1 / 128, isolated=0.25)
    n0, e0, _ = wl.SHAPES["products"]
    assert g.num_nodes == round(n0 / 128) and g.num_edges == round(e0 / 128)
    deg = torch.bincount(g.dst_index, minlength=g.num_nodes)
    frac = float((deg == 0).float().mean())
    assert 0.2 < frac < 0.3 and int(g.dst_index[-1]) == g.num_nodes - 1
    assert bool((g.dst_index[1:] >= g.dst_index[:-1]).all())
    assert int(torch.bincount(wl.power_law_graph("products", "cpu", scale=1 / 128).dst_index).min()) >= 1


def test_every_default_bench_workload_has_an_ncu_traffic_figure():
    """roofline.traffic / frac_dram of the default `bench.py` line and of its `secondary` entries come from
    profiles/traffic.json (ncu dram bytes of the dominant kernel): no default workload may be left without one, and each
    figure must lie between the compulsory and the logical bytes of its workload."""
    import json
    import os
    import bench
    import workloads as wl
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    traffic = json.load(open(os.path.join(root, "profiles", "traffic.json")))
    shapes = {"reddit": (232_965, 114_615_892), "products": (2_449_029, 61_859_140), "arxiv": (169_343, 1_166_243)}
    for name in ("reddit_gws", "reddit_index_scatter", "config1_index_scatter", "products_gs64", "products_gs64_gaps",
                 "products_gs256", "arxiv_mh_spmm"):
        assert name in traffic and traffic[name]["dram_bytes_per_launch"] > 0, name
        assert bench.ncu_traffic(name) == traffic[name]["dram_bytes_per_launch"]
        gname, op, F, H, dtype = bench.WORKLOADS[name]
        es = 2 if "bf16" in str(dtype) or "float16" in str(dtype) else 4
        if gname == "config1":
            N, E, S = 0, 1_000_000, 50_000
        else:
            N, E = shapes[gname.split("+")[0]]
            S = N
        lo = wl.bytes_compulsory(op, E, S, N, F, H, es)
        hi = wl.bytes_logical(op, E, S, N, F, H, es)
        t = traffic[name]["dram_bytes_per_launch"]
        assert 0.85 * lo <= t <= 1.05 * hi, (name, lo, t, hi)

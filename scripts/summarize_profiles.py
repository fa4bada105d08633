"""Turns one GPU session's scratch output (gpurun_out/<tag>/) into the tracked evidence under profiles/.

    python scripts/summarize_profiles.py r01 [report-name ...]

Writes profiles/<tag>_launches.md (ncu launch list: per-kernel count / total / share of the step),
profiles/<tag>_<report>.md for every gpurun_out/<tag>/<report>.ncu-rep (key `ncu --set full` counters per
launch + the hottest source lines), copies the small JSON / text results, and refreshes
profiles/traffic.json (dram bytes per launch of the dominant kernel, read by bench.py).
Bench support, not product.
"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum",
    "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum", "lts__t_sectors_srcunit_tex_op_read_evict_first_lookup_miss.sum",
    "lts__t_sectors_srcunit_tex_op_read_evict_normal_lookup_hit.sum", "lts__t_sectors_srcunit_tex_op_read_evict_normal_lookup_miss.sum",
    "derived__lts__lts2xbar_bytes.sum.per_second", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum", "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum",
    "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__lsuin_requests.avg.pct_of_peak_sustained_elapsed", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
    "lts__lts2xbar_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def launches_md(tag, src, dst):
    rows = list(csv.reader(open(src)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = {"ns": v / 1e6, "us": v / 1e3, "ms": v, "s": v * 1e3}.get(r[ui], v / 1e6)
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0, 0.0])
        a[0] += 1; a[1] += v; a[2] = max(a[2], v)
    tot = sum(a[1] for a in agg.values())
    ours = sum(a[1] for k, a in agg.items() if "geot" in k or "rowptr" in k or "index_stats" in k or "degree_stats" in k)
    with open(dst, "w") as f:
        f.write("# %s: ncu launch list of `python bench.py --steps 3 --warmup 3`\n\n" % tag)
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` (cold-cache, serialised: compare SHARES).\n"
                "Includes workload generation (torch sort / searchsorted kernels), 6 device steps and the host-buffer e2e slices.\n\n")
        f.write("| kernel | launches | total ms | max ms | share |\n|---|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:14]:
            f.write("| `%s` | %d | %.3f | %.3f | %.1f%% |\n" % (k[:100], a[0], a[1], a[2], 100 * a[1] / tot))
        f.write("\nThis library's kernels: %.1f%% of all captured device time; `segment_reduce_kernel` : `segment_fixup_kernel` = " % (100 * ours / tot))
        sr = next((a for k, a in agg.items() if "segment_reduce_kernel" in k), None)
        fx = next((a for k, a in agg.items() if "segment_fixup_kernel" in k), None)
        if sr and fx:
            f.write("%.1f%% : %.1f%% of the step (main kernel share %.3f).\n" % (100 * sr[1] / (sr[1] + fx[1]), 100 * fx[1] / (sr[1] + fx[1]),
                                                                                  sr[1] / (sr[1] + fx[1])))


def report_md(tag, name, rep, dst, traffic):
    rows = ncu_csv(rep, "raw")
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    seen = 0
    with open(dst, "w") as f:
        f.write("# %s: `ncu --set full --clock-control none --import-source on` capture `%s`\n\n" % (tag, name))
        for r in rows[2:]:
            d = {k: (r[i], units[i]) for i, k in enumerate(hdr)}
            f.write("## `%s`\n\n| counter | value |\n|---|---|\n" % d["Kernel Name"][0])
            for k in KEYS:
                if k in d and d[k][0] != "":
                    f.write("| %s | %s %s |\n" % (k, d[k][0], d[k][1]))
            stalls = sorted(((float(v[0]), k) for k, v in d.items()
                             if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v[0]),
                            reverse=True)
            if not stalls:
                stalls = sorted(((float(v[0]), k) for k, v in d.items()
                                 if "issue_stalled" in k and k.endswith("per_warp_active.pct") and v[0]), reverse=True)
            f.write("\nTop warp stall reasons: " + ", ".join("%s %.2f" % (k.split("issue_stalled_")[1].split("_per_")[0], v) for v, k in stalls[:6]) + "\n\n")
            try:
                rd = float(d["dram__bytes_read.sum"][0]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[d["dram__bytes_read.sum"][1]]
                wr = float(d["dram__bytes_write.sum"][0]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[d["dram__bytes_write.sum"][1]]
                # a capture holds the main-kernel launches of ONE step (one per src block): their bytes add up
                t = traffic.get(name) if seen else None
                seen += 1
                traffic[name] = {"kernel": d["Kernel Name"][0], "launches": seen,
                                 "dram_bytes_per_launch": int(rd + wr) + (t["dram_bytes_per_launch"] if t else 0),
                                 "dram_read": int(rd) + (t["dram_read"] if t else 0),
                                 "dram_write": int(wr) + (t["dram_write"] if t else 0),
                                 "source": "profiles/%s_%s.md" % (tag, name)}
            except Exception:
                pass
        # hottest source lines (needs -lineinfo)
        src = ncu_csv(rep, "source")
        try:
            h = next(i for i, r in enumerate(src) if "Source" in r and any("Sampl" in c for c in r))
            hdr2 = src[h]
            si = hdr2.index("Source")
            ci = next(i for i, c in enumerate(hdr2) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)")
            hot = []
            for r in src[h + 1:]:
                if r and r[0] == "Kernel Name":      # next launch: keep the first only
                    break
                try:
                    hot.append((float(r[ci]), r[si].strip()))
                except Exception:
                    continue
            hot.sort(reverse=True)
            tot = sum(v for v, _ in hot) or 1
            f.write("## Hottest lines (warp-stall samples, first launch)\n\n| share | line |\n|---:|---|\n")
            for v, s in hot[:12]:
                f.write("| %.1f%% | `%s` |\n" % (100 * v / tot, s[:140].replace("|", "\\|")))
        except StopIteration:
            pass


def main():
    tag = sys.argv[1]
    src = os.path.join(ROOT, "gpurun_out", tag)
    dst = os.path.join(ROOT, "profiles")
    os.makedirs(dst, exist_ok=True)
    if os.path.exists(os.path.join(src, "launches.csv")):
        launches_md(tag, os.path.join(src, "launches.csv"), os.path.join(dst, tag + "_launches.md"))
    tpath = os.path.join(dst, "traffic.json")
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    for fn in sorted(os.listdir(src)):
        if fn.endswith(".ncu-rep"):
            name = fn[:-8]
            report_md(tag, name, os.path.join(src, fn), os.path.join(dst, "%s_%s.md" % (tag, name)), traffic)
    # bench.py looks the workload name up: map capture names prof_<x> -> workload names
    alias = {"prof_gws": "reddit_gws", "prof_reddit_gws": "reddit_gws", "prof_reddit_index_scatter": "reddit_index_scatter",
             "prof_products_gs64": "products_gs64", "prof_products_gs256": "products_gs256"}
    for k, v in sorted(traffic.items()):
        if k in alias:
            traffic[alias[k]] = v
        elif k.startswith("prof_"):          # every other capture is named prof_<workload>
            traffic[k[5:]] = v
    json.dump(traffic, open(tpath, "w"), indent=1, sort_keys=True)
    for fn in os.listdir(src):
        if fn.endswith((".json", ".jsonl", ".txt")) and os.path.getsize(os.path.join(src, fn)) < 200_000:
            shutil.copy(os.path.join(src, fn), os.path.join(dst, "%s_%s" % (tag, fn)))
    print("profiles/:", sorted(os.listdir(dst)))


if __name__ == "__main__":
    main()

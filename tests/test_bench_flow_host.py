"""bench.py's host logic on CPU: the JSON contract at N = 1 and the sharded flow at N = 2 (gloo), with the CUDA
runtime and C-ABI calls replaced by stand-ins (tests/helpers/bench_emulation.py: the oracle plays the kernels).
Numbers are meaningless here; keys, shapes and the collective sequence of every rank are what is checked."""
import os
import socket
import subprocess
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

CONTRACT = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "parity", "secondary"]


def test_single_gpu_line_contract():
    # a fresh interpreter: the emulation monkeypatches torch.cuda and must not leak into the other tests
    code = ("import sys, json; sys.path.insert(0, %r); sys.path.insert(0, %r); from helpers import bench_emulation as be; "
            "be.install(); print(json.dumps(be.run_own('config1_index_scatter')))" % (ROOT, os.path.join(ROOT, "tests")))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert [k for k in CONTRACT if k not in line] == []
    assert line["n_gpus"] == 1 and line["higher_is_better"] is True and line["vs_baseline"] is None and line["dtype"] == "f32"
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic", "frac_logical", "frac_dram", "frac_compulsory"} <= set(line["roofline"])
    assert line["parity"]["ok"] is True and line["parity"]["rows_checked"] > 3
    assert line["e2e"]["matches_device_result"] is True and "stateless" in line["e2e"] and "resident" in line["e2e"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(line["cpu_baseline"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])
    assert "workload" in line["config"] and line["gpu_launches"] > 0


@pytest.mark.parametrize("exchange", ["bucket", "allgather", "replicated"])
def test_two_rank_flow_gloo(exchange):
    from helpers import bench_emulation as be
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=be.worker, args=(r, 2, port, exchange, "arxiv_mh_spmm", q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
    codes = [p.exitcode for p in procs]
    for p in procs:
        if p.is_alive():
            p.kill()
    assert codes == [0, 0], codes
    line = q.get(timeout=5)
    assert [k for k in CONTRACT if k not in line] == []
    assert line["n_gpus"] == 2 and line["exchange"] == exchange and line["scaling"] == "strong"
    # the N > 1 end-to-end leg ran on both ranks: bytes are summed over ranks (src rows + weights in, dst rows out)
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0 and "ms_per_step" in line["e2e"]
    assert line["parity"]["ok"] is True and line["parity"]["rows_checked"] > 6        # both ranks' rows
    if exchange == "bucket":
        got, full = line["config"]["src_rows_received_per_step_rank0"], line["config"]["src_rows_full_exchange_rank0"]
        assert 0 < got == full

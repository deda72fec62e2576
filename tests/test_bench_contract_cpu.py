"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`
times the oracle port on the host cores) prints exactly one JSON line with the keys
the driver reads, single-process and under torchrun (rank 0 alone prints)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REQUIRED = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
            "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline",
            "e2e"}


def _json_lines(text):
    out = []
    for ln in text.splitlines():
        ln = ln.strip()
        if ln.startswith("{") and ln.endswith("}"):
            try:
                out.append(json.loads(ln))
            except json.JSONDecodeError:
                pass
    return out


def _check(line, n_gpus, steps):
    assert REQUIRED <= set(line), REQUIRED - set(line)
    assert line["impl"] == "reference" and line["n_gpus"] == n_gpus and line["steps"] == steps
    assert line["metric"] == "fp64 DG grid-point RHS updates/sec"
    assert line["unit"] == "grid-point-updates/s" and line["higher_is_better"] is True
    assert line["dtype"] == "f64" and line["vs_baseline"] is None
    assert line["value"] > 0 and line["ms_per_step"] > 0
    assert "workload" in line["config"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"]
    assert cb["value"] == line["value"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["unit"] == line["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "2", "--warmup", "1", "--cpu-sample-refine", "0"],
                         capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = _json_lines(out.stdout)
    assert len(lines) == 1
    _check(lines[0], 1, 2)


def _free_port():
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_reference_arm_under_torchrun_prints_on_rank_zero_only():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                          "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
                          str(_free_port()), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "1", "--cpu-sample-refine", "0"],
                         capture_output=True, text=True, cwd=ROOT, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = _json_lines(out.stdout)
    assert len(lines) == 1
    _check(lines[0], 2, 1)
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the arm must still use every
    # host thread (round-1 defect: 1-thread CPU numbers at N >= 2)
    assert lines[0]["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert "Kerr-Schild shell" in lines[0]["cpu_baseline"]["sample"]

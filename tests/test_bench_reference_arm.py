"""The reference arm of bench.py (`--impl reference`: the CPU oracle timed on the host cores) prints the contract's JSON
line — same metric string construction as the device arm, `impl`, `cpu_baseline`, `e2e` with zero copy bytes — and, under
torchrun with two ranks, only rank 0 prints (the other exits 0 without work).  Small node counts: CPU only."""
import json
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
        "data", "config", "cpu_baseline", "e2e"}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _lines(text):
    return [json.loads(l) for l in text.splitlines() if l.startswith("{")]


def test_reference_arm_line_gsf():
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "3", "--warmup", "1", "--nodes", "1024",
                          "--cpu-max-nodes", "1024"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    (line,) = _lines(out.stdout)
    assert KEYS <= set(line) and line["impl"] == "reference"
    assert line["metric"] == "simulated-ms/sec, GSFSignature 1,024 nodes" and line["unit"] == "simulated-ms/s"
    assert line["steps"] == 3 and line["warmup"] == 1 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["value"] > 0 and abs(line["e2e"]["value"] - line["value"]) < 1e-9
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 1 and "workload" in line["config"]


def test_reference_arm_under_torchrun_prints_once():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1",
           "--nodes", "512", "--cpu-max-nodes", "512"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-1000:] + out.stderr[-2000:]
    lines = _lines(out.stdout)
    assert len(lines) == 1 and lines[0]["impl"] == "reference" and lines[0]["n_gpus"] == 2

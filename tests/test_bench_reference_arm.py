"""`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) on a small sample: runs here
without a GPU and prints ONE JSON line with the keys of the contract -- same metric / unit / config shape as
the GPU arm, `impl`, a `cpu_baseline` describing this very run and an `e2e` with no copies."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--blocks", "8", "--block-kib", "64"], capture_output=True, timeout=300,
                         cwd=ROOT)
    assert out.returncode == 0, out.stderr.decode()[-2000:]
    lines = [l for l in out.stdout.decode().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference" and d["metric"] == base["metric"] and d["unit"] == "MB/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1
    assert d["steps"] == 2 and d["warmup"] == 1 and d["ms_per_step"] > 0 and d["value"] > 0
    assert d["dtype"] == "int32" and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # a round trip on the CPU is slower than either direction alone
    assert d["value"] < min(d["compress_mb_s"], d["decompress_mb_s"])

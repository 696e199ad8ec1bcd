"""bench.py's contract, as far as it can be checked without a GPU: BASELINE.json's configuration per GPU
count (the B200 arm and the reference arm print the same `config`), and the reference arm's JSON line
(the reference's own CPU chain, timed on the host cores through oracle/_ref or the C port)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_baseline_configurations_per_gpu_count():
    import bench
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        base = json.load(f)
    assert "aggregate IQ Msamples/s" in base["metric"] and bench.METRIC == "aggregate_iq_msamples_per_s"
    cfgs = " | ".join(base["configs"])
    assert "1024" in cfgs and "8192" in cfgs and "16384" in cfgs and "65536" in cfgs
    assert bench.BASELINE_CONFIGS[1][:2] == ("am", 1024)          # the headline: AM x1024 on one GPU
    assert bench.BASELINE_CONFIGS[2][:2] == ("ssb", 16384) and bench.BASELINE_CONFIGS[4][:2] == ("ssb", 16384)
    assert bench.BASELINE_CONFIGS[8][:2] == ("mixed", 65536)
    for n, (wl, total, scaling) in bench.BASELINE_CONFIGS.items():
        assert total % n == 0 and scaling in ("weak", "strong")
        assert wl in bench.KERNELS
    cfg = bench.config_of("ssb", "x", "tone", 8192, 2, 2)
    assert cfg["channels_total"] == 16384 and cfg["channels_per_gpu"] == 8192
    assert cfg["iq_bytes_per_gpu_per_step"] == 8192 * 2 * bench.BLOCK_BYTES
    # one step's input must exceed the 126 MB L2 for every workload's default size
    from rtlsdrdiags_b200 import synth
    for wl in ("am", "fm", "wbfm", "ssb", "mixed"):
        ch = synth.WORKLOADS[wl][0]
        assert ch * bench.default_blocks(wl, ch) * bench.BLOCK_BYTES >= 512 << 20


@pytest.mark.parametrize("n_gpus,workload,total", [(1, "am", 1024), (8, "mixed", 65536)])
def test_reference_arm_line(n_gpus, workload, total):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", str(n_gpus),
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "aggregate_iq_msamples_per_s"
    assert line["unit"] == "Msamples/s" and line["higher_is_better"] is True and line["n_gpus"] == n_gpus
    assert line["config"]["workload"] == workload and line["config"]["channels_total"] == total
    assert line["value"] > 0 and line["gpu_launches"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}

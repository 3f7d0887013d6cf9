"""bench.py contract checks.  Without a GPU: the reference arm (`--impl reference`) runs the
unmodified reference binary on the host cores and prints ONE JSON line with the keys the driver
reads; under torchrun only rank 0 works.  With a GPU: the product arm's line at a reduced size."""
import json
import os
import subprocess
import sys

import pytest

from oracle import refrun

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TINY = ["--ref-mesh", "32", "--ref-block", "16", "--ref-cycles", "2"]


def run_bench(extra, env=None):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + extra, capture_output=True, text=True,
                       timeout=600, cwd=ROOT, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.strip()]


@pytest.mark.skipif(not refrun.have_ref("blast_mhd_hlld"), reason="oracle/_ref/athena_blast_mhd_hlld not built")
def test_reference_arm_prints_one_contract_line():
    lines = run_bench(["--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1"] + TINY)
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "zone-cycles/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 2 and d["n_gpus"] == 1
    assert d["warmup"] == 3      # the timing rules ask for at least three warm-up steps; bench.py clamps
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and "32^3" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "zone-cycles/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert abs(d["ms_per_step"] - 1e3 * 32 ** 3 * 2 / d["value"]) < 1e-6 * d["ms_per_step"]


@pytest.mark.skipif(not refrun.have_ref("blast_mhd_hlld"), reason="oracle/_ref/athena_blast_mhd_hlld not built")
def test_reference_arm_other_ranks_exit_quietly():
    lines = run_bench(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"] + TINY,
                      env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert lines == []


@pytest.mark.gpu
def test_product_arm_line_has_every_contract_key(cuda_lib):
    lines = run_bench(["--gpus", "1", "--steps", "3", "--warmup", "3", "--per-gpu", "128", "--block", "64",
                       "--no-cpu-baseline"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["unit"] == "zone-cycles/s" and d["dtype"] == "f64" and d["scaling"] == "weak" and d["steps"] == 3
    assert d["value"] > 1e8 and d["gpu_launches"] > 0      # a CPU fallback could not reach this
    assert abs(d["value"] - 128 ** 3 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert 0 < e["value"] < d["value"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    # the e2e number has several independent sets of MeshBlocks in flight; the one-set number stays beside it
    assert e["pipeline"]["sets_in_flight"] == 3 and e["pipeline"]["added_sets_bit_identical"] is True
    assert 0 < e["one_set_in_sequence"]["value"] < d["value"]
    assert "workload" in d["config"] and "sm_mhz" in d["clocks"]

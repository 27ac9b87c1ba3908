"""
CPU tests of the driver-facing contracts that do not need a GPU:

  * `bench.py --impl reference` (the reference arm: the oracle's C port on the host cores)
    prints ONE JSON line with the keys the driver reads;
  * the product arm has no CPU fallback: without a CUDA device `bench.py` and every compute
    entry point fail loudly;
  * `kpal_set_option` knows every switch bench.py / the tests use and rejects bad values
    (host only).
"""
import json
import os
import subprocess
import sys

import pytest

from kpal_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_bench(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args),
                          capture_output=True, text=True, cwd=ROOT, timeout=900)


def test_reference_arm_prints_the_contract_line():
    result = _run_bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert result.returncode == 0, result.stderr[-2000:]
    lines = [line for line in result.stdout.splitlines() if line.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "gbases_per_sec_counted_k12"
    assert line["unit"] == "Gbases/s" and line["higher_is_better"] is True and line["n_gpus"] == 1
    assert line["steps"] == 1 and line["value"] > 0 and line["ms_per_step"] > 0
    assert line["gpu_launches"] == 0 and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] and "sample" in line["cpu_baseline"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"],
                           "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_runs_on_rank_zero_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    result = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                             "--steps", "1", "--warmup", "0"], capture_output=True, text=True, cwd=ROOT,
                            timeout=300, env=env)
    assert result.returncode == 0 and result.stdout.strip() == ""


@pytest.mark.skipif(_cabi.device_count() > 0, reason="a GPU is present")
def test_product_arm_has_no_cpu_fallback():
    result = _run_bench("--steps", "1", "--warmup", "0")
    assert result.returncode != 0
    assert not [line for line in result.stdout.splitlines() if line.startswith("{")]


def test_set_option_knows_its_switches():
    L = _cabi.load()
    good = {"host_fasta": (0, 1), "fasta_chunks": (0, 16, 32), "fasta_split": (0, 1), "exact_div": (0, 1),
            "narrow_d2h": (0, 1, 2), "dma_share": (0, 3, 8), "count_path": (0, 1, 2, 3), "pair_flush_every": (0, 1, 6), "pair_fused": (0, 1), "gram": (0, 1), "narrow_lists": (0, 1), "tiled_finalize": (0, 1),
            "by_record_path": (0, 1), "radix_shape": (0, 1, 2), "radix_max_buckets": (1024, 2048),
            "radix_payload_bits": (0, 13, 15), "radix_debug": (0,)}
    defaults = {"host_fasta": 0, "fasta_chunks": 0, "fasta_split": 0, "exact_div": 0, "narrow_d2h": 1,
                "dma_share": 0, "count_path": 0, "pair_flush_every": 0, "pair_fused": 1, "gram": 1, "narrow_lists": 1, "tiled_finalize": 1, "by_record_path": 0, "radix_shape": 0,
                "radix_max_buckets": 2048, "radix_payload_bits": 0, "radix_debug": 0}
    bad = {"fasta_chunks": (-1, 33), "narrow_d2h": (-1, 3), "dma_share": (-1, 9), "count_path": (4,), "pair_flush_every": (-1, 7),
           "by_record_path": (2,), "radix_shape": (3,), "radix_max_buckets": (512, 4096), "radix_payload_bits": (16,)}
    try:
        for name, values in good.items():
            for value in values:
                assert L.kpal_set_option(name.encode(), value) == _cabi.KPAL_OK, (name, value)
        for name, values in bad.items():
            for value in values:
                assert L.kpal_set_option(name.encode(), value) == _cabi.KPAL_EINVAL, (name, value)
                assert name.split("_")[0] in L.kpal_last_error().decode()
        assert L.kpal_set_option(b"no_such_option", 1) == _cabi.KPAL_EINVAL
        assert L.kpal_set_option(None, 1) == _cabi.KPAL_EINVAL
    finally:
        for name, value in defaults.items():
            L.kpal_set_option(name.encode(), value)

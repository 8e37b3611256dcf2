"""CPU: the bench.py contract pieces that do not need a GPU - the reference arm's JSON line, the roofline
helpers and the committed ncu summary that feeds `roofline.traffic`."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` times the reference's CPU implementation of the path (here: the oracle port
    of its ATen calls) on a bounded sample of the product arm's workload and prints ONE JSON line."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["gpu_launches"] == 0
    assert d["metric"].startswith("rotation hypotheses scored/sec") and d["unit"] == "hyp*pairs/s"
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["config"]["pairs"] == 32 and d["config"]["hypotheses_per_gpu"] == 50000      # BASELINE configs[1]
    assert "BASELINE.json configs[1]" in d["config"]["workload"]


def test_roofline_inputs():
    import bench

    # algorithmic shared-memory bytes per (pair, hypothesis): 512 voxels x 8 taps x 16 channels x 4 B
    assert bench.GATHER_BYTES_PER_HYP == 512 * 8 * 16 * 4
    peaks = bench.load_peaks()
    assert peaks["hbm_gbs"] > 1000 and peaks["bf16_tflops"] > 100 and "source" in peaks
    traffic = bench.ncu_traffic()          # DRAM bytes per launch of the score kernel from the committed ncu capture
    assert traffic is not None and 1e5 < traffic < 1e9

"""CPU: `bench.py --impl reference` (the unmodified reference decoder timed on the host cores) needs no GPU; run it on the smallest
workload and check the contract of the line it prints."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    if not os.access(os.path.join(ROOT, "oracle", "_ref", "ref_driver"), os.X_OK):
        pytest.skip("oracle/_ref/ref_driver is not built (needs /root/reference at build time)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-800:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "decode Mpixels/s (bit-exact)" and line["unit"] == "Mpx/s"
    assert line["higher_is_better"] is True and line["steps"] == 1 and line["n_gpus"] == 1
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1

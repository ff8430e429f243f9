"""GPU: fb_encode() -- the MANIAC encoder kernel + the host container code -- through the C ABI against the oracle encoder
(byte-exact against the reference's files, tests/test_oracle_encoder.py) and against the reference's own golden files.

Named to run last and executed in ONE child process (tests/gpu_encode_child.py): the kernel was validated under the CPU
execution-model emulator only (tests/test_emu_maniac_enc.py) because the round's GPU time had been used up when it was
written, so its first run on hardware is this test."""
import json
import os
import subprocess
import sys

import pytest

from tests.cases import CASES

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = [c[0] for c in CASES] + ["approx", "approx_nosq", "pal", "pal_nosq", "noise", "synth256", "synth512", "synth1000x333"]
_results = None


def results():
    global _results
    if _results is None:
        _results = {}
        try:
            p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "gpu_encode_child.py"), *NAMES], capture_output=True, text=True, timeout=420, cwd=ROOT)
            out, err, rc = p.stdout, p.stderr, p.returncode
        except subprocess.TimeoutExpired as e:
            out = e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
            err, rc = "timed out after 420 s", -1
        for line in out.splitlines():
            if line.startswith("{"):
                r = json.loads(line)
                _results[r["case"]] = r
        _results["__process__"] = {"rc": rc, "stderr": err[-2000:]}
    return _results


@pytest.mark.parametrize("name", NAMES)
def test_encode_matches_oracle_encoder(name):
    res = results()
    r = res.get(name)
    assert r is not None, f"the encode process ended before this case: {res['__process__']}"
    assert r["ok"], r

"""GPU: the throughput shape of the MANIAC kernel (8 one-warp streams per SM, no walkers; taken by itself for batches with at
least 16 streams per SM, forced here through FB_MANIAC_SPB in a child process) on a batch of several hundred small files with
every kind of group in them -- all predictors, multi-channel groups, uncompressed groups, 14-bit planes -- against the golden
planes of the reference."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("spb", ["8", "4"])
def test_many_streams_per_sm_batch(spb):
    env = dict(os.environ, FB_MANIAC_SPB=spb)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "gpu_batch_many_child.py"), "sq128,odd,rgba14,dct,pred,unc,gray,e0", "40"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-800:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["streams"] >= 148 * int(spb), res        # enough streams for the shape to be taken
    assert res["bad"] == 0, (res, r.stderr[-600:])

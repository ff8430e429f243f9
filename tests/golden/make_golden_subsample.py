"""Regenerates tests/golden/sub_*.npz from the UNMODIFIED reference (oracle/_ref/ref_driver subsample ...).  Run in the build
container only:  python tests/golden/make_golden_subsample.py

The reference has no forward chroma subsampling (transform/subsample.h:130-133 is a stub: its subsampled inputs are JPEGs),
so the driver decimates the chroma planes itself, pushes TRANSFORM_ChromaSubsample and lets the reference undo it.
Each sub_<case>.npz holds: b (plane dump before inv_subsample), a (after), and for ratios the bitstream allows (1 or 2,
meta_subsample asserts it) fuif: the subsampled image encoded by the reference encoder.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fuif_b200.synth import synth_image, write_pnm  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from tests.cases import SUBSAMPLE_CASES  # noqa: E402


def main():
    po.build()
    assert po.have_ref(), "oracle/_ref/ref_driver missing (needs /root/reference)"
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, w, h, c, maxval, seed, params, with_file in SUBSAMPLE_CASES:
        with tempfile.TemporaryDirectory() as td:
            pnm = os.path.join(td, "in.pnm")
            write_pnm(pnm, synth_image(w, h, c, maxval, seed), maxval)
            po.ref_run("subsample", pnm, os.path.join(td, "d"), ",".join(str(p) for p in params), *(["-F"] if with_file else []))
            blob = {k: np.frombuffer(open(os.path.join(td, f"d.{k}.fbpd"), "rb").read(), dtype=np.uint8) for k in ("b", "a")}
            if with_file:
                blob["fuif"] = np.frombuffer(open(os.path.join(td, "d.fuif"), "rb").read(), dtype=np.uint8)
            np.savez_compressed(os.path.join(out_dir, "sub_" + name + ".npz"), **blob)
            print(name, {k: v.size for k, v in blob.items()})


if __name__ == "__main__":
    main()

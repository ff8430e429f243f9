"""Regenerates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/ref_driver, built by oracle/Makefile from
/root/reference).  Run in the build container only:  python tests/golden/make_golden.py

Each <case>.npz holds: pnm (input file bytes), fuif (reference-encoded file), s0..sN (plane dumps after fuif_decode and
after every single inverse transform, final one clamped), f0..fM (plane dumps after every forward transform and after
fuif_prepare_encode), r0..r4 (final pixels of the responsive decodes -R 0..4 as FBPD dumps).
"""
import glob
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fuif_b200.synth import synth_image, write_pnm  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from tests.cases import APPROX_CASES, CASES, MATCH_CASES, PALETTE_CASES, PERMUTE_CASES  # noqa: E402


def main():
    po.build()
    assert po.have_ref(), "oracle/_ref/ref_driver missing (needs /root/reference)"
    out_dir = os.path.dirname(os.path.abspath(__file__))
    which = APPROX_CASES if "approx" in sys.argv[1:] else PALETTE_CASES if "palette" in sys.argv[1:] else PERMUTE_CASES if "permute" in sys.argv[1:] else MATCH_CASES if "match" in sys.argv[1:] else CASES    # python make_golden.py [approx|palette|permute|match]
    for name, w, h, c, maxval, seed, opts in which:
        with tempfile.TemporaryDirectory() as td:
            pnm = os.path.join(td, "in.pnm")
            fuif = os.path.join(td, "x.fuif")
            pix = synth_image(w, h, c, maxval, seed)
            if name.startswith("pal"):
                pix = (pix // 64) * 64 + 17         # four levels per channel: few colours
            if name.startswith("match"):            # a noisy patch, repeated: something for the matching heuristic to find
                rng = np.random.default_rng(seed)
                patch = rng.integers(0, maxval + 1, size=(h // 2 + 3, w // 3 + 2, c)).astype(np.int32)
                pix = np.tile(patch, (2, 3, 1))[:h, :w, :]
            write_pnm(pnm, pix, maxval)
            po.ref_run("encode", pnm, fuif, *opts)
            po.ref_run("dump", fuif, os.path.join(td, "d"))
            po.ref_run("fwd", pnm, os.path.join(td, "d"), *opts)
            blob = {"pnm": np.frombuffer(open(pnm, "rb").read(), dtype=np.uint8), "fuif": np.frombuffer(open(fuif, "rb").read(), dtype=np.uint8)}
            for f in glob.glob(os.path.join(td, "d.*.fbpd")):
                key = f.split(".")[-2]
                blob[key] = np.frombuffer(open(f, "rb").read(), dtype=np.uint8)
            for r in range(5):
                for f in glob.glob(os.path.join(td, "r.*.fbpd")):
                    os.remove(f)
                po.ref_run("dump", fuif, os.path.join(td, "r"), "-R", str(r))
                dumps = sorted(glob.glob(os.path.join(td, "r.s*.fbpd")), key=lambda s: int(s.split(".")[-2][1:]))
                blob[f"r{r}"] = np.frombuffer(open(dumps[-1], "rb").read(), dtype=np.uint8)
                blob[f"r{r}s0"] = np.frombuffer(open(dumps[0], "rb").read(), dtype=np.uint8)
            np.savez_compressed(os.path.join(out_dir, name + ".npz"), **blob)
            print(name, {k: v.size for k, v in blob.items() if k in ("fuif",)}, len(blob), "entries")


if __name__ == "__main__":
    main()

"""Debug aid: packed unsqueeze on one shape, several runs; where do wrong samples sit?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fuif_b200 import api
from fuif_b200.synth import synth_image
from tests.util import default_squeeze_parameters
w, h, c = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
pix = synth_image(w, h, c, 255, seed=5 * w + h)
ctx = api.Context(0)
src = api.Image.from_pixels(pix, 255, ctx)
if c >= 3:
    src.do_transform(api.Transform(1))
src.do_transform(api.Transform(7, default_squeeze_parameters(w, h, c)))
inf = src.info(); planes = src.channels(); trs = src.transform
for run in range(4):
    g = api.Image.from_planes(inf.w, inf.h, inf.minval, inf.maxval, inf.nb_channels, inf.real_nb_channels, inf.nb_meta_channels, inf.colormodel, planes, trs, ctx)
    g.undo_transforms(0)
    out = g.pixels()
    bad = np.argwhere(out != pix)
    print("run", run, "wrong samples", len(bad), "repaired", ctx.pk_repaired, "flagged", ctx.pk_range_flagged)
    if len(bad):
        xs = bad[:, 1]
        print("  x mod 64 histogram:", np.bincount(xs % 64, minlength=64).tolist())
        print("  rows mod 64 histogram:", np.bincount(bad[:, 0] % 64, minlength=64).tolist())
        print("  first 10:", bad[:10].tolist())

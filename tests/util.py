"""Helpers shared by the tests: golden fixtures and comparisons between the CUDA path and oracle dumps."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: z[k].tobytes() for k in z.files}


def ordered(blob, prefix):
    """s0, s1, ... (or f0, f1, ...) dumps of a golden case in order."""
    keys = sorted((k for k in blob if k.startswith(prefix) and k[len(prefix):].isdigit()), key=lambda k: int(k[len(prefix):]))
    return [blob[k] for k in keys]


def gpu_plane_image(po, img):
    """fuif_b200.api.Image -> oracle.pyoracle.PlaneImage (downloads every plane)."""
    inf = img.info()
    pi = po.PlaneImage(inf.w, inf.h, inf.minval, inf.maxval, inf.nb_channels, inf.real_nb_channels, inf.nb_meta_channels, inf.colormodel)
    for i in range(inf.nb_planes):
        c = img.channel(i)
        pi.planes.append(po.Plane(c.w, c.h, c.minval, c.maxval, c.zero, c.q, c.hshift, c.vshift, c.hcshift, c.vcshift, c.component, c.data))
    pi.transforms = [(t.ID, list(t.parameters)) for t in img.transform]
    return pi


def upload_plane_image(api, pi, ctx):
    planes = [api.Channel(p.w, p.h, p.minval, p.maxval, p.zero, p.q, p.hshift, p.vshift, p.hcshift, p.vcshift, p.component, p.data) for p in pi.planes]
    trs = [api.Transform(t, ps) for t, ps in pi.transforms]
    return api.Image.from_planes(pi.w, pi.h, pi.minval, pi.maxval, pi.nb_channels, pi.real_nb_channels, pi.nb_meta_channels, pi.colormodel, planes, trs, ctx)


def default_squeeze_parameters(w, h, nb_channels, nb_meta=0, chroma_same_size=True):
    """The schedule meta_squeeze fills in at decode time (reference transform/squeeze.h:266-321).  An Image that was
    squeezed with empty parameters can only be un-squeezed in memory when the schedule is passed explicitly: the
    reference derives defaults from the *current* channel sizes (squeeze.h:364-365), which have changed by then."""
    p = []
    if nb_channels > 2 and chroma_same_size:
        p += [3, nb_meta + 1, nb_meta + 2, 2, nb_meta + 1, nb_meta + 2]
    if not (w > h) and h > 8:
        p += [0, nb_meta, nb_meta + nb_channels - 1]
        h = (h + 1) // 2
    while w > 8 or h > 8:
        if w > 8:
            p += [1, nb_meta, nb_meta + nb_channels - 1]
            w = (w + 1) // 2
        if h > 8:
            p += [0, nb_meta, nb_meta + nb_channels - 1]
            h = (h + 1) // 2
    return p

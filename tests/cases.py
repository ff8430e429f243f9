"""Parity cases shared by tests/golden/make_golden.py and the tests.

name, w, h, channels, maxval, seed, ref_driver encode options.  They cover what SURVEY.md 4 asks for: odd widths /
heights (odd-tail paths of squeeze.h:129, 217-222), 1xN / Nx1-ish strips, sizes that are not a multiple of 8 for the DCT
(repeating_edge_value, dct.h:327), 14-bit 4-channel input, quantised residuals, every predictor, no back-references,
uncompressed groups, images too small to be squeezed.
"""

CASES = [
    ("odd", 37, 29, 3, 255, 1, []),
    ("tall", 5, 131, 3, 255, 2, []),
    ("wide", 200, 3, 3, 255, 3, []),
    ("gray", 64, 48, 1, 255, 4, []),
    ("rgba14", 96, 80, 4, 16383, 9, ["-q", "12,64"]),
    ("lossyq", 130, 70, 3, 255, 5, ["-q", "12,64"]),
    ("dct", 64, 64, 3, 255, 7, ["-C", "1", "-J", "-q", "8,12"]),
    ("dctodd", 70, 45, 3, 255, 8, ["-C", "1", "-J", "-q", "8,12"]),
    ("nosq", 40, 30, 3, 255, 6, ["-S", "0"]),
    ("pred", 50, 50, 3, 255, 10, ["-P", "1456"]),
    ("e0", 60, 44, 3, 255, 11, ["-E", "0"]),
    ("unc", 33, 21, 3, 255, 12, ["-U"]),
    ("tiny", 3, 2, 3, 255, 13, []),
    ("one", 1, 1, 3, 255, 14, []),
    ("sq128", 128, 96, 3, 255, 15, []),
]

# ChromaSubsample (reference transform/subsample.h): name, w, h, channels, maxval, seed, transform parameters, also as a file?
# One abbreviated parameter = 4:2:0 / 4:2:2 / 4:4:0 / 4:1:1; four = (first channel, last channel, ratio_h, ratio_v).
# Odd sizes make the upscaled planes larger than the image; 4:1:1 and 3x3 take the box-filter branch (no file: the
# bitstream's meta step only allows ratios 1 and 2).
SUBSAMPLE_CASES = [
    ("420", 64, 48, 3, 255, 31, [0], True),
    ("420odd", 37, 29, 3, 255, 32, [0], True),
    ("422", 50, 21, 3, 255, 33, [1], True),
    ("440", 33, 40, 3, 255, 34, [2], True),
    ("411", 64, 20, 3, 255, 35, [3], False),
    ("box3", 31, 17, 3, 1023, 36, [1, 2, 3, 3], False),
    ("one_chan", 40, 30, 4, 16383, 37, [2, 2, 2, 2], True),
    ("tiny", 2, 3, 3, 255, 38, [0], True),
]

# Approximate (reference transform/approximate.h; ref_driver option -A k,q = the last k channels divided by q+1, remainders
# appended as extra channels, as fuif.cpp:504-510 does).  Same tuple layout as CASES.
APPROX_CASES = [
    ("approx", 64, 48, 3, 255, 41, ["-A", "2,3"]),
    ("approx_q", 50, 40, 3, 255, 42, ["-q", "12,64", "-A", "3,1"]),
    ("approx_nosq", 40, 30, 3, 255, 43, ["-S", "0", "-A", "3,7"]),
    ("approx_noop", 33, 31, 1, 255, 44, ["-S", "0", "-A", "1,0"]),
    ("approx14", 48, 36, 4, 16383, 45, ["-A", "4,99"]),
]

# Palette (reference transform/palette.h; ref_driver option -L n = one palette over all channels with at most n colours, after
# the colour transform, as fuif.cpp:398-407 does).  The input of these cases is the synthetic image reduced to four levels per
# channel (tests/golden/make_golden.py), so that it has few colours.  Same tuple layout as CASES.
PALETTE_CASES = [
    ("pal", 48, 40, 3, 255, 51, ["-L", "512"]),
    ("pal_nosq", 37, 29, 3, 255, 52, ["-S", "0", "-L", "1000"]),
    ("pal4", 40, 30, 4, 255, 53, ["-L", "4000"]),
    ("pal_c0", 33, 27, 3, 255, 54, ["-C", "0", "-S", "0", "-L", "300"]),
]

# Permute with explicit parameters (reference transform/permute.h; ref_driver option -M a,b,c before the palette / squeeze
# steps).  Same tuple layout as CASES.
PERMUTE_CASES = [
    ("perm", 40, 30, 3, 255, 61, ["-M", "2,0,1"]),
    ("perm_nosq", 33, 21, 4, 255, 62, ["-S", "0", "-M", "1,0,3,2"]),
    ("perm2", 36, 28, 3, 255, 63, ["-C", "0", "-M", "1,0"]),
]

# 2DMatch (reference transform/2dmatch.h; ref_driver option -D n = exact matches over all channels within offset codes 1..n,
# applied before the colour transform as fuif.cpp:440-447 does).  The input of these cases is a noisy patch repeated
# horizontally and vertically (tests/golden/make_golden.py), so that the reference's heuristic finds matches.
MATCH_CASES = [
    ("match", 60, 44, 3, 255, 71, ["-D", "1500"]),
    ("match_nosq", 52, 40, 3, 255, 72, ["-S", "0", "-D", "1200"]),
    ("match_gray", 64, 36, 1, 255, 73, ["-C", "0", "-D", "900"]),
    ("match_soft", 48, 34, 3, 255, 74, ["-S", "0", "-D", "900", "-W"]),
]

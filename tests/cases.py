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

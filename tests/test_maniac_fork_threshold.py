"""The run-ahead walkers of the GPU MANIAC decoder (fuif_b200/csrc/fb_maniac.cu, walker_main) turn a tree test of
property 12, slog(left - leftleft) > splitval, into a threshold on the not-yet-decoded pixel: leftleft <= T with
T = left - dmin(splitval).  This checks that identity (reference slog: encoding/context_predict.h:54-61) for every
difference a helped group can produce (value range <= 256) and every split value, with the same dmin() the kernel uses."""
import numpy as np


def slog(x: int) -> int:            # context_predict.h:54-61 on an int16 argument
    x = int(np.int16(x))
    if x == 0:
        return 0
    b = abs(x).bit_length()
    return b if x > 0 else -b


def dmin(sv: int) -> int:           # walker_main
    if sv >= 0:
        return 512 if sv >= 9 else (1 << sv)
    t = -(sv + 1)
    return -511 if t >= 9 else -((1 << t) - 1)


def test_slog_test_is_a_threshold_on_the_difference():
    for sv in range(-40, 41):
        dm = dmin(sv)
        for d in range(-255, 256):
            assert (slog(d) > sv) == (d >= dm), (sv, d)


def test_threshold_on_leftleft_fits_int16():
    # helped groups keep their values within [-32000, 32000] and a range <= 256, so T = left - dmin never leaves int16
    for left in (-32000, -1, 0, 1, 32000):
        for sv in (-40, -9, -1, 0, 8, 9, 40):
            t = left - dmin(sv)
            assert -32768 <= t <= 32767
            for leftleft in range(left - 255, left + 256):
                assert (slog(left - leftleft) > sv) == (leftleft <= t)

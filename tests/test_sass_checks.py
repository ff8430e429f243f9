"""CPU: static checks on the SASS of the built library (cuobjdump, no GPU needed).

ptxas 12.9 was caught splitting a packed add with the immediate 0x00010001 into per-half pieces and losing the +1 of the
low half (VIADD.16x2 R, R, 0x0 followed by a PRMT 0x7610 merge) in one unrolled instance of the packed unsqueeze kernel --
wrong pixels on the hardware, invisible to the CPU emulator.  The kernels now take their packed constants from a register
the compiler cannot see through (ps::PK); this test keeps the pattern from coming back, and records that the TMA /
16x2 instructions the design relies on are really in the binary."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fuif_b200", "libfuif_b200.so")


@pytest.fixture(scope="module")
def sass():
    if not shutil.which("cuobjdump") or not os.path.exists(LIB):
        pytest.skip("cuobjdump or the built library is not available")
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
            funcs[cur].append(ln)
    return funcs


def test_no_split_packed_immediates(sass):
    pk = {k: v for k, v in sass.items() if "k_pk_" in k}
    assert len(pk) >= 6, sorted(pk)
    for name, lines in pk.items():
        for ln in lines:
            assert not re.search(r"VIADD\.16x2 R\d+, R\d+(\.reuse)?, 0x[0-9a-f]+ ;", ln), f"{name}: packed add with an immediate operand: {ln.strip()}"


def test_blackwell_data_movement_is_in_the_binary(sass):
    h = [v for k, v in sass.items() if "k_pk_hsq" in k]
    assert h
    for lines in h:
        text = "\n".join(lines)
        assert "UTMALDG.2D" in text and "UTMASTG.2D" in text, "TMA tile loads / stores missing from the packed horizontal kernel"
        assert "SYNCS.ARRIVE.TRANS64" in text and "SYNCS.PHASECHK.TRANS64.TRYWAIT" in text, "mbarrier pipeline missing"
        assert "VIADD.16x2" in text and "VIADDMNMX.S16x2.RELU" in text and "VIMNMX.S16x2" in text, "packed 16x2 arithmetic missing"
    v = "\n".join(next(v for k, v in sass.items() if "k_pk_vsq" in k))
    assert "LDG.E.128" in v and "STG.E.128" in v and "VIADD.16x2" in v

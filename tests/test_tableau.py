"""The DOP853 coefficient headers (oracle + CUDA) against two independent sources."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADERS = [os.path.join(ROOT, "oracle", "dop853_tableau.h"),
           os.path.join(ROOT, "numbalsoda_b200", "csrc", "dop853_tableau.cuh")]
REF = "/root/reference/src/dop853_constants.f90"


def _read(path):
    out = {}
    for line in open(path):
        m = re.match(r"#define (DP_\w+)\s+(\S+)", line) if "b200_dpc[" not in line else None
        c = re.match(r"\s+(\S+),\s+// (DP_\w+)", line)
        if c:
            out[c.group(2)] = float(c.group(1))
        if m:
            out[m.group(1)] = float(m.group(2))
    return out


@pytest.mark.parametrize("path", HEADERS)
def test_against_scipy(path):
    from scipy.integrate._ivp import dop853_coefficients as co
    c = _read(path)
    assert len(c) == 154
    for k, v in c.items():
        name = k[3:]
        if name.startswith("BHH") or name.startswith("ER") or name.startswith("D"):
            continue
        if name[0] == "C":
            assert v == co.C[int(name[1:]) - 1]
        elif name[0] == "A":
            i, j = map(int, name[1:].split("_"))
            assert v == co.A[i - 1, j - 1]
        elif name[0] == "B":
            assert v == co.B[int(name[1:]) - 1]
    for j in range(1, 13):
        if co.E5[j - 1] != 0:
            assert c["DP_ER%d" % j] == co.E5[j - 1]
    for r in range(4):
        for j in range(1, 17):
            if co.D[r, j - 1] != 0:
                assert c["DP_D%d_%d" % (r + 4, j)] == co.D[r, j - 1]
    # scipy folds bhh into E3 = B - bhh (different rounding): check to 1e-16
    e3 = co.B.copy()
    e3[0] -= c["DP_BHH1"]
    e3[8] -= c["DP_BHH2"]
    e3[11] -= c["DP_BHH3"]
    assert np.allclose(e3, co.E3[:12], rtol=0, atol=1e-16)


@pytest.mark.skipif(not os.path.exists(REF), reason="reference checkout absent")
@pytest.mark.parametrize("path", HEADERS)
def test_against_reference_literals(path):
    ref = {}
    for line in open(REF):
        m = re.match(r"\s*real\(wp\),parameter :: (\w+)\s*=\s*([-+0-9.eE]+)_wp", line)
        if m:
            ref[m.group(1)] = float(m.group(2))
    mine = _read(path)

    def refname(k):
        k = k[3:]
        if k.startswith("BHH"):
            return "bhh" + k[3:]
        if k.startswith("ER"):
            return "er" + k[2:]
        if k[0] in "AD":
            a, b = k[1:].split("_")
            return k[0].lower() + a + b
        return k[0].lower() + k[1:]
    assert set(refname(k) for k in mine) == set(ref)
    for k, v in mine.items():
        assert ref[refname(k)] == v, k

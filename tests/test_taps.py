"""The generated kernel constants (csrc/sdr_q15_taps.h) equal the oracle's quantisation of
the reference's coefficient tables, and the clamp-safety thresholds are what they claim."""
import os
import re

import numpy as np

import _oracle as O

HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "rtlsdrdiags_b200", "csrc", "sdr_q15_taps.h")
ORDER = ["AM1", "AM2", "AM3", "FM_TUNER", "FM_POST", "AUDIO40", "WB_PRE", "WB_DEC1", "SSB_DELAY", "SSB_HILBERT"]
# SURVEY appendix A.1 checksums of the quantised taps
SUMS = {"AM1": 29002, "AM2": 34926, "FM_TUNER": 35938, "FM_POST": 36758, "WB_DEC1": 29126}
SUMABS = {"AM3": 48394, "AUDIO40": 66852, "WB_PRE": 54924, "SSB_HILBERT": 67250}


def parse():
    txt = open(HDR).read()
    out = {}
    for m in re.finditer(r"struct (\w+) \{.*?N = (\d+);.*?SUMABS = (\d+);.*?SAFE = (\d+);.*?t\[N\] = \{([^}]*)\}", txt, re.S):
        out[m.group(1)] = (int(m.group(2)), int(m.group(3)), int(m.group(4)),
                           np.array([int(v) for v in m.group(5).split(",")], dtype=np.int64))
    return out


def test_header_matches_oracle_quantisation():
    hdr = parse()
    assert sorted(hdr) == sorted(ORDER)
    for fid, name in enumerate(ORDER):
        n, sumabs, safe, q = hdr[name]
        oq = O.q15_taps(fid).astype(np.int64)
        assert n == oq.size and np.array_equal(q, oq), name
        assert sumabs == int(np.abs(oq).sum())
        assert safe == min((0x3FFFFFFF - 16384) // sumabs, 32768)
        # at |x| = safe no ordered partial sum can leave [-2^30, 2^30-1]; one above, it can
        assert 16384 + sumabs * safe <= 0x3FFFFFFF
        if safe < 32768:
            assert 16384 + sumabs * (safe + 1) > 0x3FFFFFFF


def test_survey_checksums():
    hdr = parse()
    for name, s in SUMS.items():
        assert int(hdr[name][3].sum()) == s
    for name, s in SUMABS.items():
        assert hdr[name][1] == s
    assert list(hdr["SSB_DELAY"][3]) == [0] * 15 + [-32768]
    assert int(hdr["SSB_HILBERT"][3].sum()) == 0


def test_int8_storage_bounds_of_the_am_cascade():
    """Stage 1 and 2 outputs of the AM/SSB cascade are stored as int8 in the kernels."""
    hdr = parse()
    s1 = (16384 + hdr["AM1"][1] * 128) >> 15
    s2 = (16384 + hdr["AM2"][1] * s1) >> 15
    assert s1 <= 127 and s2 <= 127
    # FM tuner output range that sizes the 280x280 atan2 table
    q = hdr["FM_TUNER"][3]
    assert (q > 0).all()
    assert (16384 + int(q.sum()) * 127) >> 15 == 139 and (16384 - int(q.sum()) * 128) >> 15 == -140

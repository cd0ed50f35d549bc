"""CPU: the oracle restatement (oracle/recon_oracle.c) against pictures dumped from the compiled
reference (tests/golden/*.npz, made by tests/golden/make_fixtures.py).  Bit-exact."""
import glob
import os

import numpy as np
import pytest

from losslessh264_b200.records import MB_REC, SLICE_REC, load_fixture, n_slots_needed
from oracle import oracle

FIXTURES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def test_fixtures_present():
    assert len(FIXTURES) >= 10


def test_record_sizes_match_c_header():
    lib = oracle.lib()
    assert lib.ora_sizeof_mb_rec() == MB_REC.itemsize == 128
    assert lib.ora_sizeof_slice_rec() == SLICE_REC.itemsize


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_oracle_matches_reference_pictures(path):
    pics = load_fixture(path)
    st = oracle.OracleStream(pics[0].mb_w, pics[0].mb_h, n_slots_needed(pics))
    for i, p in enumerate(pics):
        out = st.recon(p)
        assert np.array_equal(out, p.final_yuv()), f"{os.path.basename(path)} picture {i} differs from the reference"

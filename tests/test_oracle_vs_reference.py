"""The committed golden vectors are reproducible from the reference's OWN code: re-run
/root/reference's NumPy paths and CuPy kernel source strings through oracle/ref_harness (serial
C++ emulation compiled into oracle/_ref/) and compare every array with tests/golden/*.npz.

Skipped where /root/reference does not exist (the GPU box): there the fixtures are the pin."""

import os

import numpy as np
import pytest

from oracle.ref_harness import shim

pytestmark = pytest.mark.skipif(not shim.reference_available(),
                                reason="/root/reference not present")


def test_goldens_regenerate_bit_identically(tmp_path, monkeypatch):
    from oracle.ref_harness import gen_golden as gg
    committed = gg.OUT
    monkeypatch.setattr(gg, "OUT", str(tmp_path))
    gg.main()
    names = sorted(f for f in os.listdir(committed) if f.endswith(".npz"))
    assert names == sorted(os.listdir(tmp_path)), "generator and committed fixture sets differ"
    for f in names:
        a, b = np.load(tmp_path / f, allow_pickle=True), np.load(os.path.join(committed, f),
                                                                 allow_pickle=True)
        assert set(a.files) == set(b.files), f
        for k in a.files:
            if a[k].dtype.kind in "fc":
                assert np.array_equal(a[k], b[k], equal_nan=True), (f, k)
            else:
                assert np.array_equal(a[k], b[k]), (f, k)

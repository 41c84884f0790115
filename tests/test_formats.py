"""Reference file formats (sqmc_b200/formats.py): the wf_eps_var checkpoint (Fortran sequential unformatted, hci.f90:194-231,
601-625) and the deterministic-projector text file (do_walk.f90:883-1013).  CPU only.  No reference-written sample exists
(parity unpinned by the reference), so the byte layout is pinned against hand-assembled gfortran records."""
import os
import struct

import numpy as np
import pytest

from sqmc_b200 import formats


def test_wf_filename_matches_es7_2e1():
    assert formats.wf_filename(5e-4) == "wf_eps_var=5.00E-4"      # hci.f90:196-197
    assert formats.wf_filename(1e-3) == "wf_eps_var=1.00E-3"
    assert formats.wf_filename(2.5e-6) == "wf_eps_var=2.50E-6"
    assert formats.wf_filename(3.0) == "wf_eps_var=3.00E+0"
    assert formats.wf_filename(1e-10) == "wf_eps_var=*******"      # two exponent digits overflow the e1 field


def test_wf_bytes_are_gfortran_records(tmp_path):
    up = np.array([[0b1111, 0], [0b10111, 0], [1, 1 << 3]], dtype=np.uint64)   # third det uses orbital 67 (high word)
    dn = np.array([[0b1111, 0], [0b1111, 0], [0b11011, 0]], dtype=np.uint64)
    wts = np.array([[0.9, 0.1], [-0.3, 0.8], [0.05, -0.2]])
    en = np.array([-75.5, -75.25])
    p = tmp_path / "wf"
    formats.write_wf(p, up, dn, wts, en)
    raw = open(p, "rb").read()

    def rec(b):
        return struct.pack("<i", len(b)) + b + struct.pack("<i", len(b))

    def i16(lo, hi):
        return struct.pack("<QQ", lo, hi)

    expect = (rec(struct.pack("<i", 3))
              + rec(i16(15, 0) + i16(23, 0) + i16(1, 8))
              + rec(i16(15, 0) + i16(15, 0) + i16(27, 0))
              + rec(struct.pack("<6d", 0.9, -0.3, 0.05, 0.1, 0.8, -0.2))     # wts(1:n,1:n_states): column-major
              + rec(struct.pack("<2d", -75.5, -75.25)))
    assert raw == expect
    back = formats.read_wf(p)
    assert np.array_equal(back["up"], up) and np.array_equal(back["dn"], dn)
    assert np.array_equal(back["wts"], wts) and np.array_equal(back["energies"], en)


def test_wf_subrecords_round_trip(tmp_path):
    rng = np.random.default_rng(5)
    n = 37
    up = rng.integers(0, 1 << 40, (n, 2)).astype(np.uint64)
    dn = rng.integers(0, 1 << 40, (n, 2)).astype(np.uint64)
    wts = rng.normal(size=(n, 1))
    p = tmp_path / "wf_small_subrecords"
    formats.write_wf(p, up, dn, wts, [-1.5], max_sub=100)       # 592-byte records -> 6 subrecords each
    raw = open(p, "rb").read()
    # second record: first subrecord is "continued" (negative leading marker, positive trailing marker) ...
    off = 12
    assert struct.unpack("<i", raw[off:off + 4])[0] == -100 and struct.unpack("<i", raw[off + 104:off + 108])[0] == 100
    # ... the last one closes the record (positive leading marker, negative trailing marker)
    last = off + 5 * 108
    assert struct.unpack("<i", raw[last:last + 4])[0] == 92 and struct.unpack("<i", raw[last + 96:last + 100])[0] == -92
    back = formats.read_wf(p)
    assert np.array_equal(back["up"], up) and np.array_equal(back["dn"], dn) and np.array_equal(back["wts"], wts)


def test_wf_rejects_damaged_files(tmp_path):
    p = tmp_path / "wf"
    formats.write_wf(p, np.array([[3, 0]], dtype=np.uint64), np.array([[3, 0]], dtype=np.uint64), [[1.0]], [-1.0])
    raw = open(p, "rb").read()
    open(p, "wb").write(raw[:-3])
    with pytest.raises(ValueError):
        formats.read_wf(p)


GFORTRAN_STYLE = """\
           3           4  -1.2500000000000000      number of deterministic dets, number of nonzero deterministic Hamiltonian elements, ground state energy within deterministic space
       2       1       1
           1           1           2           1           2
           2           1           3           1           2
           3           1           2           2           3
                    1  -1.2500000000000000     
                    3  0.12500000000000000     
                    2 -0.75000000000000000     
                    3  -5.0000000000000003E-002
"""


def test_dtm_projector_reads_list_directed_text(tmp_path):
    """a file laid out the way gfortran's list-directed writes of do_walk.f90:971-1007 look (nup = ndn = 2)"""
    p = tmp_path / "dtm_projector.in"
    open(p, "w").write(GFORTRAN_STYLE)
    d = formats.read_dtm_projector(p, nup=2, ndn=2)
    assert d["counts"].tolist() == [2, 1, 1] and d["indices"].tolist() == [1, 3, 2, 3]
    assert d["values"].tolist() == [-1.25, 0.125, -0.75, -0.05]
    assert d["up"][:, 0].tolist() == [0b011, 0b101, 0b011] and d["dn"][:, 0].tolist() == [0b011, 0b011, 0b110]
    assert d["dtm_energy"] == -1.25


def test_dtm_projector_round_trip_with_core_orbitals(tmp_path):
    rng = np.random.default_rng(11)
    n, ncore, nup, ndn = 20, 2, 5, 4
    up = np.zeros((n, 2), dtype=np.uint64)
    dn = np.zeros((n, 2), dtype=np.uint64)
    for i in range(n):
        a = sum(1 << int(o) for o in rng.choice(np.arange(ncore, 70), nup - ncore, replace=False)) | 0b11
        b = sum(1 << int(o) for o in rng.choice(np.arange(ncore, 70), ndn - ncore, replace=False)) | 0b11
        up[i] = (a & (2 ** 64 - 1), a >> 64)
        dn[i] = (b & (2 ** 64 - 1), b >> 64)
    counts = rng.integers(1, 4, n)
    idx = np.concatenate([np.sort(rng.choice(np.arange(i + 1, n + 1), min(c, n - i), replace=False)) for i, c in enumerate(counts)])
    counts = np.array([min(c, n - i) for i, c in enumerate(counts)])
    val = rng.normal(size=len(idx))
    p = tmp_path / "dtm_projector.out"
    formats.write_dtm_projector(p, up, dn, counts, idx, val, dtm_energy=-3.5, n_core_orb=ncore)
    d = formats.read_dtm_projector(p, nup, ndn, n_core_orb=ncore)
    assert np.array_equal(d["up"], up) and np.array_equal(d["dn"], dn)
    assert np.array_equal(d["counts"], counts) and np.array_equal(d["indices"], idx)
    assert np.array_equal(d["values"], val)          # 17 significant digits: doubles survive the text round trip
    assert d["dtm_energy"] == -3.5

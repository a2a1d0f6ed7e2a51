"""CPU: the host side of the delta loader -- the plain-C FITS scanner inside libpicca_b200.so
(pb2_fits_scan, pb2_fits_cards, pb2_fits_hierarch; no GPU involved) and the per-file geometry the
loader derives from it -- against the test harness's independent FITS reader and hand-made
headers."""
import os

import numpy as np
import pytest

from picca_b200 import io
from tests.golden import cases_io
from tests.refharness import minifits

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fixtures")


def _hdu(cards, data=b""):
    text = "".join(c.ljust(80) for c in cards) + "END".ljust(80)
    text += " " * ((-len(text)) % 2880)
    return text.encode("ascii") + data + b"\0" * ((-len(data)) % 2880)


def test_scan_and_cards_on_handmade_headers():
    primary = _hdu(["SIMPLE  =                    T", "BITPIX  =                    8",
                    "NAXIS   =                    0"])
    table = np.arange(6, dtype=">f8").tobytes()
    ext = _hdu(["XTENSION= 'BINTABLE'", "BITPIX  =                    8",
                "NAXIS   =                    2", "NAXIS1  =                   16",
                "NAXIS2  =                    3", "PCOUNT  =                    0",
                "GCOUNT  =                    1", "TFIELDS =                    2",
                "TTYPE1  = 'LOGLAM  '", "TFORM1  = 'D       '", "TTYPE2  = 'DELTA   '",
                "TFORM2  = 'D       '", "EXTNAME = 'it''s    '           / escaped quote",
                "RA      =   1.2345678901234567 / radians", "DEC     = -2.5D-1",
                "Z       = 2.75", "LOS_ID  =    39627123456789012 / beyond 2**53",
                "BLINDING= 'none    '", "COMMENT   a comment card without a value",
                "HIERARCH WAVE_SOLUTION = 'lin     '", "HIERARCH DELTA_LAMBDA = 0.8"], table)
    buf = np.frombuffer(primary + ext, dtype=np.uint8)
    info = io._scan(buf)
    assert info.shape == (2, 8)
    assert list(info[0]) == [0, 2880, 0, 8, 0, 0, 0, 0]
    assert list(info[1, :3]) == [2880, 2880 + 2880, 48] and list(info[1, 3:]) == [8, 2, 16, 3, 2]
    kind, num, inum, strs = io._cards(buf, info[:, 0])
    K = io._K
    assert kind[0, K["RA"]] == 0 and kind[1, K["RA"]] == 1
    assert num[1, K["RA"]] == 1.2345678901234567 and num[1, K["DEC"]] == -0.25
    assert num[1, K["Z"]] == 2.75
    assert inum[1, K["LOS_ID"]] == 39627123456789012        # exact, not through a double
    assert kind[1, K["THING_ID"]] == 0
    assert strs[1, K["EXTNAME"]] == b"it's" and strs[1, K["BLINDING"]] == b"none"
    assert strs[1, K["TTYPE2"]] == b"DELTA" and strs[1, K["TFORM1"]] == b"D"
    assert io._hierarch(buf, int(info[1, 0]), "WAVE_SOLUTION") == "lin"
    assert io._hierarch(buf, int(info[1, 0]), "DELTA_LAMBDA") == 0.8
    with pytest.raises(KeyError):
        io._hierarch(buf, int(info[1, 0]), "NOT_THERE")


def test_truncated_and_foreign_files_raise():
    with pytest.raises(OSError):
        io._scan(np.frombuffer(b"not a fits file".ljust(2880), dtype=np.uint8))
    good = _hdu(["SIMPLE  =                    T", "BITPIX  =                  -64",
                 "NAXIS   =                    1", "NAXIS1  =                 1000"],
                np.zeros(1000, ">f8").tobytes())
    with pytest.raises(OSError):  # data runs past the end of the file
        io._scan(np.frombuffer(good[:2880 + 2880], dtype=np.uint8))


@pytest.mark.parametrize("name", ["delta-272.fits.gz", "image-delta-50.fits.gz",
                                  "delta_attributes.fits.gz"])
def test_scan_agrees_with_the_harness_reader_on_reference_files(name):
    path = os.path.join(FIX, name)
    buf = io._file_bytes(path)
    info = io._scan(buf)
    ref = minifits.FITS(path)
    assert len(info) == len(ref)
    kind, num, inum, strs = io._cards(buf, info[:, 0])
    for h, hdu in enumerate(ref):
        head = hdu.read_header()
        assert info[h, 4] == head.get("NAXIS", 0)
        if head.get("NAXIS", 0) >= 1:
            assert info[h, 5] == head["NAXIS1"]
        for key in ("RA", "DEC", "Z"):
            if key in head:
                assert num[h, io._K[key]] == head[key]
        if "EXTNAME" in head:
            assert strs[h, io._K["EXTNAME"]].decode() == str(head["EXTNAME"]).strip()


def test_bintable_geometry_and_header_values(tmp_path):
    in_dir, _ = cases_io.write_case(str(tmp_path), "sdss")
    path = sorted(os.listdir(in_dir))[0]
    part = io._open_delta_file(os.path.join(in_dir, path), 0, 10)
    ref = minifits.FITS(os.path.join(in_dir, path))
    kept = [h for h in ref[1:] if 0 < h.read_header()["Z"] < 10]
    assert part.n == len(kept) and not part.is_image
    for k, hdu in enumerate(kept):
        head = hdu.read_header()
        assert (part.ra[k], part.dec[k], part.z_qso[k]) == (head["RA"], head["DEC"], head["Z"])
        assert (part.los_id[k], part.plate[k], part.fiberid[k]) == (
            head["THING_ID"], head["PLATE"], head["FIBERID"])
        assert part.n_pix[k] == hdu.get_nrows()
        rows = np.frombuffer(part.buf[part.row0[k]:part.row0[k] + part.row_bytes[k] * part.n_pix[k]]
                             .tobytes(), dtype=">f8").reshape(-1, part.row_bytes[k] // 8)
        assert np.array_equal(rows[:, part.col_off[k, 0] // 8], hdu["LOGLAM"][:])
        assert np.array_equal(rows[:, part.col_off[k, 1] // 8], hdu["DELTA"][:])
        assert np.array_equal(rows[:, part.col_off[k, 2] // 8], hdu["WEIGHT"][:])
    # strict quasar redshift cut of io.py:359-360
    cut = io._open_delta_file(os.path.join(in_dir, path), 2.4, 3.0)
    assert cut.n == sum(2.4 < h.read_header()["Z"] < 3.0 for h in ref[1:])


def test_image_geometry_blinding_and_inclusive_cut(tmp_path):
    in_dir, _ = cases_io.write_image_case(str(tmp_path), "imageblind")
    path = os.path.join(in_dir, sorted(os.listdir(in_dir))[0])
    part = io._open_delta_file(path, 0, 10, rebin=True)
    ref = minifits.FITS(path)
    z = ref["METADATA"]["Z"][:]
    assert part.is_image and part.n == int(((z >= 0) & (z <= 10)).sum())  # data.py:602
    assert part.dwave == 0.8 and part.wave_flag
    assert part.n_lambda == ref["LAMBDA"].read_header()["NAXIS1"]
    names = [h.read_header().get("EXTNAME") for h in ref]
    assert "DELTA_BLIND" in names  # the blinded column is the one picked (data.py:546-554)
    raw = io._file_bytes(path)
    lam = np.frombuffer(raw[part.lambda_off:part.lambda_off + 8 * part.n_lambda].tobytes(), ">f8")
    assert np.array_equal(lam, ref["LAMBDA"][:])
    d = np.frombuffer(raw[part.delta_off:part.delta_off + 8 * part.n_lambda].tobytes(), ">f8")
    assert np.array_equal(d, ref["DELTA_BLIND"].read()[0])
    assert np.array_equal(part.los_id, ref["METADATA"]["LOS_ID"][:][part.rows])


def test_find_order_fallbacks(tmp_path):
    io.userprint = lambda *a, **k: None
    assert io.find_order(str(tmp_path), os.path.join(FIX, "delta_attributes.fits.gz")) == 1
    # no attributes file, no config -> None; a config file two levels up -> its order (io.py:94-103)
    deltas = tmp_path / "run" / "Delta"
    deltas.mkdir(parents=True)
    assert io.find_order(str(deltas), str(tmp_path / "missing.fits.gz")) is None
    (tmp_path / "run" / ".config.ini").write_text("[expected flux]\norder = 0\n")
    assert io.find_order(str(deltas), str(tmp_path / "missing.fits.gz")) == 0

"""On-disk contract of the fit products: {root}_{n}vcomp.fits = concatenate([parcube, errcube])
with PLANEi cards (mufasa/multi_v_fit.py:523-582), written and read back without astropy."""
import numpy as np

from mufasa_b200 import multi_v_fit as mvf
from mufasa_b200.utils import fits_io


class _FakePcube(object):
    def __init__(self, ncomp, ny=5, nx=7):
        rng = np.random.default_rng(ncomp)
        self.parcube = rng.normal(size=(4 * ncomp, ny, nx))
        self.errcube = np.abs(rng.normal(size=(4 * ncomp, ny, nx)))
        self.parcube[:, 0, 0] = 0.0
        self.errcube[:, 0, 0] = 0.0
        self.header = {"OBJECT": "synthetic", "BMAJ": (0.0088, "[deg] beam"), "CRVAL1": 266.4, "CTYPE1": "RA---TAN",
                       "CTYPE3": "VRAD", "WCSAXES": 3}


def test_fits_structure_and_roundtrip(tmp_path):
    for dtype in (np.float64, np.float32, np.int16, np.uint8):
        a = (np.arange(2 * 3 * 5).reshape(2, 3, 5) * 1.5).astype(dtype)
        path = tmp_path / ("a_%s.fits" % np.dtype(dtype).name)
        fits_io.write(path, a, {"OBJECT": "x'y", "EXPTIME": (12.5, "seconds"), "FLAG": True, "N": 7})
        raw = open(path, "rb").read()
        assert len(raw) % 2880 == 0
        assert raw[:30].decode() == "SIMPLE  =                    T"
        assert raw[80:80 + 8].decode() == "BITPIX  " and b"END" + b" " * 77 in raw[:2880]
        # NAXIS1 is the fastest axis (FITS order is the reverse of the numpy shape)
        assert b"NAXIS1  =                    5" in raw[:2880] and b"NAXIS3  =                    2" in raw[:2880]
        b, hdr = fits_io.read(path)
        assert b.dtype == np.dtype(dtype) and np.array_equal(a, b)
        assert hdr.value("OBJECT") == "x'y" and hdr.value("EXPTIME") == 12.5 and hdr["EXPTIME"][1] == "seconds"
        assert hdr.value("FLAG") is True and hdr.value("N") == 7
    # big-endian IEEE on disk
    fits_io.write(tmp_path / "one.fits", np.array([[1.0]]))
    raw = open(tmp_path / "one.fits", "rb").read()
    assert raw[2880:2888] == np.array([1.0], dtype=">f8").tobytes()


def test_save_pcube_plane_cards_and_load_model_fit(tmp_path):
    from mufasa_b200.pcube import PCube
    for ncomp in (1, 2):
        fake = _FakePcube(ncomp)
        path = str(tmp_path / ("fit_%dvcomp.fits" % ncomp))
        mvf.save_pcube(fake, path, ncomp=ncomp)
        data, hdr = fits_io.read(path)
        assert data.shape == (8 * ncomp, 5, 7)
        assert np.array_equal(data[:4 * ncomp], fake.parcube) and np.array_equal(data[4 * ncomp:], fake.errcube)
        assert hdr.value("CTYPE3") == "FITPAR" and hdr.value("CUNIT3") == "None" and hdr.value("CRPIX3") == 1
        assert hdr.value("OBJECT") == "synthetic" and hdr.value("CTYPE1") == "RA---TAN"
        if ncomp == 1:
            assert [hdr.value("PLANE%d" % i) for i in range(8)] == [
                "VELOCITY", "SIGMA", "TEX", "TAU", "eVELOCITY", "eSIGMA", "eTEX", "eTAU"]
        else:
            assert [hdr.value("PLANE%d" % i) for i in range(16)] == [
                "VELOCITY_1", "SIGMA_1", "TEX_1", "TAU_1", "VELOCITY_2", "SIGMA_2", "TEX_2", "TAU_2",
                "eVELOCITY_1", "eSIGMA_1", "eTEX_1", "eTAU_1", "eVELOCITY_2", "eSIGMA_2", "eTEX_2", "eTAU_2"]
        assert any("mufasa_b200" in h for h in hdr.history)
        # resume: a fresh PCube on the same grid picks the fit up again
        p = PCube(np.zeros((16, 5, 7), dtype=np.float32), np.arange(16.0))
        p.load_model_fit(path, npars=4 * ncomp, fittype="nh3_multi_v")
        assert np.array_equal(p.parcube, fake.parcube) and np.array_equal(p.errcube, fake.errcube)
        assert not p.has_fit[0, 0] and p.has_fit[1:, :].all()

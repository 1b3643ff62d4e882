"""healpy-free mask path: FITS BINTABLE reader + ud_grade (ConstrainedRealization.py:33-37), CPU only."""
import numpy as np
import pytest

from gibbssampler_b200 import healpix_io as H
from oracle import sht as O


def write_fits(path, cols, nside, ordering, rep=1024, fmt="E"):
    """Minimal HEALPix FITS writer (what hp.write_map produces: one BINTABLE, `rep` values per row)."""
    def card(k, v, quote=False):
        val = ("'%-8s'" % v) if quote else ("%20s" % v)
        return ("%-8s= %s" % (k, val)).ljust(80)
    def block(cards):
        s = "".join(cards) + "END".ljust(80)
        return s.ljust(((len(s) + 2879) // 2880) * 2880).encode("ascii")
    npix = 12 * nside * nside
    rep = min(rep, npix)
    nrows = npix // rep
    size = {"E": 4, "D": 8}[fmt]
    prim = block([card("SIMPLE", "T"), card("BITPIX", 8), card("NAXIS", 0), card("EXTEND", "T")])
    h = [card("XTENSION", "BINTABLE", True), card("BITPIX", 8), card("NAXIS", 2), card("NAXIS1", rep * size * len(cols)),
         card("NAXIS2", nrows), card("PCOUNT", 0), card("GCOUNT", 1), card("TFIELDS", len(cols))]
    for i in range(len(cols)):
        h += [card("TTYPE%d" % (i + 1), "COL%d" % i, True), card("TFORM%d" % (i + 1), "%d%s" % (rep, fmt), True)]
    h += [card("PIXTYPE", "HEALPIX", True), card("ORDERING", ordering, True), card("NSIDE", nside), card("INDXSCHM", "IMPLICIT", True)]
    dt = np.dtype([("c%d" % i, ">f%d" % size, (rep,)) for i in range(len(cols))])
    rec = np.zeros(nrows, dtype=dt)
    for i, c in enumerate(cols):
        rec["c%d" % i] = np.asarray(c).reshape(nrows, rep)
    raw = rec.tobytes()
    with open(path, "wb") as f:
        f.write(prim + block(h) + raw + b"\0" * ((-len(raw)) % 2880))


def test_nest2ring_is_the_healpix_permutation():
    # nside 1: identity; nside 2: the north polar ring is the top child (index 3) of base pixels 0..3
    assert np.array_equal(H.nest2ring(1, np.arange(12)), np.arange(12))
    assert list(H.nest2ring(2, [3, 7, 11, 15])) == [0, 1, 2, 3]
    for nside in (2, 4, 16, 64):
        r = H.nest2ring(nside, np.arange(12 * nside * nside))
        assert np.array_equal(np.sort(r), np.arange(12 * nside * nside))
    # hierarchy: the four NESTED children of a pixel are its neighbours on the sphere -> their mean direction is the
    # parent's centre to O(pixel size^2)
    for nside in (4, 32):
        th, ph = O.pix_angles(2 * nside)
        v = np.stack([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)])
        vn = v[:, H.nest2ring(2 * nside, np.arange(48 * nside * nside))].reshape(3, -1, 4).mean(axis=2)
        thc, phc = O.pix_angles(nside)
        vc = np.stack([np.sin(thc) * np.cos(phc), np.sin(thc) * np.sin(phc), np.cos(thc)])[:, H.nest2ring(nside, np.arange(12 * nside * nside))]
        vn /= np.linalg.norm(vn, axis=0)
        assert np.abs(vn - vc).max() < 0.6 / nside


def test_ud_grade_mean_and_unseen():
    nside = 8
    th, _ = O.pix_angles(nside)
    m = np.cos(th)
    lo = H.ud_grade(m, 4)
    th4, _ = O.pix_angles(4)
    assert lo.shape == (192,) and np.abs(lo - np.cos(th4)).max() < 0.03      # smooth field: child mean ~ parent value
    assert abs(lo.mean() - m.mean()) < 1e-15                                 # equal-area pixels: the mean is preserved
    assert np.array_equal(H.ud_grade(np.ones(768), 2), np.ones(48))
    up = H.ud_grade(lo, 8)
    assert np.allclose(H.ud_grade(up, 4), lo, rtol=0, atol=1e-15)            # replicate then average = identity
    bad = m.copy()
    kids = H.nest2ring(8, np.arange(4))                                      # the 4 children of NESTED pixel 0 at nside 4
    bad[kids[:3]] = H.UNSEEN
    assert H.ud_grade(bad, 4)[H.nest2ring(4, [0])[0]] == m[kids[3]]
    bad[kids] = H.UNSEEN
    assert H.ud_grade(bad, 4)[H.nest2ring(4, [0])[0]] == H.UNSEEN
    binary = (np.abs(np.cos(th)) > 0.3).astype(float)                        # a binary mask gets fractional edge values
    g = H.ud_grade(binary, 2)
    assert g.min() >= 0 and g.max() <= 1 and ((g > 0) & (g < 1)).any()


@pytest.mark.parametrize("ordering,fmt", [("RING", "E"), ("NESTED", "E"), ("RING", "D")])
def test_read_map_roundtrip(tmp_path, ordering, fmt):
    nside = 16
    rng = np.random.default_rng(0)
    t = rng.standard_normal(12 * nside * nside).astype(np.float32 if fmt == "E" else np.float64)
    q = rng.standard_normal(t.size).astype(t.dtype)
    disk = [H.reorder(x, r2n=True) if ordering == "NESTED" else x for x in (t, q)]
    path = str(tmp_path / "mask.fits")
    write_fits(path, disk, nside, ordering, fmt=fmt)
    assert np.array_equal(H.read_map(path), t.astype(np.float64))
    assert np.array_equal(H.read_map(path, field=1), q.astype(np.float64))
    assert np.array_equal(H.read_map(path, nest=True), H.reorder(t.astype(np.float64), r2n=True))


def test_mask_path_goes_through_the_reader(tmp_path):
    from gibbssampler_b200 import _dev
    nside = 8
    th, _ = O.pix_angles(nside)
    binary = (np.abs(np.cos(th)) > 0.3).astype(np.float32)
    path = str(tmp_path / "gal.fits")
    write_fits(path, [binary], nside, "RING")
    m = _dev.load_mask(path, 4)
    assert np.array_equal(m, H.ud_grade(binary.astype(np.float64), 4))
    with pytest.raises(ValueError):
        _dev.load_mask(None, 4, mask=np.ones(5))

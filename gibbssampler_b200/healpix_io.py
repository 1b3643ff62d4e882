"""healpy-free replacements for the two calls the reference makes when it is given a sky mask
(`hp.ud_grade(hp.read_map(mask_path), nside)`, ConstrainedRealization.py:33-37, CenteredGibbs.py:266-268):
a minimal FITS BINTABLE reader for HEALPix maps and the NESTED-children average of ud_grade.
Host-side numpy, run once at construction; nothing here is on the hot path."""
import numpy as np

UNSEEN = -1.6375e30
_JRLL = np.array([2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4], dtype=np.int64)
_JPLL = np.array([1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7], dtype=np.int64)


def _compress_bits(v):
    """Every second bit of v (bits 0, 2, 4, ...) packed into the low half."""
    v = v & 0x5555555555555555
    v = (v | (v >> 1)) & 0x3333333333333333
    v = (v | (v >> 2)) & 0x0F0F0F0F0F0F0F0F
    v = (v | (v >> 4)) & 0x00FF00FF00FF00FF
    v = (v | (v >> 8)) & 0x0000FFFF0000FFFF
    v = (v | (v >> 16)) & 0x00000000FFFFFFFF
    return v


def nest2ring(nside, ipnest):
    """RING index of NESTED pixel(s) (standard HEALPix nest2xyf + xyf2ring); nside a power of two."""
    nside = int(nside)
    if nside < 1 or nside & (nside - 1):
        raise ValueError("NESTED ordering needs nside to be a power of two")
    p = np.asarray(ipnest, dtype=np.int64)
    npface = nside * nside
    face = p // npface
    q = p % npface
    ix = _compress_bits(q)
    iy = _compress_bits(q >> 1)
    jr = _JRLL[face] * nside - ix - iy - 1
    npix, ncap = 12 * npface, 2 * nside * (nside - 1)
    north, south = jr < nside, jr > 3 * nside
    nr = np.where(north, jr, np.where(south, 4 * nside - jr, nside))
    n_before = np.where(north, 2 * nr * (nr - 1), np.where(south, npix - 2 * (nr + 1) * nr, ncap + (jr - nside) * 4 * nside))
    kshift = np.where(north | south, 0, (jr - nside) & 1)
    jp = (_JPLL[face] * nr + ix - iy + 1 + kshift) // 2
    jp = np.where(jp > 4 * nr, jp - 4 * nr, jp)
    jp = np.where(jp < 1, jp + 4 * nr, jp)
    return n_before + jp - 1


def reorder(m, r2n=False, n2r=False):
    """hp.reorder: RING -> NESTED (r2n) or NESTED -> RING (n2r)."""
    m = np.asarray(m)
    nside = int(round(np.sqrt(m.shape[-1] / 12)))
    idx = nest2ring(nside, np.arange(12 * nside * nside))
    if r2n:
        return m[..., idx]
    out = np.empty_like(m)
    out[..., idx] = m
    return out


def ud_grade(map_in, nside_out, order_in="RING", order_out=None):
    """hp.ud_grade with the default power (degrade = mean over the NESTED children, UNSEEN children ignored;
    upgrade = replication)."""
    m = np.asarray(map_in, dtype=np.float64)
    nside_in = int(round(np.sqrt(m.size / 12)))
    if 12 * nside_in * nside_in != m.size:
        raise ValueError("not a HEALPix map: %d pixels" % m.size)
    nside_out = int(nside_out)
    order_out = order_out or order_in
    nest = m if order_in.upper().startswith("NEST") else reorder(m, r2n=True)
    if nside_out < nside_in:
        if nside_in % nside_out or (nside_in // nside_out) & (nside_in // nside_out - 1):
            raise ValueError("nside ratio must be a power of two")
        k = (nside_in // nside_out) ** 2
        ch = nest.reshape(12 * nside_out * nside_out, k)
        good = ch != UNSEEN
        cnt = good.sum(axis=1)
        s = np.where(good, ch, 0.0).sum(axis=1)
        nest = np.where(cnt > 0, s / np.maximum(cnt, 1), UNSEEN)
    elif nside_out > nside_in:
        if nside_out % nside_in or (nside_out // nside_in) & (nside_out // nside_in - 1):
            raise ValueError("nside ratio must be a power of two")
        nest = np.repeat(nest, (nside_out // nside_in) ** 2)
    return nest if order_out.upper().startswith("NEST") else reorder(nest, n2r=True)


# ---------------------------------------------------------------------------------------------- FITS
_TFORM = {"E": ">f4", "D": ">f8", "J": ">i4", "K": ">i8", "I": ">i2", "B": "u1", "L": "u1"}


def _read_header(f):
    cards = {}
    while True:
        block = f.read(2880)
        if len(block) < 2880:
            raise ValueError("truncated FITS header")
        for i in range(0, 2880, 80):
            card = block[i:i + 80].decode("ascii", "replace")
            key = card[:8].strip()
            if key == "END":
                return cards
            if card[8:10] != "= ":
                continue
            val = card[10:]
            if val.lstrip().startswith("'"):
                v = val.lstrip()[1:]
                v = v[:v.find("'")] if "'" in v else v
                cards[key] = v.strip()
            else:
                v = val.split("/")[0].strip()
                try:
                    cards[key] = int(v)
                except ValueError:
                    try:
                        cards[key] = float(v.replace("D", "E"))
                    except ValueError:
                        cards[key] = v
    return cards


def read_map(path, field=0, nest=False):
    """hp.read_map(path, field): column `field` of the first binary-table extension as a float64 HEALPix map in
    RING order (NESTED if nest=True), whatever the ordering on disk."""
    with open(path, "rb") as f:
        prim = _read_header(f)
        if prim.get("SIMPLE") not in ("T", True) and str(prim.get("SIMPLE")).strip() != "T":
            raise ValueError("%s is not a FITS file" % path)
        nbytes = abs(int(prim.get("BITPIX", 8))) // 8
        for i in range(int(prim.get("NAXIS", 0))):
            nbytes *= int(prim.get("NAXIS%d" % (i + 1), 0))
        if int(prim.get("NAXIS", 0)) == 0:
            nbytes = 0
        f.seek(((nbytes + 2879) // 2880) * 2880, 1)
        hdr = _read_header(f)
        if str(hdr.get("XTENSION", "")).strip() != "BINTABLE":
            raise ValueError("first extension of %s is not a binary table" % path)
        rowlen, nrows, nf = int(hdr["NAXIS1"]), int(hdr["NAXIS2"]), int(hdr["TFIELDS"])
        fields = []
        for i in range(nf):
            t = str(hdr["TFORM%d" % (i + 1)]).strip()
            rep = "".join(ch for ch in t if ch.isdigit())
            code = t[len(rep):len(rep) + 1]
            if code not in _TFORM:
                raise ValueError("unsupported TFORM %r" % t)
            fields.append(("c%d" % i, _TFORM[code], (int(rep) if rep else 1,)))
        dt = np.dtype(fields)
        if dt.itemsize != rowlen:
            raise ValueError("row length %d does not match the column formats (%d)" % (rowlen, dt.itemsize))
        data = np.frombuffer(f.read(rowlen * nrows), dtype=dt, count=nrows)
    m = np.ascontiguousarray(data["c%d" % field]).reshape(-1).astype(np.float64)
    nside = int(hdr.get("NSIDE", int(round(np.sqrt(m.size / 12)))))
    if 12 * nside * nside != m.size:
        raise ValueError("%s: %d values are not a full-sky HEALPix map" % (path, m.size))
    on_disk_nest = str(hdr.get("ORDERING", "RING")).upper().startswith("NEST")
    if on_disk_nest and not nest:
        m = reorder(m, n2r=True)
    elif nest and not on_disk_nest:
        m = reorder(m, r2n=True)
    return m

"""oracle/sht_fast.c (the vectorised CPU stand-in for healpy that bench.py's reference arm and cpu_baseline leg time)
against oracle/sht_oracle.c (long double accuracy checker), plus its FFT against numpy; all three ISA builds of the
hot loops (AVX-512 / AVX2 / baseline) when the host supports them."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import sht as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rand_alm(lmax, rng):
    a = rng.standard_normal(O.nalm(lmax)) + 1j * rng.standard_normal(O.nalm(lmax))
    a[:lmax + 1] = a[:lmax + 1].real
    a[[0, 1, lmax + 1]] = 0
    return a


@pytest.mark.parametrize("n", [1, 2, 4, 8, 12, 20, 28, 36, 52, 64, 2036, 2048])
def test_ring_fft_matches_numpy(n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    xr, xi = np.ascontiguousarray(x.real), np.ascontiguousarray(x.imag)
    outr, outi = np.empty(n), np.empty(n)
    assert O._fast().orf_test_fft(n, O._p(xr), O._p(xi), O._p(outr), O._p(outi)) == 0
    ref = np.fft.fft(x)
    assert np.abs(outr + 1j * outi - ref).max() <= 2e-15 * max(1.0, np.abs(ref).max()) * max(1, np.log2(max(n, 2)))


@pytest.mark.parametrize("nside,lmax", [(1, 2), (2, 4), (4, 8), (4, 11), (8, 16), (16, 32), (16, 47), (32, 64), (64, 128)])
def test_fast_pair_matches_long_double_oracle(nside, lmax):
    rng = np.random.default_rng(100 * nside + lmax)
    e, b = rand_alm(lmax, rng), rand_alm(lmax, rng)
    q0, u0 = O.alm2map_spin2(e, b, nside, lmax, "ld")
    q1, u1 = O.alm2map_spin2(e, b, nside, lmax, "fast")
    assert max(np.abs(q1 - q0).max(), np.abs(u1 - u0).max()) <= 2e-12 * np.abs(q0).max()
    q, u = rng.standard_normal(12 * nside ** 2), rng.standard_normal(12 * nside ** 2)
    for adjoint in (False, True):
        e0, b0 = O.map2alm_spin2(q, u, nside, lmax, adjoint=adjoint, kind="ld")
        e1, b1 = O.map2alm_spin2(q, u, nside, lmax, adjoint=adjoint, kind="fast")
        assert max(np.abs(e1 - e0).max(), np.abs(b1 - b0).max()) <= 2e-12 * np.abs(e0).max()


def test_fast_pair_at_a_pruned_size_matches_scalar_f64_oracle():
    """NSIDE 256 / lmax 512: m-pruning (libsharp's m_lim rule) is active; the scalar oracle is unpruned."""
    nside, lmax = 256, 512
    rng = np.random.default_rng(7)
    e, b = rand_alm(lmax, rng), rand_alm(lmax, rng)
    q0, u0 = O.alm2map_spin2(e, b, nside, lmax, "f64")
    q1, u1 = O.alm2map_spin2(e, b, nside, lmax, "fast")
    assert max(np.abs(q1 - q0).max(), np.abs(u1 - u0).max()) <= 1e-11 * np.abs(q0).max()
    e0, b0 = O.map2alm_spin2(q0, u0, nside, lmax, adjoint=True, kind="f64")
    e1, b1 = O.map2alm_spin2(q0, u0, nside, lmax, adjoint=True, kind="fast")
    assert max(np.abs(e1 - e0).max(), np.abs(b1 - b0).max()) <= 1e-11 * np.abs(e0).max()
    # adjointness of the pair in the real layout
    from oracle import reference_logic as R
    x = rng.standard_normal((lmax + 1) ** 2), rng.standard_normal((lmax + 1) ** 2)
    for v in x:
        v[[0, 1, lmax + 1, lmax + 2]] = 0
    y = rng.standard_normal(12 * nside ** 2), rng.standard_normal(12 * nside ** 2)
    aq, au = R.synth_pol(x[0], x[1], nside, lmax, "fast")
    te, tb = R.adjoint_pol(y[0], y[1], nside, lmax, 0, "fast")
    lhs, rhs = aq @ y[0] + au @ y[1], x[0] @ te + x[1] @ tb
    assert abs(lhs - rhs) <= 1e-11 * abs(lhs)


@pytest.mark.parametrize("bits", [256, 128])
def test_other_isa_builds_give_the_same_numbers(bits):
    """The AVX2 and baseline builds of the hot loops (ORF_SIMD forces the dispatcher) agree with the default one."""
    code = ("import numpy as np, sys; sys.path.insert(0, %r)\n"
            "from oracle import sht as O\n"
            "rng = np.random.default_rng(3); lmax, nside = 40, 16\n"
            "e = rng.standard_normal(O.nalm(lmax)) + 1j * rng.standard_normal(O.nalm(lmax)); b = e[::-1].copy()\n"
            "q, u = O.alm2map_spin2(e, b, nside, lmax, 'fast'); ee, bb = O.map2alm_spin2(q, u, nside, lmax, kind='fast')\n"
            "print(O._fast().orf_simd_bits(), repr(float(np.abs(q).sum() + np.abs(u).sum())), repr(float(np.abs(ee).sum() + np.abs(bb).sum())))\n" % ROOT)
    outs = []
    for b in (None, bits):
        env = dict(os.environ)
        env.pop("ORF_SIMD", None)
        if b is not None:
            env["ORF_SIMD"] = str(b)
        outs.append(subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, check=True).stdout.split())
    assert int(outs[1][0]) <= bits
    for i in (1, 2):
        a, c = float(outs[0][i]), float(outs[1][i])
        assert abs(a - c) <= 1e-11 * abs(a)


def test_fused_real_layout_entry_points_equal_the_unfused_composition():
    """orf_synth_real / orf_adjoint_real (layout conversion, per-l factors and the pixel weights folded into the SHT
    calls; what the timed CPU arm uses) == real_to_complex / almxfl / N^-1 multiply / complex_to_real around the
    long-double oracle, for the PCG operator, the right-hand side and the non-centred likelihood."""
    from oracle import reference_logic as R
    from tests.test_pncp_oracle import small_problem
    P = small_problem()
    a = R.PolProblem(P["nside"], P["lmax"], P["dQ"], P["dU"], P["mask"] / P["noise0"], P["fwhm"], kind="ld")
    b = R.PolProblem(P["nside"], P["lmax"], P["dQ"], P["dU"], P["mask"] / P["noise0"], P["fwhm"], kind="fast", vectorised=True)
    rng = np.random.default_rng(0)
    n = (P["lmax"] + 1) ** 2
    xE, xB = rng.standard_normal(n), rng.standard_normal(n)
    dls = {k: R.unfold_bins(P["init"][k], P["bins"][k]) for k in ("EE", "BB")}
    ya, yb = a.apply_Q(dls["EE"], dls["BB"], xE, xB), b.apply_Q(dls["EE"], dls["BB"], xE, xB)
    assert max(np.abs(ya[0] - yb[0]).max(), np.abs(ya[1] - yb[1]).max()) <= 1e-12 * np.abs(ya[0]).max()
    s = {"EE": xE, "BB": xB}
    for l_cut in (0, 5):
        la, lb = R.nc_loglik(P["init"], P["bins"], s, a, l_cut), R.nc_loglik(P["init"], P["bins"], s, b, l_cut)
        assert abs(la - lb) <= 1e-12 * abs(la)
    xi = [rng.standard_normal(k) for k in (P["npix"], P["npix"], n, n)]
    ra, rb = a.rhs(dls["EE"], dls["BB"], *xi), b.rhs(dls["EE"], dls["BB"], *xi)
    assert max(np.abs(ra[0] - rb[0]).max(), np.abs(ra[1] - rb[1]).max()) <= 1e-12 * np.abs(ra[0]).max()

"""GPU parity of the sm_100a spherical-harmonic transforms against the CPU oracle
(oracle/sht_oracle.c, long double) on the same seeded inputs; FP64 tolerance 1e-10 relative
(BASELINE.json north_star), plus size-independent properties at the full bench size."""
import numpy as np
import pytest
import torch

from oracle import sht as O

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def rand_alm(lmax, rng, lmin=0):
    a = rng.standard_normal(O.nalm(lmax)) + 1j * rng.standard_normal(O.nalm(lmax))
    a[:lmax + 1] = a[:lmax + 1].real
    ell = np.concatenate([np.arange(m, lmax + 1) for m in range(lmax + 1)])
    a[ell < lmin] = 0
    return a


def relerr(got, ref):
    return float(np.abs(got - ref).max() / np.abs(ref).max())


def to_real(a, lmax):
    r = np.empty((lmax + 1) ** 2)
    r[:lmax + 1] = a[:lmax + 1].real
    r[lmax + 1::2] = a[lmax + 1:].real * np.sqrt(2)
    r[lmax + 2::2] = a[lmax + 1:].imag * np.sqrt(2)
    return r


def dev(x):
    return torch.as_tensor(x, device="cuda")


@pytest.mark.parametrize("nside,lmax", [(1, 2), (2, 5), (4, 8), (8, 16), (16, 47), (32, 64), (64, 128), (128, 256)])
def test_spin0_synthesis_and_analysis_vs_oracle(nside, lmax):
    from gibbssampler_b200.sht import Plan
    plan = Plan.get(nside, lmax)
    rng = np.random.default_rng(10 + nside)
    a = rand_alm(lmax, rng)
    ref = O.alm2map(a, nside, lmax)
    got = plan.alm2map(dev(a)).cpu().numpy()
    assert relerr(got, ref) < RTOL
    got_r = plan.alm2map(dev(to_real(a, lmax))).cpu().numpy()
    assert relerr(got_r, ref) < RTOL
    f = rng.standard_normal(12 * nside ** 2)
    ref_a = O.map2alm(f, nside, lmax)
    got_a = plan.map2alm(dev(f)).cpu().numpy()
    assert relerr(got_a, ref_a) < RTOL
    got_adj = plan.map2alm(dev(f), adjoint=True, real_layout=True).cpu().numpy()
    assert relerr(got_adj, to_real(O.map2alm(f, nside, lmax, adjoint=True), lmax)) < RTOL


@pytest.mark.parametrize("nside,lmax", [(1, 2), (2, 5), (4, 8), (8, 16), (16, 47), (32, 64), (64, 128), (128, 256)])
def test_spin2_synthesis_and_analysis_vs_oracle(nside, lmax):
    from gibbssampler_b200.sht import Plan
    plan = Plan.get(nside, lmax)
    rng = np.random.default_rng(20 + nside)
    e, b = rand_alm(lmax, rng, 2), rand_alm(lmax, rng, 2)
    rq, ru = O.alm2map_spin2(e, b, nside, lmax)
    q, u = plan.alm2map_spin2(dev(e), dev(b))
    assert relerr(q.cpu().numpy(), rq) < RTOL and relerr(u.cpu().numpy(), ru) < RTOL
    q2, u2 = plan.alm2map_spin2(dev(to_real(e, lmax)), dev(to_real(b, lmax)))
    assert relerr(q2.cpu().numpy(), rq) < RTOL and relerr(u2.cpu().numpy(), ru) < RTOL
    fq, fu = rng.standard_normal(12 * nside ** 2), rng.standard_normal(12 * nside ** 2)
    re_, rb_ = O.map2alm_spin2(fq, fu, nside, lmax)
    ge, gb = plan.map2alm_spin2(dev(fq), dev(fu))
    assert relerr(ge.cpu().numpy(), re_) < RTOL and relerr(gb.cpu().numpy(), rb_) < RTOL
    ae, ab = O.map2alm_spin2(fq, fu, nside, lmax, adjoint=True)
    ge, gb = plan.map2alm_spin2(dev(fq), dev(fu), adjoint=True, real_layout=True)
    assert relerr(ge.cpu().numpy(), to_real(ae, lmax)) < RTOL and relerr(gb.cpu().numpy(), to_real(ab, lmax)) < RTOL


def test_iter3_analysis_pixel_weights_and_filter_vs_oracle():
    """utils.adjoint_synthesis_hp semantics: map2alm(iter=3) (utils.py:89) with fused N^-1 and b_l."""
    from gibbssampler_b200.sht import Plan
    nside, lmax = 16, 32
    plan = Plan.get(nside, lmax)
    rng = np.random.default_rng(5)
    fq, fu = rng.standard_normal(12 * nside ** 2), rng.standard_normal(12 * nside ** 2)
    w = rng.uniform(0.0, 2.0, 12 * nside ** 2)
    bl = O.gauss_beam(np.radians(3.0), lmax)
    re_, rb_ = O.map2alm_spin2(fq * w, fu * w, nside, lmax, iter=3)
    re_, rb_ = O.almxfl(re_, bl, lmax), O.almxfl(rb_, bl, lmax)
    ge, gb = plan.map2alm_spin2(dev(fq), dev(fu), iter=3, pixw=dev(w), fl=dev(bl))
    assert relerr(ge.cpu().numpy(), re_) < RTOL and relerr(gb.cpu().numpy(), rb_) < RTOL
    ra = O.almxfl(O.map2alm(fq * w, nside, lmax, iter=3), bl, lmax)
    ga = plan.map2alm(dev(fq), iter=3, pixw=dev(w), fl=dev(bl))
    assert relerr(ga.cpu().numpy(), ra) < RTOL
    # fused almxfl on the synthesis side
    e, b = rand_alm(lmax, rng, 2), rand_alm(lmax, rng, 2)
    rq, ru = O.alm2map_spin2(O.almxfl(e, bl, lmax), O.almxfl(b, bl, lmax), nside, lmax)
    q, u = plan.alm2map_spin2(dev(e), dev(b), fl=dev(bl))
    assert relerr(q.cpu().numpy(), rq) < RTOL and relerr(u.cpu().numpy(), ru) < RTOL


def test_analytic_known_answers_on_gpu():
    from gibbssampler_b200.sht import Plan
    nside, lmax = 8, 16
    plan = Plan.get(nside, lmax)
    th, ph = O.pix_angles(nside)
    a = np.zeros(O.nalm(lmax), complex)
    a[0] = np.sqrt(4 * np.pi)
    assert np.abs(plan.alm2map(dev(a)).cpu().numpy() - 1).max() < 1e-13
    a[:] = 0
    a[O.alm_index(lmax, 1, 1)] = 0.3 + 0.7j
    ref = -2 * np.sqrt(3 / 8 / np.pi) * np.sin(th) * (0.3 * np.cos(ph) - 0.7 * np.sin(ph))
    assert np.abs(plan.alm2map(dev(a)).cpu().numpy() - ref).max() < 1e-13
    e = np.zeros(O.nalm(lmax), complex)
    e[2] = 1
    q, u = plan.alm2map_spin2(dev(e), dev(0 * e))
    assert np.abs(q.cpu().numpy() + 0.25 * np.sqrt(15 / 2 / np.pi) * np.sin(th) ** 2).max() < 1e-13
    assert np.abs(u.cpu().numpy()).max() < 1e-13


@pytest.mark.parametrize("nside,lmax", [(256, 512), (512, 1024), (1024, 2048), (2048, 4096)])
def test_adjointness_at_bench_sizes(nside, lmax):
    """<A x, y> = <x, A^T y> in the real layout; size-independent property, checked at every
    BASELINE.json size (configs #2, #3, #4; nside 2048 exercises the split ring path)."""
    from gibbssampler_b200.sht import Plan
    plan = Plan.get(nside, lmax) if nside <= 512 else Plan(nside, lmax)
    g = torch.Generator(device="cuda").manual_seed(7)
    nre, npix = (lmax + 1) ** 2, 12 * nside ** 2
    x0 = torch.randn(nre, generator=g, device="cuda", dtype=torch.float64)
    y0 = torch.randn(npix, generator=g, device="cuda", dtype=torch.float64)
    lhs = torch.dot(plan.alm2map(x0), y0).item()
    rhs = torch.dot(x0, plan.map2alm(y0, adjoint=True, real_layout=True)).item()
    assert abs(lhs - rhs) < 1e-11 * abs(lhs)
    xe = torch.randn(nre, generator=g, device="cuda", dtype=torch.float64)
    xb = torch.randn(nre, generator=g, device="cuda", dtype=torch.float64)
    # l < 2 carries no spin-2 signal: zero monopole/dipole entries {0, 1, L+1, L+2} (variance_expension.pyx:107-110)
    for v in (xe, xb):
        v[[0, 1, lmax + 1, lmax + 2]] = 0
    yq = torch.randn(npix, generator=g, device="cuda", dtype=torch.float64)
    yu = torch.randn(npix, generator=g, device="cuda", dtype=torch.float64)
    q, u = plan.alm2map_spin2(xe, xb)
    te, tb = plan.map2alm_spin2(yq, yu, adjoint=True, real_layout=True)
    lhs = (torch.dot(q, yq) + torch.dot(u, yu)).item()
    rhs = (torch.dot(xe, te) + torch.dot(xb, tb)).item()
    assert abs(lhs - rhs) < 1e-11 * abs(lhs)


def test_large_size_vs_double_oracle():
    """nside 256 / lmax 512 against the FP64 OpenMP build of the oracle (a few seconds of CPU)."""
    from gibbssampler_b200.sht import Plan
    nside, lmax = 256, 512
    plan = Plan.get(nside, lmax)
    rng = np.random.default_rng(99)
    e, b = rand_alm(lmax, rng, 2), rand_alm(lmax, rng, 2)
    rq, ru = O.alm2map_spin2(e, b, nside, lmax, kind="f64")
    q, u = plan.alm2map_spin2(dev(e), dev(b))
    assert relerr(q.cpu().numpy(), rq) < RTOL and relerr(u.cpu().numpy(), ru) < RTOL
    re_, rb_ = O.map2alm_spin2(rq, ru, nside, lmax, kind="f64")
    ge, gb = plan.map2alm_spin2(q, u)
    assert relerr(ge.cpu().numpy(), re_) < RTOL and relerr(gb.cpu().numpy(), rb_) < RTOL


def test_bench_size_vs_double_oracle():
    """BASELINE config #3 size (nside 512 / lmax 1024): spin-0 and spin-2 synthesis and analysis against the FP64 OpenMP build of
    the oracle on the same seeded inputs (adjointness alone is blind to errors shared by A and A^T)."""
    from gibbssampler_b200.sht import Plan
    nside, lmax = 512, 1024
    plan = Plan.get(nside, lmax)
    rng = np.random.default_rng(512)
    e, b = rand_alm(lmax, rng, 2), rand_alm(lmax, rng, 2)
    rq, ru = O.alm2map_spin2(e, b, nside, lmax, kind="f64")
    q, u = plan.alm2map_spin2(dev(e), dev(b))
    assert relerr(q.cpu().numpy(), rq) < RTOL and relerr(u.cpu().numpy(), ru) < RTOL
    q2, u2 = plan.alm2map_spin2(dev(to_real(e, lmax)), dev(to_real(b, lmax)))
    assert relerr(q2.cpu().numpy(), rq) < RTOL and relerr(u2.cpu().numpy(), ru) < RTOL
    fq, fu = rng.standard_normal(12 * nside ** 2), rng.standard_normal(12 * nside ** 2)
    re_, rb_ = O.map2alm_spin2(fq, fu, nside, lmax, kind="f64")
    ge, gb = plan.map2alm_spin2(dev(fq), dev(fu))
    assert relerr(ge.cpu().numpy(), re_) < RTOL and relerr(gb.cpu().numpy(), rb_) < RTOL
    ae, ab = O.map2alm_spin2(fq, fu, nside, lmax, adjoint=True, kind="f64")
    ge, gb = plan.map2alm_spin2(dev(fq), dev(fu), adjoint=True, real_layout=True)
    assert relerr(ge.cpu().numpy(), to_real(ae, lmax)) < RTOL and relerr(gb.cpu().numpy(), to_real(ab, lmax)) < RTOL
    a = rand_alm(lmax, rng)
    ref = O.alm2map(a, nside, lmax, kind="f64")
    assert relerr(plan.alm2map(dev(a)).cpu().numpy(), ref) < RTOL
    ra = O.map2alm(fq, nside, lmax, kind="f64")
    assert relerr(plan.map2alm(dev(fq)).cpu().numpy(), ra) < RTOL


@pytest.mark.parametrize("nside,lmax", [(1024, 2048), (2048, 4096)])
def test_sampled_legendre_sums_at_large_sizes(nside, lmax):
    """nside 1024 / 2048 (BASELINE config #4), where a full CPU transform is too slow for the suite: alm that live on a sampled
    set of m (0, 1, 2, 3, 17, ~lmax/4, lmax/2, 3 lmax/4, lmax - 96, lmax - 1, lmax) are synthesised by the GPU, and pixels of
    ~2400 rings (the first and last 400 rings one by one, the cap/belt boundaries, the equator, every 5th ring in between) are
    compared with direct sums over the oracle's long-double lambda_lm (tests/sampled_sht.py; the helper itself is pinned to the
    oracle's full transforms on the CPU, tests/test_oracle_sht.py).  The rings around every sampled m's pruning boundary
    m = m_lim(theta) are in the set and the deepest range-extension ladder (m = lmax next to the poles) is exercised.
    Analysis (A^T): maps supported on a few rings, every l of the sampled m."""
    from gibbssampler_b200.sht import Plan
    from tests import sampled_sht as S
    plan = Plan(nside, lmax)
    rng = np.random.default_rng(4096 + nside)
    ms = S.sampled_m(lmax)
    nring = 4 * nside - 1
    rings = sorted(set(range(1, 401)) | set(range(nring - 399, nring + 1)) | set(range(1, nring + 1, 5))
                   | set(range(nside - 3, nside + 4)) | set(range(2 * nside - 2, 2 * nside + 3)) | set(range(3 * nside - 3, 3 * nside + 4)))
    e, b, t, coef = S.sampled_alm(lmax, ms, rng)
    q, u = (x.cpu().numpy() for x in plan.alm2map_spin2(dev(e), dev(b)))
    tm = plan.alm2map(dev(t)).cpu().numpy()
    worst = S.synthesis_error(nside, lmax, ms, coef, rings, q, u, tm, rng)
    assert worst < RTOL, worst
    trings = [1, 2, 37, nside - 1, nside, 2 * nside, 3 * nside + 1, nring - 5, nring]
    fq, fu, geo = S.ring_supported_maps(nside, trings, rng)
    ge, gb = (x.cpu().numpy() for x in plan.map2alm_spin2(dev(fq), dev(fu), adjoint=True))
    gt = plan.map2alm(dev(fq), adjoint=True).cpu().numpy()
    worst = S.analysis_error(nside, lmax, ms, trings, fq, fu, geo, ge, gb, gt)
    assert worst < RTOL, worst


@pytest.mark.parametrize("nside,lmax,mcap", [(16, 32, 32), (32, 64, 64), (64, 160, 128), (128, 256, 256)])
def test_split_ring_path_matches_direct_path(nside, lmax, mcap, monkeypatch):
    """Rings too long for one CTA (nside >= 2048 in production) are transformed as 4 sub-transforms of length
    n/4 through a global scratch buffer; GS_RING_MCAP forces that path at small sizes.  Same results as the
    single-CTA path to rounding, spin 0 and spin 2, both directions."""
    from gibbssampler_b200.sht import Plan
    ref = Plan.get(nside, lmax)
    monkeypatch.setenv("GS_RING_MCAP", str(mcap))
    plan = Plan(nside, lmax)
    monkeypatch.delenv("GS_RING_MCAP")
    rng = np.random.default_rng(3 + nside)
    e, b = rand_alm(lmax, rng, lmin=2), rand_alm(lmax, rng, lmin=2)
    q0, u0 = ref.alm2map_spin2(dev(e), dev(b))
    q1, u1 = plan.alm2map_spin2(dev(e), dev(b))
    assert relerr(q1.cpu().numpy(), q0.cpu().numpy()) < 1e-12 and relerr(u1.cpu().numpy(), u0.cpu().numpy()) < 1e-12
    w = dev(rng.random(12 * nside ** 2) + 0.5)
    fq, fu = dev(rng.standard_normal(12 * nside ** 2)), dev(rng.standard_normal(12 * nside ** 2))
    e0, b0 = ref.map2alm_spin2(fq, fu, pixw=w, iter=1)
    e1, b1 = plan.map2alm_spin2(fq, fu, pixw=w, iter=1)
    assert relerr(e1.cpu().numpy(), e0.cpu().numpy()) < 1e-12 and relerr(b1.cpu().numpy(), b0.cpu().numpy()) < 1e-12
    t0, t1 = ref.alm2map(dev(e)), plan.alm2map(dev(e))
    assert relerr(t1.cpu().numpy(), t0.cpu().numpy()) < 1e-12
    a0, a1 = ref.map2alm(fq, adjoint=True), plan.map2alm(fq, adjoint=True)
    assert relerr(a1.cpu().numpy(), a0.cpu().numpy()) < 1e-12

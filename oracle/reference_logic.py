"""numpy restatement of the reference's Python logic on the hot path (TEST INFRASTRUCTURE ONLY).

Every function cites the reference lines it follows.  healpy calls go to oracle.sht; the forked
qcinv (not vendored, no version recorded -- SURVEY.md 8c) is restated from upstream dhanson/qcinv
semantics: opfilt_pp.fwd_op / calc_prep / pre_op_diag, cd_solve with tr_cg, monitor_basic.
tests/golden/make_golden.py checks these restatements against the reference's own modules
imported with stubbed third-party packages."""
import numpy as np

from . import sht


# ---------------------------------------------------------------- layouts (utils.py / variance_expension.pyx)
def real_to_complex(alms):
    """utils.py:49-60"""
    lmax = int(np.sqrt(len(alms))) - 1
    m_0 = alms[:lmax + 1] + 0j
    m_pos = alms[lmax + 1:]
    m_pos = (m_pos[::2] + 1j * m_pos[1::2]) / np.sqrt(2)
    return np.concatenate([m_0, m_pos])


def complex_to_real(alms):
    """utils.py:63-76"""
    lmax = int((-3 + np.sqrt(9 + 8 * (len(alms) - 1))) / 2)
    out = np.empty((lmax + 1) ** 2)
    out[:lmax + 1] = alms[:lmax + 1].real
    out[lmax + 1::2] = alms[lmax + 1:].real * np.sqrt(2)
    out[lmax + 2::2] = alms[lmax + 1:].imag * np.sqrt(2)
    return out


def generate_var_cl(dls):
    """utils.py:114-147 / variance_expension.pyx:8-33 (takes D_l; l = 0 copied unscaled)"""
    lmax = len(dls) - 1
    size_complex = (lmax + 1) * (lmax + 2) // 2
    alms_shape = np.zeros(size_complex)
    for l in range(lmax + 1):
        for m in range(l + 1):
            idx = m * (2 * lmax + 1 - m) // 2 + l
            alms_shape[idx] = dls[l] if l == 0 else dls[l] * 2 * np.pi / (l * (l + 1))
    var = np.zeros((lmax + 1) ** 2)
    var[:lmax + 1] = alms_shape[:lmax + 1]
    var[lmax + 1::2] = alms_shape[lmax + 1:]
    var[lmax + 2::2] = alms_shape[lmax + 1:]
    return var


def expand_per_l(x):
    """GibbsSampler.py:73 / config.py:81-83"""
    lmax = len(x) - 1
    return np.concatenate([x, np.array([c for m in range(1, lmax + 1) for c in x[m:] for _ in range(2)])])


def unfold_bins(binned, bins):
    """utils.py:150-162"""
    bins = np.asarray(bins)
    return np.repeat(binned, bins[1:] - bins[:-1])


_ELL = {}


def ell_index(lmax):
    """multipole of every entry of the real layout, cached (vectorised twin of the loops above)."""
    if lmax not in _ELL:
        _ELL[lmax] = l_of_real_layout(lmax)
    return _ELL[lmax]


def generate_var_cl_vec(dls):
    """Same values as generate_var_cl (utils.py:114-147), without the Python double loop: what the compiled Cython twin
    (variance_expension.pyx:8-33) costs.  tests/test_golden_oracle.py checks the two agree bit for bit."""
    lmax = len(dls) - 1
    ell = np.arange(lmax + 1, dtype=np.float64)
    cl = np.asarray(dls, dtype=np.float64).copy()
    cl[1:] = cl[1:] * 2 * np.pi / (ell[1:] * (ell[1:] + 1))
    return cl[ell_index(lmax)]


def expand_per_l_vec(x):
    return np.asarray(x)[ell_index(len(x) - 1)]


def safe_inv(v):
    out = np.zeros(len(v))
    out[v != 0] = 1 / v[v != 0]
    return out


# ---------------------------------------------------------------- SHT wrappers in the real layout
def synth_pol(eE, eB, nside, lmax, kind="ld"):
    return sht.alm2map_spin2(real_to_complex(eE), real_to_complex(eB), nside, lmax, kind)


def adjoint_pol(q, u, nside, lmax, iter=0, kind="ld"):
    """(Npix/4pi) map2alm_iter (utils.py:87-100 with iter=3; iter=0 is the exact transpose)"""
    npix = 12 * nside * nside
    e, b = sht.map2alm_spin2(q, u, nside, lmax, iter=iter, kind=kind)
    return complex_to_real(e) * npix / (4 * np.pi), complex_to_real(b) * npix / (4 * np.pi)


class PolProblem:
    """Inputs of the polarised masked-sky CR step (CenteredGibbs.py:243-314)."""

    def __init__(self, nside, lmax, d_Q, d_U, inv_noise_pol, fwhm_deg, kind="ld", vectorised=False):
        """vectorised=True swaps the reference's pure-Python O(L^2) loops (generate_var_cl, the bl_map expansion) for their
        numpy twins -- identical values; used where the restatement is TIMED (bench.py reference arm), so that the CPU
        baseline is not charged for interpreter loops the reference's compiled Cython twin avoids."""
        self.nside, self.lmax, self.kind = nside, lmax, kind
        self.var_cl = generate_var_cl_vec if vectorised else generate_var_cl
        self.expand = expand_per_l_vec if vectorised else expand_per_l
        self.npix = 12 * nside * nside
        self.d_Q, self.d_U = d_Q, d_U
        self.inv_noise = inv_noise_pol
        self.bl_gauss = sht.gauss_beam(np.radians(fwhm_deg), lmax)
        self.bl_map = self.expand(self.bl_gauss)
        # second_part_grad (CenteredGibbs.py:298-308)
        e, b = adjoint_pol(d_Q * inv_noise_pol, d_U * inv_noise_pol, nside, lmax, 0, kind)
        self.bdata_E, self.bdata_B = e * self.bl_map, b * self.bl_map

    def rhs(self, dl_EE, dl_BB, xi_Q, xi_U, xi_E, xi_B, fluct_iter=3):
        """CenteredGibbs.py:469-483 + the data term qcinv's calc_prep adds in chain.sample"""
        ivE, ivB = safe_inv(self.var_cl(dl_EE)), safe_inv(self.var_cl(dl_BB))
        fe, fb = adjoint_pol(xi_Q * np.sqrt(self.inv_noise), xi_U * np.sqrt(self.inv_noise), self.nside, self.lmax,
                             fluct_iter, self.kind)
        bE = fe * self.bl_map + np.sqrt(ivE) * xi_E + self.bdata_E
        bB = fb * self.bl_map + np.sqrt(ivB) * xi_B + self.bdata_B
        return bE, bB

    def apply_Q(self, dl_EE, dl_BB, xE, xB, inv_var=None):
        """qcinv opfilt_pp.fwd_op: C^-1 x + b (Npix/4pi) map2alm0(N^-1 alm2map(b x))"""
        ivE, ivB = inv_var if inv_var is not None else (safe_inv(self.var_cl(dl_EE)), safe_inv(self.var_cl(dl_BB)))
        if self.kind == "fast":   # same operator; the beam, N^-1 and the layout conversions ride inside the two SHT calls
            q, u = sht.synth_real_fast(xE, xB, self.bl_gauss, self.bl_gauss, self.nside, self.lmax)
            e, b = sht.adjoint_real_fast(q, u, self.inv_noise, self.bl_gauss, self.bl_gauss, self.nside, self.lmax)
            e += ivE * xE
            b += ivB * xB
            return e, b
        q, u = synth_pol(xE * self.bl_map, xB * self.bl_map, self.nside, self.lmax, self.kind)
        e, b = adjoint_pol(q * self.inv_noise, u * self.inv_noise, self.nside, self.lmax, 0, self.kind)
        return ivE * xE + e * self.bl_map, ivB * xB + b * self.bl_map

    def precond(self, dl_EE, dl_BB):
        """qcinv opfilt_pp.pre_op_diag: 1 / (1/C_l + b_l^2 sum(N^-1)/(4 pi))"""
        ninv = np.sum(self.inv_noise) / (4 * np.pi)
        out = []
        for dl in (dl_EE, dl_BB):
            cl = dl * np.array([2 * np.pi / (l * (l + 1)) if l else 0 for l in range(self.lmax + 1)])
            cl[0] = dl[0]
            d = safe_inv(cl) + self.bl_gauss ** 2 * ninv
            out.append(self.expand(safe_inv(d)))
        return out

    def pcg(self, dl_EE, dl_BB, bE, bB, eps=1e-5, itermax=4000, x0=None):
        """qcinv cd_solve with tr_cg (= plain PCG) and monitor_basic: stop when <r,r> <= eps^2 <r0,r0>."""
        ME, MB = self.precond(dl_EE, dl_BB)
        b = np.concatenate([bE, bB])
        M = np.concatenate([ME, MB])
        n = len(bE)
        inv_var = (safe_inv(self.var_cl(dl_EE)), safe_inv(self.var_cl(dl_BB)))   # s_cls of the chain, built once per solve
        q, tmp = np.empty(2 * n), np.empty(2 * n)

        def A(v):
            q[:n], q[n:] = self.apply_Q(dl_EE, dl_BB, v[:n], v[n:], inv_var)
            return q
        x = np.zeros(2 * n) if x0 is None else np.concatenate(x0)
        r = b - A(x) if x0 is not None else b.copy()
        d0 = r @ r
        z = M * r
        p = z.copy()
        delta = r @ z
        it, rr = 0, d0
        while it < itermax and rr > eps ** 2 * d0:
            A(p)
            alpha = delta / (p @ q)
            np.multiply(p, alpha, out=tmp); x += tmp          # x += alpha p
            np.multiply(q, alpha, out=tmp); r -= tmp          # r -= alpha q
            np.multiply(M, r, out=z)
            dn = r @ z
            p *= dn / delta; p += z                            # p = z + beta p
            delta = dn
            rr = r @ r
            it += 1
        return x[:n], x[n:], it, np.sqrt(rr / d0) if d0 > 0 else 0.0

    def dense_Q(self, dl_EE, dl_BB):
        n = (self.lmax + 1) ** 2
        Q = np.zeros((2 * n, 2 * n))
        for i in range(2 * n):
            v = np.zeros(2 * n)
            v[i] = 1
            Q[:, i] = np.concatenate(self.apply_Q(dl_EE, dl_BB, v[:n], v[n:]))
        return Q


# ---------------------------------------------------------------- diagonal CR (full sky, isotropic)
def sample_no_mask(dl, bl_map, d_alm, xi, npix, noise0):
    """CenteredGibbs.py:317-353 for one spectrum"""
    inv_var = safe_inv(generate_var_cl(dl))
    sigma = 1 / ((npix / (noise0 * 4 * np.pi)) * bl_map ** 2 + inv_var)
    r = bl_map * ((npix * (1 / noise0) / (4 * np.pi)) * d_alm)
    return sigma * r + xi * np.sqrt(sigma)


def sample_no_mask_nc(dl, bl_map, d_alm, xi, npix, noise0):
    """NonCenteredGibbs.py:138-176 (all_sph) for one spectrum"""
    var = generate_var_cl(dl)
    sigma = 1 / (1 + (1 / noise0) * bl_map ** 2 * var * npix / (4 * np.pi))
    r = np.sqrt(var) * bl_map * ((npix * (1 / noise0) / (4 * np.pi)) * d_alm)
    return sigma * r + xi * np.sqrt(sigma)


def sample_no_mask_nc_pix(dl_EE, dl_BB, bl_map, d_Q, d_U, inv_noise_pol, xi_E, xi_B, nside, lmax, kind="ld"):
    """PolarizedNonCenteredConstrainedRealization.sample_no_mask, pixel-domain branch (NonCenteredGibbs.py:141-176 with
    all_sph False, :155-160): r = sqrt(C) b (Npix/4pi) complex_to_real(map2alm([0, Q N^-1, U N^-1], iter=3 (healpy default)))."""
    npix = 12 * nside * nside
    out = []
    rE, rB = adjoint_pol(d_Q * inv_noise_pol, d_U * inv_noise_pol, nside, lmax, 3, kind)
    for dl, r, xi in ((dl_EE, rE, xi_E), (dl_BB, rB, xi_B)):
        var = generate_var_cl(dl)
        sigma = 1 / (1 + inv_noise_pol[0] * bl_map ** 2 * var * npix / (4 * np.pi))
        out.append(sigma * (np.sqrt(var) * bl_map * r) + xi * np.sqrt(sigma))
    return out


def ula_no_mask(dl_EE, dl_BB, bl_map, d_E, d_B, s_old, npix, noise0, tau, xi_E, xi_B):
    """ULA_no_mask (CenteredGibbs.py:417-446) with compute_gradient_no_mask (:355-377), compute_log_proposal_no_mask
    (:379-392) and compute_log_density_no_mask (:394-414): returns (s_new dict, log_ratio)."""
    sig, mean = {}, {}
    for pol, dl, d in (("EE", dl_EE, d_E), ("BB", dl_BB, d_B)):
        iv = safe_inv(generate_var_cl(dl))
        sig[pol] = 1 / ((npix / (noise0 * 4 * np.pi)) * bl_map ** 2 + iv)
        mean[pol] = sig[pol] * (bl_map * ((npix * (1 / noise0) / (4 * np.pi)) * d))
    grad = lambda s: {p: -(1 / sig[p]) * (s[p] - mean[p]) for p in ("EE", "BB")}
    logq = lambda to, frm: sum(-0.5 * np.sum((to[p] - frm[p] - tau * sig[p] * grad(frm)[p]) ** 2 / (2 * tau * sig[p])) for p in ("EE", "BB"))
    logd = lambda s: sum(-0.5 * np.sum((s[p] - mean[p]) ** 2 / sig[p]) for p in ("EE", "BB"))
    g = grad(s_old)
    s_new = {"EE": s_old["EE"] + tau * sig["EE"] * g["EE"] + np.sqrt(2 * tau * sig["EE"]) * xi_E,
             "BB": s_old["BB"] + tau * sig["BB"] * g["BB"] + np.sqrt(2 * tau * sig["BB"]) * xi_B}
    return s_new, logd(s_new) + logq(s_old, s_new) - (logd(s_old) + logq(s_new, s_old))


# ---------------------------------------------------------------- C_l conditional (CenteredGibbs.py:54-79)
def cls_alpha_beta(alms_real, bins, lmax):
    observed = sht.alm2cl(real_to_complex(alms_real), lmax)
    exponent = np.array([(2 * l + 1) / 2 for l in range(lmax + 1)])
    betas = np.array([(2 * l + 1) * l * (l + 1) * (c / (4 * np.pi)) for l, c in enumerate(observed)])
    ba, bb = [], []
    for i, l in enumerate(bins[:-1]):
        bb.append(np.sum(betas[l:bins[i + 1]]))
        ba.append(np.sum(exponent[l:bins[i + 1]]) - 1)
    ba[0] = 1
    return np.array(ba), np.array(bb)


def cls_sample(alms_real, bins, lmax, gamma_draws):
    """D = beta * invgamma.rvs(alpha) = beta / Gamma(alpha,1); D[:2] = 0"""
    a, b = cls_alpha_beta(alms_real, bins, lmax)
    d = b / gamma_draws
    d[:2] = 0
    return d


# ---------------------------------------------------------------- non-centred likelihood (NonCenteredGibbs.py:333-355)
def nc_loglik(binned, bins, s_nc, prob, l_cut=0):
    """NonCenteredGibbs.py:333-355.  l_cut > 0 (partially non-centred parametrisation, see PNCPPol below): the
    multipoles l < l_cut of `s_nc` are the centred coefficients and are synthesised without the sqrt(C_l) factor."""
    dlE, dlB = unfold_bins(binned["EE"], bins["EE"]), unfold_bins(binned["BB"], bins["BB"])
    var_cl = getattr(prob, "var_cl", generate_var_cl)
    vE, vB = var_cl(dlE), var_cl(dlB)
    if prob.kind == "fast":   # the filter b_l sqrt(C_l) is a per-l factor: hand it to the fused synthesis
        lm = prob.lmax
        fE, fB = np.sqrt(vE[:lm + 1]), np.sqrt(vB[:lm + 1])      # entries 0..lmax of the real layout are the m = 0 column
        fE[:l_cut], fB[:l_cut] = 1.0, 1.0
        q, u = sht.synth_real_fast(s_nc["EE"], s_nc["BB"], prob.bl_gauss * fE, prob.bl_gauss * fB, prob.nside, lm)
        return -0.5 * sht.chi2_fast(prob.d_Q, prob.d_U, q, u, prob.inv_noise)
    fE, fB = np.sqrt(vE), np.sqrt(vB)
    if l_cut > 0:
        low = ell_index(prob.lmax) < l_cut
        fE, fB = np.where(low, 1.0, fE), np.where(low, 1.0, fB)
    q, u = synth_pol(prob.bl_map * fE * s_nc["EE"], prob.bl_map * fB * s_nc["BB"], prob.nside, prob.lmax, prob.kind)
    return -0.5 * (np.sum((prob.d_Q - q) ** 2 * prob.inv_noise) + np.sum((prob.d_U - u) ** 2 * prob.inv_noise))


# ---------------------------------------------------------------- blocked Metropolis-within-Gibbs (NonCenteredGibbs.py:292-330, 401-445)
def mwg_propose(old, proposal_variances):
    """propose_dl for one spectrum (NonCenteredGibbs.py:292-309 / ClsSampler.py:79-92): truncated normal on [0, inf) centred on
    the current value, bins 0 and 1 pinned to 0; scipy draws from numpy's global stream."""
    from scipy.stats import truncnorm
    sc = np.sqrt(proposal_variances)
    return np.concatenate([np.zeros(2), truncnorm.rvs(a=-old[2:] / sc, b=np.inf, loc=old[2:], scale=sc)])


def mwg_log_proposal(x, frm, proposal_variances):
    """compute_log_proposal for one spectrum (NonCenteredGibbs.py:313-330): log q(frm -> x) per bin."""
    from scipy.stats import truncnorm
    sc = np.sqrt(proposal_variances)
    return np.concatenate([np.zeros(2), truncnorm.logpdf(x[2:], a=-frm[2:] / sc, b=np.inf, loc=frm[2:], scale=sc)])


def mwg_sweep(prob, bins, blocks, proposal_variances, s_nc, binned_old, l_cut=0, n_iter=1):
    """PolarizationNonCenteredClsSampler.sample (NonCenteredGibbs.py:401-445): propose every bin of EE then BB, then accept
    / reject block by block (EE blocks first) with  log r = sum_block [log q(new -> old) - log q(old -> new)] + lik(new) -
    lik(old);  one uniform per test from numpy's global stream.  Returns (binned, accept lists)."""
    cur = {p: np.array(binned_old[p], dtype=np.float64) for p in ("EE", "BB")}
    prop = {p: mwg_propose(cur[p], proposal_variances[p]) for p in ("EE", "BB")}
    logr = {p: mwg_log_proposal(cur[p], prop[p], proposal_variances[p]) - mwg_log_proposal(prop[p], cur[p], proposal_variances[p])
            for p in ("EE", "BB")}
    old_lik = nc_loglik(cur, bins, s_nc, prob, l_cut)
    accept = {"EE": [], "BB": []}
    for pol in ("EE", "BB"):
        bl = blocks[pol]
        for i in range(len(bl) - 1):
            b0, b1 = int(bl[i]), int(bl[i + 1])
            for _ in range(n_iter):
                cand = {p: cur[p].copy() for p in cur}
                cand[pol][b0:b1] = prop[pol][b0:b1]
                new_lik = nc_loglik(cand, bins, s_nc, prob, l_cut)
                log_r = np.sum(logr[pol][b0:b1]) + new_lik - old_lik
                if np.log(np.random.uniform()) < log_r:
                    cur, old_lik = cand, new_lik
                    accept[pol].append(1)
                else:
                    accept[pol].append(0)
    return cur, accept


class PNCPPol:
    """Polarised, masked-sky partially non-centred Gibbs iteration (BASELINE config #3).

    The reference ships PNCP only as TT / full-sky bytecode (__pycache__/PNCP.cpython-38.pyc), so there is no reference
    implementation of this sampler: it is DEFINED as the composition of reference pieces below, and this class is the
    deterministic CPU statement of that definition which gibbssampler_b200/PNCP.py must reproduce draw for draw:
      1. CR: the centred PCG draw of CenteredGibbs.py:448-491 (PolProblem.rhs + .pcg, eps 1e-5);
      2. low l: inverse-gamma draw of CenteredGibbs.py:54-79, kept only for the bins below l_cut
         (recovered PNCPClsSampler.sample_low_l);
      3. s -> mixed variable: s_l for l < l_cut, C_l^-1/2 s_l for l >= l_cut (generalises NonCenteredGibbs.py:192-194 and
         the recovered compute_var_high_low / PNCPConstrainedRealization.sample);
      4. high l: the blocked Metropolis-within-Gibbs sweep of NonCenteredGibbs.py:401-445 on the bins >= l_cut with the
         pixel-space likelihood of :333-355 (recovered PNCPClsSampler.sample_high_l).
    Random numbers come from numpy's legacy global stream in exactly this order: xi_Q, xi_U, xi_E, xi_B (normal), the
    EE then BB inverse-gamma draws (scipy), the EE then BB truncated-normal proposals (scipy), one uniform per test."""

    def __init__(self, prob, bins, blocks, proposal_variances, l_cut, n_iter=1, eps=1e-5):
        self.prob, self.bins, self.blocks, self.pv = prob, bins, blocks, proposal_variances
        self.l_cut, self.n_iter, self.eps = int(l_cut), n_iter, eps
        self.low_bins = {p: int(np.searchsorted(np.asarray(bins[p]), l_cut, side="left")) for p in ("EE", "BB")}
        self.last_pcg_iterations = 0

    def factor(self, dl, inverse):
        """per-coefficient factor of step 3 (inverse=True: C^-1/2, else C^1/2) on l >= l_cut, 1 below"""
        v = self.prob.var_cl(dl)
        f = np.sqrt(safe_inv(v)) if inverse else np.sqrt(v)
        return np.where(ell_index(self.prob.lmax) < self.l_cut, 1.0, f)

    def iteration(self, binned):
        from scipy.stats import invgamma
        p, lmax = self.prob, self.prob.lmax
        npix, nre = p.npix, (lmax + 1) ** 2
        dls = {k: unfold_bins(binned[k], self.bins[k]) for k in ("EE", "BB")}
        xi = [np.random.normal(loc=0, scale=1, size=n) for n in (npix, npix, nre, nre)]
        bE, bB = p.rhs(dls["EE"], dls["BB"], *xi)
        sE, sB, it, _ = p.pcg(dls["EE"], dls["BB"], bE, bB, eps=self.eps)
        self.last_pcg_iterations = it
        sky = {"EE": sE, "BB": sB}
        new = {}
        for pol in ("EE", "BB"):                                    # step 2
            al, be = cls_alpha_beta(sky[pol], self.bins[pol], lmax)
            draw = be * invgamma.rvs(a=al)
            draw[:2] = 0
            o = np.array(binned[pol], dtype=np.float64)
            o[:self.low_bins[pol]] = draw[:self.low_bins[pol]]
            new[pol] = o
        dls = {k: unfold_bins(new[k], self.bins[k]) for k in ("EE", "BB")}
        mixed = {k: sky[k] * self.factor(dls[k], True) for k in ("EE", "BB")}    # step 3
        out, accept = mwg_sweep(p, self.bins, self.blocks, self.pv, mixed, new, self.l_cut, self.n_iter)   # step 4
        return out, accept, sky


def nc_loglik_all_sph(binned, bins, s_nc, d_E, d_B, bl_map, inv_noise0, npix):
    """compute_log_likelihood_all_sph (NonCenteredGibbs.py:357-377): full sky, isotropic noise, harmonic-space data."""
    dlE, dlB = unfold_bins(binned["EE"], bins["EE"]), unfold_bins(binned["BB"], bins["BB"])
    vE, vB = generate_var_cl(dlE), generate_var_cl(dlB)
    w = inv_noise0 * npix / (4 * np.pi)
    return -0.5 * (np.sum((d_E - bl_map * np.sqrt(vE) * s_nc["EE"]) ** 2 * w) + np.sum((d_B - bl_map * np.sqrt(vB) * s_nc["BB"]) ** 2 * w))


# ---------------------------------------------------------------- per-l 3x3 TT/TE/EE/BB machinery
def expand_var_cl_3x3(dls):
    """variance_expension.pyx:36-61 with the :51 index bug fixed (cls_[l], not cls_[idx]): (L+1,3,3) D_l ->
    ((L+1)^2,3,3) C_l over the real alm layout; l = 0 copied unscaled."""
    lmax = len(dls) - 1
    size_complex = (lmax + 1) * (lmax + 2) // 2
    alms_shape = np.zeros((size_complex, 3, 3))
    for l in range(lmax + 1):
        for m in range(l + 1):
            idx = m * (2 * lmax + 1 - m) // 2 + l
            alms_shape[idx] = dls[l] if l == 0 else dls[l] * 2 * np.pi / (l * (l + 1))
    variance = np.zeros(((lmax + 1) ** 2, 3, 3))
    variance[:lmax + 1] = alms_shape[:lmax + 1]
    for i in range(lmax + 1, size_complex):
        variance[2 * i - (lmax + 1)] = alms_shape[i]
        variance[2 * i - (lmax + 1) + 1] = alms_shape[i]
    return variance


def compute_inverse_and_cholesky(all_cls, pix_part_variance):
    """utils.compute_inverse_and_cholesky recovered from __pycache__/utils.cpython-38.pyc: for l >= 2
    M = blockdiag(inv(C[:2,:2]), 1/C[2,2]) + diag(pix_part); Sigma = inv(M); L = chol(Sigma)."""
    lmax = len(all_cls) - 1
    sig = np.zeros((lmax + 1, 3, 3))
    cho = np.zeros((lmax + 1, 3, 3))
    for l in range(2, lmax + 1):
        m = np.zeros((3, 3))
        m[:2, :2] = np.linalg.inv(all_cls[l, :2, :2])
        m[2, 2] = 1.0 / all_cls[l, 2, 2]
        m += np.diag(pix_part_variance[l])
        sig[l] = np.linalg.inv(m)
        cho[l] = np.linalg.cholesky(sig[l])
    return sig, cho


def l_of_real_layout(lmax):
    """multipole of every entry of the real layout (utils.py:49-76 ordering)."""
    out = np.empty((lmax + 1) ** 2, dtype=np.int64)
    out[:lmax + 1] = np.arange(lmax + 1)
    pos = lmax + 1
    for m in range(1, lmax + 1):
        ls = np.arange(m, lmax + 1)
        out[pos:pos + 2 * ls.size:2] = ls
        out[pos + 1:pos + 2 * ls.size:2] = ls
        pos += 2 * ls.size
    return out


def matrix_product(mats, b):
    """utils.matrix_product recovered from bytecode: per-l (L+1,3,3) matrices applied to ((L+1)^2,3) vectors."""
    lmax = len(mats) - 1
    ell = l_of_real_layout(lmax)
    return np.einsum("iab,ib->ia", mats[ell], b)


def invwishart_bartlett(cl_tt, cl_te, cl_ee, draws):
    """IW(df = 2l-2, scale = (2l+1) Chat_l) for the (TT,TE;TE,EE) block (.ipynb_checkpoints/main-checkpoint.py:333-346)
    from supplied (chi2_df, chi2_{df-1}, N(0,1)) per l: X^-1 = (G A)(G A)^T with Psi^-1 = G G^T."""
    lmax = len(cl_tt) - 1
    out = np.zeros((lmax + 1, 2, 2))
    for l in range(2, lmax + 1):
        psi = (2 * l + 1) * np.array([[cl_tt[l], cl_te[l]], [cl_te[l], cl_ee[l]]])
        g = np.linalg.cholesky(np.linalg.inv(psi))
        a = np.array([[np.sqrt(draws[l, 0]), 0.0], [draws[l, 2], np.sqrt(draws[l, 1])]])
        h = g @ a
        out[l] = np.linalg.inv(h @ h.T)
    return out


# ---------------------------------------------------------------- temperature-only twins (CenteredGibbs.py:95-235, NonCenteredGibbs.py:17-101)
class TTProblem:
    """Inputs of the TT constrained-realization step (ConstrainedRealization.py:8-49)."""

    def __init__(self, nside, lmax, d, inv_noise, fwhm_deg, kind="ld"):
        self.nside, self.lmax, self.kind = nside, lmax, kind
        self.npix = 12 * nside * nside
        self.d, self.inv_noise = d, inv_noise
        self.bl_gauss = sht.gauss_beam(np.radians(fwhm_deg), lmax)
        self.bl_map = expand_per_l(self.bl_gauss)
        self.resc = self.npix / (4 * np.pi)
        self.bdata = self.adjoint(d * inv_noise, 0) * self.bl_map   # what qcinv's calc_prep adds in chain.sample

    def adjoint(self, m, iter):
        """(Npix/4pi) complex_to_real(map2alm(iter)) -- utils.adjoint_synthesis_hp (utils.py:101-111, iter = 3)"""
        return complex_to_real(sht.map2alm(m, self.nside, self.lmax, iter=iter, kind=self.kind)) * self.resc

    def synth(self, a):
        return sht.alm2map(real_to_complex(a), self.nside, self.lmax, self.kind)

    def rhs(self, var_cls, xi_alm, xi_pix, fluct_iter=3):
        """CenteredGibbs.py:145-147 (alm draw first) + the data term added by qcinv"""
        iv = safe_inv(var_cls)
        return xi_alm * np.sqrt(iv) + self.bl_map * self.adjoint(xi_pix * np.sqrt(self.inv_noise), fluct_iter) + self.bdata

    def apply_Q(self, var_cls, x):
        """qcinv opfilt_tt.fwd_op: C^-1 x + b A^T N^-1 A b x"""
        return safe_inv(var_cls) * x + self.bl_map * self.adjoint(self.synth(self.bl_map * x) * self.inv_noise, 0)

    def pcg(self, var_cls, b, eps=1e-6, itermax=4000, x0=None):
        cl = var_cls[:self.lmax + 1]
        M = expand_per_l(safe_inv(safe_inv(cl) + self.bl_gauss ** 2 * np.sum(self.inv_noise) / (4 * np.pi)))
        x = np.zeros(len(b)) if x0 is None else x0.copy()
        r = b - self.apply_Q(var_cls, x) if x0 is not None else b.copy()
        d0 = r @ r
        z = M * r
        p = z.copy()
        delta = r @ z
        it = 0
        while it < itermax and r @ r > eps ** 2 * d0:
            q = self.apply_Q(var_cls, p)
            alpha = delta / (p @ q)
            x += alpha * p
            r -= alpha * q
            z = M * r
            dn = r @ z
            p = z + (dn / delta) * p
            delta = dn
            it += 1
        return x, it

    def sample_no_mask(self, var_cls, xi_alm, xi_pix):
        """CenteredGibbs.py:100-127"""
        iv = safe_inv(var_cls)
        b_w = self.bl_map * self.adjoint(self.inv_noise * self.d, 3)
        b_f = xi_alm * np.sqrt(iv) + self.bl_map * self.adjoint(xi_pix * np.sqrt(self.inv_noise), 3)
        sigma = 1 / (iv + self.inv_noise[0] * self.resc * self.bl_map ** 2)
        return sigma * b_w + sigma * b_f

    def sample_no_mask_nc(self, var_cls, xi_alm, xi_pix):
        """NonCenteredGibbs.py:22-41"""
        b_w = np.sqrt(var_cls) * self.bl_map * self.adjoint(self.d * self.inv_noise, 3)
        b_f = xi_alm + np.sqrt(var_cls) * self.bl_map * self.adjoint(xi_pix * np.sqrt(self.inv_noise), 3)
        sigma = 1 / (1 + var_cls * self.inv_noise[0] * self.resc * self.bl_map ** 2)
        return sigma * b_w + sigma * b_f

    def sample_pncp(self, var_cls, xi_alm, xi_pix, l_cut):
        """PNCPConstrainedRealization.sample recovered from PNCP.cpython-38.pyc"""
        ell = l_of_real_layout(self.lmax)
        nc = ell >= l_cut
        var_low, var_high = var_cls.copy(), var_cls.copy()
        var_low[nc] = 1
        var_high[~nc] = 1
        inv_low = np.zeros(len(var_cls))
        ok = np.ones(len(var_cls), bool)
        ok[[0, 1, self.lmax + 1, self.lmax + 2]] = False
        inv_low[ok] = 1 / var_low[ok]
        b_w = np.sqrt(var_high) * self.bl_map * self.adjoint(self.d * self.inv_noise, 3)
        b_f = xi_alm * np.sqrt(inv_low) + np.sqrt(var_high) * self.bl_map * self.adjoint(xi_pix * np.sqrt(self.inv_noise), 3)
        sigma = 1 / (inv_low + var_high * self.inv_noise[0] * self.resc * self.bl_map ** 2)
        out = sigma * (b_w + b_f)
        out[[0, 1, self.lmax + 1, self.lmax + 2]] = 0
        return out

    def loglik(self, var_cls, s_nc):
        """ClsSampler.py:94-108"""
        return -0.5 * np.sum((self.d - self.synth(self.bl_map * np.sqrt(var_cls) * s_nc)) ** 2 * self.inv_noise)

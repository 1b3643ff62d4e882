"""Alternative constrained-realization kernels of PolarizedCenteredConstrainedRealization
(CenteredGibbs.py:494-825): auxiliary-variable Gibbs, its over-relaxed variant and MALA.  Each is a
composition of the spin-2 SHT pair with fused elementwise / reduction kernels; state stays in HBM.
RNG order follows the reference (Q, U pixel draws, then EE, BB alm draws)."""
import numpy as np
import torch

from . import _dev, _lib, utils
from ._dev import f64, ptr, stream
from ._lib import check

_S = {}


def _scratch(dev):
    if dev not in _S:
        _S[dev] = (torch.empty(592, dtype=torch.float64, device=dev), torch.empty(1, dtype=torch.float64, device=dev))
    return _S[dev]


def _v_step(cr, s, v, alpha):
    """v | s for Q and U; returns the maps v + N^-1 d (CenteredGibbs.py:692-705, 786-797)."""
    L = _lib.lib()
    mq, mu_ = cr.plan.alm2map_spin2(s["EE"], s["BB"], fl=cr.bl_gauss_d)
    outs = []
    for comp, m, d in (("Q", mq, cr.d_Q), ("U", mu_, cr.d_U)):
        xi = cr.rng.normal(cr.Npix)
        out = torch.empty_like(m)
        check(L.gs_aux_v_update(ptr(m), ptr(cr.inv_noise_pol), ptr(d), ptr(xi), cr.mu, alpha, ptr(v[comp]), ptr(out), cr.Npix, stream()))
        outs.append(out)
    return outs


def _s_step(cr, dle, dlb, maps, s, alpha):
    """s | v for EE and BB (CenteredGibbs.py:708-727, 763-783)."""
    L = _lib.lib()
    be, bb = cr.plan.map2alm_spin2(maps[0], maps[1], adjoint=True, fl=cr.bl_gauss_d, real_layout=True)
    w = 4 * np.pi / cr.Npix
    for pol, badj, dl in (("EE", be, dle), ("BB", bb, dlb)):
        xi = cr.rng.normal(cr.dimension_alm)
        check(L.gs_aux_s_update(ptr(badj), ptr(dl), ptr(cr.bl_gauss_d), ptr(xi), cr.mu / w, alpha, cr.lmax, ptr(s[pol]), stream()))


def sample_gibbs_change_variable(cr, all_dls, old_s):
    """Auxiliary-variable CR step (CenteredGibbs.py:676-729)."""
    cr._single_gpu_only("sample_gibbs_change_variable")
    dle, dlb = cr._dls(all_dls)
    s = {"EE": f64(old_s["EE"]).clone(), "BB": f64(old_s["BB"]).clone()}
    v = {"Q": torch.zeros(cr.Npix, dtype=torch.float64, device=cr.dev), "U": torch.zeros(cr.Npix, dtype=torch.float64, device=cr.dev)}
    for _ in range(cr.n_gibbs):
        maps = _v_step(cr, s, v, 0.0)
        _s_step(cr, dle, dlb, maps, s, 0.0)
    return cr._ret(s, old_s["EE"]), 1


def overrelaxation_sampler(cr, all_dls, old_s):
    """Over-relaxed auxiliary-variable step (CenteredGibbs.py:733-825): v|s plain, then n_gibbs x (s|v, v|s, s|v)."""
    cr._single_gpu_only("overrelaxation_sampler")
    dle, dlb = cr._dls(all_dls)
    s = {"EE": f64(old_s["EE"]).clone(), "BB": f64(old_s["BB"]).clone()}
    v = {"Q": torch.zeros(cr.Npix, dtype=torch.float64, device=cr.dev), "U": torch.zeros(cr.Npix, dtype=torch.float64, device=cr.dev)}
    maps = _v_step(cr, s, v, 0.0)
    for _ in range(cr.n_gibbs):
        _s_step(cr, dle, dlb, maps, s, cr.alpha)
        maps = _v_step(cr, s, v, cr.alpha)
        _s_step(cr, dle, dlb, maps, s, cr.alpha)
    return cr._ret(s, old_s["EE"]), 1


def _grad_and_pix(cr, dle, dlb, s):
    """compute_gradient_mala (CenteredGibbs.py:494-520): gradient of the log density and the beamed maps."""
    L = _lib.lib()
    mq, mu_ = cr.plan.alm2map_spin2(s["EE"], s["BB"], fl=cr.bl_gauss_d)
    ye, yb = cr.plan.map2alm_spin2(mq, mu_, adjoint=True, pixw=cr.inv_noise_pol, fl=cr.bl_gauss_d, real_layout=True)
    g = {}
    for pol, dl, y, bd in (("EE", dle, ye, cr.second_part_grad_E), ("BB", dlb, yb, cr.second_part_grad_B)):
        invc = utils.expand_per_l(dl, 2)
        out = torch.empty_like(y)
        check(L.gs_mala_grad(ptr(bd), ptr(invc), ptr(s[pol]), ptr(y), ptr(out), y.numel(), stream()))
        g[pol] = out
    return g, (mq, mu_)


def compute_gradient_mala(cr, all_dls, s_old):
    """compute_gradient_mala (CenteredGibbs.py:494-520) with the reference's return tuple (grad_E, grad_B, Q map, U map)."""
    dle, dlb = cr._dls(all_dls)
    s = {"EE": f64(s_old["EE"]), "BB": f64(s_old["BB"])}
    g, (mq, mu_) = _grad_and_pix(cr, dle, dlb, s)
    host = not isinstance(s_old["EE"], torch.Tensor)
    outs = (g["EE"], g["BB"], mq, mu_)
    return tuple(o.cpu().numpy() for o in outs) if host else outs


def compute_log_density(cr, all_dls, s, s_Q_pix=None, s_U_pix=None):
    """compute_log_density (CenteredGibbs.py:534-558) -> python float."""
    dle, dlb = cr._dls(all_dls)
    sd = {"EE": f64(s["EE"]), "BB": f64(s["BB"])}
    if s_Q_pix is None or s_U_pix is None:
        pix = cr.plan.alm2map_spin2(sd["EE"], sd["BB"], fl=cr.bl_gauss_d)
    else:
        pix = (f64(s_Q_pix), f64(s_U_pix))
    return _log_density(cr, dle, dlb, sd, pix)


def _dot3(a, b, c):
    scratch, out = _scratch(a.device)
    check(_lib.lib().gs_dot3(ptr(a), ptr(b), ptr(c), a.numel(), ptr(scratch), ptr(out), stream()))
    return float(out.item())


def _log_density(cr, dle, dlb, s, pix):
    """compute_log_density (CenteredGibbs.py:534-558)."""
    t = 0.0
    for pol, dl, bd in (("EE", dle, cr.second_part_grad_E), ("BB", dlb, cr.second_part_grad_B)):
        t += -0.5 * _dot3(s[pol], s[pol], utils.expand_per_l(dl, 2)) + _dot3(s[pol], bd, None)
    for m in pix:
        t += -0.5 * _dot3(m, m, cr.inv_noise_pol)
    return t


def _log_q(cr, to, frm, g_from, sigma):
    """compute_log_proposal (CenteredGibbs.py:530-532)."""
    scratch, out = _scratch(cr.dev)
    t = 0.0
    for pol in ("EE", "BB"):
        check(_lib.lib().gs_mala_logq(ptr(to[pol]), ptr(frm[pol]), ptr(g_from[pol]), ptr(sigma[pol]), cr.tau, to[pol].numel(),
                                      ptr(scratch), ptr(out), stream()))
        t += float(out.item())
    return t


def sample_mala(cr, all_dls, s_old):
    """Preconditioned MALA step (CenteredGibbs.py:560-603)."""
    cr._single_gpu_only("sample_mala")
    L = _lib.lib()
    dle, dlb = cr._dls(all_dls)
    so = {"EE": f64(s_old["EE"]), "BB": f64(s_old["BB"])}
    w = cr.Npix / (cr.noise_pol0 * 4 * np.pi)
    sigma = {}
    for pol, dl in (("EE", dle), ("BB", dlb)):
        o = torch.empty(cr.dimension_alm, dtype=torch.float64, device=cr.dev)
        check(L.gs_mala_sigma(ptr(dl), ptr(cr.bl_gauss_d), w, cr.lmax, ptr(o), stream()))
        sigma[pol] = o
    g_old, pix_old = _grad_and_pix(cr, dle, dlb, so)
    sn = {}
    for pol in ("EE", "BB"):                                   # RNG order: EE then BB (CenteredGibbs.py:524-525)
        xi = cr.rng.normal(cr.dimension_alm)
        o = torch.empty_like(so[pol])
        check(L.gs_mala_propose(ptr(so[pol]), ptr(g_old[pol]), ptr(sigma[pol]), ptr(xi), cr.tau, ptr(o), o.numel(), stream()))
        sn[pol] = o
    g_new, pix_new = _grad_and_pix(cr, dle, dlb, sn)
    log_ratio = (_log_density(cr, dle, dlb, sn, pix_new) + _log_q(cr, so, sn, g_new, sigma)
                 - _log_density(cr, dle, dlb, so, pix_old) - _log_q(cr, sn, so, g_old, sigma))
    u = float(cr.rng.uniform(2)[0].item()) if cr.rng.mode == "philox" else np.random.uniform()
    if np.log(u) < log_ratio:
        return cr._ret(sn, s_old["EE"]), 1
    return s_old, 0

"""Stage times of the PCG mat-vec (gs_profile_matvec) at NSIDE 512 / lmax 1024 for isotropic noise under different masks:
random weights (every ring on the transform path), an axisymmetric band, and the galactic-plane-like mask of bench.py."""
import ctypes as C, sys, numpy as np, torch
sys.path.insert(0, '.')
from gibbssampler_b200 import _dev, _lib
from gibbssampler_b200.sht import Plan
import bench
nside, lmax = 512, 1024
L = _lib.lib()
plan = Plan.get(nside, lmax)
nre, npix = (lmax+1)**2, 12*nside**2
g = torch.Generator(device='cuda').manual_seed(1)
xe = torch.randn(nre, generator=g, device='cuda', dtype=torch.float64); xb = torch.randn(nre, generator=g, device='cuda', dtype=torch.float64)
ye, yb = torch.empty_like(xe), torch.empty_like(xb)
bl = torch.ones(lmax+1, device='cuda', dtype=torch.float64)
ms = (C.c_float*4)()
ws = {"random": torch.rand(npix, generator=g, device='cuda', dtype=torch.float64)}
for kind in ("band", "galplane"):
    ws[kind] = torch.as_tensor(bench.make_mask(nside, 0.8, kind) * 25.0, device='cuda', dtype=torch.float64)
for name, w in ws.items():
    for const in (1, 0):
        old = L.gs_set_ring_const(const)
        for n in (3, 20):
            _lib.check(L.gs_profile_matvec(plan._h, _dev.ptr(xe), _dev.ptr(xb), _dev.ptr(bl), _dev.ptr(w), _dev.ptr(ye), _dev.ptr(yb), n, ms, _dev.stream()))
        L.gs_set_ring_const(old)
        nc, act, tot = C.c_int(), C.c_int(), C.c_int()
        L.gs_constant_rings(plan._h, C.byref(nc)); L.gs_active_ring_pairs(plan._h, C.byref(act), C.byref(tot))
        print("%-9s ring_const=%d leg_synth %.3f ring %.3f leg_anal %.3f total %.3f  constant rings %d/%d active pairs %d/%d fsky %.3f chk %.9e" % (
            name, const, ms[0], ms[1] + ms[2], ms[3], sum(ms), nc.value, 4*nside-1, act.value, tot.value, float((w != 0).double().mean()), float(ye.abs().sum())))

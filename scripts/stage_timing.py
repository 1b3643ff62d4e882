import ctypes as C, sys, numpy as np, torch
sys.path.insert(0, '.')
from gibbssampler_b200 import _dev, _lib
from gibbssampler_b200.sht import Plan
nside, lmax = 512, 1024
L = _lib.lib()
plan = Plan.get(nside, lmax)
nre, npix = (lmax+1)**2, 12*nside**2
g = torch.Generator(device='cuda').manual_seed(1)
xe = torch.randn(nre, generator=g, device='cuda', dtype=torch.float64); xb = torch.randn(nre, generator=g, device='cuda', dtype=torch.float64)
ye, yb = torch.empty_like(xe), torch.empty_like(xb)
bl = torch.ones(lmax+1, device='cuda', dtype=torch.float64); w = torch.rand(npix, generator=g, device='cuda', dtype=torch.float64)
ms = (C.c_float*4)()
for n in (3, 20):
    _lib.check(L.gs_profile_matvec(plan._h, _dev.ptr(xe), _dev.ptr(xb), _dev.ptr(bl), _dev.ptr(w), _dev.ptr(ye), _dev.ptr(yb), n, ms, _dev.stream()))
print("leg_synth %.3f ring_synth %.3f ring_anal %.3f leg_anal %.3f total %.3f  chk %.6e" % (ms[0], ms[1], ms[2], ms[3], sum(ms), float(ye.double().abs().sum().item())))

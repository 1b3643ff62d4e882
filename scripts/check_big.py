"""Size-independent checks + stage timing of the SHT at a large size (default NSIDE 2048 / lmax 4096, config #4)."""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gibbssampler_b200 import _dev, _lib  # noqa: E402
from gibbssampler_b200.sht import Plan  # noqa: E402

nside = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
lmax = int(sys.argv[2]) if len(sys.argv) > 2 else 2 * nside
t0 = time.time()
plan = Plan(nside, lmax)
print("plan create %.1f s" % (time.time() - t0), flush=True)
g = torch.Generator(device="cuda").manual_seed(1)
nre, npix = plan.nreal, plan.npix
x = [torch.randn(nre, generator=g, device="cuda", dtype=torch.float64) for _ in range(2)]
for a in x:
    a[[0, 1, lmax + 1, lmax + 2]] = 0
y = [torch.randn(npix, generator=g, device="cuda", dtype=torch.float64) for _ in range(2)]
q, u = plan.alm2map_spin2(x[0], x[1])
e, b = plan.map2alm_spin2(y[0], y[1], adjoint=True, real_layout=True)
lhs = float(torch.dot(q, y[0]) + torch.dot(u, y[1]))
rhs = float(torch.dot(x[0], e) + torch.dot(x[1], b))
print("adjointness <Ax,y> = %.15e  <x,A^T y> = %.15e  rel %.2e" % (lhs, rhs, abs(lhs - rhs) / abs(lhs)), flush=True)
# band-limited round trip: map2alm(iter=3)(alm2map(a)) -> a
e2, b2 = plan.map2alm_spin2(q, u, iter=3, real_layout=True)
print("round trip (iter=3) rel err E %.2e B %.2e" % (float((e2 - x[0]).abs().max() / x[0].abs().max()),
                                                       float((b2 - x[1]).abs().max() / x[1].abs().max())), flush=True)
L = _lib.lib()
ms4 = (C.c_float * 4)()
one = torch.ones(lmax + 1, device="cuda", dtype=torch.float64)
w = torch.ones(npix, device="cuda", dtype=torch.float64)
for nrep in (2, 5):
    _lib.check(L.gs_profile_matvec(plan._h, _dev.ptr(x[0]), _dev.ptr(x[1]), _dev.ptr(one), _dev.ptr(w), _dev.ptr(e), _dev.ptr(b), nrep, ms4,
                                   _dev.stream()))
nring = 4 * nside - 1
f2 = 26.0 * ((nring + 1) // 2) * sum(lmax - max(m, 2) + 1 for m in range(lmax + 1))
print("stage ms: leg_synth %.3f ring_synth %.3f ring_anal %.3f leg_anal %.3f | pair %.3f ms | leg TF/s synth %.1f anal %.1f" %
      (ms4[0], ms4[1], ms4[2], ms4[3], sum(ms4), f2 / ms4[0] * 1e-9, f2 / ms4[3] * 1e-9), flush=True)

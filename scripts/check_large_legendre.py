"""Where does the GPU synthesis differ most from the long-double direct sums at nside 2048 / lmax 4096?
One sampled m at a time (tests/sampled_sht.py), worst ring per m."""
import sys
import numpy as np
import torch

sys.path.insert(0, ".")
from gibbssampler_b200.sht import Plan
from oracle import sht as O
from tests import sampled_sht as S

nside, lmax = int(sys.argv[1]), int(sys.argv[2])
plan = Plan(nside, lmax)
rng = np.random.default_rng(1)
nring = 4 * nside - 1
rings = sorted(set(range(1, 401)) | set(range(1, nring + 1, 9)) | set(range(nring - 399, nring + 1)))
for m in S.sampled_m(lmax):
    e, b, t, coef = S.sampled_alm(lmax, [m], rng)
    q, u = (x.cpu().numpy() for x in plan.alm2map_spin2(torch.as_tensor(e, device="cuda"), torch.as_tensor(b, device="cuda")))
    tm = plan.alm2map(torch.as_tensor(t, device="cuda")).cpu().numpy()
    sc, sct = max(np.abs(q).max(), np.abs(u).max()), np.abs(tm).max()
    worst = (0.0, 0, "")
    for ring in rings:
        z, sth, phi0, nphi, start = O.ring_info(nside, ring)
        j = np.unique(rng.integers(0, nphi, size=min(16, nphi)))
        phi = phi0 + 2 * np.pi * j / nphi
        f1, f2, l0 = S._basis(lmax, m, z, sth)
        ce, cb, ct = coef[m]
        w = 1.0 if m == 0 else 2.0
        ph = np.exp(1j * m * phi)
        rq = w * (np.sum(-ce * f1 - 1j * cb * f2) * ph).real
        ru = w * (np.sum(-cb * f1 + 1j * ce * f2) * ph).real
        rt = w * (np.sum(ct * l0) * ph).real
        for name, got, ref, s in (("Q", q[start + j], rq, sc), ("U", u[start + j], ru, sc), ("T", tm[start + j], rt, sct)):
            err = np.abs(got - ref).max() / s
            if err > worst[0]:
                worst = (err, ring, name + " local |ref| %.3e of max %.3e" % (np.abs(ref).max(), s))
    print("m %5d worst rel err %.3e at ring %d %s" % (m, worst[0], worst[1], worst[2]), flush=True)

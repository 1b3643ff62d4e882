"""Per-CTA timeline of ring_apply_kernel (library built with -DGS_RING_DEBUG into csrc/_v_dbg)."""
import ctypes as C, os, sys, numpy as np, torch
sys.path.insert(0, '.')
os.environ["GIBBS_B200_LIB"] = os.path.abspath("gibbssampler_b200/csrc/_v_dbg/libgibbs_b200.so")
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
_lib.check(L.gs_profile_matvec(plan._h, _dev.ptr(xe), _dev.ptr(xb), _dev.ptr(bl), _dev.ptr(w), _dev.ptr(ye), _dev.ptr(yb), 5, ms, _dev.stream()))
torch.cuda.synchronize()
n = 4 * 2048
buf = (C.c_ulonglong * n)()
L.gs_ring_debug_dump.argtypes = [C.c_void_p, C.c_int]
print("dump rc", L.gs_ring_debug_dump(buf, n), "stage ms", list(ms))
a = np.array(buf, dtype=np.uint64).reshape(-1, 4)
a = a[a[:, 0] > 0]
t0 = a[:, 0].min()
start, mid, end = (a[:, 0] - t0) / 1e3, (a[:, 1] - t0) / 1e3, (a[:, 2] - t0) / 1e3
M = (a[:, 3] & 0xffffffff) >> 1
bs = a[:, 3] & 1
sm = a[:, 3] >> 32
print("CTAs", len(a), "kernel span us", end.max())
for key in sorted(set(zip(M.tolist(), bs.tolist())), reverse=True):
    sel = (M == key[0]) & (bs == key[1])
    print("M %5d bluestein %d: %4d CTAs, build %.1f us, rest %.1f us, total %.1f us (max %.1f), start range %.0f-%.0f us" % (
        key[0], key[1], sel.sum(), (mid - start)[sel].mean(), (end - mid)[sel].mean(), (end - start)[sel].mean(), (end - start)[sel].max(), start[sel].min(), start[sel].max()))
busy = np.zeros(148)
for s_, e_, k in zip(start, end, sm):
    busy[int(k) % 148] += e_ - s_
print("per-SM busy (sum of CTA durations) us: min %.0f mean %.0f max %.0f" % (busy.min(), busy.mean(), busy.max()))

import numpy as np, torch, sys, time
sys.path.insert(0, '.')
from oracle import sht as O
from gibbssampler_b200.sht import Plan
def rand_alm(lmax, rng, lmin=0):
    a = rng.standard_normal(O.nalm(lmax)) + 1j * rng.standard_normal(O.nalm(lmax))
    a[:lmax + 1] = a[:lmax + 1].real
    ell = np.concatenate([np.arange(m, lmax + 1) for m in range(lmax + 1)])
    a[ell < lmin] = 0
    return a
for nside, lmax in [(1,2),(2,5),(4,8),(8,16),(16,47),(32,64),(64,128),(128,256)]:
    rng = np.random.default_rng(1)
    plan = Plan.get(nside, lmax)
    a = rand_alm(lmax, rng)
    ref = O.alm2map(a, nside, lmax)
    got = plan.alm2map(torch.as_tensor(a, device='cuda')).cpu().numpy()
    e0 = np.abs(got-ref).max()/np.abs(ref).max()
    e, b = rand_alm(lmax, rng, 2), rand_alm(lmax, rng, 2)
    rq, ru = O.alm2map_spin2(e, b, nside, lmax)
    q, u = plan.alm2map_spin2(torch.as_tensor(e, device='cuda'), torch.as_tensor(b, device='cuda'))
    e2 = max(np.abs(q.cpu().numpy()-rq).max(), np.abs(u.cpu().numpy()-ru).max())/np.abs(rq).max()
    f = rng.standard_normal(12*nside**2); g = rng.standard_normal(12*nside**2)
    ra = O.map2alm(f, nside, lmax); ga = plan.map2alm(torch.as_tensor(f, device='cuda')).cpu().numpy()
    e0a = np.abs(ga-ra).max()/np.abs(ra).max()
    re_, rb_ = O.map2alm_spin2(f, g, nside, lmax)
    ge, gb = plan.map2alm_spin2(torch.as_tensor(f, device='cuda'), torch.as_tensor(g, device='cuda'))
    e2a = max(np.abs(ge.cpu().numpy()-re_).max(), np.abs(gb.cpu().numpy()-rb_).max())/np.abs(re_).max()
    print(f"nside {nside} lmax {lmax}: synth0 {e0:.2e} synth2 {e2:.2e} anal0 {e0a:.2e} anal2 {e2a:.2e}", flush=True)
    if e0 > 1e-8 and nside <= 4:
        # per-ring diagnostics
        for r in range(1, 4*nside):
            z, s, p0, n, st = O.ring_info(nside, r)
            print("  ring", r, n, np.abs(got[st:st+n]-ref[st:st+n]).max())
for nside, lmax in [(256,512),(512,1024)]:
    plan = Plan.get(nside, lmax)
    nre = (lmax+1)**2
    xe = torch.randn(nre, device='cuda', dtype=torch.float64); xb = torch.randn(nre, device='cuda', dtype=torch.float64)
    for _ in range(3):
        q,u = plan.alm2map_spin2(xe, xb); te,tb = plan.map2alm_spin2(q,u,adjoint=True,real_layout=True)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record(); 
    for _ in range(5): q,u = plan.alm2map_spin2(xe, xb)
    ev[1].record()
    for _ in range(5): te,tb = plan.map2alm_spin2(q,u,adjoint=True,real_layout=True)
    ev[2].record(); torch.cuda.synchronize()
    print(f"nside {nside}: synth2 {ev[0].elapsed_time(ev[1])/5:.3f} ms  anal2 {ev[1].elapsed_time(ev[2])/5:.3f} ms", flush=True)

"""m-sharded spherical-harmonic transforms over the GPUs of one node (BASELINE config #4).

One process per GPU (torch.distributed, backend nccl).  Rank r owns the m pairs {j, L - j} with
j mod world = r for the Legendre stage and the ring pairs p with p mod world = r for the ring-FFT /
pixel stage; libgibbs_b200.so transposes the ring spectra between the two partitions with one NCCL
all-to-all per transform.  alm vectors live as LOCAL real-layout shards, maps as LOCAL ring shards;
the partition itself is a pure function (gs_shard_partition_m / gs_shard_partition_rings), so it can
be computed -- and is tested -- without a GPU.

torch.distributed is used for the plumbing only: broadcasting the NCCL unique id at plan creation and
(in tests / IO helpers) gathering shards back into full arrays.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import GS_ALM_REAL, check
from .sht import Plan, _ptr, _stream


# ---------------------------------------------------------------------------- partition (host only)
def partition_m(lmax, world, rank):
    """Owned m of `rank` (ascending)."""
    L = _lib.lib()
    n = L.gs_shard_partition_m(lmax, world, rank, None)
    if n < 0:
        raise ValueError("bad (lmax, world, rank)")
    out = np.empty(n, dtype=np.int32)
    L.gs_shard_partition_m(lmax, world, rank, out.ctypes.data_as(C.c_void_p))
    return out


def partition_rings(nside, world, rank):
    """Owned rings (0-based, ascending) of `rank`."""
    L = _lib.lib()
    n = L.gs_shard_partition_rings(nside, world, rank, None)
    if n < 0:
        raise ValueError("bad (nside, world, rank)")
    out = np.empty(n, dtype=np.int32)
    L.gs_shard_partition_rings(nside, world, rank, out.ctypes.data_as(C.c_void_p))
    return out


def real_index(lmax, world, rank):
    """Index into the reference's real alm layout ((L+1)^2, utils.py:49-76) of every local alm entry."""
    L = _lib.lib()
    n = L.gs_shard_real_index(lmax, world, rank, None)
    if n < 0:
        raise ValueError("bad (lmax, world, rank)")
    out = np.empty(n, dtype=np.int64)
    L.gs_shard_real_index(lmax, world, rank, out.ctypes.data_as(C.c_void_p))
    return out


def pixel_index(nside, world, rank):
    """RING pixel number of every pixel of the local map shard."""
    L = _lib.lib()
    n = L.gs_shard_pixel_index(nside, world, rank, None)
    if n < 0:
        raise ValueError("bad (nside, world, rank)")
    out = np.empty(n, dtype=np.int64)
    L.gs_shard_pixel_index(nside, world, rank, out.ctypes.data_as(C.c_void_p))
    return out


def pack_spectra(values, ring_lists, m_list, RL, ML, nb=1):
    """Host model of the all-to-all send buffer [block][peer][comp][RL][MLb] of an m-owner: values[comp][ring][k]
    for k over m_list, the local m cut in `nb` blocks of MLb = ceil(ML / nb) so that the exchange can run block by block
    (one all-to-all of [peer][comp][RL][MLb] chunks per block).  nb = 1 is the single-exchange layout [peer][comp][RL][ML].
    Used by the CPU (gloo) test of the exchange layout; the CUDA kernels write the same positions (ShardDev in
    csrc/gs_internal.h)."""
    world = len(ring_lists)
    ncomp = values.shape[0]
    mlb = (ML + nb - 1) // nb
    buf = np.zeros((nb, world, ncomp, RL, mlb), dtype=values.dtype)
    for peer, rings in enumerate(ring_lists):
        for k in range(len(m_list)):
            buf[k // mlb, peer, :, :len(rings), k % mlb] = values[:, rings, k]
    return buf if nb > 1 else buf[0]


# ---------------------------------------------------------------------------- the plan
class ShardedPlan(Plan):
    """A Plan whose transforms take and return LOCAL shards.  `npix` / `nreal` are the local sizes;
    `npix_global` / `nreal_global` the full ones.  Collective: every rank of `group` must construct it
    and call its transforms in the same order."""

    def __init__(self, nside, lmax, group=None):
        if not torch.cuda.is_available():
            raise _lib.GibbsB200Error("gibbssampler_b200 needs a CUDA device (no CPU fallback)")
        if not dist.is_initialized():
            raise _lib.GibbsB200Error("ShardedPlan needs an initialised torch.distributed process group")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.nside, self.lmax = int(nside), int(lmax)
        L = _lib.lib()
        ident = C.create_string_buffer(128)
        if self.rank == 0:
            check(L.gs_nccl_unique_id(ident))
        box = [bytes(ident.raw)]
        src = dist.get_global_rank(group, 0) if group is not None else 0
        dist.broadcast_object_list(box, src=src, group=group)
        self._h = C.c_void_p()
        check(L.gs_plan_create_sharded(C.byref(self._h), self.nside, self.lmax, self.device.index, self.rank,
                                       self.world, box[0]))
        self._finish_init()

    @classmethod
    def local_group(cls, nside, lmax, world):
        """`world` sharded plans on the CURRENT device forming an in-process group (gs_local_group_create):
        each must be driven by its own host thread and CUDA stream (see run_local_group).  Verification
        helper for single-GPU boxes; no torch.distributed needed."""
        if not torch.cuda.is_available():
            raise _lib.GibbsB200Error("gibbssampler_b200 needs a CUDA device (no CPU fallback)")
        L = _lib.lib()
        grp = C.c_void_p()
        check(L.gs_local_group_create(C.byref(grp), int(world)))
        plans = []
        for r in range(world):
            self = cls.__new__(cls)
            self.group, self.world, self.rank = None, int(world), r
            self._lgroup = grp
            self.device = torch.device("cuda", torch.cuda.current_device())
            self.nside, self.lmax = int(nside), int(lmax)
            self._h = C.c_void_p()
            check(L.gs_plan_create_sharded_local(C.byref(self._h), self.nside, self.lmax, self.device.index, r, int(world), grp))
            self._finish_init()
            plans.append(self)
        return plans

    def _finish_init(self):
        L = _lib.lib()
        self.npix_global = 12 * self.nside ** 2
        self.nreal_global = (self.lmax + 1) ** 2
        self.npix = int(L.gs_plan_npix_local(self._h))
        self.nreal = int(L.gs_plan_nreal_local(self._h))
        self.nalm = None  # the complex layout is not available on sharded plans
        self.real_index = torch.as_tensor(real_index(self.lmax, self.world, self.rank), device=self.device)
        self.pixel_index = torch.as_tensor(pixel_index(self.nside, self.world, self.rank), device=self.device)
        assert self.real_index.numel() == self.nreal and self.pixel_index.numel() == self.npix

    # ---- shards <-> full arrays ---------------------------------------------------------------
    def local_map(self, full):
        """Local ring shard of a full RING map (any float64 CUDA tensor of 12 nside^2 pixels)."""
        if full.numel() == self.npix and self.npix != self.npix_global:
            return full
        assert full.numel() == self.npix_global
        return full.reshape(-1)[self.pixel_index].contiguous()

    def local_alm(self, full_real):
        if full_real.numel() == self.nreal and self.nreal != self.nreal_global:
            return full_real
        assert full_real.numel() == self.nreal_global
        return full_real.reshape(-1)[self.real_index].contiguous()

    def _gather(self, local, index_fn, nglobal, npad):
        pad = torch.zeros(npad, dtype=local.dtype, device=self.device)
        pad[: local.numel()] = local
        parts = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(parts, pad, group=self.group)
        out = torch.zeros(nglobal, dtype=local.dtype, device=self.device)
        for r in range(self.world):
            idx = torch.as_tensor(index_fn(r), device=self.device)
            out[idx] = parts[r][: idx.numel()]
        return out

    def gather_alm(self, local):
        """Full real-layout alm on every rank (test / IO helper, not on the hot path)."""
        npad = max(real_index(self.lmax, self.world, r).size for r in range(self.world))
        return self._gather(local, lambda r: real_index(self.lmax, self.world, r), self.nreal_global, npad)

    def gather_map(self, local):
        npad = max(pixel_index(self.nside, self.world, r).size for r in range(self.world))
        return self._gather(local, lambda r: pixel_index(self.nside, self.world, r), self.npix_global, npad)

    def allreduce_sum(self, value):
        """Sum of a python float over the ranks (through the plan's own communicator)."""
        t = torch.tensor([float(value)], dtype=torch.float64, device=self.device)
        check(_lib.lib().gs_shard_allreduce_sum(self._h, _ptr(t), 1, _stream()))
        return float(t.item())

    # ---- layout helpers used by the sampler classes ---------------------------------------------
    def expand_per_l(self, x, mode=0):
        x = torch.as_tensor(x, dtype=torch.float64, device=self.device).contiguous()
        out = torch.empty(self.nreal, dtype=torch.float64, device=self.device)
        check(_lib.lib().gs_shard_expand_per_l(self._h, _ptr(x), int(mode), _ptr(out), _stream()))
        return out

    def alm2cl(self, alm_local):
        cl = torch.empty(self.lmax + 1, dtype=torch.float64, device=self.device)
        check(_lib.lib().gs_shard_alm2cl(self._h, _ptr(alm_local.contiguous()), _ptr(cl), _stream()))
        return cl

    # ---- Plan overrides ----------------------------------------------------------------------------
    def _alm_in(self, a):
        assert a.dtype == torch.float64 and a.numel() == self.nreal, "sharded plans take local real-layout alm shards"
        return a.contiguous(), GS_ALM_REAL

    def _alm_out(self, layout):
        if layout != GS_ALM_REAL:
            raise _lib.GibbsB200Error("sharded plans return real-layout alm shards only (pass real_layout=True)")
        return torch.empty(self.nreal, dtype=torch.float64, device=self.device)


def run_local_group(plans, fn):
    """Runs fn(plan) for every plan of an in-process group concurrently, one host thread and one CUDA stream
    per rank (the group's collectives are host barriers); returns the results in rank order."""
    import threading
    out, err = [None] * len(plans), [None] * len(plans)
    dev = plans[0].device

    def work(i):
        try:
            torch.cuda.set_device(dev)
            with torch.cuda.stream(torch.cuda.Stream(device=dev)):
                out[i] = fn(plans[i])
                torch.cuda.current_stream().synchronize()
        except BaseException as e:  # noqa: BLE001 -- reported below; a dead rank would deadlock the others' barriers
            err[i] = e
            import os
            import sys
            import traceback
            msg = "run_local_group: rank %d failed\n%s" % (i, traceback.format_exc())
            try:   # pytest captures fd 2 and os._exit drops the capture: keep a copy where it survives
                with open(os.environ.get("GS_CRASH_LOG", "/tmp/gs_local_group_crash.log"), "a") as f:
                    f.write(msg)
            except OSError:
                pass
            sys.stderr.write(msg)
            sys.stderr.flush()
            os._exit(86)
    torch.cuda.synchronize()
    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(plans))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return out

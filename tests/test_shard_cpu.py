"""CPU tests of the m-sharded transform's host logic: the partition functions of the
C ABI (pure host code) and, over world_size-2 gloo, the [peer][comp][RL][ML] all-to-all layout that
transposes ring spectra between the m partition and the ring partition."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.mark.parametrize("lmax", [2, 3, 7, 64, 511, 1024])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_m_partition_is_exact_and_balanced(lmax, world):
    from gibbssampler_b200 import sharded as S
    if world > (lmax + 2) // 2:
        pytest.skip("more ranks than m pairs")
    lists = [S.partition_m(lmax, world, r) for r in range(world)]
    allm = np.concatenate(lists)
    assert sorted(allm.tolist()) == list(range(lmax + 1))          # every m exactly once
    for l in lists:
        assert np.all(np.diff(l) > 0)                                # ascending
    cost = np.array([np.sum(lmax + 1 - l) for l in lists], dtype=float)  # Legendre work ~ sum (L - m + 1)
    assert cost.max() - cost.min() <= lmax + 2                       # balanced to within one m pair


@pytest.mark.parametrize("nside", [1, 2, 8, 64])
@pytest.mark.parametrize("world", [1, 2, 4])
def test_ring_partition_keeps_north_south_pairs_together(nside, world):
    from gibbssampler_b200 import sharded as S
    if world > 2 * nside:
        pytest.skip("more ranks than ring pairs")
    nring = 4 * nside - 1
    lists = [S.partition_rings(nside, world, r) for r in range(world)]
    assert sorted(np.concatenate(lists).tolist()) == list(range(nring))
    for l in lists:
        s = set(l.tolist())
        assert all((nring - 1 - r) in s for r in s)                  # mirror ring owned by the same rank


@pytest.mark.parametrize("lmax,nside,world", [(5, 2, 2), (16, 8, 3), (47, 16, 4), (128, 64, 8)])
def test_shard_index_maps_are_bijections_onto_the_reference_layouts(lmax, nside, world):
    from gibbssampler_b200 import sharded as S
    ri = [S.real_index(lmax, world, r) for r in range(world)]
    assert sorted(np.concatenate(ri).tolist()) == list(range((lmax + 1) ** 2))
    pi = [S.pixel_index(nside, world, r) for r in range(world)]
    assert sorted(np.concatenate(pi).tolist()) == list(range(12 * nside ** 2))
    # local alm order: owned m ascending; m = 0 block is l = 0..L, m > 0 blocks are (re, im) interleaved
    for r in range(world):
        pos = 0
        for m in S.partition_m(lmax, world, r):
            base = m * (2 * lmax + 1 - m) // 2
            for l in range(m, lmax + 1):
                if m == 0:
                    assert ri[r][pos] == l
                    pos += 1
                else:
                    o = 2 * (base + l) - (lmax + 1)  # utils.py:49-76
                    assert ri[r][pos] == o and ri[r][pos + 1] == o + 1
                    pos += 2
        assert pos == ri[r].size
    # local map order: owned rings ascending, pixels of a ring contiguous
    for r in range(world):
        assert np.all(np.diff(pi[r]) > 0)


def _f(comp, ring, m):
    return 1000.0 * ring + m + 0.25 * comp


def _worker(rank, world, port, nside, lmax, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gibbssampler_b200 import sharded as S
    nring = 4 * nside - 1
    ring_lists = [S.partition_rings(nside, world, r) for r in range(world)]
    m_lists = [S.partition_m(lmax, world, r) for r in range(world)]
    RL, ML = max(len(x) for x in ring_lists), max(len(x) for x in m_lists)
    # m-owner side (output of the Legendre synthesis): F for my m, ALL rings
    mine = m_lists[rank]
    vals = np.array([[[_f(c, r, m) for m in mine] for r in range(nring)] for c in range(2)])
    send = torch.from_numpy(S.pack_spectra(vals, ring_lists, mine, RL, ML))
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send)        # chunk per peer = [comp][RL][ML], as gs_shard_exchange
    # ring-owner side: element (ring, m) must sit at [m_owner][comp][ring_loc][m_loc]
    ok = True
    for src in range(world):
        for k, m in enumerate(m_lists[src]):
            for j, r in enumerate(ring_lists[rank]):
                for c in range(2):
                    ok &= recv[src, c, j, k].item() == _f(c, r, m)
    # and back (analysis direction): the transpose of the transpose returns the send buffer
    back = torch.empty_like(send)
    dist.all_to_all_single(back, recv)
    ok &= torch.equal(back, send)
    # blocked layout (pipelined exchange: one all-to-all per block of local m): element (ring, m) must sit at
    # [m_loc // MLb][m_owner][comp][ring_loc][m_loc % MLb]
    for nb in (2, 3):
        mlb = (ML + nb - 1) // nb
        sendb = torch.from_numpy(S.pack_spectra(vals, ring_lists, mine, RL, ML, nb=nb))
        recvb = torch.empty_like(sendb)
        for b in range(nb):
            dist.all_to_all_single(recvb[b], sendb[b])
        for src in range(world):
            for k, m in enumerate(m_lists[src]):
                for j, r in enumerate(ring_lists[rank]):
                    for c in range(2):
                        ok &= recvb[k // mlb, src, c, j, k % mlb].item() == _f(c, r, m)
    res = torch.tensor([1 if ok else 0])
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(int(res.item()))
    dist.destroy_process_group()


def test_world_size_2_ring_m_transpose_layout_over_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 4, 9, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok == 1

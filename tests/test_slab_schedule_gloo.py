"""CPU, world_size 2, gloo: the N > 1 path's host logic. Each rank holds one z-slab (+ 4 ghost
planes) as numpy arrays and replays the slab V-cycle schedule published by the package
(`slab_schedule`, the host mirror of csrc/mg_engine.cuh::slab_twogrid): halo planes travel by
torch.distributed send/recv over gloo, the first replicated level by all_gather, the replicated
coarse V-cycle runs in the oracle. The gathered result must equal the oracle's single-domain
V-cycles bit for bit -- i.e. ghost depths, exchange points and the replicated threshold of the
schedule are sufficient. (The CUDA kernels themselves are checked on a GPU: test_gpu_slabs.py,
tests/mgpu_check.py.)"""
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
sys.path.insert(0, os.path.join(ROOT, "tests"))
G = 4


def nsum(u):
    """((((xl+xr)+yl)+yr)+zl)+zr with zero outside the array (SURVEY 8(a'))."""
    p = np.pad(u, 1)
    return ((((p[1:-1, 1:-1, :-2] + p[1:-1, 1:-1, 2:]) + p[1:-1, :-2, 1:-1]) + p[1:-1, 2:, 1:-1]) + p[:-2, 1:-1, 1:-1]) + p[2:, 1:-1, 1:-1]


def exchange(a, depth, rank, world, nz):
    """ghost planes of depth `depth` from the two neighbours (what ncclSend/ncclRecv do)."""
    reqs = []
    lo_recv = up_recv = None
    if rank > 0:
        reqs.append(dist.isend(torch.from_numpy(a[G:G + depth].copy()), rank - 1))
        lo_recv = torch.empty((depth,) + a.shape[1:], dtype=torch.float64)
        reqs.append(dist.irecv(lo_recv, rank - 1))
    if rank < world - 1:
        reqs.append(dist.isend(torch.from_numpy(a[G + nz - depth:G + nz].copy()), rank + 1))
        up_recv = torch.empty((depth,) + a.shape[1:], dtype=torch.float64)
        reqs.append(dist.irecv(up_recv, rank + 1))
    for r in reqs:
        r.wait()
    if lo_recv is not None:
        a[G - depth:G] = lo_recv.numpy()
    if up_recv is not None:
        a[G + nz:G + nz + depth] = up_recv.numpy()


def worker(rank, world, port, size, cycles, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import load_package
    import oracle as O
    pkg = load_package()
    part = {x["L"]: x for x in pkg.slab_partition(size, world)}
    orc = O.Oracle(size, "double", 3)          # replicated coarse levels live in here (persistent Vs)
    f_full, psi_full = orc.f.copy(), orc.psi.copy()
    slabs = {}
    for L, x in part.items():
        if x["distributed"]:
            nz = L // world
            slabs[L] = {k: np.zeros((nz + 2 * G, L, L)) for k in ("u", "f")}
    nz = size // world
    slabs[size]["u"][G:G + nz] = psi_full[rank * nz:(rank + 1) * nz]
    slabs[size]["f"][G:G + nz] = f_full[rank * nz:(rank + 1) * nz]

    def dom_mask(L):  # planes of the local array that lie inside the global grid
        n = L // world
        g = np.arange(n + 2 * G) - G + rank * n
        return ((g >= 0) & (g < L))[:, None, None]

    for _ in range(cycles):
        h = 1.0 / size
        hs = {}
        L = size
        while L >= 1:
            hs[L] = h
            h *= 2
            L //= 2
        for op, L, d in pkg.slab_schedule(size, world):
            if op in ("exchange_u", "exchange_V"):
                exchange(slabs[L]["u"], d, rank, world, L // world)
            elif op in ("exchange_f", "exchange_R"):
                exchange(slabs[L]["f"], d, rank, world, L // world)
            elif op == "pass":
                u, f, hh, m = slabs[L]["u"], slabs[L]["f"], hs[L], dom_mask(L)
                n = L // world
                if d["pro"]:
                    coarse = part[L // 2]
                    if coarse["distributed"]:
                        V = slabs[L // 2]["u"]
                        gc = (np.arange(n + 2 * G) - G + rank * n) // 2 - rank * (n // 2) + G
                    else:
                        V = orc.buffer(O.BUF_V, L // 2)
                        gc = (np.arange(n + 2 * G) - G + rank * n) // 2
                    ok = (gc >= 0) & (gc < V.shape[0]) & m[:, 0, 0]
                    Vp = np.zeros_like(u)
                    Vp[ok] = np.repeat(np.repeat(V[gc[ok]], 2, 1), 2, 2)
                    u[...] = np.where(m, u + Vp, 0.0)
                for _s in range(d["sweeps"]):
                    u[...] = np.where(m, (f - nsum(u) / hh**2) / (-6 / hh**2), 0.0)
                if d["res"]:
                    r = np.where(m, f - (nsum(u) / hh**2 + (-6 / hh**2) * u), 0.0)[G:G + n]
                    s = r[0::2, 0::2, 0::2] + r[0::2, 0::2, 1::2]
                    for dz, dy, dx in ((0, 1, 0), (0, 1, 1), (1, 0, 0), (1, 0, 1), (1, 1, 0), (1, 1, 1)):
                        s = s + r[dz::2, dy::2, dx::2]
                    Rc = .125 * s
                    if part[L // 2]["distributed"]:
                        slabs[L // 2]["f"][G:G + n // 2] = Rc
                    else:
                        orc.buffer(O.BUF_R, L // 2)[rank * (n // 2):(rank + 1) * (n // 2)] = Rc
            elif op == "allgather_R":
                R = orc.buffer(O.BUF_R, L)
                n = L // world
                parts = [torch.empty((n, L, L), dtype=torch.float64) for _ in range(world)]
                dist.all_gather(parts, torch.from_numpy(R[rank * n:(rank + 1) * n].copy()))
                for r_, p_ in enumerate(parts):
                    R[r_ * n:(r_ + 1) * n] = p_.numpy()
            elif op == "replicated_vcycle":
                orc.two_grid(hs[L], orc.buffer(O.BUF_V, L), orc.buffer(O.BUF_R, L), L)
    mine = torch.from_numpy(slabs[size]["u"][G:G + nz].copy())
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    if rank == 0:
        np.save(out, torch.cat(parts, 0).numpy())
    dist.destroy_process_group()


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_schedule_description(mgp, monkeypatch):
    monkeypatch.delenv("MGPOISSON_SLAB_MIN_PLANES", raising=False)
    ops = mgp.slab_schedule(1024, 8)            # default threshold: 32 planes per rank
    assert sorted({L for op, L, _ in ops if op == "pass"}, reverse=True) == [1024, 512, 256]
    assert ("allgather_R", 128, None) in ops and ("replicated_vcycle", 128, None) in ops
    assert sum(1 for op, _, _ in ops if op.startswith("exchange")) == 17
    monkeypatch.setenv("MGPOISSON_SLAB_MIN_PLANES", "8")
    ops = mgp.slab_schedule(1024, 8)
    lv = [L for op, L, _ in ops if op == "pass"]
    assert sorted(set(lv), reverse=True) == [1024, 512, 256, 128, 64]
    assert ("allgather_R", 32, None) in ops and ("replicated_vcycle", 32, None) in ops
    assert mgp.plan_passes(7, 4, True) == [4, 3] and mgp.plan_passes(7, 4, False) == [4, 3]
    assert mgp.plan_passes(7, 3, True) == [2, 2, 3] and mgp.plan_passes(7, 2, False) == [2, 2, 2, 1]
    # every pass is preceded by a ghost refresh at least as deep as its pipeline
    for i, (op, L, d) in enumerate(ops):
        if op == "pass":
            assert ops[i - 1][0] == "exchange_u" and ops[i - 1][2] >= d["sweeps"] + (1 if d["res"] else 0)
    # 5 levels x 4 passes + 4 x Rs + 4 x Vs (+ f once) halo exchanges per V-cycle at 1024^3 on 8 GPUs
    assert sum(1 for op, _, _ in ops if op.startswith("exchange")) == 29


@pytest.mark.timeout(600)
def test_two_rank_gloo_slab_vcycle_matches_oracle(tmp_path, orc, monkeypatch):
    monkeypatch.setenv("MGPOISSON_SLAB_MIN_PLANES", "8")   # two distributed levels (128, 64) below a replicated 32
    size, world, cycles = 128, 2, 2
    out = str(tmp_path / "psi.npy")
    mp.spawn(worker, args=(world, free_port(), size, cycles, out), nprocs=world, join=True)
    got = np.load(out)
    o = orc.Oracle(size, "double", 3, nthreads=8)
    for _ in range(cycles):
        o.vcycle()
    assert got.tobytes() == o.psi.tobytes()

"""Multi-GPU parity check, run under torchrun with one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/mgpu_check.py [size]

Every rank runs its slab through NCCL halo exchange; rank 0 gathers the slabs and compares
them, bit for bit, with a single-GPU solver of the whole grid (which the -m gpu tests tie to
the oracle). Prints one line per check and exits non-zero on any mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    only = sys.argv[2] if len(sys.argv) > 2 else ""     # "batch": only the host-batch check (short GPU calls)
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = load_package()
    ok = True
    # (real kind, fused transport, smooth): smooth = 4 and 8 give an odd number of ping-pong passes per level visit
    # unless the slab schedule pads them; the default 7 gives an even one
    for real, p2p, smooth in (("float", True, 7), ("double", True, 7), ("float", False, 7), ("float", True, 4), ("float", True, 8)):
        if only:
            break
        s = pkg.create_distributed(size, real, dim=3, p2p=p2p, smooth=smooth)
        errs = [s.step() for _ in range(3)]
        rn = s.residual_norm()
        mine = torch.from_numpy(s.psi.download()).cuda()
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        if rank == 0:
            full = torch.cat(parts, 0).cpu().numpy()
            one = pkg.MultigridCUDA(size, real, dim=3, out=False, device=local, smooth=smooth)
            one.set_tuning(tb=4)
            one.set_option("stream_min_L", 64)
            ref_errs = [one.step() for _ in range(3)]
            same = full.tobytes() == one.psi.download().tobytes()
            eok = all(abs(a - b) <= 1e-9 * abs(b) for a, b in zip(errs, ref_errs))
            eok = eok and abs(rn - one.residual_norm()) <= 1e-9 * rn
            print(f"[mgpu_check] {world} GPUs, {size}^3 {real}, smooth={smooth}, {'fused peer-store' if p2p else 'NCCL send/recv'} halos: psi bit-identical to 1 GPU: {same}; "
                  f"err {errs} vs {ref_errs}: {eok}; slab info {s.slab_info()}", flush=True)
            ok = ok and same and eok
            one.close()
        s.close()
    # conjugate gradient on slabs (mg_cg): ghost planes of p by ncclSend/ncclRecv, scalars all-reduced; compared with the
    # single-GPU run (same algorithm, different summation order: iteration counts equal, histories to 1e-9)
    for real in ("double",):
        if only:
            break
        s = pkg.create_distributed(size, real, dim=3)
        errs, linf = s.conjgrad(max_iter=60, epsilon=1e-6)
        if rank == 0:
            one = pkg.MultigridCUDA(size, real, dim=3, out=False, device=local)
            e1, l1 = one.conjgrad(max_iter=60, epsilon=1e-6)
            cok = len(errs) == len(e1) and all(abs(a - b) <= 1e-9 * abs(b) for a, b in zip(errs, e1)) and \
                all(abs(a - b) <= 1e-9 * abs(b) for a, b in zip(linf, l1))
            print(f"[mgpu_check] {world} GPUs, {size}^3 {real}, conjugate gradient on slabs: {len(errs)} iterations, err {errs[-1]:.3e} "
                  f"vs 1 GPU {len(e1)} iterations, err {e1[-1]:.3e}: bit-identical to 1 GPU: n/a, histories equal to 1e-9: {cok}", flush=True)
            ok = ok and cok
            one.close()
        # and the smoother still works afterwards (psi's ghost planes are refreshed)
        e_after = s.step()
        ok = ok and bool(np.isfinite(e_after))
        s.close()
    # pipelined host batches on slabs (mg_step_host_batch: every rank passes its own planes of each problem): against
    # the single-GPU solver fed the whole fields through mg_step_host, one problem at a time
    def gather(a):
        mine = torch.from_numpy(np.ascontiguousarray(a)).cuda()
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        return torch.cat(parts, 0).cpu().numpy()
    s = pkg.create_distributed(size, "float", dim=3)
    f0, p0 = s.f.download(), s.psi.download()
    nprob = 4
    fs = [pkg.PinnedArray(f0.shape, np.float32) for _ in range(nprob)]
    ps = [pkg.PinnedArray(f0.shape, np.float32) for _ in range(nprob)]
    for k in range(nprob):
        fs[k].array[...] = f0 * np.float32(1 + 0.25 * k)
        ps[k].array[...] = p0 * np.float32(1 - 0.125 * k)
    gf = [gather(a.array) for a in fs]
    gp = [gather(a.array) for a in ps]
    errs = s.step_host_batch([a.array for a in fs], [a.array for a in ps])
    got = [gather(a.array) for a in ps]
    if rank == 0:
        one = pkg.MultigridCUDA(size, "float", dim=3, out=False, device=local)
        one.set_tuning(tb=4)
        one.set_option("stream_min_L", 64)
        bok, e1 = True, []
        for k in range(nprob):
            fh, ph = pkg.PinnedArray(gf[k].shape, np.float32), pkg.PinnedArray(gf[k].shape, np.float32)
            fh.array[...] = gf[k]; ph.array[...] = gp[k]
            e1.append(one.step_host(fh.array, ph.array))
            bok = bok and ph.array.tobytes() == got[k].tobytes() and abs(errs[k] - e1[-1]) <= 1e-9 * abs(e1[-1])
            fh.free(); ph.free()
        print(f"[mgpu_check] {world} GPUs, {size}^3 float, {nprob} host problems pipelined through mg_step_host_batch on slabs: "
              f"err {errs} vs {e1}; psi bit-identical to 1 GPU: {bok}", flush=True)
        ok = ok and bok
        one.close()
    for a in fs + ps:
        a.free()
    s.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()

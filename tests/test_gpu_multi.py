"""-m gpu, needs >= 2 GPUs in the box (skipped otherwise): the multi-GPU transports against the single-GPU solver.

* one process per GPU (torchrun + NCCL bootstrap, CUDA IPC peer memory, in-kernel handshakes, whole cycle in a CUDA
  graph): tests/mgpu_check.py, spawned here so that the pytest tier covers it;
* one process, one GPU per slab (mg_create_slab_multi): what a single LuaJIT host would use."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("size", [256])
def test_one_process_per_gpu_matches_single_gpu(size):
    n = ngpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "mgpu_check.py"), str(size)],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if "[mgpu_check]" in l]
    if r.returncode != 0:      # keep the whole transcript where a gpurun call brings it back
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        open(os.path.join(ROOT, "gpurun_out", "mgpu_check_failed.log"), "w").write(r.stdout + "\n---- stderr ----\n" + r.stderr)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert len(lines) >= 5 and all("bit-identical to 1 GPU: True" in l or "histories equal to 1e-9: True" in l for l in lines), lines


@pytest.mark.parametrize("real,smooth", [("float", 7), ("double", 7), ("float", 4)])
def test_one_process_many_gpus_matches_single_gpu(mgp, real, smooth):
    n = ngpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    devs = list(range(8 if n >= 8 else (4 if n >= 4 else 2)))
    size = 256
    one = mgp.MultigridCUDA(size, real, dim=3, device=0, out=False, smooth=smooth)
    many = mgp.MultigridCUDA(size, real, dim=3, out=False, smooth=smooth, devices=devs)
    try:
        for cyc in range(3):
            e1, e2 = one.step(), many.step()
            assert abs(e1 - e2) <= 1e-9 * abs(e1), (cyc, e1, e2)
        a, b = one.psi.download(), many.psi.download()
        assert a.tobytes() == b.tobytes(), f"{int((a != b).sum())} values differ"
        # upload / download round trip and a re-initialised run behave like the single solver
        rng = np.random.default_rng(3)
        u = rng.uniform(-1, 1, a.shape).astype(a.dtype)
        one.psi.upload(u); many.psi.upload(u)
        assert abs(one.step() - many.step()) <= 1e-9
        assert one.psi.download().tobytes() == many.psi.download().tobytes()
        r1, r2 = one.residual_norm(), many.residual_norm()
        assert abs(r1 - r2) <= 1e-9 * r1
        tr = many.slab_traffic()
        assert tr["peer_store_bytes"] > 0
    finally:
        one.close()
        many.close()

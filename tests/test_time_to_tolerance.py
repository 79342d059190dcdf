"""North-star clause "the same number of V-cycles to reach a given residual".

cpu-raw.lua's iteration (coarse corrections carried over) never reaches a small residual; cpu.lua's
(corrections re-zeroed every cycle, cpu.lua:138 -- the variant test/converge-multigrid-vs-krylov.lua drives, pinned
bit for bit by tests/golden/refcpu_*.npz) does. Cycles until ||f - A psi|| <= 1e-8 ||f - A psi_0||, residual checked
after every cycle:

  CPU tier: the oracle's counts for that variant (pins the convergence behaviour itself);
  GPU tier: the CUDA path needs exactly the same number of cycles as the oracle and ends at the same residual.
"""
import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402

spec = importlib.util.spec_from_file_location("time_to_tolerance", os.path.join(ROOT, "tools", "time_to_tolerance.py"))
ttt = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ttt)

COUNTS = {(2, 16): 70, (2, 32): 255, (2, 64): 949, (3, 8): 20, (3, 16): 66, (3, 32): 234}


@pytest.mark.parametrize("dim,size", sorted(COUNTS))
def test_oracle_cycles_to_1e8_residual(dim, size):
    r = ttt.solve_oracle(O, dim, size, 1e-8, 1, 5000, 4)
    assert r["cycles"] == COUNTS[(dim, size)] and r["residual_rel"] <= 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("dim,size", [(2, 32), (2, 64), (3, 16), (3, 32)])
def test_cuda_needs_the_same_number_of_cycles(mgp, dim, size):
    want = ttt.solve_oracle(O, dim, size, 1e-8, 1, 5000, 4)
    got = ttt.solve_gpu(mgp, dim, size, 1e-8, 1, 5000)
    assert got["cycles"] == want["cycles"] == COUNTS[(dim, size)]
    assert abs(got["residual_rel"] - want["residual_rel"]) <= 1e-9 * want["residual_rel"]

"""Generates tests/golden/*.npz from the CPU oracle (oracle/mg_oracle.c).

These fixtures pin the oracle to ITSELF at the commit that generated them (including the 3-D
cases and cycle counts the reference never runs), so that later edits to the oracle cannot
silently change the numbers the CUDA path is compared against. The oracle's tie to the
REFERENCE is separate: tests/golden/ref_2d_*.npz, produced by executing the reference's own
cpu-raw.lua (oracle/run_reference.py), and the hand-derived known answers in
tests/test_oracle_known_answers.py.

Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle as O  # noqa: E402

CASES = [  # (name, dim, size, real, cycles)
    ("c1_2d_64_f64", 2, 64, "double", 6),       # BASELINE config 1: the reference's own case
    ("2d_32_f64", 2, 32, "double", 4),          # test/test.lua:45 runs log2size = 5
    ("2d_64_f32", 2, 64, "float", 4),
    ("2d_64_f32a64", 2, 64, "float_acc64", 4),
    ("3d_16_f64", 3, 16, "double", 4),
    ("3d_32_f32", 3, 32, "float", 3),
    ("3d_16_f32a64", 3, 16, "float_acc64", 3),
]


def main():
    for name, dim, size, real, cycles in CASES:
        o = O.Oracle(size, real, dim)
        errs, psis = [], []
        for _ in range(cycles):
            errs.append(o.step())
            psis.append(o.psi.copy())
        lv = {}
        L = size // 2
        while L >= 1:
            lv[f"R{L}"] = o.buffer(O.BUF_R, L).copy()
            lv[f"V{L}"] = o.buffer(O.BUF_V, L).copy()
            L //= 2
        np.savez_compressed(os.path.join(HERE, name + ".npz"), errs=np.array(errs),
                            psi_first=psis[0], psi_last=psis[-1], residual_rms=o.residual_rms(),
                            meta=np.array([dim, size, O.REAL_NAMES[real], cycles]), **lv)
        print(name, errs)


if __name__ == "__main__":
    main()

"""helpers for the -m gpu parity tests (torch is only device-memory plumbing here)."""
import numpy as np
import torch

KINDS = ["double", "float", "float_acc64"]


def to_dev(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def to_host(t: torch.Tensor) -> np.ndarray:
    return t.cpu().numpy()


def bits(a: np.ndarray) -> np.ndarray:
    return a.view(np.uint64 if a.dtype == np.float64 else np.uint32)


def assert_bits_equal(got: np.ndarray, want: np.ndarray, what=""):
    """Bit-for-bit equality (signed zeros and NaN payloads included)."""
    assert got.shape == want.shape and got.dtype == want.dtype, (what, got.shape, want.shape, got.dtype, want.dtype)
    neq = bits(np.ascontiguousarray(got)) != bits(np.ascontiguousarray(want))
    if neq.any():
        idx = np.argwhere(neq)
        first = tuple(idx[0])
        with np.errstate(invalid="ignore"):
            maxabs = np.nanmax(np.abs(got.astype(np.float64) - want.astype(np.float64)))
        raise AssertionError(
            f"{what}: {neq.sum()} of {neq.size} values differ bitwise; first at {first}: "
            f"got {got[first]!r} want {want[first]!r}; max |diff| = {maxabs:g}")


def rand_field(rng, dim, L, dtype):
    return rng.uniform(-1, 1, (L,) * dim).astype(dtype)


def err_rtol(n):
    """Relative tolerance for the per-cycle `err`: the reference (and the oracle) add the n
    squared differences sequentially in double (cpu-raw.lua:250-253), the CUDA path adds them
    as a tree; the sequential sum carries up to ~n*2^-53 relative rounding error."""
    return 1e-12 + n * 2.0 ** -53 * 0.25

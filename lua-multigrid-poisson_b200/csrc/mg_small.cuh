// mg_small.cuh -- (K-d) persistent small-level kernel: the whole V-cycle below a level
// threshold -- every pre-sweep, residual+restriction, the L = 1 solve (cpu-raw.lua:190-196),
// prolongation+add and every post-sweep of every level L <= Ltop -- in ONE launch of ONE CTA,
// with __syncthreads() where the reference has kernel boundaries. At 512^3 the levels L <= 32
// are 162 of the reference's 290 enqueues per cycle; here they are one.
//
// The fields stay in global memory and are served by L1/L2 (the whole sub-hierarchy below
// 32^3 fp32 is < 1 MB). Global data written inside this kernel is only ever re-read by the
// same CTA after a __syncthreads(), through ordinary (coherent) loads -- hence no
// __restrict__ / ld.global.nc on these pointers.
//
// Per-point arithmetic: mg_math.cuh, identical to every other kernel.
#pragma once
#include "mg_fused_simple.cuh"
#include "mg_math.cuh"

namespace mg {

constexpr int SMALL_MAX_LEVELS = 9;  // L = 1 .. 256

template <typename R, typename A> struct SmallArgs {
    R *u[SMALL_MAX_LEVELS];        // correction / solution at level lv (Vs[L], or the caller's u at top)
    const R *f[SMALL_MAX_LEVELS];  // right-hand side at level lv (Rs[L], or the caller's f at top)
    R *w[SMALL_MAX_LEVELS];        // ping-pong partner
    Coef<A> coef[SMALL_MAX_LEVELS];
    int top;                       // log2 of the top level width
    int smooth;
};

template <typename R, typename A, int DIM, bool PROLONG>
__device__ __forceinline__ void cta_sweep(R *dst, const R *src, const R *f, const R *V, int lg,
                                          const Coef<A> &c)
{
    const int L = 1 << lg;
    const int n = DIM == 3 ? (L * L * L) : (L * L);
    const size_t sL = (size_t)L, sLL = sL * sL;
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
        int i = idx & (L - 1), j = (idx >> lg) & (L - 1), k = DIM == 3 ? (idx >> (2 * lg)) : 0;
        A S;
        if (!PROLONG) {
            S = stencil_sum<DIM, R, A>(src, i, j, k, L, (size_t)idx);
        } else {
            A xl = i > 0 ? corrected<R, A, DIM>(src, V, i - 1, j, k, L, idx - 1) : (A)0;
            A xr = i < L - 1 ? corrected<R, A, DIM>(src, V, i + 1, j, k, L, idx + 1) : (A)0;
            A yl = j > 0 ? corrected<R, A, DIM>(src, V, i, j - 1, k, L, idx - sL) : (A)0;
            A yr = j < L - 1 ? corrected<R, A, DIM>(src, V, i, j + 1, k, L, idx + sL) : (A)0;
            S = Ar<A>::add(Ar<A>::add(Ar<A>::add(xl, xr), yl), yr);
            if (DIM == 3) {
                A zl = k > 0 ? corrected<R, A, DIM>(src, V, i, j, k - 1, L, idx - sLL) : (A)0;
                A zr = k < L - 1 ? corrected<R, A, DIM>(src, V, i, j, k + 1, L, idx + sLL) : (A)0;
                S = Ar<A>::add(Ar<A>::add(S, zl), zr);
            }
        }
        A out = jacobi_point<DIM, A>(S, (A)f[idx], c);
        if (c.weighted) out = relax<A>(out, PROLONG ? corrected<R, A, DIM>(src, V, i, j, k, L, (size_t)idx) : (A)src[idx], c);
        dst[idx] = (R)out;
    }
}

template <typename R, typename A, int DIM>
__device__ __forceinline__ void cta_residual_restrict(R *Rc, const R *f, const R *u, int lg,
                                                      const Coef<A> &c)
{
    const int L = 1 << lg, lg2 = lg - 1, L2 = L >> 1;
    const int n2 = DIM == 3 ? (L2 * L2 * L2) : (L2 * L2);
    const size_t sL = (size_t)L, sLL = sL * sL;
    for (int cidx = threadIdx.x; cidx < n2; cidx += blockDim.x) {
        int I = cidx & (L2 - 1), J = (cidx >> lg2) & (L2 - 1), K = DIM == 3 ? (cidx >> (2 * lg2)) : 0;
        A s = (A)0;
        bool first = true;
#pragma unroll
        for (int dk = 0; dk < (DIM == 3 ? 2 : 1); ++dk)
#pragma unroll
            for (int dj = 0; dj < 2; ++dj)
#pragma unroll
                for (int di = 0; di < 2; ++di) {
                    int i = 2 * I + di, j = 2 * J + dj, k = 2 * K + dk;
                    size_t idx = (size_t)i + sL * j + sLL * k;
                    A S = stencil_sum<DIM, R, A>(u, i, j, k, L, idx);
                    A rv = (A)(R)residual_point<A>(S, (A)f[idx], (A)u[idx], c);
                    s = first ? rv : Ar<A>::add(s, rv);
                    first = false;
                }
        Rc[cidx] = (R)Ar<A>::mul(DIM == 3 ? (A).125 : (A).25, s);
    }
}

template <typename R, typename A, int DIM>
__global__ void __launch_bounds__(1024, 1) k_small_vcycle(SmallArgs<R, A> a)
{
    const int smooth = a.smooth;
    const int par = smooth & 1;  // after `smooth` ping-pong sweeps the field sits in w if odd
    // ---- descend: pre-smooth, residual, restrict (cpu-raw.lua:198-218)
    for (int lv = a.top; lv >= 1; --lv) {
        R *src = a.u[lv], *dst = a.w[lv];
        for (int s = 0; s < smooth; ++s) {
            cta_sweep<R, A, DIM, false>(dst, src, a.f[lv], (const R *)nullptr, lv, a.coef[lv]);
            __syncthreads();
            R *t = src; src = dst; dst = t;
        }
        cta_residual_restrict<R, A, DIM>(const_cast<R *>(a.f[lv - 1]), a.f[lv], src, lv, a.coef[lv]);
        __syncthreads();
    }
    // ---- L = 1: one smoother call (cpu-raw.lua:190-196): every neighbour is out of range
    if (threadIdx.x == 0) {
        A S = Ar<A>::add(Ar<A>::add(Ar<A>::add((A)0, (A)0), (A)0), (A)0);
        if (DIM == 3) S = Ar<A>::add(Ar<A>::add(S, (A)0), (A)0);
        a.u[0][0] = (R)relax<A>(jacobi_point<DIM, A>(S, (A)a.f[0][0], a.coef[0]), (A)a.u[0][0], a.coef[0]);
    }
    __syncthreads();
    // ---- ascend: prolong, add, post-smooth (cpu-raw.lua:221-236)
    for (int lv = 1; lv <= a.top; ++lv) {
        R *src = par ? a.w[lv] : a.u[lv];
        R *dst = par ? a.u[lv] : a.w[lv];
        const R *V = a.u[lv - 1];
        if (smooth == 0) {
            const int L = 1 << lv;
            const int n = DIM == 3 ? (L * L * L) : (L * L);
            for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
                int i = idx & (L - 1), j = (idx >> lv) & (L - 1), k = DIM == 3 ? (idx >> (2 * lv)) : 0;
                src[idx] = (R)corrected<R, A, DIM>(src, V, i, j, k, L, (size_t)idx);
            }
            __syncthreads();
            continue;
        }
        cta_sweep<R, A, DIM, true>(dst, src, a.f[lv], V, lv, a.coef[lv]);
        __syncthreads();
        { R *t = src; src = dst; dst = t; }
        for (int s = 1; s < smooth; ++s) {
            cta_sweep<R, A, DIM, false>(dst, src, a.f[lv], (const R *)nullptr, lv, a.coef[lv]);
            __syncthreads();
            R *t = src; src = dst; dst = t;
        }
        // 2*smooth sweeps in total at this level => the field is back in a.u[lv]
    }
}

}  // namespace mg

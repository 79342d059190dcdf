// mg_fused_simple.cuh -- untiled fused passes. Generic in level size (any L >= 2), used
//  * at levels too small for the tiled/streaming kernels but above the persistent kernel,
//  * as the tb=1 baseline the temporally blocked kernels are measured against.
// Same per-point arithmetic as the reference sequence (mg_math.cuh), but
//  - ping-pong (src -> dst) instead of Jacobi-into-tmpU + copy-back (cpu-raw.lua:181-182),
//  - prolongation + addTo folded into the loads of the first post-sweep
//    (cpu-raw.lua:225-230 then :233), so vs[L] is never written,
//  - residual + restriction in one kernel (cpu-raw.lua:211,218), so rs[L] is never written.
#pragma once
#include "mg_math.cuh"
#include "mg_ops_ref.cuh"

namespace mg {

// value of the corrected field u + prolong(V) at (i,j,k), rounded to storage as addTo does
template <typename R, typename A, int DIM>
__device__ __forceinline__ A corrected(const R *u, const R *V, int i, int j, int k, int L, size_t idx)
{
    const int L2 = L >> 1;
    R vv = V[(size_t)(i >> 1) + (size_t)L2 * ((size_t)(j >> 1) + (size_t)L2 * (DIM == 3 ? (k >> 1) : 0))];
    return (A)(R)Ar<A>::add((A)u[idx], (A)vv);
}

// one Jacobi sweep src -> dst; with PROLONG the source field is (src + prolong(V)).
template <typename R, typename A, int DIM, bool PROLONG>
__global__ void k_sweep_pp(R *__restrict__ dst, const R *__restrict__ src, const R *__restrict__ f,
                           const R *__restrict__ V, int L, Coef<A> c)
{
    pdl_enter();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    int k = DIM == 3 ? blockIdx.z : 0;
    if (i >= L || j >= L) return;
    const size_t sL = (size_t)L, sLL = sL * sL;
    size_t idx = (size_t)i + sL * j + sLL * k;
    A S;
    if (!PROLONG) {
        S = stencil_sum<DIM, R, A>(src, i, j, k, L, idx);
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
    if (c.weighted) out = relax<A>(out, PROLONG ? corrected<R, A, DIM>(src, V, i, j, k, L, idx) : (A)src[idx], c);
    dst[idx] = (R)out;
}

// u += prolong(V) in place (only needed when a level visit has zero post-sweeps)
template <typename R, typename A, int DIM>
__global__ void k_prolong_add(R *__restrict__ u, const R *__restrict__ V, int L)
{
    pdl_enter();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    int k = DIM == 3 ? blockIdx.z : 0;
    if (i >= L || j >= L) return;
    size_t idx = (size_t)i + (size_t)L * ((size_t)j + (size_t)L * k);
    u[idx] = (R)corrected<R, A, DIM>(u, V, i, j, k, L, idx);
}

// R[I,J,K] = restrict(f - A u): one thread per coarse cell, children in the reference's
// order; each child residual is rounded to storage as if it had been written to rs[L].
template <typename R, typename A, int DIM>
__global__ void k_residual_restrict(R *__restrict__ Rc, const R *__restrict__ f,
                                    const R *__restrict__ u, int L, Coef<A> c)
{
    pdl_enter();
    const int L2 = L >> 1;
    int I = blockIdx.x * blockDim.x + threadIdx.x;
    int J = blockIdx.y * blockDim.y + threadIdx.y;
    int K = DIM == 3 ? blockIdx.z : 0;
    if (I >= L2 || J >= L2) return;
    const size_t sL = (size_t)L, sLL = sL * sL;
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
    Rc[(size_t)I + (size_t)L2 * ((size_t)J + (size_t)L2 * K)] =
        (R)Ar<A>::mul(DIM == 3 ? (A).125 : (A).25, s);
}

}  // namespace mg

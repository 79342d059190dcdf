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

// ---------------------------------------------------------------------------------------------------------------
// (K-d2) the same, for the levels between the streaming kernels and the one-CTA kernel (3-D 32^3 and 64^3; 2-D 128^2
// and 256^2 if asked): ONE launch of ONE thread-block cluster. Every CTA of the cluster takes every ncta-th block of
// points of each sweep; a hardware cluster barrier (release / acquire at cluster scope, which also makes the other
// CTAs' global-memory writes visible: the fields stay in L2) stands where the reference has kernel boundaries. The
// levels at and below `a.top1` are handled by the cluster's first CTA alone, with __syncthreads(), exactly like
// k_small_vcycle. At 512^3 this replaces 30 launches (7+7 sweeps + the transfer kernels at 64^3 and 32^3) and the
// one-CTA launch by a single launch.
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_cta_rank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cluster_num_ctas()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}

// one Jacobi sweep of a level by a team of threads -- a warp, a CTA or a whole cluster: the calling thread takes the
// points start, start + step, ...
template <typename R, typename A, int DIM, bool PROLONG>
__device__ __forceinline__ void team_sweep(R *dst, const R *src, const R *f, const R *V, int lg, const Coef<A> &c,
                                           int start, int step)
{
    const int L = 1 << lg;
    const int n = DIM == 3 ? (L * L * L) : (L * L);
    const size_t sL = (size_t)L, sLL = sL * sL;
    for (int idx = start; idx < n; idx += step) {
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
__device__ __forceinline__ void team_residual_restrict(R *Rc, const R *f, const R *u, int lg, const Coef<A> &c,
                                                       int start, int step)
{
    const int L = 1 << lg, lg2 = lg - 1, L2 = L >> 1;
    const int n2 = DIM == 3 ? (L2 * L2 * L2) : (L2 * L2);
    const size_t sL = (size_t)L, sLL = sL * sL;
    for (int cidx = start; cidx < n2; cidx += step) {
        int I = cidx & (L2 - 1), J = (cidx >> lg2) & (L2 - 1), K = DIM == 3 ? (cidx >> (2 * lg2)) : 0;
        A s = (A)0;
        bool firstc = true;
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
                    s = firstc ? rv : Ar<A>::add(s, rv);
                    firstc = false;
                }
        Rc[cidx] = (R)Ar<A>::mul(DIM == 3 ? (A).125 : (A).25, s);
    }
}

template <typename R, typename A> struct ClusterArgs {
    SmallArgs<R, A> s;   // s.top = widest level (log2) handled here
    int top1;            // levels <= top1 (log2) are done by the first CTA alone
};

template <typename R, typename A, int DIM> __device__ void small_vcycle_body(const SmallArgs<R, A> &a, int top);

template <typename R, typename A, int DIM>
__global__ void __launch_bounds__(1024, 1) k_cluster_vcycle(ClusterArgs<R, A> ca)
{
    const SmallArgs<R, A> &a = ca.s;
    const unsigned me = cluster_cta_rank(), nc = cluster_num_ctas();
    const int t0 = (int)(me * blockDim.x + threadIdx.x), tn = (int)(nc * blockDim.x);
    const int smooth = a.smooth;
    const int par = smooth & 1;
    // ---- descend through the wide levels, every CTA working (cpu-raw.lua:198-218)
    for (int lv = a.top; lv > ca.top1; --lv) {
        R *src = a.u[lv], *dst = a.w[lv];
        for (int s = 0; s < smooth; ++s) {
            team_sweep<R, A, DIM, false>(dst, src, a.f[lv], (const R *)nullptr, lv, a.coef[lv], t0, tn);
            cluster_sync_all();
            R *t = src; src = dst; dst = t;
        }
        team_residual_restrict<R, A, DIM>(const_cast<R *>(a.f[lv - 1]), a.f[lv], src, lv, a.coef[lv], t0, tn);
        cluster_sync_all();
    }
    // ---- the narrow levels: one CTA, __syncthreads() only
    if (me == 0) small_vcycle_body<R, A, DIM>(a, ca.top1);
    cluster_sync_all();
    // ---- ascend (cpu-raw.lua:221-236)
    for (int lv = ca.top1 + 1; lv <= a.top; ++lv) {
        R *src = par ? a.w[lv] : a.u[lv];
        R *dst = par ? a.u[lv] : a.w[lv];
        const R *V = a.u[lv - 1];
        if (smooth == 0) {
            const int L = 1 << lv;
            const int n = DIM == 3 ? (L * L * L) : (L * L);
            for (int idx = t0; idx < n; idx += tn) {
                int i = idx & (L - 1), j = (idx >> lv) & (L - 1), k = DIM == 3 ? (idx >> (2 * lv)) : 0;
                src[idx] = (R)corrected<R, A, DIM>(src, V, i, j, k, L, (size_t)idx);
            }
            cluster_sync_all();
            continue;
        }
        team_sweep<R, A, DIM, true>(dst, src, a.f[lv], V, lv, a.coef[lv], t0, tn);
        cluster_sync_all();
        { R *t = src; src = dst; dst = t; }
        for (int s = 1; s < smooth; ++s) {
            team_sweep<R, A, DIM, false>(dst, src, a.f[lv], (const R *)nullptr, lv, a.coef[lv], t0, tn);
            cluster_sync_all();
            R *t = src; src = dst; dst = t;
        }
    }
}

template <typename R, typename A, int DIM>
__global__ void __launch_bounds__(1024, 1) k_small_vcycle(SmallArgs<R, A> a)
{
    small_vcycle_body<R, A, DIM>(a, a.top);
}

// the V-cycle over the levels top .. 1 by one team: the whole CTA (__syncthreads) or its first warp (__syncwarp)
template <typename R, typename A, int DIM, bool WARP>
__device__ __forceinline__ void team_vcycle(const SmallArgs<R, A> &a, int top, int bottom)
{
    const int t0 = WARP ? (int)(threadIdx.x & 31) : (int)threadIdx.x, tn = WARP ? 32 : (int)blockDim.x;
    auto sync = [] { if (WARP) __syncwarp(); else __syncthreads(); };
    const int smooth = a.smooth;
    const int par = smooth & 1;  // after `smooth` ping-pong sweeps the field sits in w if odd
    // ---- descend: pre-smooth, residual, restrict (cpu-raw.lua:198-218)
    for (int lv = top; lv > bottom; --lv) {
        R *src = a.u[lv], *dst = a.w[lv];
        for (int s = 0; s < smooth; ++s) {
            team_sweep<R, A, DIM, false>(dst, src, a.f[lv], (const R *)nullptr, lv, a.coef[lv], t0, tn);
            sync();
            R *t = src; src = dst; dst = t;
        }
        team_residual_restrict<R, A, DIM>(const_cast<R *>(a.f[lv - 1]), a.f[lv], src, lv, a.coef[lv], t0, tn);
        sync();
    }
    if (bottom == 0) {
        // ---- L = 1: one smoother call (cpu-raw.lua:190-196): every neighbour is out of range
        if (t0 == 0) {
            A S = Ar<A>::add(Ar<A>::add(Ar<A>::add((A)0, (A)0), (A)0), (A)0);
            if (DIM == 3) S = Ar<A>::add(Ar<A>::add(S, (A)0), (A)0);
            a.u[0][0] = (R)relax<A>(jacobi_point<DIM, A>(S, (A)a.f[0][0], a.coef[0]), (A)a.u[0][0], a.coef[0]);
        }
        sync();
    } else if constexpr (!WARP) {
        // ---- the levels below `bottom`: the first warp alone, warp-synchronous (a block barrier costs ~0.45 us per
        // phase with 32 warps; 15 phases per level). Everybody else waits for it at one block barrier.
        if (threadIdx.x < 32) team_vcycle<R, A, DIM, true>(a, bottom, 0);
        __syncthreads();
    }
    // ---- ascend: prolong, add, post-smooth (cpu-raw.lua:221-236)
    for (int lv = bottom + 1; lv <= top; ++lv) {
        R *src = par ? a.w[lv] : a.u[lv];
        R *dst = par ? a.u[lv] : a.w[lv];
        const R *V = a.u[lv - 1];
        if (smooth == 0) {
            const int L = 1 << lv;
            const int n = DIM == 3 ? (L * L * L) : (L * L);
            for (int idx = t0; idx < n; idx += tn) {
                int i = idx & (L - 1), j = (idx >> lv) & (L - 1), k = DIM == 3 ? (idx >> (2 * lv)) : 0;
                src[idx] = (R)corrected<R, A, DIM>(src, V, i, j, k, L, (size_t)idx);
            }
            sync();
            continue;
        }
        team_sweep<R, A, DIM, true>(dst, src, a.f[lv], V, lv, a.coef[lv], t0, tn);
        sync();
        { R *t = src; src = dst; dst = t; }
        for (int s = 1; s < smooth; ++s) {
            team_sweep<R, A, DIM, false>(dst, src, a.f[lv], (const R *)nullptr, lv, a.coef[lv], t0, tn);
            sync();
            R *t = src; src = dst; dst = t;
        }
        // 2*smooth sweeps in total at this level => the field is back in a.u[lv]
    }
}

// every level <= top (log2) by ONE CTA; the levels with at most MG_SMALL_WARP_POINTS points by its first warp alone.
// MEASURED (2-D 64^2 fp64, one launch per V-cycle): 256 points -> 53 us against 47 us without (a warp walks 8 points per
// lane through L1 latency; 32 warps hide it, and their barrier costs less than that).
#ifndef MG_SMALL_WARP_LOG2_POINTS
#define MG_SMALL_WARP_LOG2_POINTS 5
#endif
template <typename R, typename A, int DIM> __device__ void small_vcycle_body(const SmallArgs<R, A> &a, int top)
{
    int wl = 0;                                          // widest warp-synchronous level (log2)
    while (DIM * (wl + 1) <= MG_SMALL_WARP_LOG2_POINTS && wl + 1 <= top) ++wl;   // points = 2^(DIM * lv)
    if (wl >= top) {
        if (threadIdx.x < 32) team_vcycle<R, A, DIM, true>(a, top, 0);
        return;
    }
    team_vcycle<R, A, DIM, false>(a, top, wl);
}

}  // namespace mg

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
    pdl_enter();
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
    pdl_enter();
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

// ---------------------------------------------------------------------------------------------------------------
// (K-d3) the one-CTA kernel with the whole sub-hierarchy in SHARED MEMORY (round 2). k_small_vcycle above walks global
// memory: every barrier phase is an L2 round trip (a store invalidates the line in L1) plus ~45 instructions per point
// of index arithmetic and boundary selects -- 0.45-1.2 us per phase, 47 us per cycle at 2-D 64^2 fp64, which is the whole
// of BASELINE config [0] and 16 % of config [4]. Here
//  * every level's u, ping-pong partner and right-hand side live in shared memory for the whole launch, stored with a
//    one-cell border of +0: a neighbour outside the grid reads 0 (cpu-raw.lua:36-39) without any select;
//  * the level width is a template parameter (recursion over levels): index arithmetic is shifts and constant strides;
//  * a thread keeps the right-hand side of its points in registers for all sweeps of a level visit;
//  * a level is worked on by a TEAM of min(points, 1024) threads synchronised by a named barrier of exactly that many
//    threads (a warp: __syncwarp), the rest of the CTA waits at the team barrier of the level above;
//  * prolongation + add is an in-place phase of its own (u + prolong(V) rounded to storage, exactly addTo,
//    cpu-raw.lua:83-85), then `smooth` plain sweeps.
// Global memory is touched twice: all u (the persistent corrections Vs, cpu-raw.lua:164) and the top right-hand side
// come in at the start; all u and the restricted right-hand sides Rs go back at the end.
// Per-point arithmetic: mg_math.cuh as everywhere; summation order ((((xl+xr)+yl)+yr)+zl)+zr.
template <int DIM, int LG> struct SmallGeom {
    static constexpr int L = 1 << LG, P = L + 2;
    static constexpr int N = DIM == 3 ? L * L * L : L * L;            // points
    static constexpr int NP = DIM == 3 ? P * P * P : P * P;           // padded elements
    static constexpr int NPP = (NP + 3) / 4 * 4;                      // one array, rounded up
    static constexpr int TEAM = N >= 1024 ? 1024 : (N >= 32 ? N : 32);
    static constexpr int PT = (N + TEAM - 1) / TEAM;                  // points per thread
    static constexpr int NC = DIM == 3 ? N / 8 : N / 4;               // coarse cells below
    static constexpr int CPT = (NC + TEAM - 1) / TEAM;
};
// element offset of level lg's three arrays (u, w, f): below it lie the levels 0 .. lg-1
template <int DIM> __host__ __device__ constexpr int small_off(int lg)
{
    int o = 0;
    for (int k = 0; k < lg; ++k) {
        const int P = (1 << k) + 2, np = DIM == 3 ? P * P * P : P * P;
        o += 3 * ((np + 3) / 4 * 4);
    }
    return o;
}
template <int DIM> static inline size_t small_smem_bytes(int top, size_t elem) { return (size_t)small_off<DIM>(top + 1) * elem; }
static inline int small_team(int dim, int lg)
{
    const long n = 1L << (dim * lg);
    return n >= 1024 ? 1024 : (n >= 32 ? (int)n : 32);
}

template <int TEAM, int ID> __device__ __forceinline__ void team_sync()
{
    if (TEAM <= 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(TEAM) : "memory");
}
// padded index of point idx = i + L (j + L k)
template <int DIM, int LG> __device__ __forceinline__ int small_pidx(int idx)
{
    constexpr int L = 1 << LG, P = L + 2;
    const int i = idx & (L - 1), j = (idx >> LG) & (L - 1);
    int p = (i + 1) + P * (j + 1);
    if (DIM == 3) p += P * P * ((idx >> (2 * LG)) + 1);
    return p;
}
template <typename R, typename A, int DIM, int LG>
__device__ __forceinline__ A small_stencil(const R *src, int p)
{
    constexpr int P = (1 << LG) + 2;
    A S = Ar<A>::add(Ar<A>::add(Ar<A>::add((A)src[p - 1], (A)src[p + 1]), (A)src[p - P]), (A)src[p + P]);
    if (DIM == 3) S = Ar<A>::add(Ar<A>::add(S, (A)src[p - P * P]), (A)src[p + P * P]);
    return S;
}

template <typename R, typename A, int DIM, int LG> __device__ __noinline__ void small_level_smem(const SmallArgs<R, A> &a)
{
    typedef SmallGeom<DIM, LG> G;
    extern __shared__ __align__(16) unsigned char small_smem_raw[];
    R *const u = reinterpret_cast<R *>(small_smem_raw) + small_off<DIM>(LG), *const w = u + G::NPP, *const f = w + G::NPP;
    const int tid = (int)threadIdx.x;
    const Coef<A> c = a.coef[LG];
    if constexpr (LG == 0) {
        // L = 1: one smoother call (cpu-raw.lua:190-196); every neighbour is the +0 border
        if (tid == 0) {
            const int p = small_pidx<DIM, 0>(0);
            const A S = small_stencil<R, A, DIM, 0>(u, p);
            u[p] = (R)relax<A>(jacobi_point<DIM, A>(S, (A)f[p], c), (A)u[p], c);
        }
        __syncwarp();
    } else {
        constexpr int ID = 1 + LG;
        const int smooth = a.smooth;
        int pp[G::PT];
        A fv[G::PT];
#pragma unroll
        for (int m = 0; m < G::PT; ++m) {
            const int idx = tid + m * G::TEAM;
            pp[m] = small_pidx<DIM, LG>(idx < G::N ? idx : 0);
            fv[m] = (A)f[pp[m]];
        }
        const bool mine = G::N >= G::TEAM || tid < G::N;
        R *src = u, *dst = w;
        // all points of the thread together: their loads are issued before anything is stored (src and dst never
        // alias, which the compiler cannot know) and the division guard of mg_math.cuh is one branch for the group
        auto sweep = [&]() {
            A num[G::PT], out[G::PT];
#pragma unroll
            for (int m = 0; m < G::PT; ++m) num[m] = jacobi_num<A>(small_stencil<R, A, DIM, LG>(src, pp[m]), fv[m], c);
            div_adiag_group<DIM, A, G::PT>(num, out, c);
            if (c.weighted) {
#pragma unroll
                for (int m = 0; m < G::PT; ++m) out[m] = relax<A>(out[m], (A)src[pp[m]], c);
            }
            if (mine) {
#pragma unroll
                for (int m = 0; m < G::PT; ++m) dst[pp[m]] = (R)out[m];
            }
            team_sync<G::TEAM, ID>();
            R *t = src; src = dst; dst = t;
        };
        // ---- pre-smoothing, residual, restriction (cpu-raw.lua:198-218)
        for (int s = 0; s < smooth; ++s) sweep();
        {
            typedef SmallGeom<DIM, LG - 1> GC;
            R *const fc = reinterpret_cast<R *>(small_smem_raw) + small_off<DIM>(LG - 1) + 2 * GC::NPP;
            constexpr int P = G::P;
#pragma unroll
            for (int m = 0; m < G::CPT; ++m) {
                const int cidx = tid + m * G::TEAM;
                if (cidx < G::NC) {
                    constexpr int L2 = G::L / 2, LG2 = LG - 1;
                    const int I = cidx & (L2 - 1), J = (cidx >> LG2) & (L2 - 1), K = DIM == 3 ? (cidx >> (2 * LG2)) : 0;
                    const int p0 = (2 * I + 1) + P * (2 * J + 1) + (DIM == 3 ? P * P * (2 * K + 1) : 0);
                    A sacc = (A)0;
                    bool firstc = true;
#pragma unroll
                    for (int dk = 0; dk < (DIM == 3 ? 2 : 1); ++dk)
#pragma unroll
                        for (int dj = 0; dj < 2; ++dj)
#pragma unroll
                            for (int di = 0; di < 2; ++di) {
                                const int p = p0 + di + P * dj + P * P * dk;
                                const A S = small_stencil<R, A, DIM, LG>(src, p);
                                const A rv = (A)(R)residual_point<A>(S, (A)f[p], (A)src[p], c);
                                sacc = firstc ? rv : Ar<A>::add(sacc, rv);
                                firstc = false;
                            }
                    fc[small_pidx<DIM, LG - 1>(cidx)] = (R)Ar<A>::mul(DIM == 3 ? (A).125 : (A).25, sacc);
                }
            }
            team_sync<G::TEAM, ID>();
            // ---- the level below, by its own (smaller or equal) team; everybody else waits for it here
            if (tid < GC::TEAM) small_level_smem<R, A, DIM, LG - 1>(a);
            team_sync<G::TEAM, ID>();
            // ---- prolongation + add, in place (cpu-raw.lua:221-233): every thread corrects its own points
            const R *const V = reinterpret_cast<R *>(small_smem_raw) + small_off<DIM>(LG - 1);
#pragma unroll
            for (int m = 0; m < G::PT; ++m) {
                const int idx = tid + m * G::TEAM;
                if (idx < G::N) {
                    const int i = idx & (G::L - 1), j = (idx >> LG) & (G::L - 1), k = DIM == 3 ? (idx >> (2 * LG)) : 0;
                    const int pc = ((i >> 1) + 1) + GC::P * ((j >> 1) + 1) + (DIM == 3 ? GC::P * GC::P * ((k >> 1) + 1) : 0);
                    src[pp[m]] = (R)Ar<A>::add((A)src[pp[m]], (A)V[pc]);
                }
            }
            team_sync<G::TEAM, ID>();
        }
        // ---- post-smoothing (cpu-raw.lua:234-236): 2 * smooth sweeps in total => the field is back in u
        for (int s = 0; s < smooth; ++s) sweep();
    }
}

template <typename R, typename A, int DIM, int LGMAX> struct SmallDispatch {
    static __device__ __forceinline__ void run(const SmallArgs<R, A> &a)
    {
        if (a.top == LGMAX) small_level_smem<R, A, DIM, LGMAX>(a);
        else if constexpr (LGMAX > 0) SmallDispatch<R, A, DIM, LGMAX - 1>::run(a);
    }
};
// widest level (log2) the shared-memory kernel is instantiated for: 2-D 64^2, 3-D 16^3 (3 arrays per level with borders:
// 143 KB / 163 KB for 8-byte reals)
template <int DIM> struct SmallSmemMax { static constexpr int LG = DIM == 3 ? 4 : 6; };

template <typename R, typename A, int DIM>
__global__ void __launch_bounds__(1024, 1) k_small_vcycle_smem(const __grid_constant__ SmallArgs<R, A> a)
{
    pdl_enter();
    extern __shared__ __align__(16) unsigned char small_smem_raw[];
    R *const sm = reinterpret_cast<R *>(small_smem_raw);
    const int tid = (int)threadIdx.x, nt = (int)blockDim.x, top = a.top;
    const int total = small_off<DIM>(top + 1);
    for (int e = tid; e < total; e += nt) sm[e] = (R)0;     // borders (and everything else) +0
    __syncthreads();
    // ---- in: every level's u (the corrections persist between cycles), the top level's right-hand side
    // (the loads of a thread are all requested before the first of them is stored: one round trip, not one per level.
    // Only the widest level the kernel is built for has more than one point per thread.)
    {
        constexpr int LGM = SmallSmemMax<DIM>::LG, PTM = SmallGeom<DIM, LGM>::PT;
        R u1[LGM + 1], ux[PTM > 1 ? PTM - 1 : 1], fx[PTM];
#pragma unroll
        for (int lg = 0; lg <= LGM; ++lg) u1[lg] = (lg <= top && tid < (1 << (DIM * lg))) ? a.u[lg][tid] : (R)0;
#pragma unroll
        for (int m = 1; m < PTM; ++m) ux[m - 1] = top == LGM ? a.u[LGM][tid + m * nt] : (R)0;
#pragma unroll
        for (int m = 0; m < PTM; ++m) fx[m] = tid + m * nt < (1 << (DIM * top)) ? a.f[top][tid + m * nt] : (R)0;
        auto put = [&](int lg, int idx, R uval, bool with_f, R fval) {
            const int L = 1 << lg, P = L + 2;
            const int np = ((DIM == 3 ? P * P * P : P * P) + 3) / 4 * 4;
            const int i = idx & (L - 1), j = (idx >> lg) & (L - 1), k = DIM == 3 ? (idx >> (2 * lg)) : 0;
            const int p = (i + 1) + P * (j + 1) + (DIM == 3 ? P * P * (k + 1) : 0);
            R *const us = sm + small_off<DIM>(lg);
            us[p] = uval;
            if (with_f) us[2 * np + p] = fval;
        };
#pragma unroll
        for (int lg = 0; lg <= LGM; ++lg)
            if (lg <= top && tid < (1 << (DIM * lg))) put(lg, tid, u1[lg], lg == top, fx[0]);
#pragma unroll
        for (int m = 1; m < PTM; ++m)
            if (top == LGM) put(LGM, tid + m * nt, ux[m - 1], true, fx[m]);
    }
    __syncthreads();
    SmallDispatch<R, A, DIM, SmallSmemMax<DIM>::LG>::run(a);
    __syncthreads();
    // ---- out: every u, and the restricted right-hand sides below the top (Rs, as the global-memory kernel leaves them)
    for (int lg = 0; lg <= top; ++lg) {
        const int L = 1 << lg, P = L + 2, n = DIM == 3 ? L * L * L : L * L;
        const int np = ((DIM == 3 ? P * P * P : P * P) + 3) / 4 * 4;
        const R *const us = sm + small_off<DIM>(lg);
        for (int idx = tid; idx < n; idx += nt) {
            const int i = idx & (L - 1), j = (idx >> lg) & (L - 1), k = DIM == 3 ? (idx >> (2 * lg)) : 0;
            const int p = (i + 1) + P * (j + 1) + (DIM == 3 ? P * P * (k + 1) : 0);
            a.u[lg][idx] = us[p];
            if (lg < top) const_cast<R *>(a.f[lg])[idx] = us[2 * np + p];
        }
    }
}


}  // namespace mg

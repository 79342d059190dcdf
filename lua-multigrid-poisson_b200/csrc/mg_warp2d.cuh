// mg_warp2d.cuh -- (K-a/K-b/K-c, 2-D) temporally blocked Jacobi smoother, warp-synchronous.
//
// 2-D is the reference's own dimension (cpu-raw.lua indexes i + L*j throughout). One launch
// performs S <= 7 Jacobi sweeps (cpu-raw.lua:34-44), optionally starting from
// src + prolong(V) (PRO: cpu-raw.lua:65-73,83-85) and optionally followed by
// Rout = restrict(f - A u) (RES: cpu-raw.lua:46-63) -- a whole pre- or post-smoothing leg of
// twoGrid (cpu-raw.lua:198-218 / 225-236) in ONE pass over the level.
//
// A warp owns a strip of 128 columns (4 per lane, one 128-bit load per lane per row) and
// streams down the rows. The S sweeps (+ the residual stage) are a register pipeline: a row
// leaves stage s and enters stage s+1 in the same step, in registers; x-neighbours across
// lanes come from __shfl_up/down, the y-1 neighbour is the previous row (register `prev`),
// the y+1 neighbour completes the pending sum `acc` one step later. No __syncthreads(), no
// intermediate field ever touches L2/HBM. The summation order ((xl+xr)+yl)+yr of the reference is
// preserved => bit-identical to one sweep per launch.
// 4-byte reals: the rows a warp reads from global memory (f, the source, PRO's coarse values) arrive
// through per-warp shared-memory rings filled by cp.async three rows ahead (MG_W2D_RING below); every
// lane reads back only what it copied itself, so the rings need no barrier either. 8-byte reals read
// their rows with plain (L1-cached) loads: the rings measured slower there.
// fp32 arithmetic is issued as packed FADD2 / FFMA2 / FMUL2 (two IEEE-rn operations per issue slot); the
// rows where every stage is active and inside the grid run a predicate-free body.
// (MEASURED: making the stages of a step independent -- stage s+1 consuming what stage s emitted a
// step earlier, as the 3-D kernel does -- costs 60-120 more registers and halves the resident warps:
// 4096^2 fp32 fell from 1391 to 1150-1232 V-cycles/s. The chained form with 16 warps per SM stays.)
//
// Columns/rows outside the grid stay exactly 0 at every stage (Dirichlet rule,
// cpu-raw.lua:36-39). Lanes near the strip edge compute garbage that never reaches the
// interior TXU = 128 - 2*HX columns the warp stores.
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

#include "mg_math.cuh"

#ifndef MG_WARP2D_MIN_CTAS
#define MG_WARP2D_MIN_CTAS 4   // resident CTAs (of 4 warps) per SM the register allocation must allow
#endif
#ifndef MG_PACKED_F32
#define MG_PACKED_F32 1       // fp32 stage arithmetic with Blackwell's packed FADD2/FFMA2/FMUL2
#endif

// Cache prefetch of the rows a warp will need a few steps from now (f for stage 1 -- its first touch is a DRAM/L2 miss on
// the row-to-row critical path -- and, where the source row is not fetched ahead in registers (fp64), the source row):
// 0 off, 1 prefetch.global.L1, 2 prefetch.global.L2. MG_W2D_PFD = how many rows ahead.
#ifndef MG_W2D_PF
#define MG_W2D_PF 0
#endif
#ifndef MG_W2D_PFD
#define MG_W2D_PFD 2
#endif

// f rows through a per-warp shared-memory ring filled by cp.async (LDGSTS): every f row is fetched from global memory ONCE,
// MG_W2D_PFROWS rows before stage 1 needs it and without holding registers, and the NST stages read it back with LDS (a
// lane only ever reads the 16-byte chunks it copied itself, so cp.async.wait_group is all the synchronisation there is).
// Before: NST global loads of the row (one miss + NST - 1 L1 hits), the miss on the row-to-row critical path.
// Where the source row is not fetched ahead in registers (8-byte accumulators) it takes the same route.
#ifndef MG_W2D_RING
#define MG_W2D_RING 1         // 0 off, 1 4-byte reals only, 2 every real kind
#endif
#ifndef MG_W2D_SRCRING
#define MG_W2D_SRCRING 2      // source rows (and, PRO, their coarse values) through rings too: 0 never, 1 where they are not
                              // fetched ahead in registers (8-byte accumulators), 2 always (no register prefetch at all)
#endif
#ifndef MG_W2D_SRC64
#define MG_W2D_SRC64 0        // 8-byte reals: 1 = the source rows (and PRO's coarse values) through a ring although f stays on L1 loads
#endif
#ifndef MG_W2D_PFROWS
#define MG_W2D_PFROWS 3
#endif

namespace mg {

constexpr int w2d_pow2ceil(int x) { int p = 1; while (p < x) p <<= 1; return p; }

// 16 bytes global -> shared, asynchronously; nbytes = 0 writes zeros (rows / columns outside the grid)
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void *g, uint32_t nbytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(g), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t saddr, const void *g, uint32_t nbytes)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(saddr), "l"(g), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// one lane's four values of a ring row: fp32 one 16-byte chunk at lane*16; fp64 two, the second 512 bytes further on
// (both layouts are bank-conflict free for 128-bit accesses)
template <typename R> __device__ __forceinline__ void ring_copy(uint32_t saddr, const R *g, bool ok)
{
    cp_async16(saddr, g, ok ? 16u : 0u);
    if (sizeof(R) == 8) cp_async16(saddr + 512u, g + 2, ok ? 16u : 0u);
}
__device__ __forceinline__ void ring_read(uint32_t saddr, float *o)
{
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]) : "r"(saddr));
}
__device__ __forceinline__ void ring_read(uint32_t saddr, double *o)
{
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(o[0]), "=d"(o[1]) : "r"(saddr));
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+512];" : "=d"(o[2]), "=d"(o[3]) : "r"(saddr));
}

// a lane's two coarse values (PRO): 8 bytes (fp32) or 16 (fp64) per lane and coarse row
__device__ __forceinline__ void ring_copy2(uint32_t saddr, const float *g, bool ok) { cp_async8(saddr, g, ok ? 8u : 0u); }
__device__ __forceinline__ void ring_copy2(uint32_t saddr, const double *g, bool ok) { cp_async16(saddr, g, ok ? 16u : 0u); }
__device__ __forceinline__ void ring_read2(uint32_t saddr, float *o)
{
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(o[0]), "=f"(o[1]) : "r"(saddr));
}
__device__ __forceinline__ void ring_read2(uint32_t saddr, double *o)
{
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(o[0]), "=d"(o[1]) : "r"(saddr));
}

__device__ __forceinline__ void w2d_prefetch(const void *p)
{
#if MG_W2D_PF == 1
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#elif MG_W2D_PF == 2
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

template <typename T> __device__ __forceinline__ T shfl_up1(T v) { return __shfl_up_sync(0xffffffffu, v, 1); }
template <typename T> __device__ __forceinline__ T shfl_dn1(T v) { return __shfl_down_sync(0xffffffffu, v, 1); }

template <typename R> __device__ __forceinline__ void load4(const R *p, R *o);
template <> __device__ __forceinline__ void load4<float>(const float *p, float *o)
{
    float4 v = *(const float4 *)p;
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <> __device__ __forceinline__ void load4<double>(const double *p, double *o)
{
    double2 a = *(const double2 *)p, b = *(const double2 *)(p + 2);
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}
template <typename R> __device__ __forceinline__ void store4(R *p, const R *o);
template <> __device__ __forceinline__ void store4<float>(float *p, const float *o)
{
    *(float4 *)p = make_float4(o[0], o[1], o[2], o[3]);
}
template <> __device__ __forceinline__ void store4<double>(double *p, const double *o)
{
    *(double2 *)p = make_double2(o[0], o[1]);
    *(double2 *)(p + 2) = make_double2(o[2], o[3]);
}

template <int S, bool RES> struct Warp2DCfg {
    static constexpr int NST = S + (RES ? 1 : 0);
    static constexpr int H = NST;
    static constexpr int HX = (H + 3) / 4 * 4;
    static constexpr int TXU = 128 - 2 * HX;  // columns stored per warp
    // shared-memory rings (MG_W2D_RING): f rows q - NST .. q - 1 + PF are live at step q, source rows q .. q + PF
    static constexpr int PF = MG_W2D_PFROWS;
    static constexpr int FRING = w2d_pow2ceil(NST + PF);
    static constexpr int SRING = w2d_pow2ceil(PF + 1);
    template <typename R> static constexpr bool f_ring() { return MG_W2D_RING == 2 || (MG_W2D_RING == 1 && sizeof(R) == 4); }
    template <typename R, typename A> static constexpr bool src_ring()
    {
        return (f_ring<R>() && (MG_W2D_SRCRING == 2 || (MG_W2D_SRCRING == 1 && sizeof(A) != 4))) ||
               (MG_W2D_SRC64 != 0 && sizeof(R) == 8);
    }
    template <typename R, typename A> static constexpr int warp_bytes()
    {
        // f rows, source rows, and (PRO) the coarse values of the source rows: 2 per lane and row
        return (f_ring<R>() ? FRING * 128 * (int)sizeof(R) : 0) +
               (src_ring<R, A>() ? SRING * (128 + 64) * (int)sizeof(R) : 0);
    }
    template <typename R, typename A> static constexpr int smem_bytes() { return 4 * warp_bytes<R, A>(); }
};

template <typename R, typename A, int S, bool PRO, bool RES>
__global__ void __launch_bounds__(128, MG_WARP2D_MIN_CTAS)
k_warp2d(R *__restrict__ dst, const R *__restrict__ src, const R *__restrict__ f, const R *__restrict__ Vp,
         R *__restrict__ Rout, int L, int TY, int nstrips, int nitems, Coef<A> cf)
{
    pdl_enter();
    typedef Warp2DCfg<S, RES> C;
    constexpr int NST = C::NST, H = C::H;
    constexpr bool PACKED = MG_PACKED_F32 != 0 && std::is_same<A, float>::value && std::is_same<R, float>::value;
    const int item = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (item >= nitems) return;  // warp-uniform
    const int bx = item % nstrips, by = item / nstrips;
    const int x0 = bx * C::TXU, y0 = by * TY;
    const int y1 = min(y0 + TY, L);
    const int gx0 = x0 - C::HX + 4 * lane;
    const bool xin = gx0 >= 0 && gx0 < L;  // L % 4 == 0: the four columns are in or out together
    const bool xst = xin && (4 * lane >= C::HX) && (4 * lane < C::HX + C::TXU);
    const int yb = y0 - H, nin = (y1 - y0) + 2 * H;
    const size_t sL = (size_t)L;
    const int L2 = L >> 1;
    // the strip (with its halo columns) lies inside the grid: no column masks
    const bool strip_inner = (x0 - C::HX >= 0) && (x0 - C::HX + 128 <= L);

    constexpr bool RING = C::template f_ring<R>();
    constexpr bool SRCRING = C::template src_ring<R, A>();
    constexpr int ROWB = 128 * (int)sizeof(R), PF = C::PF;
    extern __shared__ __align__(16) unsigned char w2d_smem[];
    // this lane's chunk of row slot 0 of its warp's f ring; the source ring follows the f ring
    constexpr bool ANYRING = RING || SRCRING;
    const uint32_t fring = ANYRING ? (uint32_t)__cvta_generic_to_shared(w2d_smem) +
                                         (uint32_t)((threadIdx.x >> 5) * C::template warp_bytes<R, A>() + lane * 16) : 0u;
    const uint32_t sring = fring + (uint32_t)(RING ? C::FRING * ROWB : 0);
    constexpr int VROWB = 64 * (int)sizeof(R);
    const uint32_t vring = sring + (uint32_t)(C::SRING * ROWB) - (uint32_t)(lane * 16) + (uint32_t)(lane * 2 * (int)sizeof(R));
    // request f row `qf` (and, SRCRING, source row `qs`) as one cp.async group
    auto request = [&](const int qf, const int qs) {
        if (ANYRING) {
            if (RING) {
                const bool okf = xin && qf >= 0 && qf < L;
                ring_copy<R>(fring + (uint32_t)((qf & (C::FRING - 1)) * ROWB), okf ? f + (size_t)gx0 + sL * (size_t)qf : f, okf);
            }
            if (SRCRING) {
                const bool oks = xin && qs >= 0 && qs < L;
                ring_copy<R>(sring + (uint32_t)((qs & (C::SRING - 1)) * ROWB), oks ? src + (size_t)gx0 + sL * (size_t)qs : src, oks);
                if (PRO)
                    ring_copy2(vring + (uint32_t)((qs & (C::SRING - 1)) * VROWB),
                               oks ? Vp + (size_t)(gx0 >> 1) + (size_t)L2 * (size_t)(qs >> 1) : Vp, oks);
            }
            cp_async_commit();
        }
    };
    // before step 0: f rows yb - 1 .. yb - 2 + PF and source rows yb .. yb + PF - 1 (step t then requests yb + t - 1 + PF / yb + t + PF)
#pragma unroll
    for (int k = 0; k < PF; ++k) request(yb - 1 + k, yb + k);

    A acc[NST][4], prev[NST][4];
#pragma unroll
    for (int s = 0; s < NST; ++s)
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc[s][i] = (A)0; prev[s][i] = (A)0; }
    A rpart[2] = {(A)0, (A)0};

    // Without a source ring and with fp32 arithmetic the source row (and, PRO, its two coarse values) of the NEXT
    // step is fetched while this step computes, so the global-load latency leaves the row-to-row critical path
    // (+3 % at 4096^2, +10 % on the PRO pass; superseded by the rings in the default build). With 8-byte
    // accumulators the extra registers spill (-5 % at 2048^2 fp64): there the row is fetched at the start of its own step.
    constexpr bool PREFETCH = sizeof(A) == 4 && !C::template src_ring<R, A>();
    R pre[4] = {(R)0, (R)0, (R)0, (R)0}, pv[2] = {(R)0, (R)0};
    auto fetch = [&](const int q, const bool ok) {
#pragma unroll
        for (int i = 0; i < 4; ++i) pre[i] = (R)0;
        pv[0] = pv[1] = (R)0;
        if (SRCRING) {                                                                    // zero-filled outside the grid
            ring_read(sring + (uint32_t)((q & (C::SRING - 1)) * ROWB), pre);
            if (PRO) ring_read2(vring + (uint32_t)((q & (C::SRING - 1)) * VROWB), pv);
        } else if (ok) {
            load4<R>(src + (size_t)gx0 + sL * (size_t)q, pre);
            if (PRO) {
                const R *vp = Vp + (size_t)(gx0 >> 1) + (size_t)L2 * (size_t)(q >> 1);
                pv[0] = vp[0]; pv[1] = vp[1];
            }
        }
    };
    if (PREFETCH) fetch(yb, xin && yb >= 0 && yb < L);

    // One row step. ST: every stage is fed and emits and every row touched lies inside the grid (and, for
    // the stages that store, inside [y0, y1)); MK: columns may lie outside the grid.
    auto step = [&](auto steady_tag, auto masked_tag, const int t) {
        constexpr bool ST = decltype(steady_tag)::value, MK = decltype(masked_tag)::value;
        const int q = yb + t;
        if (MG_W2D_PF != 0 && xin) {
            const int qf = q - 1 + MG_W2D_PFD;               // stage 1 completes row q - 1 now
            if (qf >= 0 && qf < L) w2d_prefetch(f + (size_t)gx0 + sL * (size_t)qf);
            const int qs = q + MG_W2D_PFD + (PREFETCH ? 1 : 0);
            if (qs >= 0 && qs < L) w2d_prefetch(src + (size_t)gx0 + sL * (size_t)qs);
        }
        if (ANYRING) {
            request(q - 1 + PF, q + PF);
            cp_async_wait<PF>();          // everything but the PF newest groups has landed: f row q - 1, source row q
        }
        if (!PREFETCH) fetch(q, ST ? (!MK || xin) : (xin && q >= 0 && q < L));
        R row[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) row[i] = PRO ? (R)Ar<A>::add((A)pre[i], (A)pv[i >> 1]) : pre[i];   // outside the grid: 0 (+ 0)
        if (PREFETCH) fetch(q + 1, ST ? (!MK || xin) : (xin && q + 1 >= 0 && q + 1 < L));
#pragma unroll
        for (int sidx = 0; sidx < NST; ++sidx) {
            const int s = sidx + 1;
            if (!ST && t < 2 * sidx) break;      // stage not fed yet (warp-uniform)
            const bool emit = ST ? true : (t >= 2 * s);
            const int p = q - s;                 // row this stage completes now
            const bool pin = ST ? true : (p >= 0 && p < L);
            const bool keep = pin && (MK ? xin : true);
            const bool is_res = RES && s == NST;
            R fv[4] = {(R)0, (R)0, (R)0, (R)0};
            if (RING) {
                if (emit) ring_read(fring + (uint32_t)((p & (C::FRING - 1)) * ROWB), fv);   // zeros outside the grid
            } else if (emit && keep) {
                load4<R>(f + (size_t)gx0 + sL * (size_t)p, fv);
            }
            const R lft = shfl_up1(row[3]), rgt = shfl_dn1(row[0]);
            A o[4];
            if constexpr (PACKED) {
                auto P = [](float a, float b) { return make_float2(a, b); };
                // xl + xr pairs operands one element apart: scalar adds that land in aligned pairs
                const float2 sx[2] = {P(__fadd_rn(lft, row[1]), __fadd_rn(row[0], row[2])),
                                      P(__fadd_rn(row[1], row[3]), __fadd_rn(row[2], rgt))};
                const float2 Cc[2] = {P(row[0], row[1]), P(row[2], row[3])};
                const float2 NINV = P(-cf.inv_h2, -cf.inv_h2), INV = P(cf.inv_h2, cf.inv_h2), AD = P(cf.adiag, cf.adiag);
                const float2 CN = P(cf.cneg, cf.cneg), M1 = P(-1.f, -1.f);
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const float2 PRV = P(prev[sidx][2 * k], prev[sidx][2 * k + 1]);
                    const float2 Tt = __fadd2_rn(P(acc[sidx][2 * k], acc[sidx][2 * k + 1]), Cc[k]);   // pending row gets its y+1
                    const float2 F = P(fv[2 * k], fv[2 * k + 1]);
                    float2 ov;
                    if (is_res) {
                        const float2 au = __fadd2_rn(__fmul2_rn(Tt, INV), __fmul2_rn(AD, PRV));
                        ov = __ffma2_rn(au, M1, F);                                                    // f - au
                    } else {
                        ov = __fmul2_rn(__ffma2_rn(Tt, NINV, F), CN);    // RN(f - S/h^2) * (-h^2/4): exact scaling (mg_math.cuh)
                    }
                    o[2 * k] = ov.x; o[2 * k + 1] = ov.y;
                    const float2 NA = __fadd2_rn(sx[k], PRV);                                          // (xl+xr)+yl of this row
                    acc[sidx][2 * k] = NA.x; acc[sidx][2 * k + 1] = NA.y;
                    prev[sidx][2 * k] = Cc[k].x; prev[sidx][2 * k + 1] = Cc[k].y;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const A xl = (A)(i == 0 ? lft : row[i - 1]), xr = (A)(i == 3 ? rgt : row[i + 1]);
                    const A c = (A)row[i];
                    const A tot = Ar<A>::add(acc[sidx][i], c);                       // pending row gets its y+1
                    o[i] = is_res ? residual_point<A>(tot, (A)fv[i], prev[sidx][i], cf)
                                  : jacobi_point<2, A>(tot, (A)fv[i], cf);
                    acc[sidx][i] = Ar<A>::add(Ar<A>::add(xl, xr), prev[sidx][i]);    // (xl+xr)+yl of this row
                    prev[sidx][i] = c;
                }
            }
            if (!emit) break;
            R outv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) outv[i] = (ST && !MK) ? (R)o[i] : (keep ? (R)o[i] : (R)0);
            if (!is_res) {
                if (s == S && (ST || (p >= y0 && p < y1)) && xst) store4<R>(dst + (size_t)gx0 + sL * (size_t)p, outv);
#pragma unroll
                for (int i = 0; i < 4; ++i) row[i] = outv[i];
            } else if (ST || (p >= y0 && p < y1)) {
                // restriction, children in the reference's order (cpu-raw.lua:62)
                if ((p & 1) == 0) {
                    rpart[0] = Ar<A>::add((A)outv[0], (A)outv[1]);
                    rpart[1] = Ar<A>::add((A)outv[2], (A)outv[3]);
                } else if (xst) {
                    R *rp = Rout + (size_t)(gx0 >> 1) + (size_t)L2 * (size_t)(p >> 1);
                    rp[0] = (R)Ar<A>::mul((A).25, Ar<A>::add(Ar<A>::add(rpart[0], (A)outv[0]), (A)outv[1]));
                    rp[1] = (R)Ar<A>::mul((A).25, Ar<A>::add(Ar<A>::add(rpart[1], (A)outv[2]), (A)outv[3]));
                }
            }
        }
    };

    // Steady range: every stage emits (t >= 2*NST), the last stage's row q - NST is inside the grid, the
    // source row q AND the prefetched row q + 1 are inside the grid, and the storing stages' rows are inside
    // [y0, y1): the last Jacobi stage's row q - S < y1 (its lower bound and the residual stage's follow from
    // t >= 2*NST and H = NST).
    const int t_lo = max(2 * NST, NST - yb);
    const int t_hi = min(min(nin - 1, L - 2 - yb), (y1 - 1) + S - yb);
    int t = 0;
    for (; t < min(t_lo, nin); ++t) step(std::false_type{}, std::true_type{}, t);
    if (strip_inner) {
        for (; t <= t_hi; ++t) step(std::true_type{}, std::false_type{}, t);
    } else {
        for (; t <= t_hi; ++t) step(std::true_type{}, std::true_type{}, t);
    }
    for (; t < nin; ++t) step(std::false_type{}, std::true_type{}, t);
    if (ANYRING) cp_async_wait<0>();
}

}  // namespace mg

// mg_stream3d.cuh -- (K-a/K-b/K-c) temporally blocked 3-D Jacobi smoother for sm_100a.
//
// One launch performs S Jacobi sweeps (cpu-raw.lua:34-44 semantics, 3-D rules of SURVEY 8(a'))
// on a level, optionally
//   PRO: starting from  src + prolong(V)  (expandResidual + addTo, cpu-raw.lua:65-73,83-85)
//   RES: followed by    Rout = restrict(f - A u)  (calcResidual + reduceResidual, :46-63)
// reading src once and writing dst once: HBM traffic per launch ~ 3 words/point instead of
// 3*S (+6 for the transfer operators).
//
// Structure ("3.5-D blocking"): a CTA owns an in-plane tile TX x TY plus halo and streams
// along z. Plane a_t of the source arrives in shared memory by TMA (cp.async.bulk.tensor,
// mbarrier-signalled, NSLOT-deep ring); TMA's out-of-bounds ZERO FILL is the reference's
// Dirichlet rule "a neighbour outside the grid reads 0" (cpu-raw.lua:36-39), so domain edges
// need no branches on load. The S sweeps (+1 residual stage) form a software pipeline of
// NST stages; at step t stage s consumes plane a_t - 2(s-1) of stage s-1's output and emits its
// own plane one below it. A thread owns VX x 2 columns for the whole launch and keeps, per stage:
//   prev  = centre value of the previous plane            (the z-1 neighbour)
//   acc   = ((xl+xr)+yl)+yr + zl of the pending plane      (waiting for its z+1 neighbour)
//   carry = its own output of that stage from the last step (fp32): the next stage's centre rows
// so of each plane of each stage only the rows ABOVE and BELOW the unit (other threads' values)
// are read back from the shared-memory ring (2 x LDS.128 per 8 points), x-neighbours come from
// warp shuffles, and ONE __syncthreads() per step serves all stages.
// The summation order ((((xl+xr)+yl)+yr)+zl)+zr is preserved, so results are bit-identical
// to the one-sweep-per-launch kernels (all arithmetic from mg_math.cuh).
//
// Cells outside the grid must stay exactly 0 at every stage: in-plane via a per-thread
// bit mask, whole planes via a CTA-uniform test. Values near the tile edge that lack a
// neighbour are garbage by construction and never reach the H-deep interior (H = NST).
//
// Also in this kernel (details at the code):
//  * f planes travel through their own TMA ring (2*NST+1 slots); zero fill makes them predicate free.
//    The f plane and the source plane fetched in the same step are counted on ONE mbarrier: a step
//    has a single wait.
//  * fill / steady / drain are separate loops; the steady body has no per-stage predicates and,
//    on tiles inside the grid, no masks.
//  * the fp32 tile is 56 x 40 (+ halo = 64 wide): 16 vectors per row, so a warp holds whole rows of
//    units, 128-bit accesses are bank-conflict free and the shuffles need no edge patch.
//  * all shared-memory traffic of the hot loop is addressed as 32-bit shared address + uniform slot
//    offset + immediate; the global stores follow running pointers.
//  * fp32 arithmetic is issued as Blackwell packed FADD2 / FFMA2 / FMUL2 (bit-identical per lane).
//  * work is dealt to one CTA per SM in equal shares of tile x plane-pair units (run_chunk), so
//    tiles need not divide the grid and there is no tail wave.
//  * multi-GPU: slab views (local vs global planes), boundary planes stored straight into the
//    neighbours' ghost planes over NVLink peer memory, and the per-pass handshake
//    (acquire on entry, last CTA publishes) -- see mg_slab.cuh.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "mg_math.cuh"

#ifndef MG_STREAM_SHFL
#define MG_STREAM_SHFL 1      // x-neighbours across units via warp shuffle (0: 4-byte shared loads)
#endif
#ifndef MG_STREAM_CARRY
#define MG_STREAM_CARRY 1     // fp32: a stage's own output rows reach the next stage through registers
#endif
#ifndef MG_PACKED_F32
#define MG_PACKED_F32 1       // fp32 stage arithmetic with Blackwell's packed FADD2/FFMA2/FMUL2
#endif
#ifndef MG_STREAM_DEFER
#define MG_STREAM_DEFER 1     // fp32: ring stores of a step are issued together at its end (from the carry registers), so
                              // that no shared-memory store separates one stage's loads from the previous stage's work
#endif
#ifndef MG_STREAM_EARLY
#define MG_STREAM_EARLY 0     // fp32: the next stage's inputs are requested before the current stage's division guard
                              // (1: neighbour rows + shuffles, 2: also its f rows)
#endif
#ifndef MG_F_TMEM
#define MG_F_TMEM 0           // 4-byte reals: the f planes a thread needs live in TENSOR MEMORY (thread-private columns,
                              // tcgen05.st / tcgen05.ld) instead of a shared-memory ring fed by TMA. MEASURED (round 2, 512^3
                              // fp32, S = 4 pass): shared-memory wavefronts 97 M -> 68 M as intended, bit-identical, but
                              // 0.47 -> 0.64 ms: an LDTM.x8 costs the SM ~16 cycles (64 B/clk) against 8 for the two LDS.128
                              // it replaces, and it does not overlap them; requesting a stage ahead did not help. Off.
#endif
#ifndef MG_SLAB_DEFER_HI
#define MG_SLAB_DEFER_HI 0    // multi-GPU: the wait for the UPPER neighbour is deferred to the first request of a plane near the
                              // top of the slab (0: both neighbours are awaited on entry)
#endif
#ifndef MG_STEADY_UNROLL
#define MG_STEADY_UNROLL 2    // unroll factor of the steady-state step loop
#endif

namespace mg {

constexpr int kSteadyUnroll = MG_STEADY_UNROLL;
constexpr bool kPackedF32 = MG_PACKED_F32 != 0;

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_inval(uint32_t bar)
{
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// orders this CTA's earlier generic-proxy accesses to shared memory (made visible to the
// calling thread by a preceding barrier) before later async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap *map, int x, int y, int z, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_dst),
        "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar)
        : "memory");
}

// ---- tensor memory (TMEM) as thread-private storage: a warp owns the 32 lanes of its quarter (warp % 4), thread i of
// the warp lane 32 * (warp % 4) + i; the .32x32b shape moves N consecutive 32-bit columns of that lane to / from N registers
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish()
{
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float *o)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]), "=f"(o[4]), "=f"(o[5]), "=f"(o[6]), "=f"(o[7]) : "r"(taddr));
}
// the loaded registers are only valid after this wait: they pass through it so that no use can be scheduled before it
__device__ __forceinline__ void tmem_wait_ld8(float *o)
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(o[0]), "+f"(o[1]), "+f"(o[2]), "+f"(o[3]), "+f"(o[4]), "+f"(o[5]), "+f"(o[6]), "+f"(o[7])::"memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float *o)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "f"(o[0]), "f"(o[1]),
                 "f"(o[2]), "f"(o[3]), "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7]) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t, double *) {}
__device__ __forceinline__ void tmem_wait_ld8(double *) {}
__device__ __forceinline__ void tmem_st8(uint32_t, const double *) {}

// predicated shared-memory load: returns *p if pred, else old (no branch, no access when !pred)
__device__ __forceinline__ float lds_if(bool pred, uint32_t addr, float old)
{
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\n@q ld.shared.f32 %0, [%1];\n}\n"
                 : "+f"(old) : "r"(addr), "r"((unsigned)pred));
    return old;
}
__device__ __forceinline__ double lds_if(bool pred, uint32_t addr, double old)
{
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\n@q ld.shared.f64 %0, [%1];\n}\n"
                 : "+d"(old) : "r"(addr), "r"((unsigned)pred));
    return old;
}
__device__ __forceinline__ float lds1(uint32_t addr, float)
{
    float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); return v;
}
__device__ __forceinline__ double lds1(uint32_t addr, double)
{
    double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr)); return v;
}

// ------------------------------------------------------------------ vector access
#ifndef MG_F32_VX
#define MG_F32_VX 4           // fp32 points per thread and row: 4 (128-bit accesses, 8 points per thread) or 2 (64-bit, 4 points)
#endif
template <typename R> struct Vec;
#if MG_F32_VX == 2
template <> struct Vec<float> {
    static constexpr int N = 2;
    typedef float2 T;
    static __device__ __forceinline__ void unpack(const T &v, float *o) { o[0] = v.x; o[1] = v.y; }
    static __device__ __forceinline__ T pack(const float *o) { return make_float2(o[0], o[1]); }
    static __device__ __forceinline__ T zero() { return make_float2(0.f, 0.f); }
    static __device__ __forceinline__ void lds(uint32_t addr, float *o)
    {
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(o[0]), "=f"(o[1]) : "r"(addr));
    }
    static __device__ __forceinline__ void sts(uint32_t addr, const float *o)
    {
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(o[0]), "f"(o[1]) : "memory");
    }
};
#else
template <> struct Vec<float> {
    static constexpr int N = 4;
    typedef float4 T;
    static __device__ __forceinline__ void unpack(const T &v, float *o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
    static __device__ __forceinline__ T pack(const float *o) { return make_float4(o[0], o[1], o[2], o[3]); }
    static __device__ __forceinline__ T zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
    // 128-bit shared-memory access by 32-bit shared-window address: LDS/STS [reg + imm], no
    // generic-to-shared conversion and no re-derivation of the window base per access
    static __device__ __forceinline__ void lds(uint32_t addr, float *o)
    {
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]) : "r"(addr));
    }
    static __device__ __forceinline__ void sts(uint32_t addr, const float *o)
    {
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3]) : "memory");
    }
};
#endif
template <> struct Vec<double> {
    static constexpr int N = 2;
    typedef double2 T;
    static __device__ __forceinline__ void unpack(const T &v, double *o) { o[0] = v.x; o[1] = v.y; }
    static __device__ __forceinline__ T pack(const double *o) { return make_double2(o[0], o[1]); }
    static __device__ __forceinline__ T zero() { return make_double2(0., 0.); }
    static __device__ __forceinline__ void lds(uint32_t addr, double *o)
    {
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(o[0]), "=d"(o[1]) : "r"(addr));
    }
    static __device__ __forceinline__ void sts(uint32_t addr, const double *o)
    {
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(o[0]), "d"(o[1]) : "memory");
    }
};

constexpr int S3_MAX_RANKS = 8;
// slab handshake words in the arena header (u64 indices): passes completed by this rank / published by the lower and
// upper neighbour, CTAs done, all-gather epochs (mg_slab.cuh), peer time-out flag
enum { HS_DONE = 0, HS_FROM_LO = 16, HS_FROM_HI = 32, HS_CTAS = 48, HS_AG_FROM = 72, HS_AG_EPOCH = 80, HS_TIMEOUT = 81 };
constexpr unsigned long long HS_TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;   // a neighbour that stays silent this long is dead

__device__ __forceinline__ unsigned long long s3_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ unsigned long long s3_ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void s3_st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// spin until *p >= n (acquire, system scope). A peer that stays silent for HS_TIMEOUT_NS is taken for dead: the time-out
// word of the own header is raised (the host reports MG_ESTATE at its next synchronisation) and the wait is given up.
__device__ __forceinline__ void s3_wait_counter(const unsigned long long *p, unsigned long long n, unsigned long long *hs)
{
    if (p == nullptr) return;
    unsigned long long t0 = 0;
    unsigned int spins = 0;
    while (s3_ld_acquire_sys(p) < n) {
        if ((++spins & 1023u) == 0) {
            const unsigned long long now = s3_globaltimer();
            if (t0 == 0) t0 = now;
            else if (now - t0 > HS_TIMEOUT_NS) { hs[HS_TIMEOUT] = 1ull; __threadfence_system(); return; }
        }
    }
}

template <typename R, int S, bool RES, int TX, int TY> struct Stream3DCfg {
    static constexpr int VX = Vec<R>::N;
    static constexpr int NST = S + (RES ? 1 : 0);
    static constexpr int H = NST;
    // x halo rounded to 16 bytes: the TMA box stays 16-byte aligned and (4-byte reals) 64 wide whatever VX is
    static constexpr int HXQ = (VX * (int)sizeof(R) >= 16) ? VX : 16 / (int)sizeof(R);
    static constexpr int HX = (H + HXQ - 1) / HXQ * HXQ;
    static constexpr int HY = RES ? (H + 1) / 2 * 2 : H;
    static constexpr int WX = TX + 2 * HX, WY = TY + 2 * HY;
    static constexpr int UX = WX / VX, UY = WY / 2;
    static constexpr int NT = UX * UY;
    // a warp holds whole rows of units: lanes 0 / 31 sit on a tile edge, where the missing x-neighbour
    // lies in the garbage zone anyway, so the shuffles need no shared-memory patch
    static constexpr bool EDGE_FREE = (32 % UX) == 0;
    static constexpr int NTHREADS = (NT + 31) / 32 * 32;
    static constexpr int PLANE = WX * WY;
    static constexpr int PLANE_BYTES = PLANE * (int)sizeof(R);
    static constexpr int SLOT_BYTES = (PLANE_BYTES + 127) / 128 * 128;
    static constexpr int SLOT_ELEMS = SLOT_BYTES / (int)sizeof(R);
    static constexpr int NSLOT = 3;         // source-plane ring (TMA prefetch distance NSLOT-1 = 2 steps, like f)
    static constexpr int NRING = NST - 1;   // intermediate stage outputs, double buffered
    // f planes. 4-byte reals: in tensor memory, FTD planes of 2*VX columns per thread (a plane is stored at its own step
    // and read by stage s 2s-1 steps later: 2*NST-1 steps of life, 8 is the next power of two), three warps share
    // a lane quarter -> 3 * 64 = 192 of the 256 allocated columns. 8-byte reals: a shared-memory ring fed by TMA.
    static constexpr bool FT = MG_F_TMEM && sizeof(R) == 4 && VX == 4;
    static constexpr int FTD = 8, FT_COLS = 256;
    static constexpr int NF = FT ? 0 : 2 * NST + 1;  // f-plane ring: a plane lives 2*NST-1 steps, fetched 2 ahead
    static constexpr int NSLOTS_TOTAL = NSLOT + 2 * NRING + NF;
    static constexpr int SMEM_BYTES = NSLOTS_TOTAL * SLOT_BYTES + NSLOT * 8 + 16;   // slots + the source ring's mbarriers + TMEM base
    static_assert(NSLOT == 3, "source plane t+2 and f plane t+1 are fetched together and share an mbarrier");
    static_assert(TX % VX == 0 && TY % 2 == 0 && WY % 2 == 0, "tile shape");
    static_assert(NTHREADS <= 1024, "too many threads");
    static_assert(SMEM_BYTES <= 232448, "shared memory budget (227 KB)");
    // two CTAs per SM when the shared memory (plus the 1 KB the system reserves per CTA) allows it
    static constexpr int MIN_CTAS = !FT && 2 * (SMEM_BYTES + 1024) <= 233472 && 2 * NTHREADS <= 1024 ? 2 : 1;
    static_assert(!FT || (NST <= 4 && (NTHREADS + 127) / 128 * 2 * VX * FTD <= FT_COLS), "tensor-memory budget of the f planes");
};

template <typename R> struct Stream3DArgs {
    R *dst;           // u after S sweeps
    const R *Vp;      // PRO: coarse correction Vs[L/2]
    R *Rout;          // RES: Rs[L/2]
    int L;            // level width
    int flags;        // debug: bit 0 = never take the steady-state body, bit 1 = always mask, bit 2 = always re-run chunks guarded
    // slab view (multi-GPU): the arrays hold planes [0, nplanes) of which [nz_lo, nz_hi) are
    // owned (written) by this rank; the global grid occupies local planes [zdom0, zdom1).
    // Single GPU: nz_lo = zdom0 = 0, nz_hi = zdom1 = L, rz_off = vz_off = 0.
    int nz_lo, nz_hi, zdom0, zdom1;
    int rz_off;       // RES: coarse local plane = ((p - nz_lo) >> 1) + rz_off
    int vz_off;       // PRO: coarse local plane = ((p - zdom0) >> 1) + vz_off
    // Fused halo exchange (multi-GPU): the thread that writes one of this rank's G boundary
    // planes also stores it into the neighbour's ghost planes, directly over NVLink (the peers'
    // arenas are mapped with CUDA IPC and have the same layout, so the same element offset
    // applies). peer_lo / peer_hi = the dst field in the lower / upper neighbour (or null);
    // rpeer_lo / rpeer_hi = Rout there (RES, coarse level distributed). ghost = G.
    R *peer_lo, *peer_hi, *rpeer_lo, *rpeer_hi;
    int ghost;
    // Handshake of the fused exchange, folded into the kernel (one process per GPU only):
    // hs = this rank's arena header {[0] passes completed, [16] / [32] passes the lower / upper
    // neighbour has published, [48] CTAs of the current pass that are done}; hs_lo / hs_hi = the
    // neighbours' headers (null at the ends of the chain, and everywhere when hs is null).
    unsigned long long *hs, *hs_lo, *hs_hi;
    // Work partition (all zero = balanced shares of the tile x plane-pair work for every CTA). Lock step: the tile
    // columns [0, ncol) are processed WHOLE over the owned planes [nz_lo, nz_lo + zcol), CTA b taking columns b, b + grid,
    // b + 2 grid, ...: neighbouring columns then march through z together and the halo rows and partly used sectors
    // they share are served by L2 instead of HBM. What is left -- the tiles from rem_tile0 on, planes from nz_lo + rem_z0
    // on -- is dealt out in equal contiguous shares to the CTAs rem_cta0 .. grid-1.
    //   fewer tiles than CTAs : ncol = all tiles, zcol < owned planes, the spare CTAs (rem_cta0 = ncol) share the top planes
    //   more tiles than CTAs  : ncol = whole waves of grid tiles, zcol = all planes, every CTA shares the last partial wave
    int ncol, zcol, rem_cta0, rem_tile0, rem_z0;
    // S3_FAST / S3_RERUN: {flag "repeat this pass with the guarded code", CTA arrival counter} (zero between passes)
    unsigned int *redo;
    const R *fsrc;    // the right-hand side field itself (f planes in tensor memory are fetched by plain loads, not TMA)
    // RES into a REPLICATED coarse level (multi-GPU): the restricted residual is also stored into every other rank's
    // copy of the coarse cube -- the all-gather of the first replicated level, done by the producing threads.
    R *rall[S3_MAX_RANKS];
    int rall_n;
    // Timeline of the slab passes (option "slab_trace", measurement only; null = off): word 0 counts the passes recorded,
    // record i = words 8 + 4 i .. : {globaltimer on entry, after the wait for the lower neighbour, ns CTA 0 spent waiting
    // for the upper neighbour, globaltimer when the last CTA finished}. Written by CTA 0 / the last CTA only.
    unsigned long long *trace;
};
constexpr unsigned int S3_TRACE_CAP = 4096;   // records in the timeline buffer (it wraps)

// MODE: how the division guard of mg_math.cuh (a tiny but non-zero numerator needs IEEE division) is handled.
//   S3_GUARDED  every stage tests its group of numerators and branches to IEEE division (all arithmetic types);
//   S3_FAST     (fp32) Markstein sequence for every point, NO branch in the pipeline: the guard is only a sticky
//               minimum over the numerators' keys; a launch that trips the threshold raises a.redo[0];
//   S3_RERUN    (fp32) launched right behind the S3_FAST kernel with the same arguments: returns at once unless
//               a.redo[0] is raised, else performs the whole pass again with the guarded code (src is never
//               written: the passes ping-pong). Never happens on ordinary data; bit-identical either way.
// The per-stage branch cost 10-20 % of a pass (ptxas does not move the next stage's shared-memory loads across it),
// and keeping both code versions in ONE kernel made the register allocator spill in the fast one.
// Multi-GPU handshake: the entry wait is done by S3_GUARDED / S3_FAST, the publication by S3_GUARDED / S3_RERUN.
enum { S3_GUARDED = 0, S3_FAST = 1, S3_RERUN = 2 };

template <typename R, typename A, int S, bool PRO, bool RES, int TX, int TY, int MODE = S3_GUARDED>
__global__ void __launch_bounds__((Stream3DCfg<R, S, RES, TX, TY>::NTHREADS), (Stream3DCfg<R, S, RES, TX, TY>::MIN_CTAS))
k_stream3d(const __grid_constant__ CUtensorMap src_map, const __grid_constant__ CUtensorMap f_map,
           Stream3DArgs<R> a, Coef<A> cf)
{
    static_assert(MODE == S3_GUARDED || (kPackedF32 && std::is_same<A, float>::value && std::is_same<R, float>::value),
                  "the branch-free division is implemented for fp32 arithmetic");
    static_assert(MG_STREAM_SHFL || !PRO, "the register-path prolongation needs the shuffled x-neighbours");
    pdl_enter();
    constexpr bool GUARDED = MODE != S3_FAST;
    typedef Stream3DCfg<R, S, RES, TX, TY> C;
    typedef typename Vec<R>::T VT;
    constexpr int VX = C::VX, NST = C::NST, H = C::H, NP = 2 * VX, NSLOT = C::NSLOT, NF = C::NF;

    // layout: [NSLOT source slots][NRING*2 stage slots][NF f slots][mbarriers]. The dynamic
    // shared window starts at offset 0 of the CTA's shared memory (no static __shared__ in
    // this kernel), so every slot is 128-byte aligned as TMA requires. All shared-memory
    // traffic of the hot loop is addressed by 32-bit shared-window addresses
    // (per-thread base + slot offset + immediate): one LDS/STS each, no address rebuilding.
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    constexpr uint32_t SB = (uint32_t)C::SLOT_BYTES, ROWB = (uint32_t)(C::WX * sizeof(R));
    constexpr uint32_t RING0 = NSLOT * SB, F0 = (NSLOT + 2 * C::NRING) * SB;
    const uint32_t sbase = smem_u32(smem_raw);
    // One mbarrier per source slot. The f plane fetched together with a source plane (same step, needed at
    // the same later step) is counted on the SAME mbarrier, so a step has a single wait.
    const uint32_t mb_u = sbase + (uint32_t)C::NSLOTS_TOTAL * SB;

    const int tid = threadIdx.x, lane = tid & 31;
    const bool worker = tid < C::NT;   // the last warp's spare threads mirror unit 0 and never store
    const int ux = worker ? tid % C::UX : 0, uy = worker ? tid / C::UX : 0;
    const int L = a.L;
    const int zdom0 = a.zdom0, zdom1 = a.zdom1;
    bool first_chunk = true;

    // Before touching ghost planes (ours, by TMA; the neighbours', by peer stores): both
    // neighbours must have finished as many passes as we have. Every CTA checks for itself.
    bool skip = false;   // S3_RERUN with nothing to redo
    if (MODE == S3_RERUN) {
        // every CTA reads the flag BEFORE it counts itself in; the last one in resets flag and counter for the next pass
        int redo = 0;
        if (threadIdx.x == 0) {
            redo = (int)*reinterpret_cast<volatile unsigned int *>(a.redo);
            __threadfence();
            if (atomicAdd(a.redo + 1, 1u) == gridDim.x - 1) { a.redo[1] = 0u; a.redo[0] = 0u; }
        }
        skip = __syncthreads_or(redo) == 0;
    } else if (a.hs != nullptr) {
        // The LOWER neighbour is needed at once (every column starts at the bottom of the slab: our lower ghost planes
        // are read, its upper ghost planes written, within the first steps). The UPPER neighbour is only needed where a
        // column reaches the top G planes of the slab -- its ghost planes are read, ours stored into it, nowhere else --
        // so that wait is deferred to the first TMA request at or above plane nz_hi - G (hi_wait below): a pass no
        // longer starts late because the rank above finished the previous one late. Thread 0 issues every TMA request
        // and every other thread meets it at the step barrier before it can store anything of those planes.
        if (threadIdx.x == 0) {
            const bool tr = a.trace != nullptr && blockIdx.x == 0;
            unsigned long long *rec = nullptr;
            if (tr) { rec = a.trace + 8 + 4 * (a.trace[0] & (S3_TRACE_CAP - 1)); rec[0] = s3_globaltimer(); rec[2] = 0ull; }
            const unsigned long long n = s3_ld_acquire_sys(a.hs + HS_DONE);
            s3_wait_counter(a.hs_lo != nullptr ? a.hs + HS_FROM_LO : nullptr, n, a.hs);
            if (!MG_SLAB_DEFER_HI) s3_wait_counter(a.hs_hi != nullptr ? a.hs + HS_FROM_HI : nullptr, n, a.hs);
            if (tr) rec[1] = s3_globaltimer();
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(mb_u + NSLOT * 8 + 8), "r"(0u) : "memory");   // "upper neighbour awaited"
        }
        __syncthreads();
    }
    // thread 0, before it requests source plane `plane` (local index): see above. HS_DONE does not change during a pass
    // (the last CTA to finish bumps it), so it is read again here instead of living in a register.
    auto hi_wait = [&](int plane) {
        if (!MG_SLAB_DEFER_HI || MODE == S3_RERUN || a.hs_hi == nullptr || plane < a.nz_hi - a.ghost) return;
        unsigned int waited;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(waited) : "r"(mb_u + NSLOT * 8 + 8) : "memory");
        if (waited) return;
        const bool tr = a.trace != nullptr && blockIdx.x == 0;
        const unsigned long long t0 = tr ? s3_globaltimer() : 0ull;
        s3_wait_counter(a.hs + HS_FROM_HI, s3_ld_acquire_sys(a.hs + HS_DONE), a.hs);
        if (tr) a.trace[8 + 4 * (a.trace[0] & (S3_TRACE_CAP - 1)) + 2] = s3_globaltimer() - t0;
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(mb_u + NSLOT * 8 + 8), "r"(1u) : "memory");
    };

    // f planes in tensor memory: warp 0 allocates FT_COLS columns for the CTA; every thread then owns 2*VX*FTD columns
    // of its lane: lane 32 * (warp % 4) + laneid (implied by the warp), columns (warp / 4) * 64 onwards
    constexpr bool FT = C::FT;
    uint32_t tmem_base = 0, tf0 = 0;
    if (FT && !skip) {
        const uint32_t tslot = mb_u + NSLOT * 8;
        if (tid < 32) { tmem_alloc(tslot, (uint32_t)C::FT_COLS); tmem_relinquish(); }
        tmem_fence_before_sync();
        __syncthreads();
        tmem_fence_after_sync();
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tslot));
        tf0 = tmem_base + ((uint32_t)((tid >> 5) & 3) << 21) + (uint32_t)((tid >> 7) * 2 * VX * C::FTD);
    }

    // One chunk = tile (x0, y0) streamed through the owned planes [z0, z1). A CTA may process
    // several chunks (balanced persistent partition, see the end of the kernel).
    unsigned int mall = 0xffffffffu;   // S3_FAST: sticky minimum of the numerator keys over the whole launch
    auto run_chunk = [&](const int x0, const int y0, const int z0, const int z1) {
    const int TZ = z1 - z0;
    const int zb = z0 - H;          // plane index of input step 0
    const int nin = TZ + 2 * H;     // input planes
    const int T = TZ + 3 * H - 1;   // steps

    // ---- per-thread geometry (constant over the launch)
    const int gx0 = x0 - C::HX + VX * ux;          // global x of the first owned point
    const int gy0 = y0 - C::HY + 2 * uy;           // global y of the first owned row
    const int off0 = (2 * uy) * C::WX + VX * ux;   // element offset of row 0 of the unit inside a slot
    const uint32_t a0 = sbase + (uint32_t)off0 * (uint32_t)sizeof(R);   // row 0 of the unit in slot 0; row 1 = + ROWB
    const uint32_t aU = uy == 0 ? a0 : a0 - ROWB;                       // row above (clamped: garbage zone)
    const uint32_t aD = uy == C::UY - 1 ? a0 + ROWB : a0 + 2 * ROWB;    // row below
    const uint32_t aL = ux == 0 ? a0 : a0 - (uint32_t)sizeof(R);        // left neighbour of the first point
    const uint32_t aR = a0 + (uint32_t)((ux == C::UX - 1 ? VX - 1 : VX) * sizeof(R));   // right neighbour of the last
    const bool xin = gx0 >= 0 && gx0 < L;                          // L % VX == 0: whole group in or out
    const bool yin0 = gy0 >= 0 && gy0 < L, yin1 = gy0 + 1 >= 0 && gy0 + 1 < L;
    const bool in0 = worker && xin && yin0, in1 = worker && xin && yin1;
    // interior of the output tile (what this CTA is responsible for writing)
    const bool xint = ux >= C::HX / VX && ux < (C::HX + TX) / VX;
    const bool st0 = in0 && xint && (2 * uy >= C::HY) && (2 * uy < C::HY + TY);
    const bool st1 = in1 && xint && (2 * uy + 1 >= C::HY) && (2 * uy + 1 < C::HY + TY);
    const size_t sL = (size_t)L, sLL = sL * sL;
    R *const dst0 = a.dst + ((size_t)gx0 + sL * (size_t)gy0);      // dereferenced only when in-domain
    int stm = (st0 ? 1 : 0) | (st1 ? 2 : 0);                       // which rows of the unit this thread writes
    asm volatile("" : "+r"(stm));   // opaque: kept in a register instead of being re-derived from tid every step
    const bool has_peer = a.peer_lo != nullptr || a.peer_hi != nullptr;
    // running output pointer of the last Jacobi stage: it emits plane zb + t - 2*(S-1) - 1 at step t
    R *dcur = reinterpret_cast<R *>(reinterpret_cast<intptr_t>(dst0) +
                                    (intptr_t)sizeof(R) * (intptr_t)sLL * (intptr_t)(zb - 2 * (S - 1) - 1));
    // RES: running offset of the unit's coarse row inside Rout (first store: fine plane z0 + 1)
    size_t rcur = (size_t)(gx0 >> 1) + (size_t)(L >> 1) * ((size_t)(gy0 >> 1) +
                  (size_t)(L >> 1) * (size_t)(((z0 - a.nz_lo) >> 1) + a.rz_off));
    // the whole (tile + halo) footprint lies inside the grid in x and y: no in-plane masking
    const bool cta_inner = (x0 - C::HX >= 0) && (x0 + TX + C::HX <= L) && (y0 - C::HY >= 0) && (y0 + TY + C::HY <= L);

    __syncthreads();  // every thread is done with the previous chunk's shared memory
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < NSLOT; ++k) {
            if (!first_chunk) mbar_inval(mb_u + 8 * k);   // all of its transfers were awaited
            mbar_init(mb_u + 8 * k, 1);
        }
        mbar_fence_init();
    }
    first_chunk = false;
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < NSLOT - 1; ++k)
            if (k < nin) {      // nin >= 4: planes 0 and 1 always exist
                hi_wait(zb + k);
                mbar_expect_tx(mb_u + 8 * k, k == 1 && !FT ? 2 * C::PLANE_BYTES : C::PLANE_BYTES);
                tma_load_3d(sbase + k * SB, &src_map, x0 - C::HX, y0 - C::HY, zb + k, mb_u + 8 * k);
            }
        // f plane j (global z = zb + j) is first needed at step j + 1 (stage 1) and last at step
        // j + 2*NST - 1 (stage NST); it is fetched at step j - 1 together with source plane j + 1, when
        // the slot's previous tenant j - NF retired (step j - 2), and it is counted on that source
        // plane's mbarrier (awaited at step j + 1). Plane 0 is never used but is fetched here, with
        // source plane 1, so that every slot sees its planes in order.
        if (!FT) tma_load_3d(sbase + F0, &f_map, x0 - C::HX, y0 - C::HY, zb, mb_u + 8);
    }
    // (FT) f plane j is fetched with plain 128-bit loads at the top of step j, parked in tensor memory at its end,
    // and read back by stage s at step j + 2s - 1
    const R *fcur = reinterpret_cast<const R *>(reinterpret_cast<intptr_t>(a.fsrc) + (intptr_t)sizeof(R) *
                    ((intptr_t)gx0 + (intptr_t)sL * (intptr_t)gy0 + (intptr_t)sLL * (intptr_t)zb));
    R fnew[NP];
    auto f_fetch = [&](int t, bool masked) {
        const int p = zb + t;
        const bool pin = p >= zdom0 && p < zdom1;     // f outside the grid is never used (those planes are masked to 0)
        VT v0 = Vec<R>::zero(), v1 = Vec<R>::zero();
        if (pin && (!masked || in0)) v0 = __ldg(reinterpret_cast<const VT *>(fcur));
        if (pin && (!masked || in1)) v1 = __ldg(reinterpret_cast<const VT *>(fcur + sL));
        Vec<R>::unpack(v0, fnew); Vec<R>::unpack(v1, fnew + VX);
    };

    // PRO: stage 1 reads  src + prolong(V)  (expandResidual + addTo, cpu-raw.lua:65-73,83-85): the coarse values
    // under the unit's two rows and the rows above and below are fetched at the START of the step and added, in
    // registers, to what stage 1 loads from the source slot at the END of the step (stage 1 runs last). The four
    // fine rows gy0-1 .. gy0+2 lie over three coarse rows when gy0 is even and over two when it is odd (gy0 has the
    // parity of the y halo). Outside the grid the source is +0 (TMA zero fill) and V reads as +0: 0 + 0 = +0.
    constexpr bool YODD = (C::HY & 1) != 0;
    constexpr int NVR = YODD ? 2 : 3;                        // coarse rows fetched
    constexpr int VS_UP = 0, VS_C0 = YODD ? 0 : 1, VS_C1 = 1, VS_DN = YODD ? 1 : 2;   // fine row -> fetched coarse row
    R vq[NVR][VX / 2 > 0 ? VX / 2 : 1];
    R vxl[2] = {(R)0, (R)0}, vxr[2] = {(R)0, (R)0};   // tiles that patch lanes 0 / 31 from shared memory: the coarse
                                                      // values left / right of the unit, for its two rows
    // offsets of the coarse rows inside a coarse plane and whether they exist (constant over the chunk)
    int voff[NVR];
    bool vin[NVR];
#pragma unroll
    for (int r = 0; r < NVR; ++r) {
        const int gy = gy0 + (r == 0 ? -1 : (YODD ? 1 : (r == 1 ? 0 : 2)));
        vin[r] = PRO && worker && xin && gy >= 0 && gy < L;
        voff[r] = vin[r] ? (gx0 >> 1) + (L >> 1) * (gy >> 1) : 0;
    }
    auto v_load = [&](int t) {
        if (!PRO) return;
        const int p = zb + t;
        const bool pin = p >= zdom0 && p < zdom1;
        const int pc = pin ? ((p - zdom0) >> 1) + a.vz_off : 0;    // coarse plane (local index)
        const R *vp = a.Vp + (size_t)(L >> 1) * (size_t)(L >> 1) * (size_t)pc;
#pragma unroll
        for (int r = 0; r < NVR; ++r) {
            const bool in = pin && vin[r];
            if (VX == 4 && sizeof(R) == 4) {       // the two coarse values of a row are one aligned 8-byte load
                float2 v = make_float2(0.f, 0.f);
                if (in) v = __ldg(reinterpret_cast<const float2 *>(vp + voff[r]));
                vq[r][0] = (R)v.x; vq[r][VX / 2 - 1] = (R)v.y;
            } else {
#pragma unroll
                for (int i = 0; i < VX / 2; ++i) vq[r][i] = in ? __ldg(vp + voff[r] + i) : (R)0;
            }
        }
        if (!C::EDGE_FREE) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int k = r == 0 ? VS_C0 : VS_C1;
                vxl[r] = (pin && vin[k] && lane == 0 && gx0 - 1 >= 0) ? __ldg(vp + voff[k] - 1) : (R)0;
                vxr[r] = (pin && vin[k] && lane == 31 && gx0 + VX < L) ? __ldg(vp + voff[k] + VX / 2) : (R)0;
            }
        }
    };

    A acc[NST][NP], prev[NST][NP];
#pragma unroll
    for (int s = 0; s < NST; ++s)
#pragma unroll
        for (int i = 0; i < NP; ++i) { acc[s][i] = (A)0; prev[s][i] = (A)0; }
    // carry[s] = the unit's own output of stage s+1 from the previous step: the centre rows of
    // stage s+2 this step. Only the rows ABOVE and BELOW the unit (other threads' values) are read
    // back from the shared-memory ring.
    // (fp32 arithmetic only: with 8-byte accumulators the extra registers would spill.)
    constexpr bool CARRY = MG_STREAM_CARRY && sizeof(A) == 4;
    constexpr bool DEFER = CARRY && MG_STREAM_DEFER;
    constexpr int EARLY = CARRY ? MG_STREAM_EARLY : 0;   // 0: none, 1: neighbour rows + shuffles, 2: also f
    R carry[CARRY && NST > 1 ? NST - 1 : 1][NP];
#pragma unroll
    for (int s = 0; s < (CARRY && NST > 1 ? NST - 1 : 1); ++s)
#pragma unroll
        for (int i = 0; i < NP; ++i) carry[s][i] = (R)0;
    A rpart[VX / 2 > 0 ? VX / 2 : 1];  // RES: restriction partial sums of the even plane
#pragma unroll
    for (int i = 0; i < VX / 2; ++i) rpart[i] = (A)0;

    // ring cursors, advanced once per step (no integer division in the loop); bu / bfq are the
    // same cursors as byte offsets
    int su = 0, pu = 0;        // source slot of step t = t % NSLOT, and its mbarrier parity
    int sf = NF - 1;           // slot of f plane j = t - 1 (stage 1's plane); j = -1 at t = 0 -> becomes 0 at t = 1
    uint32_t bu = 0, bfq = (NF - 1) * SB;

    // One pipeline step. STEADY: every stage is active, emits, and every emitted plane is inside
    // the grid and (for the last Jacobi stage) inside [z0, z1): no per-stage predicates.
    // MASKED: the tile footprint crosses the grid boundary in x or y.
    auto step = [&](auto steady_tag, auto masked_tag, const int t) {
        constexpr bool ST = decltype(steady_tag)::value, MK = decltype(masked_tag)::value;
        // (1) refill: the source slot consumed at step t-1 and the f slot retired at step t-1
        if (tid == 0) {
            const int k = t + NSLOT - 1;                      // source plane needed at step t + NSLOT - 1
            if (k < nin) {
                const int ks = su == 0 ? NSLOT - 1 : su - 1;  // (t + NSLOT - 1) % NSLOT
                const int j = t + 1;                          // f plane stage 1 needs at step t + 2
                int ksf = sf + 2; if (ksf >= NF) ksf -= NF;   // (t + 1) % NF  (sf = (t - 1) % NF)
                if constexpr (!ST) hi_wait(zb + k);   // (the steady-state loop is split around that plane instead, see below)
                // The slots being refilled were last READ through the generic proxy before the barrier that ended
                // step t-1 (the values are in registers); nothing writes them through the generic proxy.
                mbar_expect_tx(mb_u + 8 * ks, FT ? C::PLANE_BYTES : 2 * C::PLANE_BYTES);
                tma_load_3d(sbase + ks * SB, &src_map, x0 - C::HX, y0 - C::HY, zb + k, mb_u + 8 * ks);
                if (!FT) tma_load_3d(sbase + F0 + ksf * SB, &f_map, x0 - C::HX, y0 - C::HY, zb + j, mb_u + 8 * ks);
            }
        }
        if (PRO && (ST || t < nin)) v_load(t);
        if (FT) {
            tmem_wait_st();                              // last step's plane is in tensor memory
            if (ST || t < nin) f_fetch(t, MK);
        }
        // (2) source plane of this step
        if (ST || t < nin) mbar_wait(mb_u + 8 * su, (uint32_t)pu);
        // (3) the pipeline stages. The stages of one step are independent of each other (each reads
        // what was written a step earlier), so they may run in any order, except that stage s+1 must
        // take its centre rows from the carry registers before stage s refills them.
        struct In { R c0[VX], c1[VX], up[VX], dn[VX], l0, l1, r0, r1, fv[NP]; };
        auto stage_p = [&](int sidx) { return zb + t - 2 * sidx - 1; };      // plane emitted (q - 1)
        auto stage_active = [&](int sidx) { return ST ? true : ((t >= 3 * sidx) && (t <= nin + sidx - 1)); };
        auto stage_emit = [&](int sidx) { return ST ? true : (t >= 3 * sidx + 2); };
        // everything stage sidx+1 reads this step: centre rows (source slot or carry), the rows above and
        // below from the plane stage sidx wrote into the ring a step ago, x-neighbours by shuffle, and f
        auto load_f = [&](int sidx, bool emit, In &q) {
            // f of the emitted plane, from the f ring (TMA zero fill covers everything outside the grid)
            if (FT) {
                if (emit) tmem_ld8(tf0 + (uint32_t)(((t - 1 - 2 * sidx) & (C::FTD - 1)) * 2 * VX), q.fv);
                else {
#pragma unroll
                    for (int i = 0; i < NP; ++i) q.fv[i] = (R)0;
                }
            } else if (emit) {
                uint32_t fb = bfq - (uint32_t)(2 * sidx) * SB;     // slot (t - 1 - 2*sidx) % NF
                if ((int)fb < 0) fb += NF * SB;
                Vec<R>::lds(a0 + F0 + fb, q.fv);
                Vec<R>::lds(a0 + F0 + fb + ROWB, q.fv + VX);
            } else {
#pragma unroll
                for (int i = 0; i < NP; ++i) q.fv[i] = (R)0;
            }
        };
        auto load_nb = [&](int sidx, In &q) {
            // input plane: this step's source slot, or the ring slot stage s-1 wrote during step t-1
            const uint32_t ib = sidx == 0 ? bu : RING0 + (uint32_t)(2 * (sidx - 1) + ((t - 1) & 1)) * SB;
            if (sidx == 0 || !CARRY) {
                Vec<R>::lds(a0 + ib, q.c0);
                Vec<R>::lds(a0 + ib + ROWB, q.c1);
            } else {
#pragma unroll
                for (int i = 0; i < VX; ++i) { q.c0[i] = carry[sidx > 0 ? sidx - 1 : 0][i]; q.c1[i] = carry[sidx > 0 ? sidx - 1 : 0][VX + i]; }
            }
            Vec<R>::lds(aU + ib, q.up);
            Vec<R>::lds(aD + ib, q.dn);
            if (PRO && sidx == 0 && kPackedF32 && std::is_same<A, float>::value && std::is_same<R, float>::value) {
#pragma unroll
                for (int k = 0; k < VX / 2; ++k) {   // u + prolong(V) as packed adds (both points of a pair share the coarse value)
                    const float2 vu = make_float2((float)vq[VS_UP][k], (float)vq[VS_UP][k]), v0 = make_float2((float)vq[VS_C0][k], (float)vq[VS_C0][k]);
                    const float2 v1 = make_float2((float)vq[VS_C1][k], (float)vq[VS_C1][k]), vd = make_float2((float)vq[VS_DN][k], (float)vq[VS_DN][k]);
                    const float2 tu = __fadd2_rn(make_float2((float)q.up[2 * k], (float)q.up[2 * k + 1]), vu);
                    const float2 t0 = __fadd2_rn(make_float2((float)q.c0[2 * k], (float)q.c0[2 * k + 1]), v0);
                    const float2 t1 = __fadd2_rn(make_float2((float)q.c1[2 * k], (float)q.c1[2 * k + 1]), v1);
                    const float2 td = __fadd2_rn(make_float2((float)q.dn[2 * k], (float)q.dn[2 * k + 1]), vd);
                    q.up[2 * k] = (R)tu.x; q.up[2 * k + 1] = (R)tu.y; q.c0[2 * k] = (R)t0.x; q.c0[2 * k + 1] = (R)t0.y;
                    q.c1[2 * k] = (R)t1.x; q.c1[2 * k + 1] = (R)t1.y; q.dn[2 * k] = (R)td.x; q.dn[2 * k + 1] = (R)td.y;
                }
            } else if (PRO && sidx == 0) {   // u + prolong(V), rounded to storage like addTo (cpu-raw.lua:83-85)
#pragma unroll
                for (int i = 0; i < VX; ++i) {
                    q.up[i] = (R)Ar<A>::add((A)q.up[i], (A)vq[VS_UP][i >> 1]);
                    q.c0[i] = (R)Ar<A>::add((A)q.c0[i], (A)vq[VS_C0][i >> 1]);
                    q.c1[i] = (R)Ar<A>::add((A)q.c1[i], (A)vq[VS_C1][i >> 1]);
                    q.dn[i] = (R)Ar<A>::add((A)q.dn[i], (A)vq[VS_DN][i >> 1]);
                }
            }
            // x-neighbours across units come from the adjacent lane (warp shuffle) instead of a
            // 4-byte shared load at 16-byte stride (4-way bank conflict). Where the adjacent lane is a
            // different row (ux = 0 or UX-1) the value is garbage, exactly in the garbage zone of the
            // tile edge. Lanes 0 / 31 have no such lane: unless they sit on a tile edge themselves
            // (EDGE_FREE) they patch the value with a one-lane predicated load, no branch.
            if (MG_STREAM_SHFL) {
                q.l0 = __shfl_up_sync(0xffffffffu, q.c0[VX - 1], 1); q.l1 = __shfl_up_sync(0xffffffffu, q.c1[VX - 1], 1);
                q.r0 = __shfl_down_sync(0xffffffffu, q.c0[0], 1); q.r1 = __shfl_down_sync(0xffffffffu, q.c1[0], 1);
                if (!C::EDGE_FREE) {
                    if (PRO && sidx == 0) {   // the patched values come from the raw source slot: they need their + prolong(V) too
                        const R pl0 = (R)Ar<A>::add((A)lds_if(lane == 0, aL + ib, (R)0), (A)vxl[0]), pl1 = (R)Ar<A>::add((A)lds_if(lane == 0, aL + ib + ROWB, (R)0), (A)vxl[1]);
                        const R pr0 = (R)Ar<A>::add((A)lds_if(lane == 31, aR + ib, (R)0), (A)vxr[0]), pr1 = (R)Ar<A>::add((A)lds_if(lane == 31, aR + ib + ROWB, (R)0), (A)vxr[1]);
                        q.l0 = lane == 0 ? pl0 : q.l0; q.l1 = lane == 0 ? pl1 : q.l1;
                        q.r0 = lane == 31 ? pr0 : q.r0; q.r1 = lane == 31 ? pr1 : q.r1;
                    } else {
                        q.l0 = lds_if(lane == 0, aL + ib, q.l0); q.l1 = lds_if(lane == 0, aL + ib + ROWB, q.l1);
                        q.r0 = lds_if(lane == 31, aR + ib, q.r0); q.r1 = lds_if(lane == 31, aR + ib + ROWB, q.r1);
                    }
                }
            } else {
                q.l0 = lds1(aL + ib, (R)0); q.l1 = lds1(aL + ib + ROWB, (R)0);
                q.r0 = lds1(aR + ib, (R)0); q.r1 = lds1(aR + ib + ROWB, (R)0);
            }
        };
        auto load_inputs = [&](int sidx, bool emit, In &q) { load_f(sidx, emit, q); load_nb(sidx, q); };
        // what a Jacobi stage does with its new plane o[]: masks, carry + ring for the next stage, or (last
        // sweep) the global store, plus the neighbours' ghost planes
        auto emit_jacobi = [&](int sidx, const A *o) {
            const int s = sidx + 1, p = stage_p(sidx);
            const bool pin = ST ? true : (p >= zdom0 && p < zdom1);
            R outv[NP];
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                const bool keep = MK ? (((i < VX) ? in0 : in1) && pin) : pin;
                outv[i] = (ST && !MK) ? (R)o[i] : (keep ? (R)o[i] : (R)0);
            }
            if (s < NST) {  // feed the next stage: own rows in registers, for the neighbours in the ring
                if (CARRY) {
#pragma unroll
                    for (int i = 0; i < NP; ++i) carry[sidx < NST - 1 ? sidx : 0][i] = outv[i];
                }
                if (worker && !DEFER) {
                    const uint32_t ob = RING0 + (uint32_t)(2 * sidx + (t & 1)) * SB;
                    Vec<R>::sts(a0 + ob, outv);
                    Vec<R>::sts(a0 + ob + ROWB, outv + VX);
                }
            }
            if (s == S && (ST || (p >= z0 && p < z1))) {
                R *d = dcur;                                   // = dst0 + sLL * p (running pointer)
                if (stm & 1) *(VT *)d = Vec<R>::pack(outv);
                if (stm & 2) *(VT *)(d + sL) = Vec<R>::pack(outv + VX);
                // boundary planes also land in the neighbours' ghost planes (peer stores)
                if (has_peer) {
                    const int nown = a.nz_hi - a.nz_lo;
                    if (a.peer_lo != nullptr && p - a.nz_lo < a.ghost) {       // -> lower rank's upper ghost
                        R *pd = a.peer_lo + (d - a.dst) + sLL * (size_t)nown;
                        if (stm & 1) *(VT *)pd = Vec<R>::pack(outv);
                        if (stm & 2) *(VT *)(pd + sL) = Vec<R>::pack(outv + VX);
                    }
                    if (a.peer_hi != nullptr && a.nz_hi - p <= a.ghost) {      // -> upper rank's lower ghost
                        R *pd = a.peer_hi + (d - a.dst) - sLL * (size_t)nown;
                        if (stm & 1) *(VT *)pd = Vec<R>::pack(outv);
                        if (stm & 2) *(VT *)(pd + sL) = Vec<R>::pack(outv + VX);
                    }
                }
            }
        };
        // the residual stage's plane o[]: rounded to storage like the reference's rs[L], then restricted;
        // children in the order i fastest, then j, then k (SURVEY 8(a'))
        auto emit_res = [&](const A *o) {
            const int p = stage_p(NST - 1);
            const bool pin = ST ? true : (p >= zdom0 && p < zdom1);
            R outv[NP];
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                const bool keep = MK ? (((i < VX) ? in0 : in1) && pin) : pin;
                outv[i] = (ST && !MK) ? (R)o[i] : (keep ? (R)o[i] : (R)0);
            }
            if (!((ST || (p >= z0 && p < z1)) && stm == 3)) return;
            const int L2 = L >> 1;
            if (((p - a.nz_lo) & 1) == 0) {
#pragma unroll
                for (int cidx = 0; cidx < VX / 2; ++cidx) {
                    A sacc = Ar<A>::add((A)outv[2 * cidx], (A)outv[2 * cidx + 1]);
                    sacc = Ar<A>::add(sacc, (A)outv[VX + 2 * cidx]);
                    rpart[cidx] = Ar<A>::add(sacc, (A)outv[VX + 2 * cidx + 1]);
                }
            } else {
                // coarse row of this unit in the coarse plane ((p - nz_lo) >> 1) + rz_off: a running
                // offset, one coarse plane further after every store (chunks start on plane pairs)
                const size_t cb = rcur;
                rcur += (size_t)L2 * (size_t)L2;
#pragma unroll
                for (int cidx = 0; cidx < VX / 2; ++cidx) {
                    A sacc = Ar<A>::add(rpart[cidx], (A)outv[2 * cidx]);
                    sacc = Ar<A>::add(sacc, (A)outv[2 * cidx + 1]);
                    sacc = Ar<A>::add(sacc, (A)outv[VX + 2 * cidx]);
                    sacc = Ar<A>::add(sacc, (A)outv[VX + 2 * cidx + 1]);
                    const R rv = (R)Ar<A>::mul((A).125, sacc);
                    a.Rout[cb + cidx] = rv;
                    const int qc = (p - a.nz_lo) >> 1, nc = (a.nz_hi - a.nz_lo) >> 1;   // coarse owned index / count
                    if (a.rpeer_lo != nullptr && qc < a.ghost) a.rpeer_lo[cb + cidx + (size_t)L2 * L2 * (size_t)nc] = rv;
                    if (a.rpeer_hi != nullptr && qc >= nc - a.ghost) a.rpeer_hi[cb + cidx - (size_t)L2 * L2 * (size_t)nc] = rv;
                    for (int r = 0; r < a.rall_n; ++r)          // replicated coarse level: same offset in every rank's cube
                        if (a.rall[r] != nullptr) a.rall[r][cb + cidx] = rv;
                }
            }
        };

        if constexpr (kPackedF32 && std::is_same<A, float>::value && std::is_same<R, float>::value) {
            // Blackwell packed fp32, one stage after the other (LAST STAGE FIRST), one guard per stage
            auto P = [](float a, float b) { return make_float2(a, b); };
            const float2 INV = P(cf.inv_h2, cf.inv_h2), NINV = P(-cf.inv_h2, -cf.inv_h2), AD = P(cf.adiag, cf.adiag);
            const float2 NAD = P(cf.nadiag, cf.nadiag), Y = P(cf.yneg, cf.yneg), M1 = P(-1.f, -1.f);
            In qbuf[2];   // EARLY: double buffered stage inputs
            auto load_early = [&](int sidx, In &q) {
                if (EARLY >= 2) load_f(sidx, stage_emit(sidx), q);
                load_nb(sidx, q);
            };
            if (EARLY && stage_active(NST - 1)) load_early(NST - 1, qbuf[0]);
#pragma unroll
            for (int srev = 0; srev < NST; ++srev) {
                const int sidx = NST - 1 - srev;
                if (!stage_active(sidx)) {
                    if (EARLY && sidx > 0 && stage_active(sidx - 1)) load_early(sidx - 1, qbuf[(srev + 1) & 1]);
                    continue;
                }
                const bool emit = stage_emit(sidx);
                const bool is_res = RES && sidx == NST - 1;
                In &q = qbuf[srev & 1];
                if (EARLY < 2) load_f(sidx, emit, q);
                if (!EARLY) load_nb(sidx, q);
                // xl + xr pairs operands one element apart, which would cost register moves to pair up:
                // these sums stay scalar and land directly in aligned pairs
                constexpr int NPK = NP / 2, HPK = VX / 2;     // pairs per unit, pairs per row
                float2 sxx[NPK], Cc[NPK], YL[NPK], YR[NPK];
#pragma unroll
                for (int k = 0; k < HPK; ++k) {
                    const float a0l = k == 0 ? q.l0 : q.c0[2 * k - 1], a0r = k == HPK - 1 ? q.r0 : q.c0[2 * k + 2];
                    const float a1l = k == 0 ? q.l1 : q.c1[2 * k - 1], a1r = k == HPK - 1 ? q.r1 : q.c1[2 * k + 2];
                    sxx[k] = P(__fadd_rn(a0l, q.c0[2 * k + 1]), __fadd_rn(q.c0[2 * k], a0r));
                    sxx[HPK + k] = P(__fadd_rn(a1l, q.c1[2 * k + 1]), __fadd_rn(q.c1[2 * k], a1r));
                    Cc[k] = P(q.c0[2 * k], q.c0[2 * k + 1]);
                    Cc[HPK + k] = P(q.c1[2 * k], q.c1[2 * k + 1]);
                }
#pragma unroll
                for (int k = 0; k < HPK; ++k) {
                    YL[k] = P(q.up[2 * k], q.up[2 * k + 1]); YL[HPK + k] = Cc[k];
                    YR[k] = Cc[HPK + k]; YR[HPK + k] = P(q.dn[2 * k], q.dn[2 * k + 1]);
                }
                float2 T[NPK], F[NPK], AU[NPK];
                float o[NP];
#pragma unroll
                for (int k = 0; k < NPK; ++k) {
                    const float2 part = __fadd2_rn(__fadd2_rn(sxx[k], YL[k]), YR[k]);
                    const float2 PRV = P(prev[sidx][2 * k], prev[sidx][2 * k + 1]);
                    T[k] = __fadd2_rn(P(acc[sidx][2 * k], acc[sidx][2 * k + 1]), Cc[k]);   // pending plane gets its z+1
                    if (is_res) AU[k] = __fadd2_rn(__fmul2_rn(T[k], INV), __fmul2_rn(AD, PRV));
                    const float2 NA = __fadd2_rn(part, PRV);                               // this plane gets its z-1
                    acc[sidx][2 * k] = NA.x; acc[sidx][2 * k + 1] = NA.y;
                    prev[sidx][2 * k] = Cc[k].x; prev[sidx][2 * k + 1] = Cc[k].y;
                }
                if (FT && emit) tmem_wait_ld8(q.fv);   // f of the emitted plane (tensor memory) has arrived
#pragma unroll
                for (int k = 0; k < NPK; ++k) {
                    F[k] = P(q.fv[2 * k], q.fv[2 * k + 1]);
                    if (is_res) {
                        const float2 rv = __ffma2_rn(AU[k], M1, F[k]);                     // f - au
                        o[2 * k] = rv.x; o[2 * k + 1] = rv.y;
                    }
                }
                // the next stage's inputs do not depend on anything this step produces: request them now, so that
                // their shared-memory / shuffle latency runs under this stage's division
                if (EARLY && sidx > 0 && stage_active(sidx - 1)) load_early(sidx - 1, qbuf[(srev + 1) & 1]);
                if (!emit) continue;
                if (is_res) { emit_res(o); continue; }
                float2 N[NPK];
#pragma unroll
                for (int k = 0; k < NPK; ++k) N[k] = __ffma2_rn(T[k], NINV, F[k]);         // RN(f - S/h^2)
                // Division guard (mg_math.cuh): the Markstein sequence is exact unless a numerator is tiny
                // but non-zero; one integer min-chain per group (the key ranks exact zeros highest), IEEE
                // division for the group otherwise. (A floating-point pre-test of min |n| with FMNMX3 was
                // measured 10 % slower per V-cycle.)
                unsigned int m = 0xffffffffu;
#ifndef MG_STREAM_NOGUARD
#pragma unroll
                for (int k = 0; k < NPK; ++k) m = min(m, min(Ar<float>::guard_key(N[k].x), Ar<float>::guard_key(N[k].y)));
#endif
                if (!GUARDED) { mall = min(mall, m); m = 0xffffffffu; }
                if (m >= Ar<float>::guard_threshold()) {
#pragma unroll
                    for (int k = 0; k < NPK; ++k) {
                        const float2 q1 = __fmul2_rn(N[k], Y);
                        const float2 rr = __ffma2_rn(NAD, q1, N[k]);
                        const float2 q2 = __ffma2_rn(rr, Y, q1);
                        o[2 * k] = q2.x; o[2 * k + 1] = q2.y;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < NPK; ++k) {
                        o[2 * k] = Ar<float>::div(N[k].x, cf.adiag);
                        o[2 * k + 1] = Ar<float>::div(N[k].y, cf.adiag);
                    }
                }
                emit_jacobi(sidx, o);
            }
        } else {
            // generic arithmetic (8-byte accumulators): one stage after the other, LAST STAGE FIRST
#pragma unroll
            for (int srev = 0; srev < NST; ++srev) {
                const int sidx = NST - 1 - srev;
                if (!stage_active(sidx)) continue;
                const bool emit = stage_emit(sidx);
                const bool is_res = RES && sidx == NST - 1;
                In q;
                load_inputs(sidx, emit, q);
                if (FT && emit) tmem_wait_ld8(q.fv);
                A tot[NP], o[NP];
#pragma unroll
                for (int i = 0; i < VX; ++i) {
                    {   // row 0 of the unit
                        const A xl = (A)(i == 0 ? q.l0 : q.c0[i - 1]), xr = (A)(i == VX - 1 ? q.r0 : q.c0[i + 1]);
                        const A yl = (A)q.up[i], yr = (A)q.c1[i], c = (A)q.c0[i];
                        const A part = Ar<A>::add(Ar<A>::add(Ar<A>::add(xl, xr), yl), yr);
                        tot[i] = Ar<A>::add(acc[sidx][i], c);                      // pending plane gets its z+1
                        if (is_res) o[i] = residual_point<A>(tot[i], (A)q.fv[i], prev[sidx][i], cf);
                        acc[sidx][i] = Ar<A>::add(part, prev[sidx][i]);            // this plane gets its z-1
                        prev[sidx][i] = c;
                    }
                    {   // row 1 of the unit
                        const int j = VX + i;
                        const A xl = (A)(i == 0 ? q.l1 : q.c1[i - 1]), xr = (A)(i == VX - 1 ? q.r1 : q.c1[i + 1]);
                        const A yl = (A)q.c0[i], yr = (A)q.dn[i], c = (A)q.c1[i];
                        const A part = Ar<A>::add(Ar<A>::add(Ar<A>::add(xl, xr), yl), yr);
                        tot[j] = Ar<A>::add(acc[sidx][j], c);
                        if (is_res) o[j] = residual_point<A>(tot[j], (A)q.fv[j], prev[sidx][j], cf);
                        acc[sidx][j] = Ar<A>::add(part, prev[sidx][j]);
                        prev[sidx][j] = c;
                    }
                }
                if (!emit) continue;
                if (is_res) {
                    emit_res(o);
                } else {
                    A num[NP];
#pragma unroll
                    for (int i = 0; i < NP; ++i) num[i] = jacobi_num<A>(tot[i], (A)q.fv[i], cf);
                    div_adiag_group<3, A, NP>(num, o, cf);
                    emit_jacobi(sidx, o);
                }
            }
        }

        if (FT && (ST || t < nin)) tmem_st8(tf0 + (uint32_t)((t & (C::FTD - 1)) * 2 * VX), fnew);
        if (DEFER && worker) {
#pragma unroll
            for (int sidx = 0; sidx < NST - 1; ++sidx)
                if (stage_active(sidx) && stage_emit(sidx)) {
                    const uint32_t ob = RING0 + (uint32_t)(2 * sidx + (t & 1)) * SB;
                    Vec<R>::sts(a0 + ob, carry[sidx]);
                    Vec<R>::sts(a0 + ob + ROWB, carry[sidx] + VX);
                }
        }
        __syncthreads();
        // advance the ring cursors
        bu += SB; bfq += SB;
        dcur += sLL; fcur += sLL;
        if (++su == NSLOT) { su = 0; bu = 0; pu ^= 1; }
        if (++sf == NF) { sf = 0; bfq = 0; }
    };

    // Steady range: all stages active and emitting, every emitted plane inside the grid. The
    // three phases (fill, steady, drain) are separate loops so that the per-stage registers
    // (acc, prev) keep one assignment per loop instead of being shuffled at a merge point
    // every step; fill and drain share one copy of the generic body.
    const int t_lo = 3 * NST - 1;
    int t_hi = min(nin - 1, zdom1 - zb);   // stage 1 emits plane zb + t - 1 <= zdom1 - 1
    if ((a.flags & 1) || t_hi < t_lo) t_hi = t_lo - 1;            // no steady phase
    if (MODE == S3_RERUN) {   // the (rare) guarded re-run: one compact copy of the generic body
#pragma unroll 1
        for (int t = 0; t < T; ++t) step(std::false_type{}, std::true_type{}, t);
        return;
    }
#pragma unroll 1
    for (int phase = 0; phase < 3; ++phase) {
        if (phase == 1) {
            // Multi-GPU: the step that requests the first source plane at or above nz_hi - G (plane zb + t + NSLOT - 1) must
            // come after the wait for the upper neighbour (hi_wait). The steady-state body itself stays free of it (a spin
            // loop inside the body cost 19 % of a pass, measured): the loop runs in two parts around that step.
            int t_mid = t_hi + 1;
            if (MG_SLAB_DEFER_HI && a.hs_hi != nullptr) t_mid = max(t_lo, min(t_hi + 1, a.nz_hi - a.ghost - zb - (NSLOT - 1)));
#pragma unroll 1
            for (int part = 0; part < 2; ++part) {
                const int ta = part == 0 ? t_lo : t_mid, tb = part == 0 ? t_mid - 1 : t_hi;
                if (part == 1 && tid == 0 && ta <= tb) hi_wait(a.nz_hi);
                if (cta_inner && !(a.flags & 2)) {
#pragma unroll kSteadyUnroll
                    for (int t = ta; t <= tb; ++t) step(std::true_type{}, std::false_type{}, t);
                } else {
#pragma unroll kSteadyUnroll
                    for (int t = ta; t <= tb; ++t) step(std::true_type{}, std::true_type{}, t);
                }
            }
        } else {
            const int ta = phase == 0 ? 0 : t_hi + 1, tb = phase == 0 ? min(t_lo, T) : T;
#pragma unroll 1
            for (int t = ta; t < tb; ++t) step(std::false_type{}, std::true_type{}, t);
        }
    }
    };  // run_chunk

    // Persistent partition: the launch is ntiles x (owned planes / 2) plane pairs of work for the gridDim.x CTAs (one
    // per SM). A share of the balanced part may end one tile column and begin the next; each piece is a chunk with its
    // own halo.
    // Lock-step columns first (Stream3DArgs::ncol), then this CTA's share of the remainder.
    {
        const int ntx = (L + TX - 1) / TX, nty = (L + TY - 1) / TY, ntiles = ntx * nty;
        const int nb = (int)gridDim.x, b = (int)blockIdx.x;
        int col = b;                                   // next whole column of this CTA
        long long lo = 0, hi = 0, npair = 1;           // its share of the remainder, in plane pairs
        const int zrem = a.nz_lo + a.rem_z0;
        if (b >= a.rem_cta0) {
            npair = (a.nz_hi - zrem) >> 1;
            const long long W2 = (long long)(ntiles - a.rem_tile0) * npair;
            lo = W2 * (b - a.rem_cta0) / (nb - a.rem_cta0);
            hi = W2 * (b - a.rem_cta0 + 1) / (nb - a.rem_cta0);
        }
        while (true) {
            int tile, z0, z1;
            if (col < a.ncol) {
                tile = col; z0 = a.nz_lo; z1 = a.nz_lo + a.zcol;
                col += nb;
            } else if (lo < hi) {
                const long long tl = lo / npair, zp = lo - tl * npair;
                const long long n = min(npair - zp, hi - lo);
                tile = a.rem_tile0 + (int)tl; z0 = zrem + 2 * (int)zp; z1 = zrem + 2 * (int)(zp + n);
                lo += n;
            } else {
                break;
            }
            if (!skip) run_chunk((tile % ntx) * TX, (tile / ntx) * TY, z0, z1);
        }
    }

    if (FT && !skip) {   // every tcgen05.ld was awaited by its consumer
        tmem_fence_before_sync();
        __syncthreads();
        if (tid < 32) tmem_dealloc(tmem_base, (uint32_t)C::FT_COLS);
    }
    if (MODE == S3_FAST) {   // flags bit 2 (debug): always ask for the guarded re-run
        const bool bad = mall < Ar<float>::guard_threshold() || (a.flags & 4);
        if (__syncthreads_or(bad ? 1 : 0) && threadIdx.x == 0) atomicOr(a.redo, 1u);
    } else if (a.hs != nullptr) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            const unsigned int done = atomicAdd(reinterpret_cast<unsigned int *>(a.hs + HS_CTAS), 1u);
            if (done == gridDim.x - 1) {
                *reinterpret_cast<volatile unsigned int *>(a.hs + HS_CTAS) = 0u;
                if (a.trace != nullptr) {
                    const unsigned long long i = a.trace[0];
                    a.trace[8 + 4 * (i & (S3_TRACE_CAP - 1)) + 3] = s3_globaltimer();
                    a.trace[0] = i + 1;
                }
                const unsigned long long n = s3_ld_acquire_sys(a.hs + HS_DONE) + 1;
                s3_st_release_sys(a.hs + HS_DONE, n);
                if (a.hs_lo != nullptr) s3_st_release_sys(a.hs_lo + HS_FROM_HI, n);   // we are its upper neighbour
                if (a.hs_hi != nullptr) s3_st_release_sys(a.hs_hi + HS_FROM_LO, n);   // and its lower neighbour
            }
        }
    }
}

}  // namespace mg

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
// NST stages; at step t stage s consumes plane a_t - 2(s-1) of stage s-1's output from
// shared memory and emits its own plane one below it. A thread owns VX x 2 columns for the
// whole launch and keeps, per stage, two registers per column:
//   prev = centre value of the previous plane            (the z-1 neighbour)
//   acc  = ((xl+xr)+yl)+yr + zl of the pending plane      (waiting for its z+1 neighbour)
// so each plane of each stage is read from shared memory exactly once (4 x LDS.128 +
// 4 x LDS.32 per 8 points) and ONE __syncthreads() per step serves all stages.
// The summation order ((((xl+xr)+yl)+yr)+zl)+zr is preserved, so results are bit-identical
// to the one-sweep-per-launch kernels (all arithmetic from mg_math.cuh).
//
// Cells outside the grid must stay exactly 0 at every stage: in-plane via a per-thread
// bit mask, whole planes via a CTA-uniform test. Values near the tile edge that lack a
// neighbour are garbage by construction and never reach the H-deep interior (H = NST).
//
// Also in this kernel (details at the code):
//  * f planes travel through their own TMA ring (2*NST+1 slots); zero fill makes them predicate free.
//  * fill / steady / drain are separate loops; the steady body has no per-stage predicates and,
//    on tiles inside the grid, no masks.
//  * shared-memory rows are 96 floats (fp32 tile 88 x 24): 128-bit accesses are bank-conflict
//    free; x-neighbours across threads come from warp shuffles.
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
#ifndef MG_PACKED_F32
#define MG_PACKED_F32 1       // fp32 stage arithmetic with Blackwell's packed FADD2/FFMA2/FMUL2
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
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_inval(uint64_t *bar)
{
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// orders this CTA's earlier generic-proxy accesses to shared memory (made visible to the
// calling thread by a preceding barrier) before later async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, int x, int y, int z,
                                            uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(smem_dst)),
        "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
}

// predicated shared-memory load: returns *p if pred, else old (no branch, no access when !pred)
__device__ __forceinline__ float lds_if(bool pred, const float *p, float old)
{
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\n@q ld.shared.f32 %0, [%1];\n}\n"
                 : "+f"(old) : "r"(smem_u32(p)), "r"((unsigned)pred) : "memory");
    return old;
}
__device__ __forceinline__ double lds_if(bool pred, const double *p, double old)
{
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\n@q ld.shared.f64 %0, [%1];\n}\n"
                 : "+d"(old) : "r"(smem_u32(p)), "r"((unsigned)pred) : "memory");
    return old;
}

// ------------------------------------------------------------------ vector access
template <typename R> struct Vec;
template <> struct Vec<float> {
    static constexpr int N = 4;
    typedef float4 T;
    static __device__ __forceinline__ void unpack(const T &v, float *o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
    static __device__ __forceinline__ T pack(const float *o) { return make_float4(o[0], o[1], o[2], o[3]); }
    static __device__ __forceinline__ T zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
};
template <> struct Vec<double> {
    static constexpr int N = 2;
    typedef double2 T;
    static __device__ __forceinline__ void unpack(const T &v, double *o) { o[0] = v.x; o[1] = v.y; }
    static __device__ __forceinline__ T pack(const double *o) { return make_double2(o[0], o[1]); }
    static __device__ __forceinline__ T zero() { return make_double2(0., 0.); }
};

template <typename R, int S, bool RES, int TX, int TY> struct Stream3DCfg {
    static constexpr int VX = Vec<R>::N;
    static constexpr int NST = S + (RES ? 1 : 0);
    static constexpr int H = NST;
    static constexpr int HX = (H + VX - 1) / VX * VX;
    static constexpr int HY = RES ? (H + 1) / 2 * 2 : H;
    static constexpr int WX = TX + 2 * HX, WY = TY + 2 * HY;
    static constexpr int UX = WX / VX, UY = WY / 2;
    static constexpr int NT = UX * UY;
    static constexpr int NTHREADS = (NT + 31) / 32 * 32;
    static constexpr int PLANE = WX * WY;
    static constexpr int PLANE_BYTES = PLANE * (int)sizeof(R);
    static constexpr int SLOT_BYTES = (PLANE_BYTES + 127) / 128 * 128;
    static constexpr int SLOT_ELEMS = SLOT_BYTES / (int)sizeof(R);
    static constexpr int NSLOT = 3;         // source-plane ring (TMA prefetch distance NSLOT-1 steps)
    static constexpr int NRING = NST - 1;   // intermediate stage outputs, double buffered
    static constexpr int NF = 2 * NST + 1;  // f-plane ring: a plane lives 2*NST-1 steps, fetched 2 ahead
    static constexpr int NSLOTS_TOTAL = NSLOT + 2 * NRING + NF;
    static constexpr int SMEM_BYTES = NSLOTS_TOTAL * SLOT_BYTES + (NSLOT + NF) * 8;
    static_assert(TX % VX == 0 && TY % 2 == 0 && WY % 2 == 0, "tile shape");
    static_assert(NTHREADS <= 1024, "too many threads");
    static_assert(SMEM_BYTES <= 232448, "shared memory budget (227 KB)");
};

template <typename R> struct Stream3DArgs {
    R *dst;           // u after S sweeps
    const R *Vp;      // PRO: coarse correction Vs[L/2]
    R *Rout;          // RES: Rs[L/2]
    int L;            // level width
    int flags;        // debug: bit 0 = never take the steady-state body, bit 1 = always mask
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
};

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

template <typename R, typename A, int S, bool PRO, bool RES, int TX, int TY>
__global__ void __launch_bounds__((Stream3DCfg<R, S, RES, TX, TY>::NTHREADS), 1)
k_stream3d(const __grid_constant__ CUtensorMap src_map, const __grid_constant__ CUtensorMap f_map,
           Stream3DArgs<R> a, Coef<A> cf)
{
    typedef Stream3DCfg<R, S, RES, TX, TY> C;
    typedef typename Vec<R>::T VT;
    constexpr int VX = C::VX, NST = C::NST, H = C::H, NP = 2 * VX, NSLOT = C::NSLOT, NF = C::NF;

    // layout: [NSLOT source slots][NRING*2 stage slots][NF f slots][mbarriers]. The dynamic
    // shared window starts at offset 0 of the CTA's shared memory (no static __shared__ in
    // this kernel), so every slot is 128-byte aligned as TMA requires. Pointers are derived
    // from the __shared__ array itself so that accesses compile to LDS/STS.
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    R *const sm = reinterpret_cast<R *>(smem_raw);
    auto in_slot = [&](int k) -> R * { return sm + k * C::SLOT_ELEMS; };
    auto ring_slot = [&](int s, int par) -> R * { return sm + (NSLOT + 2 * s + par) * C::SLOT_ELEMS; };
    auto f_slot = [&](int k) -> R * { return sm + (NSLOT + 2 * C::NRING + k) * C::SLOT_ELEMS; };
    uint64_t *const mbar_u = reinterpret_cast<uint64_t *>(smem_raw + (size_t)C::NSLOTS_TOTAL * C::SLOT_BYTES);
    uint64_t *const mbar_f = mbar_u + NSLOT;

    const int tid = threadIdx.x, lane = tid & 31;
    const bool worker = tid < C::NT;   // the last warp's spare threads mirror unit 0 and never store
    const int ux = worker ? tid % C::UX : 0, uy = worker ? tid / C::UX : 0;
    const int L = a.L;
    const int zdom0 = a.zdom0, zdom1 = a.zdom1;
    bool first_chunk = true;

    // Before touching ghost planes (ours, by TMA; the neighbours', by peer stores): both
    // neighbours must have finished as many passes as we have. Every CTA checks for itself.
    if (a.hs != nullptr) {
        if (threadIdx.x == 0) {
            const unsigned long long n = s3_ld_acquire_sys(a.hs);
            if (a.hs_lo != nullptr) while (s3_ld_acquire_sys(a.hs + 16) < n) { }
            if (a.hs_hi != nullptr) while (s3_ld_acquire_sys(a.hs + 32) < n) { }
        }
        __syncthreads();
    }

    // One chunk = tile (x0, y0) streamed through the owned planes [z0, z1). A CTA may process
    // several chunks (balanced persistent partition, see the end of the kernel).
    auto run_chunk = [&](const int x0, const int y0, const int z0, const int z1) {
    const int TZ = z1 - z0;
    const int zb = z0 - H;          // plane index of input step 0
    const int nin = TZ + 2 * H;     // input planes
    const int T = TZ + 3 * H - 1;   // steps

    // ---- per-thread geometry (constant over the launch)
    const int gx0 = x0 - C::HX + VX * ux;          // global x of the first owned point
    const int gy0 = y0 - C::HY + 2 * uy;           // global y of the first owned row
    const int off0 = (2 * uy) * C::WX + VX * ux;   // smem offset of row 0 of the unit
    const int off1 = off0 + C::WX;
    const int offU = uy == 0 ? off0 : off0 - C::WX;               // row above (clamped: garbage zone)
    const int offD = uy == C::UY - 1 ? off1 : off1 + C::WX;       // row below
    const int dl = ux == 0 ? 0 : -1;                               // left neighbour of the first point
    const int dr = ux == C::UX - 1 ? VX - 1 : VX;                  // right neighbour of the last point
    const bool xin = gx0 >= 0 && gx0 < L;                          // L % VX == 0: whole group in or out
    const bool yin0 = gy0 >= 0 && gy0 < L, yin1 = gy0 + 1 >= 0 && gy0 + 1 < L;
    const bool in0 = worker && xin && yin0, in1 = worker && xin && yin1;
    // interior of the output tile (what this CTA is responsible for writing)
    const bool xint = ux >= C::HX / VX && ux < (C::HX + TX) / VX;
    const bool st0 = in0 && xint && (2 * uy >= C::HY) && (2 * uy < C::HY + TY);
    const bool st1 = in1 && xint && (2 * uy + 1 >= C::HY) && (2 * uy + 1 < C::HY + TY);
    const size_t sL = (size_t)L, sLL = sL * sL;
    R *const dst0 = a.dst + ((size_t)gx0 + sL * (size_t)gy0);      // dereferenced only when in-domain
    // the whole (tile + halo) footprint lies inside the grid in x and y: no in-plane masking
    const bool cta_inner = (x0 - C::HX >= 0) && (x0 + TX + C::HX <= L) && (y0 - C::HY >= 0) && (y0 + TY + C::HY <= L);

    __syncthreads();  // every thread is done with the previous chunk's shared memory
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < NSLOT + NF; ++k) {
            if (!first_chunk) mbar_inval(&mbar_u[k]);   // all of its transfers were awaited
            mbar_init(&mbar_u[k], 1);                   // mbar_f follows mbar_u in memory
        }
        mbar_fence_init();
    }
    first_chunk = false;
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < NSLOT - 1; ++k)
            if (k < nin) {
                mbar_expect_tx(&mbar_u[k], C::PLANE_BYTES);
                tma_load_3d(in_slot(k), &src_map, x0 - C::HX, y0 - C::HY, zb + k, &mbar_u[k]);
            }
        // f plane j (global z = zb + j) is first needed at step j + 1 (stage 1) and last at step
        // j + 2*NST - 1 (stage NST); it is fetched at step j - 1, when the slot's previous tenant
        // j - NF retired (step j - 2). Plane 0 is never used but is fetched here so that every
        // slot sees its planes in order (uniform mbarrier phases).
        mbar_expect_tx(&mbar_f[0], C::PLANE_BYTES);
        tma_load_3d(f_slot(0), &f_map, x0 - C::HX, y0 - C::HY, zb, &mbar_f[0]);
    }

    // PRO: add prolong(V) to the own points of an arrived source slot, in place. The coarse
    // values are fetched at the START of the step (fix_load) and applied at its end (fix_apply),
    // so their global-memory latency hides behind the step's stage work.
    R vpre[2][VX / 2 > 0 ? VX / 2 : 1];
    bool vpre_on = false;
    auto fix_load = [&](int t) {
        vpre_on = false;
        if (!PRO) return;
        const int p = zb + t;
        if (p < zdom0 || p >= zdom1) return;                       // plane outside the grid stays 0
        vpre_on = true;
        const int pc = ((p - zdom0) >> 1) + a.vz_off;              // coarse plane (local index)
        const int L2 = L >> 1;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const bool in = r == 0 ? in0 : in1;
            const size_t crow = (size_t)(gx0 >> 1) + (size_t)L2 * ((size_t)((gy0 + r) >> 1) + (size_t)L2 * (size_t)pc);
#pragma unroll
            for (int i = 0; i < VX / 2; ++i) vpre[r][i] = in ? a.Vp[crow + i] : (R)0;
        }
    };
    auto fix_apply = [&](int slot) {
        if (!PRO || !vpre_on) return;
        R *sl = in_slot(slot);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const bool in = r == 0 ? in0 : in1;
            if (!in) continue;
            R u[VX];
            Vec<R>::unpack(*(const VT *)(sl + (r == 0 ? off0 : off1)), u);
#pragma unroll
            for (int i = 0; i < VX; ++i) u[i] = (R)Ar<A>::add((A)u[i], (A)vpre[r][i >> 1]);
            *(VT *)(sl + (r == 0 ? off0 : off1)) = Vec<R>::pack(u);
        }
    };

    A acc[NST][NP], prev[NST][NP];
#pragma unroll
    for (int s = 0; s < NST; ++s)
#pragma unroll
        for (int i = 0; i < NP; ++i) { acc[s][i] = (A)0; prev[s][i] = (A)0; }
    A rpart[VX / 2 > 0 ? VX / 2 : 1];  // RES: restriction partial sums of the even plane
#pragma unroll
    for (int i = 0; i < VX / 2; ++i) rpart[i] = (A)0;

    if (PRO) {
        fix_load(0);
        mbar_wait(&mbar_u[0], 0);
        fix_apply(0);
        __syncthreads();
    }

    // ring cursors, advanced once per step (no integer division in the loop)
    int su = 0, pu = 0;        // source slot of step t = t % NSLOT, and its mbarrier parity
    int sf = NF - 1, pf = 1;   // slot/parity of f plane j = t - 1 (stage 1's plane); j = -1 at t = 0
    // (sf, pf) track j = t - 1: j = -1 -> slot NF-1 of "phase -1" (parity 1); becomes (0, 0) at t = 1

    // One pipeline step. STEADY: every stage is active, emits, and every emitted plane is inside
    // the grid and (for the last Jacobi stage) inside [z0, z1): no per-stage predicates.
    // MASKED: the tile footprint crosses the grid boundary in x or y.
    auto step = [&](auto steady_tag, auto masked_tag, const int t) {
        constexpr bool ST = decltype(steady_tag)::value, MK = decltype(masked_tag)::value;
        // (1) refill: the source slot consumed at step t-1 and the f slot retired at step t-1
        if (tid == 0) {
            const int k = t + NSLOT - 1;
            if (k < nin) {
                const int ks = su == 0 ? NSLOT - 1 : su - 1;  // (t + NSLOT - 1) % NSLOT
                fence_proxy_async_smem();
                mbar_expect_tx(&mbar_u[ks], C::PLANE_BYTES);
                tma_load_3d(in_slot(ks), &src_map, x0 - C::HX, y0 - C::HY, zb + k, &mbar_u[ks]);
            }
            const int j = t + 1;                              // f plane stage 1 needs at step t + 2
            if (j <= nin - 2) {
                int ksf = sf + 2; if (ksf >= NF) ksf -= NF;   // (t + 1) % NF  (sf = (t - 1) % NF)
                mbar_expect_tx(&mbar_f[ksf], C::PLANE_BYTES);
                tma_load_3d(f_slot(ksf), &f_map, x0 - C::HX, y0 - C::HY, zb + j, &mbar_f[ksf]);
            }
        }
        if (PRO && t + 1 < nin) fix_load(t + 1);
        // (2) source plane of this step (PRO: it was awaited and fixed up during step t-1)
        if (!PRO) {
            if (ST || t < nin) mbar_wait(&mbar_u[su], (uint32_t)pu);
        }
        // (3) the pipeline stages; stage s = sidx + 1
#pragma unroll
        for (int sidx = 0; sidx < NST; ++sidx) {
            const int s = sidx + 1;
            if (!ST) {
                const bool active = (t >= 3 * sidx) && (t <= nin + s - 2);
                if (!active) continue;
            }
            const bool emit = ST ? true : (t >= 3 * s - 1);
            const int p = zb + t - 2 * sidx - 1;                   // plane emitted (q - 1)
            const R *in = sidx == 0 ? in_slot(su) : ring_slot(sidx - 1, (t - 1) & 1);
            const bool is_res = RES && s == NST;
            const bool last_jacobi = s == S;
            const bool pin = ST ? true : (p >= zdom0 && p < zdom1);

            // f of the emitted plane, from the f ring (TMA zero fill covers everything outside the grid)
            R fv[NP];
            if (emit) {
                int kf = sf - 2 * sidx; if (kf < 0) kf += NF;      // (t - 1 - 2*sidx) % NF
                if (sidx == 0) mbar_wait(&mbar_f[kf], (uint32_t)pf);
                const R *fs = f_slot(kf);
                Vec<R>::unpack(*(const VT *)(fs + off0), fv);
                Vec<R>::unpack(*(const VT *)(fs + off1), fv + VX);
            } else {
#pragma unroll
                for (int i = 0; i < NP; ++i) fv[i] = (R)0;
            }

            R c0[VX], c1[VX], up[VX], dn[VX];
            Vec<R>::unpack(*(const VT *)(in + off0), c0);
            Vec<R>::unpack(*(const VT *)(in + off1), c1);
            Vec<R>::unpack(*(const VT *)(in + offU), up);
            Vec<R>::unpack(*(const VT *)(in + offD), dn);
            // x-neighbours across units come from the adjacent lane (warp shuffle) instead of a
            // 4-byte shared load at 16-byte stride (4-way bank conflict). Lanes 0 / 31 have no such
            // lane and read shared memory; where the adjacent lane is a different row (ux = 0 or
            // UX-1) the value is garbage, exactly in the garbage zone of the tile edge.
            R l0, l1, r0, r1;
            if (MG_STREAM_SHFL) {
                l0 = __shfl_up_sync(0xffffffffu, c0[VX - 1], 1); l1 = __shfl_up_sync(0xffffffffu, c1[VX - 1], 1);
                r0 = __shfl_down_sync(0xffffffffu, c0[0], 1); r1 = __shfl_down_sync(0xffffffffu, c1[0], 1);
                // lanes 0 / 31 have no such lane: one-lane predicated loads, no branch
                l0 = lds_if(lane == 0, in + off0 + dl, l0); l1 = lds_if(lane == 0, in + off1 + dl, l1);
                r0 = lds_if(lane == 31, in + off0 + dr, r0); r1 = lds_if(lane == 31, in + off1 + dr, r1);
            } else {
                l0 = in[off0 + dl]; l1 = in[off1 + dl]; r0 = in[off0 + dr]; r1 = in[off1 + dr];
            }

            A tot[NP], o[NP];
            if constexpr (kPackedF32 && std::is_same<A, float>::value && std::is_same<R, float>::value) {
                // Blackwell packed fp32 (FADD2 / FFMA2 / FMUL2): two IEEE-rn operations per issue
                // slot, bit-identical to the scalar form. Points are paired along x inside a vector.
                auto P = [](float a, float b) { return make_float2(a, b); };
                // xl + xr pairs operands one element apart, which would cost register moves to pair up:
                // these four sums per row stay scalar and land directly in aligned pairs
                float2 sxx[4];
                sxx[0] = P(__fadd_rn(l0, c0[1]), __fadd_rn(c0[0], c0[2])); sxx[1] = P(__fadd_rn(c0[1], c0[3]), __fadd_rn(c0[2], r0));
                sxx[2] = P(__fadd_rn(l1, c1[1]), __fadd_rn(c1[0], c1[2])); sxx[3] = P(__fadd_rn(c1[1], c1[3]), __fadd_rn(c1[2], r1));
                const float2 C[4] = {P(c0[0], c0[1]), P(c0[2], c0[3]), P(c1[0], c1[1]), P(c1[2], c1[3])};
                const float2 YL[4] = {P(up[0], up[1]), P(up[2], up[3]), C[0], C[1]};
                const float2 YR[4] = {C[2], C[3], P(dn[0], dn[1]), P(dn[2], dn[3])};
                const float2 INV = P(cf.inv_h2, cf.inv_h2), NINV = P(-cf.inv_h2, -cf.inv_h2), AD = P(cf.adiag, cf.adiag);
                const float2 NAD = P(cf.nadiag, cf.nadiag), Y = P(cf.yneg, cf.yneg), M1 = P(-1.f, -1.f);
                float2 T[4], F[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 part = __fadd2_rn(__fadd2_rn(sxx[k], YL[k]), YR[k]);
                    const float2 PRV = P(prev[sidx][2 * k], prev[sidx][2 * k + 1]);
                    T[k] = __fadd2_rn(P(acc[sidx][2 * k], acc[sidx][2 * k + 1]), C[k]);   // pending plane gets its z+1
                    F[k] = P(fv[2 * k], fv[2 * k + 1]);
                    if (is_res) {
                        const float2 au = __fadd2_rn(__fmul2_rn(T[k], INV), __fmul2_rn(AD, PRV));
                        const float2 rv = __ffma2_rn(au, M1, F[k]);                        // f - au
                        o[2 * k] = rv.x; o[2 * k + 1] = rv.y;
                    }
                    const float2 NA = __fadd2_rn(part, PRV);                               // this plane gets its z-1
                    acc[sidx][2 * k] = NA.x; acc[sidx][2 * k + 1] = NA.y;
                    prev[sidx][2 * k] = C[k].x; prev[sidx][2 * k + 1] = C[k].y;
                    tot[2 * k] = T[k].x; tot[2 * k + 1] = T[k].y;
                }
                if (!emit) continue;
                if (!is_res) {
                    float2 N[4];
                    unsigned int m = 0xffffffffu;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        N[k] = __ffma2_rn(T[k], NINV, F[k]);                               // RN(f - S/h^2)
                        const unsigned int ka = Ar<float>::guard_key(N[k].x), kb = Ar<float>::guard_key(N[k].y);
                        m = min(m, min(ka, kb));
                    }
                    if (m >= Ar<float>::guard_threshold()) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float2 q1 = __fmul2_rn(N[k], Y);
                            const float2 rr = __ffma2_rn(NAD, q1, N[k]);
                            const float2 q2 = __ffma2_rn(rr, Y, q1);
                            o[2 * k] = q2.x; o[2 * k + 1] = q2.y;
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            o[2 * k] = Ar<float>::div(N[k].x, cf.adiag);
                            o[2 * k + 1] = Ar<float>::div(N[k].y, cf.adiag);
                        }
                    }
                }
            } else {
#pragma unroll
            for (int i = 0; i < VX; ++i) {
                {   // row 0 of the unit
                    const A xl = (A)(i == 0 ? l0 : c0[i - 1]), xr = (A)(i == VX - 1 ? r0 : c0[i + 1]);
                    const A yl = (A)up[i], yr = (A)c1[i], c = (A)c0[i];
                    const A part = Ar<A>::add(Ar<A>::add(Ar<A>::add(xl, xr), yl), yr);
                    tot[i] = Ar<A>::add(acc[sidx][i], c);                      // pending plane gets its z+1
                    if (is_res) o[i] = residual_point<A>(tot[i], (A)fv[i], prev[sidx][i], cf);
                    acc[sidx][i] = Ar<A>::add(part, prev[sidx][i]);            // this plane gets its z-1
                    prev[sidx][i] = c;
                }
                {   // row 1 of the unit
                    const int j = VX + i;
                    const A xl = (A)(i == 0 ? l1 : c1[i - 1]), xr = (A)(i == VX - 1 ? r1 : c1[i + 1]);
                    const A yl = (A)c0[i], yr = (A)dn[i], c = (A)c1[i];
                    const A part = Ar<A>::add(Ar<A>::add(Ar<A>::add(xl, xr), yl), yr);
                    tot[j] = Ar<A>::add(acc[sidx][j], c);
                    if (is_res) o[j] = residual_point<A>(tot[j], (A)fv[j], prev[sidx][j], cf);
                    acc[sidx][j] = Ar<A>::add(part, prev[sidx][j]);
                    prev[sidx][j] = c;
                }
            }
            if (!emit) continue;
            if (!is_res) {
                A num[NP];
#pragma unroll
                for (int i = 0; i < NP; ++i) num[i] = jacobi_num<A>(tot[i], (A)fv[i], cf);
                div_adiag_group<3, A, NP>(num, o, cf);
            }
            }
            R outv[NP];
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                const bool keep = MK ? (((i < VX) ? in0 : in1) && pin) : pin;
                outv[i] = (ST && !MK) ? (R)o[i] : (keep ? (R)o[i] : (R)0);
            }

            if (!is_res) {
                if (s < NST && worker) {  // feed the next stage
                    R *out = ring_slot(sidx, t & 1);
                    *(VT *)(out + off0) = Vec<R>::pack(outv);
                    *(VT *)(out + off1) = Vec<R>::pack(outv + VX);
                }
                if (last_jacobi && (ST || (p >= z0 && p < z1))) {
                    R *d = dst0 + sLL * (size_t)p;
                    if (st0) *(VT *)d = Vec<R>::pack(outv);
                    if (st1) *(VT *)(d + sL) = Vec<R>::pack(outv + VX);
                    // boundary planes also land in the neighbours' ghost planes (peer stores)
                    const int nown = a.nz_hi - a.nz_lo;
                    if (a.peer_lo != nullptr && p - a.nz_lo < a.ghost) {       // -> lower rank's upper ghost
                        R *pd = a.peer_lo + (dst0 - a.dst) + sLL * (size_t)(p + nown);
                        if (st0) *(VT *)pd = Vec<R>::pack(outv);
                        if (st1) *(VT *)(pd + sL) = Vec<R>::pack(outv + VX);
                    }
                    if (a.peer_hi != nullptr && a.nz_hi - p <= a.ghost) {      // -> upper rank's lower ghost
                        R *pd = a.peer_hi + (dst0 - a.dst) + sLL * (size_t)(p - nown);
                        if (st0) *(VT *)pd = Vec<R>::pack(outv);
                        if (st1) *(VT *)(pd + sL) = Vec<R>::pack(outv + VX);
                    }
                }
            } else if ((ST || (p >= z0 && p < z1)) && st0 && st1) {
                // restriction: children in the order i fastest, then j, then k (SURVEY 8(a'))
                const int L2 = L >> 1;
                if (((p - a.nz_lo) & 1) == 0) {
#pragma unroll
                    for (int cidx = 0; cidx < VX / 2; ++cidx) {
                        A sacc = Ar<A>::add((A)outv[2 * cidx], (A)outv[2 * cidx + 1]);
                        sacc = Ar<A>::add(sacc, (A)outv[VX + 2 * cidx]);
                        rpart[cidx] = Ar<A>::add(sacc, (A)outv[VX + 2 * cidx + 1]);
                    }
                } else {
                    const size_t cb = (size_t)(gx0 >> 1) +
                                      (size_t)L2 * ((size_t)(gy0 >> 1) + (size_t)L2 * (size_t)(((p - a.nz_lo) >> 1) + a.rz_off));
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
                    }
                }
            }
        }

        // (PRO) prepare next step's source plane in place
        if (PRO && t + 1 < nin) {
            const int sn = su + 1 == NSLOT ? 0 : su + 1;
            mbar_wait(&mbar_u[sn], (uint32_t)(sn == 0 ? pu ^ 1 : pu));
            fix_apply(sn);
        }
        __syncthreads();
        // advance the ring cursors
        if (++su == NSLOT) { su = 0; pu ^= 1; }
        if (++sf == NF) { sf = 0; pf ^= 1; }
    };

    // f plane 0 is fetched only to keep the ring's mbarrier phases uniform; it must still have
    // landed before this CTA may exit (no TMA transfer may outlive its shared memory)
    mbar_wait(&mbar_f[0], 0);

    // Steady range: all stages active and emitting, every emitted plane inside the grid. The
    // three phases (fill, steady, drain) are separate loops so that the per-stage registers
    // (acc, prev) keep one assignment per loop instead of being shuffled at a merge point
    // every step; fill and drain share one copy of the generic body.
    const int t_lo = 3 * NST - 1;
    int t_hi = min(nin - 1, zdom1 - zb);   // stage 1 emits plane zb + t - 1 <= zdom1 - 1
    if ((a.flags & 1) || t_hi < t_lo) t_hi = t_lo - 1;            // no steady phase
#pragma unroll 1
    for (int phase = 0; phase < 3; ++phase) {
        if (phase == 1) {
            if (cta_inner && !(a.flags & 2)) {
#pragma unroll kSteadyUnroll
                for (int t = t_lo; t <= t_hi; ++t) step(std::true_type{}, std::false_type{}, t);
            } else {
#pragma unroll kSteadyUnroll
                for (int t = t_lo; t <= t_hi; ++t) step(std::true_type{}, std::true_type{}, t);
            }
        } else {
            const int ta = phase == 0 ? 0 : t_hi + 1, tb = phase == 0 ? min(t_lo, T) : T;
#pragma unroll 1
            for (int t = ta; t < tb; ++t) step(std::false_type{}, std::true_type{}, t);
        }
    }
    };  // run_chunk

    // Balanced persistent partition: the launch is ntiles x (owned planes / 2) plane pairs of
    // work, dealt out in equal contiguous shares to the gridDim.x CTAs (one per SM). A share
    // may end one tile column and begin the next; each piece is a chunk with its own halo.
    {
        const int ntx = (L + TX - 1) / TX, nty = (L + TY - 1) / TY;
        const long long npair = (a.nz_hi - a.nz_lo) >> 1;
        const long long W2 = (long long)ntx * nty * npair;
        long long lo = W2 * blockIdx.x / gridDim.x;
        const long long hi = W2 * (blockIdx.x + 1) / gridDim.x;
        while (lo < hi) {
            const long long tile = lo / npair, zp = lo - tile * npair;
            const long long n = min(npair - zp, hi - lo);
            run_chunk((int)(tile % ntx) * TX, (int)(tile / ntx) * TY, a.nz_lo + 2 * (int)zp, a.nz_lo + 2 * (int)(zp + n));
            lo += n;
        }
    }

    // The last CTA of the pass publishes "pass n+1 done" to both neighbours: all our stores,
    // local and into their ghost planes, are ordered before it.
    if (a.hs != nullptr) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            const unsigned int done = atomicAdd(reinterpret_cast<unsigned int *>(a.hs + 48), 1u);
            if (done == gridDim.x - 1) {
                *reinterpret_cast<volatile unsigned int *>(a.hs + 48) = 0u;
                const unsigned long long n = s3_ld_acquire_sys(a.hs) + 1;
                s3_st_release_sys(a.hs, n);
                if (a.hs_lo != nullptr) s3_st_release_sys(a.hs_lo + 32, n);   // we are its upper neighbour
                if (a.hs_hi != nullptr) s3_st_release_sys(a.hs_hi + 16, n);   // and its lower neighbour
            }
        }
    }
}

}  // namespace mg

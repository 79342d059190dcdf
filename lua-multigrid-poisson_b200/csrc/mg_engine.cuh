// mg_engine.cuh -- context (grid-hierarchy arena, streams, graph) and the typed engine that
// schedules kernels for one (storage type, arithmetic type, dimension) combination.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/mgpoisson.h"
#include "mg_block3d.cuh"
#include "mg_fused_simple.cuh"
#include "mg_krylov.cuh"
#include "mg_math.cuh"
#include "mg_ops_ref.cuh"
#include "mg_slab.cuh"
#include "mg_small.cuh"
#include "mg_stream3d.cuh"
#include "mg_warp2d.cuh"

// In-plane tile of the streaming smoother for 4-byte reals (see launch_stream3d_t).
#ifndef MG_TILE_X
#define MG_TILE_X 56
#endif
#ifndef MG_WARP2D_MIN_TY
#define MG_WARP2D_MIN_TY 2    // fewest rows per work item of the 2-D smoother: the mid-size levels are latency-bound (rows
                              // streamed per warp), not work-bound. Measured 16 / 4 / 2 rows: 2002 / 2287 / 2319 V-cycles/s at
                              // 4096^2 fp32 and 2285 / 3221 / 3367 at 2048^2 fp64.
#endif
#ifndef MG_TILE_Y
#define MG_TILE_Y 40
#endif

namespace mg {

constexpr int MAX_LEVELS = 24;
constexpr size_t ARENA_ALIGN = 1024;
constexpr size_t ARENA_HEADER = 1024;   // slab handshake words (Stream3DArgs::hs): [0] passes done,
                                        // [16]/[32] passes published by the lower/upper neighbour, [48] CTAs done
constexpr size_t ARENA_REDO_OFF = 512;  // Stream3DArgs::redo: {guarded re-run requested, CTA arrival counter}

struct TraceRec {
    char name;
    int L;
    std::vector<unsigned char> data;
};

// one record per kernel launch of a profiled (ungraphed) V-cycle
struct ProfRec {
    int kind;    // MG_K_* below
    int L;       // level width
    int sweeps;  // Jacobi sweeps performed by the launch
    cudaEvent_t e0, e1;
    float ms;
};
enum { MG_K_SWEEP = 0, MG_K_RESID_RESTRICT = 1, MG_K_SMALL = 2, MG_K_PROLONG_ADD = 3, MG_K_COPY = 4,
       MG_K_SWEEP_PROLONG = 5, MG_K_SWEEP_RESTRICT = 6 };

struct Engine;

}  // namespace mg

#define MG_CK(ctx, call)                                             \
    do {                                                             \
        cudaError_t e_ = (call);                                     \
        if (e_ != cudaSuccess) return (ctx)->fail_cuda(e_, #call);   \
    } while (0)

extern int g_slab_min_planes;   // mg_api.cu

struct mg_ctx {
    int dim = 0, size = 0, real_kind = 0, smooth = 7, device = 0, nlevels = 0, rank = 0, nranks = 1;
    size_t elem = 0, N = 0;
    int mode = MG_MODE_FUSED, tb = 4, small_L = 0, use_graph = 1;
    int tb2 = 7;             // 2-D: sweeps per pass of the warp-streaming smoother (0 = untiled)
    int warp2d_min_L = 128;  // 2-D: smallest level width handled by the warp-streaming smoother
    int ty_override = 0;     // 2-D: rows per warp work item (0 = cost model)
    int stream_min_L = 128;  // smallest level width handled by the streaming (TMA) smoother
    int num_sms = 148;       // SM count of the device (queried at init)
    int block_max_L = 0;     // 3-D: widest level handled by the block-temporal kernel (mg_block3d.cuh); 0 = off, the default.
                             // MEASURED (round 2, 512^3 fp32, V-cycles/s inside the CUDA graph): off 405.1, 32^3 only 402.5,
                             // 32^3 + 64^3 398.8 -- 20 fewer launches per cycle do not pay for the 2-4.5 x redundant halo work.
    int small_smem_opt = 1;  // levels <= small_L by the shared-memory one-CTA kernel (0: the global-memory walker)
    int pdl_opt = 0;         // kernels of a V-cycle launched with programmatic stream serialization (mg_math.cuh, pdl_enter).
                             // MEASURED (round 2, same box): 512^3 fp32 397.1 -> 393.3 V-cycles/s, 2-D 4096^2 fp32 and 2048^2 fp64
                             // 4-5 % slower (with the faster small-level kernel in the same build): the early-scheduled CTAs of
                             // the next kernel buy nothing here (big kernels fill every SM with one CTA; small ones are ~2 us) and the
                             // programmatic graph edges cost more than plain ones. Off; bit-identical either way (tests).
    int colparts_opt = 1;    // deep passes: tile columns cut into k equal z-parts when that beats equal shares for every SM
    int lockstep_opt = 1;    // lock-step partition of the streaming smoother: whole columns below zsplit + helper CTAs above
    int tma_promo = 0;       // L2 promotion of the TMA descriptors: 0 none (least DRAM over-fetch), 1 64 B, 2 128 B, 3 256 B
    int stream_flags = 0;    // debug switches of the streaming smoother (see Stream3DArgs::flags)
    double omega = 1.0;      // relaxation weight of the Jacobi smoother: 1 = the reference (mg_set_omega)
    int fastdiv_opt = 1;     // fp32 streaming smoother: branch-free division kernel + guarded re-run kernel (0: guarded kernel only)
    int fast_min_L = 128;    // ... at levels at least this wide (128^3: 176 -> 157 us per level visit, measured in round 2)
    int tz_override = 0;     // planes per CTA of the streaming smoother (0 = cost model)
    int ncta_override = 0;   // CTAs per launch of the streaming smoother (0 = one per SM); debug
    int cluster_L = 0;       // widest level handled by the one-cluster kernel (0 = off, the default: see init)
    int cluster_ctas = 0;    // CTAs of that cluster (0 = the widest the device schedules: 16 or 8)
    // TMA descriptors of the source fields, keyed by (pointer, level width, box x, box y)
    std::map<std::tuple<const void *, int, int, int, int>, CUtensorMap> tmaps;
    int tensor_map(const void *base, int L, int nplanes, int box_x, int box_y, const CUtensorMap **out);
    int copy_in(int which, int lv, const void *host, size_t bytes);
    int copy_out(int which, int lv, void *host, size_t bytes);

    // ---- grid-hierarchy arena (K-f): one allocation, zero-filled once (cpu-raw.lua:159-171)
    void *arena = nullptr;
    size_t arena_bytes = 0;
    void *f = nullptr, *psi = nullptr, *psiOld = nullptr;
    void *R[mg::MAX_LEVELS] = {}, *V[mg::MAX_LEVELS] = {}, *W[mg::MAX_LEVELS] = {};
    // buffers only the reference sequence / debug dumps need; allocated on first use
    void *debug_arena = nullptr;
    size_t debug_arena_bytes = 0;
    void *errorBuf = nullptr, *tmpU = nullptr;
    void *r[mg::MAX_LEVELS] = {}, *v[mg::MAX_LEVELS] = {};

    // ---- reductions
    double *d_partial = nullptr, *d_scalar = nullptr, *h_scalar = nullptr;
    void *cg_tmp = nullptr;   // scratch of mg_cg (one field + reduction space), allocated on first use
    int npartial = 0;

    // ---- streams / graph
    cudaStream_t own_stream = nullptr, stream = nullptr, cap_stream = nullptr;
    bool borrowed_stream = false, capturing = false;
    cudaGraph_t graph = nullptr, rep_graph = nullptr;       // rep_*: replicated coarse sub-cycle of a slab
    cudaGraphExec_t gexec = nullptr, rep_gexec = nullptr;
    size_t graph_nodes = 0, rep_nodes = 0;
    uint64_t launches = 0;
    // bytes this rank's kernels stored into other GPUs' memory over NVLink (fused halo exchange and all-gather);
    // a graph replay adds what its capture counted
    uint64_t nvl_bytes = 0, cap_nvl_bytes = 0, graph_nvl_bytes = 0;

    bool prof_on = false;
    std::vector<mg::ProfRec> prof;

    std::string err;
    bool trace_on = false;
    std::vector<mg::TraceRec> trace;
    mg::Engine *eng = nullptr;

    // ---- slab decomposition (mg_slab.cuh); single GPU: G = 0, nothing distributed
    int G = 0;                          // ghost planes on each side of a distributed level
    bool dist[mg::MAX_LEVELS] = {};     // level is cut across the ranks
    int nzl[mg::MAX_LEVELS] = {};       // planes owned by each rank at a distributed level
    mg::SlabGroup *group = nullptr;
    bool owns_group = false, f_ghost_dirty = true, u_ghost_dirty = true;
    // fused halo exchange over peer memory: the neighbours' arenas (same layout as ours), mapped
    // with CUDA IPC (other process) or simply their pointers (same process). p2p = use them.
    char *peer_lo = nullptr, *peer_hi = nullptr;
    char *peer[mg::S3_MAX_RANKS] = {};   // every rank's arena as seen from this rank (peer[rank] = arena); null = not mapped
    bool peer_ipc = false, p2p = false, slab_graph_opt = true;
    unsigned long long *slab_trace = nullptr;   // timeline of the slab passes (option "slab_trace"; Stream3DArgs::trace)
    size_t Ntop = 0;                    // elements allocated for a top-level field (incl. ghosts)

    // ---- pipelined host-buffer batches (mg_step_host_batch): the two copy engines run beside the cycle's stream
    struct HostPipe {
        cudaStream_t up = nullptr, down = nullptr;              // host -> device, device -> host
        void *in_f[2] = {}, *in_u[2] = {}, *out_u[2] = {};      // two staging slots for each direction
        cudaEvent_t up_done[2] = {}, in_free[2] = {}, comp_done[2] = {}, out_free[2] = {};
        double *h_sum = nullptr;                                // pinned: sum (psi - psiOld)^2 of every problem of the batch
        int cap = 0;
        size_t nb = 0;
    } *pipe = nullptr;
    int pipe_ensure(size_t nb, int n);
    void pipe_release();

    int planes(int lv) const { return dist[lv] ? nzl[lv] + 2 * G : (dim == 3 ? (1 << lv) : 1); }
    size_t plane_elems(int lv) const { size_t L = (size_t)1 << lv; return L * L; }
    size_t level_elems(int lv) const { return plane_elems(lv) * (size_t)planes(lv); }
    size_t level_bytes(int lv) const { return level_elems(lv) * elem; }
    size_t own_off_elems(int lv) const { return dist[lv] ? (size_t)G * plane_elems(lv) : 0; }
    size_t own_elems(int lv) const { return dist[lv] ? (size_t)nzl[lv] * plane_elems(lv) : level_elems(lv); }
    size_t arena_off(const void *p) const { return (size_t)((const char *)p - (const char *)arena); }

    int fail(int code, const char *msg)
    {
        err = msg;
        return code;
    }
    int fail_cuda(cudaError_t e, const char *what)
    {
        err = std::string(what) + ": " + cudaGetErrorString(e);
        return MG_ECUDA;
    }
    int activate()
    {
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess || cur != device) MG_CK(this, cudaSetDevice(device));
        return MG_OK;
    }
    int sync();          // this handle's stream (MULTI groups: every slab's), then the peer time-out words
    int check_peers();
    void count_launch(int n = 1)
    {
        if (!capturing) launches += (uint64_t)n;
    }
    // bracket one launch with events when profiling (never while capturing a graph)
    void prof_begin(int kind, int L, int sweeps)
    {
        if (!prof_on || capturing) return;
        mg::ProfRec r{kind, L, sweeps, nullptr, nullptr, 0.f};
        cudaEventCreate(&r.e0);
        cudaEventCreate(&r.e1);
        cudaEventRecord(r.e0, stream);
        prof.push_back(r);
    }
    void prof_end()
    {
        if (!prof_on || capturing || prof.empty()) return;
        cudaEventRecord(prof.back().e1, stream);
    }

    int init(int dim_, int size_, int real_kind_, int smooth_, int device_, int rank_, int nranks_);
    void release();
    int ensure_debug_arena();
    void *buffer(int which, int level, size_t *cap);
    void drop_graph();
    int vcycle();
    int step(double *err_out);
    int trace_rec(char name, int L, const void *dev, size_t bytes);
};

namespace mg {

struct Engine {
    virtual ~Engine() {}
    virtual int init_cells(mg_ctx *c) = 0;
    virtual int jacobi(mg_ctx *c, int L, void *dest, const void *u, const void *f, double h) = 0;
    virtual int residual(mg_ctx *c, int L, void *r, const void *f, const void *u, double h) = 0;
    virtual int restrict_(mg_ctx *c, int L2, void *R, const void *r) = 0;
    virtual int prolong(mg_ctx *c, int L2, void *v, const void *V) = 0;
    virtual int add_to(mg_ctx *c, size_t n, void *u, const void *v) = 0;
    virtual int smooth(mg_ctx *c, int lv, void *u, const void *f, double h, int n) = 0;
    virtual int pre_fused(mg_ctx *c, int lv, void *u, const void *f, double h, int n, void *R) = 0;
    virtual int post_fused(mg_ctx *c, int lv, void *u, const void *f, double h, int n, const void *V) = 0;
    virtual int twogrid_refseq(mg_ctx *c, double h, void *u, const void *f, int lv) = 0;
    virtual int twogrid_fused(mg_ctx *c, double h, void *u, const void *f, int lv) = 0;
    virtual int frob_err(mg_ctx *c, double *err, bool materialise) = 0;
    virtual int residual_norm(mg_ctx *c, double *rms) = 0;
    virtual int slab_vcycle(SlabGroup *g) = 0;
    virtual int cg(mg_ctx *c, int max_iter, double epsilon, double *err_hist, double *linf_hist, int *n_done) = 0;
    virtual int linf_norm(mg_ctx *c, const void *field, size_t n, double *out) = 0;
    virtual int frob_partial_sum(mg_ctx *c, double *sum) = 0;
};

static inline dim3 grid_for(int dim, int L, dim3 b)
{
    return dim3((unsigned)((L + b.x - 1) / b.x), (unsigned)((L + b.y - 1) / b.y), dim == 3 ? (unsigned)L : 1u);
}
static inline dim3 block_for(int L)
{
    if (L >= 128) return dim3(128, 2, 1);
    if (L >= 32) return dim3(32, 8, 1);
    return dim3(8, 8, 1);
}

#define MG_LAUNCH_CHECK(c)                                   \
    do {                                                     \
        cudaError_t e_ = cudaGetLastError();                 \
        if (e_ != cudaSuccess) return (c)->fail_cuda(e_, "kernel launch"); \
        (c)->count_launch();                                 \
    } while (0)

// Launch of a V-cycle kernel (every one of them starts with pdl_enter()): with the "pdl" option the launch carries
// cudaLaunchAttributeProgrammaticStreamSerialization, which also survives stream capture as a programmatic graph edge.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k2(mg_ctx *c, bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(mg_ctx *c, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args &&...args)
{
    return launch_k2(c, c->pdl_opt == 1, kern, grid, block, smem, std::forward<Args>(args)...);
}

template <typename R, typename A, int DIM> struct EngineT : Engine {
    // ------------------------------------------------------------ reference operators
    int init_cells(mg_ctx *c) override
    {
        const int top = c->nlevels - 1;
        dim3 b = block_for(c->size), g = grid_for(DIM, c->size, b);
        int plane0 = 0, k0 = 0;
        if (c->dist[top]) {  // owned planes only: ghosts outside the grid must stay +0
            g.z = (unsigned)c->nzl[top];
            plane0 = c->G;
            k0 = c->rank * c->nzl[top];
        }
        k_init_cells<R, A, DIM><<<g, b, 0, c->stream>>>((R *)c->f, (R *)c->psi, c->size, plane0, k0);
        MG_LAUNCH_CHECK(c);
        c->f_ghost_dirty = c->u_ghost_dirty = true;
        return MG_OK;
    }
    int jacobi(mg_ctx *c, int L, void *dest, const void *u, const void *f, double h) override
    {
        dim3 b = block_for(L), g = grid_for(DIM, L, b);
        k_jacobi<R, A, DIM><<<g, b, 0, c->stream>>>((R *)dest, (const R *)u, (const R *)f, L,
                                                    make_coef<A>(DIM, h, c->omega));
        MG_LAUNCH_CHECK(c);
        return MG_OK;
    }
    int residual(mg_ctx *c, int L, void *r, const void *f, const void *u, double h) override
    {
        dim3 b = block_for(L), g = grid_for(DIM, L, b);
        k_residual<R, A, DIM><<<g, b, 0, c->stream>>>((R *)r, (const R *)f, (const R *)u, L,
                                                      make_coef<A>(DIM, h));
        MG_LAUNCH_CHECK(c);
        return MG_OK;
    }
    int restrict_(mg_ctx *c, int L2, void *Rc, const void *r) override
    {
        dim3 b = block_for(L2), g = grid_for(DIM, L2, b);
        k_restrict<R, A, DIM><<<g, b, 0, c->stream>>>((R *)Rc, (const R *)r, L2);
        MG_LAUNCH_CHECK(c);
        return MG_OK;
    }
    int prolong(mg_ctx *c, int L2, void *v, const void *V) override
    {
        int L = 2 * L2;
        dim3 b = block_for(L), g = grid_for(DIM, L, b);
        k_prolong<R, DIM><<<g, b, 0, c->stream>>>((R *)v, (const R *)V, L);
        MG_LAUNCH_CHECK(c);
        return MG_OK;
    }
    int add_to(mg_ctx *c, size_t n, void *u, const void *v) override
    {
        unsigned nb = (unsigned)((n + 255) / 256);
        k_add_to<R, A><<<nb, 256, 0, c->stream>>>((R *)u, (const R *)v, n);
        MG_LAUNCH_CHECK(c);
        return MG_OK;
    }

    // cpu-raw.lua:176-184 inPlaceIterativeSolver, literally: Jacobi into tmpU, copy back
    int in_place_solver(mg_ctx *c, int lv, R *u, const R *f, double h)
    {
        int rc = jacobi(c, 1 << lv, c->tmpU, u, f, h);
        if (rc) return rc;
        MG_CK(c, cudaMemcpyAsync(u, c->tmpU, c->level_bytes(lv), cudaMemcpyDeviceToDevice, c->stream));
        return MG_OK;
    }

    // cpu-raw.lua:186-237 twoGrid, one launch per reference operator, trace at the show sites
    int twogrid_refseq(mg_ctx *c, double h, void *u_, const void *f_, int lv) override
    {
        R *u = (R *)u_;
        const R *f = (const R *)f_;
        const int L = 1 << lv;
        const size_t nb = c->level_bytes(lv);
        int rc;
        if (L == 1) {
            if ((rc = c->trace_rec('f', L, f, nb))) return rc;
            if ((rc = in_place_solver(c, lv, u, f, h))) return rc;
            return c->trace_rec('u', L, u, nb);
        }
        for (int i = 1; i <= c->smooth; ++i) {
            if (L == c->size && (rc = c->trace_rec('f', L, f, nb))) return rc;
            if ((rc = in_place_solver(c, lv, u, f, h))) return rc;
            if ((rc = c->trace_rec('u', L, u, nb))) return rc;
        }
        R *r = (R *)c->r[lv];
        if ((rc = c->trace_rec('f', L, f, nb))) return rc;
        if ((rc = c->trace_rec('u', L, u, nb))) return rc;
        if ((rc = residual(c, L, r, f, u, h))) return rc;
        if ((rc = c->trace_rec('r', L, r, nb))) return rc;
        const int L2 = L / 2;
        const size_t nb2 = c->level_bytes(lv - 1);
        R *Rc = (R *)c->R[lv - 1];
        if ((rc = restrict_(c, L2, Rc, r))) return rc;
        if ((rc = c->trace_rec('R', L2, Rc, nb2))) return rc;
        R *Vc = (R *)c->V[lv - 1];
        if ((rc = twogrid_refseq(c, 2 * h, Vc, Rc, lv - 1))) return rc;
        if ((rc = c->trace_rec('V', L2, Vc, nb2))) return rc;
        R *v = (R *)c->v[lv];
        if ((rc = prolong(c, L2, v, Vc))) return rc;
        if ((rc = c->trace_rec('v', L, v, nb))) return rc;
        if ((rc = add_to(c, c->level_elems(lv), u, v))) return rc;
        if ((rc = c->trace_rec('u', L, u, nb))) return rc;
        for (int i = 1; i <= c->smooth; ++i) {
            if ((rc = in_place_solver(c, lv, u, f, h))) return rc;
            if ((rc = c->trace_rec('u', L, u, nb))) return rc;
        }
        return MG_OK;
    }

    // ------------------------------------------------------------ fused passes
    // n sweeps starting from `cur` (ping-pong with `oth`), the first optionally reading
    // cur + prolong(Vp), followed optionally by Rout = restrict(f - A cur).
    // ---- streaming (TMA, temporally blocked) smoother passes, 3-D only
    static constexpr int kTileX = sizeof(R) == 4 ? MG_TILE_X : 32;
    static constexpr int kTileY = sizeof(R) == 4 ? MG_TILE_Y : 32;
    // History of the partition (512^3 fp32, V-cycles/s): whole columns only, 88 x 22 tile (144 columns for 148 SMs):
    // 242; balanced shares of tile x plane work (neighbouring columns staggered in z, halo rows re-fetched from HBM):
    // 260 -> 326 after the round-1 tuning; round 2: whole columns below zsplit + helper CTAs above (see below).
    template <int S, bool PRO, bool RES>
    int launch_stream3d(mg_ctx *c, int lv, R *dst, const R *src, const R *f, const R *Vp, R *Rout, const Coef<A> &cf)
    {
        return launch_stream3d_t<S, PRO, RES, kTileY>(c, lv, dst, src, f, Vp, Rout, cf);
    }
    template <int S, bool PRO, bool RES, int TY>
    int launch_stream3d_t(mg_ctx *c, int lv, R *dst, const R *src, const R *f, const R *Vp, R *Rout, const Coef<A> &cf)
    {
        const int L = 1 << lv;
        if (int rc0 = c->activate()) return rc0;   // MULTI groups: every slab has its own device
        // In-plane tile. 4-byte reals: 56 x 40 (+ halo = 64 wide): 16 vectors per row and a 256-byte row
        // pitch, so a warp holds two whole rows of units: every quarter-warp of a 128-bit shared-memory
        // access stays inside one row (bank-conflict free) and lanes 0 / 31 sit on tile edges, where the
        // shuffled x-neighbour needs no patch. Tiles need not divide the grid (masks + partition).
        // (88 x 24, the round-1 tile, measures the same V-cycles/s but needs the edge patches.)
        // 8-byte reals: 32 x 32 (shared-memory budget).
        constexpr int TX = kTileX;
        typedef Stream3DCfg<R, S, RES, TX, TY> C;
        const CUtensorMap *map = nullptr, *fmap = nullptr;
        const int nplanes = c->planes(lv);
        int rc = c->tensor_map(src, L, nplanes, C::WX, C::WY, &map);
        if (rc) return rc;
        if ((rc = c->tensor_map(f, L, nplanes, C::WX, C::WY, &fmap))) return rc;
        // slab view of this level (single GPU / replicated level: the whole cube)
        int nz_lo = 0, nz_hi = L, zdom0 = 0, rz_off = 0, vz_off = 0;
        if (c->dist[lv]) {
            const int zg0 = c->rank * c->nzl[lv];        // global index of the first owned plane
            nz_lo = c->G; nz_hi = c->G + c->nzl[lv]; zdom0 = c->G - zg0;
            rz_off = c->dist[lv - 1] ? c->G : zg0 / 2;   // coarse slab array, or replicated cube
            vz_off = c->dist[lv - 1] ? c->G - zg0 / 2 : 0;
        }
        const int nown = nz_hi - nz_lo;
        // fp32 arithmetic, L >= fast_min_L: the branch-free kernel followed by its (normally empty) guarded re-run
        // kernel (mg_stream3d.cuh, MODE); everything else: the guarded kernel alone.
        constexpr bool HAS_FAST = kPackedF32 && std::is_same<R, float>::value && std::is_same<A, float>::value;
        const bool fast = HAS_FAST && c->fastdiv_opt != 0 && L >= c->fast_min_L;
        typedef void (*Kern)(const CUtensorMap, const CUtensorMap, Stream3DArgs<R>, Coef<A>);
        Kern kern = k_stream3d<R, A, S, PRO, RES, TX, TY, S3_GUARDED>, kern2 = nullptr;
        if constexpr (HAS_FAST) {
            if (fast) {
                kern = k_stream3d<R, A, S, PRO, RES, TX, TY, S3_FAST>;
                kern2 = k_stream3d<R, A, S, PRO, RES, TX, TY, S3_RERUN>;
                MG_CK(c, cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
                MG_CK(c, cudaFuncSetAttribute(kern2, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            }
        }
        MG_CK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        // One CTA per SM, each given an equal share of the tile x plane-pair work (balanced
        // persistent partition inside the kernel); "tz" asks for ~tz planes per share instead.
        const long tiles = (long)((L + TX - 1) / TX) * ((L + TY - 1) / TY);
        const long work = tiles * nown;                    // tile-planes
        int occ = 1;   // resident CTAs per SM (shared memory / registers decide; 1 for S >= 3)
        MG_CK(c, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        MG_CK(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::NTHREADS, C::SMEM_BYTES));
        if (occ < 1) occ = 1;
        long ncta;
        if (c->tz_override > 0) {
            ncta = (work + c->tz_override - 1) / c->tz_override;
        } else if (S <= 2) {
            // Shallow passes are HBM-bound: keep whole tile columns marching through z in lock
            // step, so that the halo rows neighbouring tiles share are still in L2 when the
            // neighbour reads them. Columns are cut into nz equal chunks by a small cost model
            // (waves x planes streamed per CTA).
            long best = -1;
            ncta = tiles;
            for (int cand = nown; cand >= 8; cand >>= 1) {
                const long n = tiles * ((nown + cand - 1) / cand);
                const long cost = ((n + (long)c->num_sms * occ - 1) / ((long)c->num_sms * occ)) * (cand + 3 * C::H);
                if (best < 0 || cost < best) { best = cost; ncta = n; }
            }
        } else {
            // Deep passes: one CTA per SM, no tail wave. A share of `pl` planes costs pl + OV steps per chunk it touches
            // (OV = 3 NST - 1 fill/drain steps of the pipeline). Candidates: equal shares of the whole tile x plane work
            // for every SM (a share that does not line up with the columns touches two of them), or every column cut into
            // k equal parts with k x tiles CTAs -- the parts then line up, march through z together (halo rows shared
            // through L2) and cost nown / k + OV. 256^3: 35 columns x 4 parts = 140 CTAs; the 64- and 32-plane slabs of the
            // 8-GPU run: 130 whole columns, 35 x 4 parts of 8 planes.
            const long nsm = (long)c->num_sms * occ;
            const double OV = 3 * C::H - 1;
            const double pl = (double)work / nsm;
            double best = pl + (std::floor(pl / nown) + 2.0) * OV;
            ncta = nsm;
            // (The parts need not be equal: with k x tiles CTAs the kernel's balanced shares of the plane PAIRS line up with
            // the column ends, and inside a column they differ by at most one pair. 128^3: 12 columns x 12 parts of 5-6 pairs.)
            const long npair = nown / 2;
            for (long k = 1; c->colparts_opt != 0 && k * tiles <= nsm && k <= npair; ++k) {
                if (npair / k < 2) continue;
                const double cst = 2.0 * (double)((npair + k - 1) / k) + OV;
                if (cst < best) { best = cst; ncta = k * tiles; }
            }
        }
        if (c->ncta_override > 0) ncta = c->ncta_override;   // debug: exercises the partitions on small grids
        if (ncta < 1) ncta = 1;
        if (ncta > work / 2) ncta = work / 2 > 0 ? work / 2 : 1;
        // Lock-step partition ("lockstep" option, default on; mg_stream3d.cuh, Stream3DArgs::ncol): whole tile columns
        // march through z together, so the halo rows and partly used sectors neighbouring tiles share are served by L2
        // instead of HBM (512^3: DRAM reads 2.2 -> 1.28 GB per pass, compulsory 1.07).
        //  * fewer tile columns than CTAs (but >= 70 %): every column whole below zcol, the spare CTAs share the planes
        //    above it. zcol balances the two groups: a column costs zs + OV steps, a helper CTA its planes plus OV per
        //    chunk (OV = fill/drain steps of the pipeline).
        //  * more columns than CTAs (1024^2 planes: 494): whole waves of ncta columns, then the last partial wave is
        //    dealt out in equal shares to everybody.
        int ncol = 0, zcol = 0, rem_cta0 = 0, rem_tile0 = 0, rem_z0 = 0;
        if (c->lockstep_opt != 0 && c->tz_override <= 0 && S >= 3 && nown >= 32) {
            const int OV = 3 * C::H - 1;
            if (tiles < ncta && tiles * 10 >= ncta * 7) {
                const long nh = ncta - tiles;
                double bestc = 1e30;
                int zsplit = 0;
                for (int zs = nown; zs >= nown / 2; zs -= 2) {
                    const double planes = (double)tiles * (nown - zs) / nh;
                    const double chunks = zs == nown ? 0.0 : (double)tiles / nh + 1.0;
                    const double cost = std::max((double)(zs + OV), planes + chunks * OV);
                    if (cost < bestc) { bestc = cost; zsplit = zs; }
                }
                if (zsplit > 0 && zsplit < nown) { ncol = (int)tiles; zcol = zsplit; rem_cta0 = (int)tiles; rem_tile0 = 0; rem_z0 = zsplit; }
            } else if (tiles >= ncta) {
                ncol = (int)(tiles / ncta * ncta); zcol = nown; rem_cta0 = 0; rem_tile0 = ncol; rem_z0 = 0;
            }
        }
        dim3 grid((unsigned)ncta, 1, 1);
        Stream3DArgs<R> a{dst, Vp, Rout, L, c->stream_flags, nz_lo, nz_hi, zdom0, zdom0 + L, rz_off, vz_off,
                          nullptr, nullptr, nullptr, nullptr, c->G, nullptr, nullptr, nullptr, ncol, zcol, rem_cta0, rem_tile0, rem_z0,
                          (unsigned int *)((char *)c->arena + ARENA_REDO_OFF), f, {}, 0, nullptr};
        if (c->dist[lv] && c->p2p) {  // fused halo exchange: same offsets inside the neighbours' arenas
            const size_t doff = c->arena_off(dst);
            if (c->peer_lo) a.peer_lo = (R *)(c->peer_lo + doff);
            if (c->peer_hi) a.peer_hi = (R *)(c->peer_hi + doff);
            if (Rout && c->dist[lv - 1]) {
                const size_t roff = c->arena_off(Rout);
                if (c->peer_lo) a.rpeer_lo = (R *)(c->peer_lo + roff);
                if (c->peer_hi) a.rpeer_hi = (R *)(c->peer_hi + roff);
            }
            if (Rout && !c->dist[lv - 1]) {   // first replicated level: all-gather by the producing threads
                const size_t roff = c->arena_off(Rout);
                a.rall_n = c->nranks;
                for (int r = 0; r < c->nranks; ++r)
                    a.rall[r] = (r != c->rank && c->peer[r]) ? (R *)(c->peer[r] + roff) : nullptr;
            }
            {   // NVLink accounting: G boundary planes per neighbour (fine and, with RES, coarse), or the all-gather
                const uint64_t nnb = (c->peer_lo ? 1 : 0) + (c->peer_hi ? 1 : 0);
                uint64_t b = nnb * (uint64_t)std::min(c->G, nown) * c->plane_elems(lv) * c->elem;
                if (Rout && c->dist[lv - 1]) b += nnb * (uint64_t)std::min(c->G, nown / 2) * c->plane_elems(lv - 1) * c->elem;
                if (Rout && !c->dist[lv - 1]) b += (uint64_t)(c->nranks - 1) * (uint64_t)(nown / 2) * c->plane_elems(lv - 1) * c->elem;
                (c->capturing ? c->cap_nvl_bytes : c->nvl_bytes) += b;
            }
            if (c->group && c->group->concurrent()) {  // the slabs run at the same time: handshake inside the kernel
                a.hs = (unsigned long long *)c->arena;
                a.hs_lo = (unsigned long long *)c->peer_lo;
                a.hs_hi = (unsigned long long *)c->peer_hi;
                a.trace = c->slab_trace;
            }
        }
        c->prof_begin(PRO ? MG_K_SWEEP_PROLONG : (RES ? MG_K_SWEEP_RESTRICT : MG_K_SWEEP), L, S);
        MG_CK(c, launch_k(c, kern, grid, dim3(C::NTHREADS), C::SMEM_BYTES, *map, *fmap, a, cf));
        if (kern2) {
            MG_LAUNCH_CHECK(c);
            MG_CK(c, launch_k2(c, c->pdl_opt != 0, kern2, grid, dim3(C::NTHREADS), C::SMEM_BYTES, *map, *fmap, a, cf));   // "pdl" = 2: this edge only
        }
        c->prof_end();
        MG_LAUNCH_CHECK(c);
        return MG_OK;
    }
    int stream3d_pass(mg_ctx *c, int lv, int S, bool pro, bool res, R *dst, const R *src, const R *f,
                      const R *Vp, R *Rout, const Coef<A> &cf)
    {
#define MG_S3D(S_, P_, R_) return launch_stream3d<S_, P_, R_>(c, lv, dst, src, f, Vp, Rout, cf)
        if (!pro && !res) {
            switch (S) { case 1: MG_S3D(1, false, false); case 2: MG_S3D(2, false, false);
                         case 3: MG_S3D(3, false, false); case 4: MG_S3D(4, false, false); }
        } else if (pro && !res) {
            switch (S) { case 1: MG_S3D(1, true, false); case 2: MG_S3D(2, true, false);
                         case 3: MG_S3D(3, true, false); case 4: MG_S3D(4, true, false); }
        } else if (!pro && res) {
            switch (S) { case 1: MG_S3D(1, false, true); case 2: MG_S3D(2, false, true);
                         case 3: MG_S3D(3, false, true); }
        }
#undef MG_S3D
        return c->fail(MG_EINVAL, "stream3d_pass: unsupported combination");
    }
    // split n sweeps into passes of <= tb sweeps; with a fused residual stage the last pass has <= 3.
    // extra_pass: one pass more than necessary (the slab schedule uses it to make the number of
    // ping-pong passes of a level visit even, so that the result lands in u without a copy).
    static std::vector<int> plan_passes(int n, int tb, bool has_res, bool extra_pass = false)
    {
        std::vector<int> plan;
        if (tb < 1) tb = 1;
        int rem = n, last = 0;
        if (has_res) { last = n < 3 ? n : 3; if (last > tb) last = tb; rem = n - last; }
        if (rem > 0) {
            int k = (rem + tb - 1) / tb;
            if (extra_pass && k < rem) ++k;
            const int base = rem / k, extra = rem % k;
            for (int i = 0; i < k; ++i) plan.push_back(base + (i < extra ? 1 : 0));
        }
        if (has_res) plan.push_back(last);
        return plan;
    }
    int sweeps_stream3d(mg_ctx *c, int lv, R *&cur, R *&oth, const R *f, const Coef<A> &cf, int n,
                        const R *Vp, R *Rout)
    {
        const std::vector<int> plan = plan_passes(n, c->tb, Rout != nullptr);
        const int np = (int)plan.size();
        for (int i = 0; i < np; ++i) {
            int rc = stream3d_pass(c, lv, plan[i], i == 0 && Vp, i == np - 1 && Rout, oth, cur, f, Vp, Rout, cf);
            if (rc) return rc;
            R *t = cur; cur = oth; oth = t;
        }
        return MG_OK;
    }

    // ---- slab V-cycle (mg_slab.cuh): the same schedule on every rank, halo planes in between
    static char *at(mg_ctx *c, size_t off) { return (char *)c->arena + off; }
    int slab_exchange(SlabGroup *g, size_t off, int lv, int depth, bool force = false)
    {
        mg_ctx *c0 = g->m[0];
        if (c0->p2p && !force) return MG_OK;   // the producing kernel already stored the ghosts
        const size_t pb = c0->plane_elems(lv) * c0->elem;
        const int G = c0->G, nz = c0->nzl[lv];
        const size_t nb = (size_t)depth * pb;
        if (g->nccl) {
            mg_ctx *c = c0;
            char *b = at(c, off);
            NcclApi *api = g->api;
            int e = api->GroupStart();
            if (c->rank > 0) {
                if (!e) e = api->Send(b + (size_t)G * pb, nb, NcclApi::kInt8, c->rank - 1, g->comm, c->stream);
                if (!e) e = api->Recv(b + (size_t)(G - depth) * pb, nb, NcclApi::kInt8, c->rank - 1, g->comm, c->stream);
            }
            if (c->rank < g->nranks - 1) {
                if (!e) e = api->Send(b + (size_t)(G + nz - depth) * pb, nb, NcclApi::kInt8, c->rank + 1, g->comm, c->stream);
                if (!e) e = api->Recv(b + (size_t)(G + nz) * pb, nb, NcclApi::kInt8, c->rank + 1, g->comm, c->stream);
            }
            int e2 = api->GroupEnd();
            if (e || e2) return c->fail(MG_ECUDA, api->GetErrorString(e ? e : e2));
        } else {
            // LOCAL: one device, one stream. MULTI: every slab on its own device and stream -- the sources must be
            // final before the copies and the ghosts in place before any neighbour's next kernel: group-wide syncs
            // (this path only runs after initCells / an upload, or with the fused exchange switched off).
            int rc;
            if (g->multi && (rc = group_sync(g))) return rc;
            for (int r = 0; r < g->nranks; ++r) {
                mg_ctx *c = g->m[r];
                if ((rc = c->activate())) return rc;
                if (r > 0)
                    MG_CK(c, cudaMemcpyAsync(at(c, off) + (size_t)(G - depth) * pb,
                                             at(g->m[r - 1], off) + (size_t)(G + nz - depth) * pb, nb,
                                             cudaMemcpyDefault, c->stream));
                if (r < g->nranks - 1)
                    MG_CK(c, cudaMemcpyAsync(at(c, off) + (size_t)(G + nz) * pb, at(g->m[r + 1], off) + (size_t)G * pb,
                                             nb, cudaMemcpyDefault, c->stream));
            }
            if (g->multi && (rc = group_sync(g))) return rc;
        }
        g->exchanges++;
        g->exchanged_bytes += 2 * nb;
        return MG_OK;
    }
    static int group_sync(SlabGroup *g)
    {
        for (mg_ctx *c : g->m) {
            int rc = c->activate();
            if (rc) return rc;
            MG_CK(c, cudaStreamSynchronize(c->stream));
        }
        return MG_OK;
    }
    // fused transport: the RES pass has stored this rank's part of the coarse residual into every rank's cube;
    // publish the epoch to everybody, then wait for everybody's (mg_slab.cuh)
    int slab_allgather_fence(SlabGroup *g)
    {
        if (!g->concurrent()) return MG_OK;   // LOCAL: one stream orders everything
        for (mg_ctx *c : g->m) {
            int rc = c->activate();
            if (rc) return rc;
            SlabPeers sp;
            memset(&sp, 0, sizeof(sp));
            sp.nranks = c->nranks; sp.rank = c->rank;
            for (int r = 0; r < c->nranks; ++r) sp.hs[r] = (unsigned long long *)c->peer[r];
            k_slab_signal_all<<<1, 32, 0, c->stream>>>(sp);
            MG_LAUNCH_CHECK(c);
            k_slab_wait_all<<<1, 32, 0, c->stream>>>(sp);
            MG_LAUNCH_CHECK(c);
        }
        return MG_OK;
    }
    // replicated level lv: every rank has produced planes [r*L/P, (r+1)*L/P) of the cube
    int slab_allgather(SlabGroup *g, size_t off, int lv)
    {
        mg_ctx *c0 = g->m[0];
        const size_t part = c0->plane_elems(lv) * c0->elem * (size_t)((1 << lv) / g->nranks);
        if (g->nccl) {
            char *b = at(c0, off);
            int e = g->api->AllGather(b + (size_t)c0->rank * part, b, part, NcclApi::kInt8, g->comm, c0->stream);
            if (e) return c0->fail(MG_ECUDA, g->api->GetErrorString(e));
        } else {
            int rc;
            if (g->multi && (rc = group_sync(g))) return rc;
            for (int r = 0; r < g->nranks; ++r) {
                if ((rc = g->m[r]->activate())) return rc;
                for (int o = 0; o < g->nranks; ++o)
                    if (o != r)
                        MG_CK(c0, cudaMemcpyAsync(at(g->m[r], off) + (size_t)o * part, at(g->m[o], off) + (size_t)o * part,
                                                  part, cudaMemcpyDefault, g->m[r]->stream));
            }
            if (g->multi && (rc = group_sync(g))) return rc;
        }
        return MG_OK;
    }
    int slab_twogrid(SlabGroup *g, double h, size_t u_off, size_t f_off, int lv)
    {
        mg_ctx *c0 = g->m[0];
        int rc;
        if (!c0->dist[lv]) {  // replicated: every rank runs the ordinary single-GPU path
            if (g->concurrent() && c0->use_graph && !c0->capturing && !c0->prof_on) {
                // eager slab cycle on concurrent slabs: replay the whole coarse sub-cycle (dozens of small launches,
                // no communication inside) from a CUDA graph per slab
                for (mg_ctx *c : g->m) {
                    if ((rc = c->activate())) return rc;
                    if (!c->rep_gexec) {
                        cudaStream_t saved = c->stream;
                        MG_CK(c, cudaStreamBeginCapture(c->cap_stream, cudaStreamCaptureModeThreadLocal));
                        c->stream = c->cap_stream; c->capturing = true;
                        rc = twogrid_fused(c, h, at(c, u_off), at(c, f_off), lv);
                        c->stream = saved; c->capturing = false;
                        cudaError_t e = cudaStreamEndCapture(c->cap_stream, &c->rep_graph);
                        if (rc || e != cudaSuccess) {
                            if (c->rep_graph) cudaGraphDestroy(c->rep_graph);
                            c->rep_graph = nullptr;
                            return rc ? rc : c->fail_cuda(e, "cudaStreamEndCapture (replicated levels)");
                        }
                        MG_CK(c, cudaGraphInstantiate(&c->rep_gexec, c->rep_graph, 0));
                        size_t nn = 0;
                        MG_CK(c, cudaGraphGetNodes(c->rep_graph, nullptr, &nn));
                        c->rep_nodes = nn;
                    }
                    MG_CK(c, cudaGraphLaunch(c->rep_gexec, c->stream));
                    c0->launches += c->rep_nodes;
                }
                return MG_OK;
            }
            for (mg_ctx *c : g->m) {
                if ((rc = c->activate())) return rc;
                if ((rc = twogrid_fused(c, h, at(c, u_off), at(c, f_off), lv))) return rc;
            }
            return MG_OK;
        }
        if constexpr (DIM == 3) {
            const Coef<A> cf = make_coef<A>(DIM, h);
            const size_t W_off = c0->arena_off(c0->W[lv]);
            const size_t Rc_off = c0->arena_off(c0->R[lv - 1]), Vc_off = c0->arena_off(c0->V[lv - 1]);
            size_t cur = u_off, oth = W_off;
            // The passes ping-pong u <-> W[lv]. With the fused halo exchange the ghost planes of the
            // field a pass writes are filled by the NEIGHBOURS' kernels, so the result must end up in u
            // by itself: a trailing copy W -> u would read ghost planes the neighbour may not have
            // written yet (no handshake covers a memcpy). An odd pass count (smooth = 4 or 8 with
            // tb = 4) therefore gets one extra post-smoothing pass.
            const std::vector<int> pre = plan_passes(c0->smooth, c0->tb, true);
            std::vector<int> post = plan_passes(c0->smooth, c0->tb, false);
            if ((pre.size() + post.size()) & 1) post = plan_passes(c0->smooth, c0->tb, false, true);
            // pre-smoothing, residual + restriction fused into the last pass (cpu-raw.lua:198-218)
            int np = (int)pre.size();
            for (int i = 0; i < np; ++i) {
                const bool res = i == np - 1;
                if ((rc = slab_exchange(g, cur, lv, pre[i] + (res ? 1 : 0)))) return rc;
                for (mg_ctx *c : g->m)
                    if ((rc = stream3d_pass(c, lv, pre[i], false, res, (R *)at(c, oth), (const R *)at(c, cur),
                                            (const R *)at(c, f_off), nullptr, res ? (R *)at(c, Rc_off) : nullptr, cf)))
                        return rc;
                size_t t = cur; cur = oth; oth = t;
            }
            // the restricted residual is the next level's right-hand side
            if (c0->dist[lv - 1]) rc = slab_exchange(g, Rc_off, lv - 1, c0->G);
            else if (c0->p2p) rc = slab_allgather_fence(g);   // the RES pass stored into every rank's cube
            else rc = slab_allgather(g, Rc_off, lv - 1);
            if (rc) return rc;
            if ((rc = slab_twogrid(g, 2 * h, Vc_off, Rc_off, lv - 1))) return rc;   // cpu-raw.lua:221-222
            if (c0->dist[lv - 1] && (rc = slab_exchange(g, Vc_off, lv - 1, 2))) return rc;
            // prolongation + add fused into the first post-smoothing pass (cpu-raw.lua:225-236)
            np = (int)post.size();
            for (int i = 0; i < np; ++i) {
                if ((rc = slab_exchange(g, cur, lv, post[i]))) return rc;
                for (mg_ctx *c : g->m)
                    if ((rc = stream3d_pass(c, lv, post[i], i == 0, false, (R *)at(c, oth), (const R *)at(c, cur),
                                            (const R *)at(c, f_off), i == 0 ? (const R *)at(c, Vc_off) : nullptr, nullptr, cf)))
                        return rc;
                size_t t = cur; cur = oth; oth = t;
            }
            if (cur != u_off) {
                // only reachable with smooth = 1 pass plans that cannot be padded: copy the OWNED planes and have
                // the ghosts refreshed by an explicit exchange before the next pass reads them
                for (mg_ctx *c : g->m) {
                    const size_t o = c->own_off_elems(lv) * c->elem;
                    if ((rc = c->activate())) return rc;
                    MG_CK(c, cudaMemcpyAsync(at(c, u_off) + o, at(c, cur) + o, c->own_elems(lv) * c->elem,
                                             cudaMemcpyDeviceToDevice, c->stream));
                }
                if ((rc = slab_exchange(g, u_off, lv, c0->G, true))) return rc;
            }
            return MG_OK;
        } else {
            return c0->fail(MG_EUNSUPPORTED, "slab decomposition is 3-D only");
        }
    }
    int slab_vcycle(SlabGroup *g) override
    {
        mg_ctx *c0 = g->m[0];
        const int top = c0->nlevels - 1;
        if (c0->f_ghost_dirty) {  // f is static between uploads: its ghosts are exchanged once
            int rc = slab_exchange(g, c0->arena_off(c0->f), top, c0->G, true);
            if (rc) return rc;
            for (mg_ctx *c : g->m) c->f_ghost_dirty = false;
        }
        if (c0->u_ghost_dirty && c0->p2p) {  // after initCells / an upload; afterwards every pass keeps them fresh
            int rc = slab_exchange(g, c0->arena_off(c0->psi), top, c0->G, true);
            if (rc) return rc;
            for (mg_ctx *c : g->m) c->u_ghost_dirty = false;
        }
        return slab_twogrid(g, 1.0 / c0->size, c0->arena_off(c0->psi), c0->arena_off(c0->f), top);
    }

    // ---- 2-D warp-streaming smoother passes
    template <int S, bool PRO, bool RES>
    int launch_warp2d(mg_ctx *c, int L, R *dst, const R *src, const R *f, const R *Vp, R *Rout, const Coef<A> &cf)
    {
        typedef Warp2DCfg<S, RES> C;
        const int nstrips = (L + C::TXU - 1) / C::TXU;
        int TY = c->ty_override;
        if (TY <= 0) {  // rows per work item: waves of 148 SMs x 16 resident warps, times rows streamed
            long best = -1;
            for (int cand = L; cand >= MG_WARP2D_MIN_TY; cand >>= 1) {
                long items = (long)nstrips * ((L + cand - 1) / cand);
                long cost = ((items + 148 * 16 - 1) / (148 * 16)) * (cand + 2 * C::H);
                if (best < 0 || cost < best) { best = cost; TY = cand; }
            }
        }
        if (TY > L) TY = L;
        TY &= ~1;
        const int nitems = nstrips * ((L + TY - 1) / TY);
        c->prof_begin(PRO ? MG_K_SWEEP_PROLONG : (RES ? MG_K_SWEEP_RESTRICT : MG_K_SWEEP), L, S);
        auto kern = k_warp2d<R, A, S, PRO, RES>;
        constexpr int smem = C::template smem_bytes<R, A>();   // the warps' f (and source) row rings
        if (smem > 48 * 1024) MG_CK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        MG_CK(c, launch_k(c, kern, dim3((unsigned)((nitems + 3) / 4)), dim3(128), (size_t)smem, dst, src, f, Vp, Rout, L, TY, nstrips, nitems, cf));
        c->prof_end();
        MG_LAUNCH_CHECK(c);
        return MG_OK;
    }
    int warp2d_pass(mg_ctx *c, int L, int S, bool pro, bool res, R *dst, const R *src, const R *f,
                    const R *Vp, R *Rout, const Coef<A> &cf)
    {
#define MG_W2D(S_) \
    case S_: \
        if (pro) return launch_warp2d<S_, true, false>(c, L, dst, src, f, Vp, Rout, cf); \
        if (res) return launch_warp2d<S_, false, true>(c, L, dst, src, f, Vp, Rout, cf); \
        return launch_warp2d<S_, false, false>(c, L, dst, src, f, Vp, Rout, cf);
        switch (S) { MG_W2D(1) MG_W2D(2) MG_W2D(3) MG_W2D(4) MG_W2D(5) MG_W2D(6) MG_W2D(7) }
#undef MG_W2D
        return c->fail(MG_EINVAL, "warp2d_pass: unsupported sweep count");
    }
    int sweeps_warp2d(mg_ctx *c, int L, R *&cur, R *&oth, const R *f, const Coef<A> &cf, int n,
                      const R *Vp, R *Rout)
    {
        const int tb = c->tb2;
        int np = (n + tb - 1) / tb, base = n / np, extra = n % np;
        for (int i = 0; i < np; ++i) {
            int S = base + (i < extra ? 1 : 0);
            int rc = warp2d_pass(c, L, S, i == 0 && Vp, i == np - 1 && Rout, oth, cur, f, Vp, Rout, cf);
            if (rc) return rc;
            R *t = cur; cur = oth; oth = t;
        }
        return MG_OK;
    }

    // ---- block-temporal passes of the mid levels (mg_block3d.cuh), 3-D only
    template <int S, bool PRO, int BX, int BY, int BZ>
    int launch_block3d(mg_ctx *c, int L, R *dst, const R *src, const R *f, const R *Vp, const Coef<A> &cf)
    {
        typedef Block3DCfg<S, BX, BY, BZ> C;
        auto kern = k_block3d<R, A, S, PRO, BX, BY, BZ>;
        constexpr int smem = C::template smem_bytes<R>();
        MG_CK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        const unsigned nblk = (unsigned)((L / BX) * (L / BY) * (L / BZ));
        c->prof_begin(PRO ? MG_K_SWEEP_PROLONG : MG_K_SWEEP, L, S);
        MG_CK(c, launch_k(c, kern, dim3(nblk), dim3(C::NTHREADS), (size_t)smem, dst, src, f, Vp, L, cf));
        c->prof_end();
        MG_LAUNCH_CHECK(c);
        return MG_OK;
    }
    int block3d_pass(mg_ctx *c, int L, int S, bool pro, R *dst, const R *src, const R *f, const R *Vp, const Coef<A> &cf)
    {
        if constexpr (DIM == 3) {
#define MG_B3D(S_, P_) \
    return L <= 32 ? launch_block3d<S_, P_, 8, 8, 8>(c, L, dst, src, f, Vp, cf) : launch_block3d<S_, P_, 16, 16, 8>(c, L, dst, src, f, Vp, cf)
            if (pro) { switch (S) { case 1: MG_B3D(1, true); case 2: MG_B3D(2, true); case 3: MG_B3D(3, true); case 4: MG_B3D(4, true); } }
            else { switch (S) { case 1: MG_B3D(1, false); case 2: MG_B3D(2, false); case 3: MG_B3D(3, false); case 4: MG_B3D(4, false); } }
#undef MG_B3D
        }
        return c->fail(MG_EINVAL, "block3d_pass: unsupported combination");
    }

    int sweeps(mg_ctx *c, int lv, R *&cur, R *&oth, const R *f, double h, int n, const R *Vp, R *Rout)
    {
        const int L = 1 << lv;
        const Coef<A> cf = make_coef<A>(DIM, h, c->omega);
        const bool ref_omega = c->omega == 1.0;   // the fused multi-sweep kernels implement the reference's omega = 1 only
        if constexpr (DIM == 2) {
            if (ref_omega && c->tb2 >= 1 && L >= c->warp2d_min_L && n >= 1 && !(Vp && Rout && n <= c->tb2))
                return sweeps_warp2d(c, L, cur, oth, f, cf, n, Vp, Rout);
        }
        // 3-D mid levels that are not cut into slabs: up to 4 sweeps per launch in shared memory (mg_block3d.cuh)
        const bool use_block = DIM == 3 && ref_omega && n >= 1 && L >= 32 && L <= c->block_max_L && !c->dist[lv];
        if constexpr (DIM == 3) {
            if (!use_block && ref_omega && c->tb >= 1 && L >= c->stream_min_L && n >= 1 && !(Vp && Rout && n <= c->tb))
                return sweeps_stream3d(c, lv, cur, oth, f, cf, n, Vp, Rout);
        }
        dim3 b = block_for(L), g = grid_for(DIM, L, b);
        bool blocked = false;
        if constexpr (DIM == 3) {
            if (use_block) {
                const std::vector<int> plan = plan_passes(n, 4, false);
                for (size_t i = 0; i < plan.size(); ++i) {
                    int rc = block3d_pass(c, L, plan[i], i == 0 && Vp, oth, cur, f, Vp, cf);
                    if (rc) return rc;
                    R *t = cur; cur = oth; oth = t;
                }
                blocked = true;
            }
        }
        for (int s = 0; s < n && !blocked; ++s) {
            c->prof_begin(s == 0 && Vp ? MG_K_SWEEP_PROLONG : MG_K_SWEEP, L, 1);
            if (s == 0 && Vp)
                MG_CK(c, launch_k(c, k_sweep_pp<R, A, DIM, true>, g, b, 0, oth, (const R *)cur, f, Vp, L, cf));
            else
                MG_CK(c, launch_k(c, k_sweep_pp<R, A, DIM, false>, g, b, 0, oth, (const R *)cur, f, (const R *)nullptr, L, cf));
            c->prof_end();
            MG_LAUNCH_CHECK(c);
            R *t = cur; cur = oth; oth = t;
        }
        if (n == 0 && Vp) {
            c->prof_begin(MG_K_PROLONG_ADD, L, 0);
            MG_CK(c, launch_k(c, k_prolong_add<R, A, DIM>, g, b, 0, cur, Vp, L));
            c->prof_end();
            MG_LAUNCH_CHECK(c);
        }
        if (Rout) {
            const int L2 = L / 2;
            dim3 b2 = block_for(L2), g2 = grid_for(DIM, L2, b2);
            c->prof_begin(MG_K_RESID_RESTRICT, L, 0);
            MG_CK(c, launch_k(c, k_residual_restrict<R, A, DIM>, g2, b2, 0, Rout, f, (const R *)cur, L, cf));
            c->prof_end();
            MG_LAUNCH_CHECK(c);
        }
        return MG_OK;
    }
    int settle(mg_ctx *c, int lv, R *u, R *cur)
    {
        if (cur != u) {
            c->prof_begin(MG_K_COPY, 1 << lv, 0);
            MG_CK(c, cudaMemcpyAsync(u, cur, c->level_bytes(lv), cudaMemcpyDeviceToDevice, c->stream));
            c->prof_end();
        }
        return MG_OK;
    }
    int smooth(mg_ctx *c, int lv, void *u_, const void *f, double h, int n) override
    {
        R *u = (R *)u_;
        if (c->mode == MG_MODE_REFSEQ) {
            int rc = c->ensure_debug_arena();
            for (int i = 0; i < n && !rc; ++i) rc = in_place_solver(c, lv, u, (const R *)f, h);
            return rc;
        }
        R *cur = u, *oth = (R *)c->W[lv];
        int rc = sweeps(c, lv, cur, oth, (const R *)f, h, n, nullptr, nullptr);
        if (rc) return rc;
        return settle(c, lv, u, cur);
    }
    int pre_fused(mg_ctx *c, int lv, void *u_, const void *f, double h, int n, void *Rout) override
    {
        R *u = (R *)u_, *cur = u, *oth = (R *)c->W[lv];
        int rc = sweeps(c, lv, cur, oth, (const R *)f, h, n, nullptr, (R *)Rout);
        if (rc) return rc;
        return settle(c, lv, u, cur);
    }
    int post_fused(mg_ctx *c, int lv, void *u_, const void *f, double h, int n, const void *Vp) override
    {
        R *u = (R *)u_, *cur = u, *oth = (R *)c->W[lv];
        int rc = sweeps(c, lv, cur, oth, (const R *)f, h, n, (const R *)Vp, nullptr);
        if (rc) return rc;
        return settle(c, lv, u, cur);
    }

    // (K-d) every level <= lv in one launch of the persistent kernel
    int small_vcycle(mg_ctx *c, double h, R *u, const R *f, int lv)
    {
        SmallArgs<R, A> a;
        memset(&a, 0, sizeof(a));
        a.top = lv;
        a.smooth = c->smooth;
        double hl = h;
        for (int l = lv; l >= 0; --l, hl *= 2) {
            a.u[l] = l == lv ? u : (R *)c->V[l];
            a.f[l] = l == lv ? f : (const R *)c->R[l];
            a.w[l] = (R *)c->W[l];
            a.coef[l] = make_coef<A>(DIM, hl, c->omega);
        }
        size_t n = c->level_elems(lv);
        int threads = n >= 1024 ? 1024 : (n >= 256 ? 256 : 64);
        c->prof_begin(MG_K_SMALL, 1 << lv, 2 * c->smooth);
        // the whole sub-hierarchy in shared memory (mg_small.cuh, K-d3) when it fits; else the global-memory walker
        const size_t sm_bytes = small_smem_bytes<DIM>(lv, sizeof(R));
        if (c->small_smem_opt != 0 && lv <= SmallSmemMax<DIM>::LG && sm_bytes <= 227 * 1024) {
            auto kern = k_small_vcycle_smem<R, A, DIM>;
            MG_CK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_bytes));
            MG_CK(c, launch_k(c, kern, dim3(1), dim3((unsigned)small_team(DIM, lv)), sm_bytes, a));
            c->prof_end();
            MG_LAUNCH_CHECK(c);
            return MG_OK;
        }
        MG_CK(c, launch_k(c, k_small_vcycle<R, A, DIM>, dim3(1), dim3((unsigned)threads), 0, a));
        c->prof_end();
        MG_LAUNCH_CHECK(c);
        return MG_OK;
    }

    // (K-d2) every level <= lv in one launch of one thread-block cluster (mg_small.cuh, k_cluster_vcycle); the levels
    // <= small_L inside it by the cluster's first CTA alone
    int cluster_vcycle(mg_ctx *c, double h, R *u, const R *f, int lv)
    {
        ClusterArgs<R, A> ca;
        memset(&ca, 0, sizeof(ca));
        SmallArgs<R, A> &a = ca.s;
        a.top = lv;
        a.smooth = c->smooth;
        double hl = h;
        for (int l = lv; l >= 0; --l, hl *= 2) {
            a.u[l] = l == lv ? u : (R *)c->V[l];
            a.f[l] = l == lv ? f : (const R *)c->R[l];
            a.w[l] = (R *)c->W[l];
            a.coef[l] = make_coef<A>(DIM, hl, c->omega);
        }
        ca.top1 = 0;
        while ((2 << ca.top1) <= c->small_L && ca.top1 + 1 < lv) ++ca.top1;
        auto kern = k_cluster_vcycle<R, A, DIM>;
        if (c->cluster_ctas == 0) {   // the widest cluster the device schedules: 16 (non-portable size) or 8
            c->cluster_ctas = 8;
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
                cudaLaunchConfig_t q = {};
                q.gridDim = dim3(16); q.blockDim = dim3(1024);
                cudaLaunchAttribute qa[1];
                qa[0].id = cudaLaunchAttributeClusterDimension;
                qa[0].val.clusterDim.x = 16; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
                q.attrs = qa; q.numAttrs = 1;
                int ncl = 0;
                if (cudaOccupancyMaxActiveClusters(&ncl, kern, &q) == cudaSuccess && ncl >= 1) c->cluster_ctas = 16;
            }
            cudaGetLastError();
        }
        if (c->cluster_ctas > 8) MG_CK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)c->cluster_ctas); cfg.blockDim = dim3(1024); cfg.stream = c->stream;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)c->cluster_ctas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = c->pdl_opt == 1 ? 2 : 1;
        c->prof_begin(MG_K_SMALL, 1 << lv, 2 * c->smooth);
        MG_CK(c, cudaLaunchKernelEx(&cfg, kern, ca));
        c->prof_end();
        MG_LAUNCH_CHECK(c);
        return MG_OK;
    }

    int twogrid_fused(mg_ctx *c, double h, void *u_, const void *f_, int lv) override
    {
        R *u = (R *)u_;
        const R *f = (const R *)f_;
        if (lv == 0 || (1 << lv) <= c->small_L) return small_vcycle(c, h, u, f, lv);
        if ((1 << lv) <= c->cluster_L && lv < SMALL_MAX_LEVELS) return cluster_vcycle(c, h, u, f, lv);
        R *cur = u, *oth = (R *)c->W[lv];
        R *Rc = (R *)c->R[lv - 1], *Vc = (R *)c->V[lv - 1];
        int rc;
        if ((rc = sweeps(c, lv, cur, oth, f, h, c->smooth, nullptr, Rc))) return rc;
        if ((rc = twogrid_fused(c, 2 * h, Vc, Rc, lv - 1))) return rc;
        if ((rc = sweeps(c, lv, cur, oth, f, h, c->smooth, Vc, nullptr))) return rc;
        return settle(c, lv, u, cur);
    }

    // ------------------------------------------------------------ norms
    int reduce_to_host(mg_ctx *c, double *out)
    {
        k_final_sum<<<1, 1024, 0, c->stream>>>(c->d_partial, c->npartial, c->d_scalar);
        MG_LAUNCH_CHECK(c);
        MG_CK(c, cudaMemcpyAsync(c->h_scalar, c->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        MG_CK(c, cudaStreamSynchronize(c->stream));
        *out = *c->h_scalar;
        return MG_OK;
    }
    // cpu-raw.lua:249-254
    int frob_err(mg_ctx *c, double *err, bool materialise) override
    {
        if (materialise) {
            int rc = c->ensure_debug_arena();
            if (rc) return rc;
        }
        if (c->group) {  // slabs: every rank sums its owned planes, then the sums are added
            SlabGroup *g = c->group;
            double tot = 0;
            if (g->nccl) {
                int rc = frob_partial_sum(c, nullptr);
                if (rc) return rc;
                int e = g->api->AllReduce(c->d_scalar, c->d_scalar, 1, NcclApi::kFloat64, NcclApi::kSum, g->comm, c->stream);
                if (e) return c->fail(MG_ECUDA, g->api->GetErrorString(e));
                MG_CK(c, cudaMemcpyAsync(c->h_scalar, c->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
                MG_CK(c, cudaStreamSynchronize(c->stream));
                tot = *c->h_scalar;
            } else {
                for (mg_ctx *m : g->m) {
                    double part;
                    int rc = m->activate();
                    if (rc) return rc;
                    rc = frob_partial_sum(m, &part);
                    if (rc) return rc;
                    tot += part;
                }
            }
            *err = std::sqrt(tot / (double)c->N);
            return MG_OK;
        }
        k_frob_partial<R, A><<<c->npartial, 256, 0, c->stream>>>(
            (const R *)c->psi, (const R *)c->psiOld, materialise ? (R *)c->errorBuf : nullptr, c->N,
            c->d_partial);
        MG_LAUNCH_CHECK(c);
        double s;
        int rc = reduce_to_host(c, &s);
        if (rc) return rc;
        *err = std::sqrt(s / (double)c->N);
        return MG_OK;
    }
    // sum of (psi - psiOld)^2 over the planes this rank owns; sum == nullptr leaves it in d_scalar
    int frob_partial_sum(mg_ctx *c, double *sum) override
    {
        const int top = c->nlevels - 1;
        const size_t o = c->own_off_elems(top), n = c->own_elems(top);
        k_frob_partial<R, A><<<c->npartial, 256, 0, c->stream>>>((const R *)c->psi + o, (const R *)c->psiOld + o,
                                                                 nullptr, n, c->d_partial);
        MG_LAUNCH_CHECK(c);
        if (sum) return reduce_to_host(c, sum);
        k_final_sum<<<1, 1024, 0, c->stream>>>(c->d_partial, c->npartial, c->d_scalar);
        MG_LAUNCH_CHECK(c);
        return MG_OK;
    }
    // ------------------------------------------------------------ Krylov comparator (mg_krylov.cuh)
    int linf_norm(mg_ctx *c, const void *field, size_t n, double *out) override
    {
        k_absmax_partial<R><<<c->npartial, 256, 0, c->stream>>>((const R *)field, n, c->d_partial);
        MG_LAUNCH_CHECK(c);
        k_cg_reduce<<<1, 1024, 0, c->stream>>>(c->d_partial, c->npartial, c->d_scalar, 0, 1);
        MG_LAUNCH_CHECK(c);
        MG_CK(c, cudaMemcpyAsync(c->h_scalar, c->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        MG_CK(c, cudaStreamSynchronize(c->stream));
        *out = *c->h_scalar;
        return MG_OK;
    }
    // x = psi (initial guess as found, the experiment uses -f = initCells' psi), b = f.
    // Slabs (one process per GPU): every rank iterates on the planes it owns; the ghost planes of p are refreshed by
    // ncclSend/ncclRecv before every operator application and the three scalars of an iteration are all-reduced.
    int cg(mg_ctx *c, int max_iter, double epsilon, double *err_hist, double *linf_hist, int *n_done) override
    {
        SlabGroup *g = c->group;
        if (g && !g->nccl) return c->fail(MG_EUNSUPPORTED, "mg_cg on slabs: one process per GPU (mg_create_slab) only");
        if (g && DIM != 3) return c->fail(MG_EUNSUPPORTED, "slab decomposition is 3-D only");
        const int top = c->nlevels - 1, L = c->size;
        const size_t n = g ? c->own_elems(top) : c->N, off = g ? c->own_off_elems(top) : 0;
        const int slab = g ? 1 : 0;
        const size_t field_bytes = (c->Ntop * c->elem + 7) / 8 * 8;   // the double partials behind the field stay 8-byte aligned
        if (!c->cg_tmp) {
            MG_CK(c, cudaMalloc(&c->cg_tmp, field_bytes + 2 * sizeof(double) * (size_t)c->npartial + sizeof(double) * CG_NSCAL));
            MG_CK(c, cudaMemsetAsync(c->cg_tmp, 0, field_bytes, c->stream));
        }
        R *x = (R *)c->psi + off, *r = (R *)c->W[top] + off, *p = (R *)c->psiOld + off, *Ap = (R *)c->cg_tmp + off;
        const R *b = (const R *)c->f + off;
        double *part2 = (double *)((char *)c->cg_tmp + field_bytes), *scal = part2 + 2 * c->npartial;
        double *partA = c->d_partial, *partB = part2;
        const A inv_h2 = make_coef<A>(DIM, 1.0 / L).inv_h2;
        const int nb = c->npartial;
        int rc;
        auto reduce = [&](double *part, int slot, int is_max) -> int {
            k_cg_reduce<<<1, 1024, 0, c->stream>>>(part, nb, scal, slot, is_max);
            c->count_launch();
            if (g) {
                int e = g->api->AllReduce(scal + slot, scal + slot, 1, NcclApi::kFloat64, is_max ? NcclApi::kMax : NcclApi::kSum, g->comm, c->stream);
                if (e) return c->fail(MG_ECUDA, g->api->GetErrorString(e));
            }
            return MG_OK;
        };
        auto ghosts = [&](const void *field) -> int {   // planes next to the slab, from the two neighbours
            if (!g) return MG_OK;
            const bool was = c->p2p;
            c->p2p = false;                 // an explicit exchange, whatever the smoother's transport is
            int e = slab_exchange(g, c->arena_off(field), top, 1, true);
            c->p2p = was;
            return e;
        };
        if ((rc = ghosts(c->psi))) return rc;
        k_cg_init<R, A, DIM><<<nb, 256, 0, c->stream>>>(r, p, x, b, L, inv_h2, partA, partB, n, slab);
        MG_LAUNCH_CHECK(c);
        if ((rc = reduce(partA, CG_RR, 0))) return rc;
        if ((rc = reduce(partB, CG_BB, 0))) return rc;
        double h[CG_NSCAL];
        MG_CK(c, cudaMemcpyAsync(h, scal, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        MG_CK(c, cudaStreamSynchronize(c->stream));
        const double bb = h[CG_BB] == 0 ? 1.0 : h[CG_BB];
        int it = 0;
        double err = std::sqrt(h[CG_RR] / bb);
        if (!(err < epsilon)) {
            for (it = 1; it <= max_iter; ++it) {
                if ((rc = ghosts(c->psiOld))) return rc;
                k_cg_apply<R, A, DIM><<<nb, 256, 0, c->stream>>>(Ap, p, L, inv_h2, partA, n, slab);
                MG_LAUNCH_CHECK(c);
                if ((rc = reduce(partA, CG_PAP, 0))) return rc;
                k_cg_update<R, A><<<nb, 256, 0, c->stream>>>(x, r, p, Ap, n, scal, partA, partB);
                MG_LAUNCH_CHECK(c);
                if ((rc = reduce(partA, CG_RRNEW, 0))) return rc;
                if ((rc = reduce(partB, CG_XMAX, 1))) return rc;
                MG_CK(c, cudaMemcpyAsync(h, scal, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
                MG_CK(c, cudaStreamSynchronize(c->stream));
                err = std::sqrt(h[CG_RRNEW] / bb);
                if (err_hist) err_hist[it - 1] = err;
                if (linf_hist) linf_hist[it - 1] = h[CG_XMAX];
                if (err < epsilon || !std::isfinite(err)) break;
                k_cg_direction<R, A><<<nb, 256, 0, c->stream>>>(p, r, n, scal);
                MG_LAUNCH_CHECK(c);
                k_cg_shift<<<1, 1, 0, c->stream>>>(scal);
                c->count_launch();
            }
            if (it > max_iter) it = max_iter;
        }
        if (g) { c->u_ghost_dirty = true; }   // psi changed behind the smoother's back: its ghosts are stale
        if (n_done) *n_done = it;
        return MG_OK;
    }

    int residual_norm(mg_ctx *c, double *rms) override
    {
        // uses the partner of psi as scratch for r; W[top] is free outside a V-cycle
        const int top = c->nlevels - 1;
        int rc;
        if (c->group) {
            // slabs: every rank computes r on the planes it owns (psi's ghost planes are kept fresh by the smoother
            // passes; after initCells / an upload they are refreshed first), sums its squares, and the sums are added
            if constexpr (DIM == 3) {
                SlabGroup *g = c->group;
                mg_ctx *c0 = g->m[0];
                if (c0->u_ghost_dirty || !c0->p2p) {
                    if ((rc = slab_exchange(g, c0->arena_off(c0->psi), top, c0->G, true))) return rc;
                    if (c0->p2p) for (mg_ctx *m : g->m) m->u_ghost_dirty = false;
                }
                const Coef<A> cf = make_coef<A>(DIM, 1.0 / c->size);
                double tot = 0;
                for (mg_ctx *m : g->m) {
                    if ((rc = m->activate())) return rc;
                    dim3 b = block_for(m->size), gr = grid_for(DIM, m->size, b);
                    gr.z = (unsigned)m->nzl[top];
                    if (g->concurrent() && m->p2p) {   // the neighbours' last pass fills our ghost planes: wait for it
                        k_slab_wait_neighbours<<<1, 32, 0, m->stream>>>((unsigned long long *)m->arena, m->peer_lo != nullptr, m->peer_hi != nullptr);
                        MG_LAUNCH_CHECK(m);
                    }
                    k_residual_slab<R, A><<<gr, b, 0, m->stream>>>((R *)m->W[top], (const R *)m->f, (const R *)m->psi, m->size, m->G, cf);
                    MG_LAUNCH_CHECK(m);
                    k_sumsq_partial<R><<<m->npartial, 256, 0, m->stream>>>((const R *)m->W[top] + m->own_off_elems(top), m->own_elems(top), m->d_partial);
                    MG_LAUNCH_CHECK(m);
                    if (g->nccl) {
                        k_final_sum<<<1, 1024, 0, m->stream>>>(m->d_partial, m->npartial, m->d_scalar);
                        MG_LAUNCH_CHECK(m);
                        int e = g->api->AllReduce(m->d_scalar, m->d_scalar, 1, NcclApi::kFloat64, NcclApi::kSum, g->comm, m->stream);
                        if (e) return m->fail(MG_ECUDA, g->api->GetErrorString(e));
                        MG_CK(m, cudaMemcpyAsync(m->h_scalar, m->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, m->stream));
                        MG_CK(m, cudaStreamSynchronize(m->stream));
                        tot = *m->h_scalar;
                    } else {
                        double part;
                        if ((rc = reduce_to_host(m, &part))) return rc;
                        tot += part;
                    }
                }
                *rms = std::sqrt(tot / (double)c->N);
                return MG_OK;
            } else {
                return c->fail(MG_EUNSUPPORTED, "slab decomposition is 3-D only");
            }
        }
        R *scratch = (R *)c->W[top];
        rc = residual(c, c->size, scratch, c->f, c->psi, 1.0 / c->size);
        if (rc) return rc;
        k_sumsq_partial<R><<<c->npartial, 256, 0, c->stream>>>(scratch, c->N, c->d_partial);
        MG_LAUNCH_CHECK(c);
        double s;
        if ((rc = reduce_to_host(c, &s))) return rc;
        *rms = std::sqrt(s / (double)c->N);
        return MG_OK;
    }
};

}  // namespace mg

// ======================================================================= mg_ctx methods
inline int mg_ctx::init(int dim_, int size_, int real_kind_, int smooth_, int device_, int rank_, int nranks_)
{
    dim = dim_; size = size_; real_kind = real_kind_; rank = rank_; nranks = nranks_;
    smooth = smooth_ > 0 ? smooth_ : 7;  // cpu-raw.lua:123
    elem = real_kind == MG_REAL_F64 ? 8 : 4;
    nlevels = 0;
    while ((1 << nlevels) < size) ++nlevels;
    ++nlevels;
    if (nlevels > mg::MAX_LEVELS) return fail(MG_EINVAL, "too many levels");
    if (nranks > 1) {  // slab decomposition along z (mg_slab.cuh)
        if (dim != 3) return fail(MG_EUNSUPPORTED, "slab decomposition is 3-D only");
        if (nranks > 8 || (nranks & (nranks - 1))) return fail(MG_EINVAL, "nranks must be 2, 4 or 8");
        G = 4;
        // a level is cut across the ranks while every rank keeps at least this many planes
        // (>= 8 = two ghost depths); thinner levels are replicated. 32 measured best on 8 GPUs
        // (1024^3: 1733 vs 1666 units/s with 8); MGPOISSON_SLAB_MIN_PLANES tunes it.
        int min_planes = g_slab_min_planes;   // mg_set_global_option("slab_min_planes"); the environment overrides it
        if (const char *e = getenv("MGPOISSON_SLAB_MIN_PLANES")) min_planes = atoi(e) < 8 ? 8 : atoi(e);
        for (int lv = 0; lv < nlevels; ++lv) {
            const int L = 1 << lv;
            if (L >= 64 && L / nranks >= min_planes) { dist[lv] = true; nzl[lv] = L / nranks; }
        }
        if (!dist[nlevels - 1]) return fail(MG_EINVAL, "grid too small to be cut into slabs (need size >= 64 and size/nranks >= slab_min_planes, 32 by default)");
        // (replicated levels run the single-GPU path with the single-GPU thresholds: 64^3 through the one-sweep launches
        // is 36 us per cycle cheaper than through the streaming kernel, measured at N = 1; distributed levels always stream)
        tb = 4;
    }
    N = (size_t)size * size * (dim == 3 ? (size_t)size : 1);
    Ntop = level_elems(nlevels - 1);
    small_L = dim == 3 ? 16 : 64;
    cluster_L = 0;   // MEASURED (512^3 fp32): with 64^3 and 32^3 in the one-cluster kernel 371 V-cycles/s against 386 with the 30 separate
                     // launches -- 16 SMs reading L2 (the acquire of every cluster barrier empties L1) lose to 148 SMs plus launch
                     // gaps; with 32^3 only: 385 (no change). Kept as an option ("cluster_L"), off by default.
    tb2 = real_kind == MG_REAL_F32 ? 7 : 4;  // double arithmetic: 8 pipeline stages would spill
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(MG_ECUDA, "no CUDA device: libmgpoisson has no CPU fallback");
    if (device_ < 0) MG_CK(this, cudaGetDevice(&device_));
    device = device_;
    MG_CK(this, cudaSetDevice(device));
    MG_CK(this, cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device));
    MG_CK(this, cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking));
    MG_CK(this, cudaStreamCreateWithFlags(&cap_stream, cudaStreamNonBlocking));
    stream = own_stream;

    // arena layout: f, psi, psiOld, then per level R, V (below the top) and the ping-pong partner W
    auto up = [](size_t x) { return (x + mg::ARENA_ALIGN - 1) / mg::ARENA_ALIGN * mg::ARENA_ALIGN; };
    const int top = nlevels - 1;
    size_t off = mg::ARENA_HEADER;
    size_t o_f = off; off += up(Ntop * elem);
    size_t o_psi = off; off += up(Ntop * elem);
    size_t o_old = off; off += up(Ntop * elem);
    size_t o_R[mg::MAX_LEVELS], o_V[mg::MAX_LEVELS], o_W[mg::MAX_LEVELS];
    for (int lv = 0; lv <= top; ++lv) {
        if (lv < top) {
            o_R[lv] = off; off += up(level_bytes(lv));
            o_V[lv] = off; off += up(level_bytes(lv));
        }
        o_W[lv] = off; off += up(level_bytes(lv));
    }
    arena_bytes = off;
    e = cudaMalloc(&arena, arena_bytes);
    if (e != cudaSuccess) {
        arena = nullptr;
        cudaGetLastError();
        return fail(MG_ENOMEM, "cudaMalloc of the grid-hierarchy arena failed");
    }
    MG_CK(this, cudaMemsetAsync(arena, 0, arena_bytes, stream));
    char *base = (char *)arena;
    f = base + o_f; psi = base + o_psi; psiOld = base + o_old;
    for (int lv = 0; lv <= top; ++lv) {
        if (lv < top) { R[lv] = base + o_R[lv]; V[lv] = base + o_V[lv]; }
        W[lv] = base + o_W[lv];
    }
    npartial = 1184;  // 8 blocks per SM on 148 SMs
    MG_CK(this, cudaMalloc(&d_partial, sizeof(double) * (size_t)(npartial + 8)));
    d_scalar = d_partial + npartial;
    MG_CK(this, cudaMallocHost(&h_scalar, sizeof(double) * 8));

    switch (real_kind * 10 + dim) {
    case MG_REAL_F64 * 10 + 2: eng = new mg::EngineT<double, double, 2>(); break;
    case MG_REAL_F64 * 10 + 3: eng = new mg::EngineT<double, double, 3>(); break;
    case MG_REAL_F32 * 10 + 2: eng = new mg::EngineT<float, float, 2>(); break;
    case MG_REAL_F32 * 10 + 3: eng = new mg::EngineT<float, float, 3>(); break;
    case MG_REAL_F32_ACC64 * 10 + 2: eng = new mg::EngineT<float, double, 2>(); break;
    case MG_REAL_F32_ACC64 * 10 + 3: eng = new mg::EngineT<float, double, 3>(); break;
    default: return fail(MG_EINVAL, "bad real_kind/dim");
    }
    int rc = eng->init_cells(this);  // cpu-raw.lua:173
    if (rc) return rc;
    return sync();
}

inline void mg_ctx::release()
{
    if (device >= 0) cudaSetDevice(device);
    if (peer_ipc) {
        for (int r = 0; r < mg::S3_MAX_RANKS; ++r)
            if (peer[r] && r != rank) cudaIpcCloseMemHandle(peer[r]);
    }
    for (int r = 0; r < mg::S3_MAX_RANKS; ++r) peer[r] = nullptr;
    peer_lo = peer_hi = nullptr;
    if (group && owns_group) {
        mg::SlabGroup *g = group;
        if (own_stream) cudaStreamSynchronize(stream);
        for (size_t r = 1; r < g->m.size(); ++r) {  // LOCAL: the other slabs belong to rank 0's handle
            g->m[r]->group = nullptr;
            if (!g->multi) g->m[r]->own_stream = nullptr;   // LOCAL: shared with rank 0 (MULTI: every slab owns its own)
            g->m[r]->release();
            delete g->m[r];
        }
        if (g->nccl && g->comm) g->api->CommDestroy(g->comm);
        delete g;
        group = nullptr;
    }
    drop_graph();
    if (own_stream) cudaStreamSynchronize(own_stream);
    pipe_release();
    if (arena) cudaFree(arena);
    if (debug_arena) cudaFree(debug_arena);
    if (d_partial) cudaFree(d_partial);
    if (cg_tmp) cudaFree(cg_tmp);
    cg_tmp = nullptr;
    if (slab_trace) cudaFree(slab_trace);
    slab_trace = nullptr;
    if (h_scalar) cudaFreeHost(h_scalar);
    if (own_stream) cudaStreamDestroy(own_stream);
    if (cap_stream) cudaStreamDestroy(cap_stream);
    delete eng;
    arena = debug_arena = nullptr; d_partial = nullptr; h_scalar = nullptr;
    own_stream = cap_stream = nullptr; eng = nullptr;
}

inline int mg_ctx::ensure_debug_arena()
{
    if (debug_arena) return MG_OK;
    auto up = [](size_t x) { return (x + mg::ARENA_ALIGN - 1) / mg::ARENA_ALIGN * mg::ARENA_ALIGN; };
    const int top = nlevels - 1;
    size_t off = 0;
    if (nranks > 1) return fail(MG_EUNSUPPORTED, "the reference-sequence mode is single-GPU only");
    size_t o_err = off; off += up(N * elem);
    size_t o_tmp = off; off += up(N * elem);
    size_t o_Rt = off; off += up(N * elem);
    size_t o_Vt = off; off += up(N * elem);
    size_t o_r[mg::MAX_LEVELS], o_v[mg::MAX_LEVELS];
    for (int lv = 0; lv <= top; ++lv) {
        o_r[lv] = off; off += up(level_bytes(lv));
        o_v[lv] = off; off += up(level_bytes(lv));
    }
    cudaError_t e = cudaMalloc(&debug_arena, off);
    if (e != cudaSuccess) {
        debug_arena = nullptr;
        cudaGetLastError();
        return fail(MG_ENOMEM, "cudaMalloc of the reference-sequence arena failed");
    }
    debug_arena_bytes = off;
    MG_CK(this, cudaMemsetAsync(debug_arena, 0, off, stream));
    char *base = (char *)debug_arena;
    errorBuf = base + o_err; tmpU = base + o_tmp;
    R[top] = base + o_Rt; V[top] = base + o_Vt;  // allocated by the reference, never used (cpu-raw.lua:162,164)
    for (int lv = 0; lv <= top; ++lv) { r[lv] = base + o_r[lv]; v[lv] = base + o_v[lv]; }
    return MG_OK;
}

inline void *mg_ctx::buffer(int which, int level, size_t *cap)
{
    int lv = nlevels - 1;
    if (which >= MG_BUF_r) {
        if (level < 1 || (level & (level - 1)) || level > size) return nullptr;
        lv = 0;
        while ((1 << lv) < level) ++lv;
    }
    if (cap) *cap = level_bytes(lv);
    switch (which) {
    case MG_BUF_F: return f;
    case MG_BUF_PSI: return psi;
    case MG_BUF_PSIOLD: return psiOld;
    case MG_BUF_ERRORBUF: return errorBuf;
    case MG_BUF_TMPU: return tmpU;
    case MG_BUF_r: return r[lv];
    case MG_BUF_R: return R[lv];
    case MG_BUF_v: return v[lv];
    case MG_BUF_V: return V[lv];
    }
    return nullptr;
}

// TMA descriptor of a dense L^3 field with a (box_x, box_y, 1) box; out-of-bounds elements are
// zero-filled, which is the reference's Dirichlet rule (cpu-raw.lua:36-39).
inline int mg_ctx::tensor_map(const void *base, int L, int nplanes, int box_x, int box_y, const CUtensorMap **out)
{
    auto key = std::make_tuple(base, L, nplanes, box_x, box_y);
    auto it = tmaps.find(key);
    if (it != tmaps.end()) { *out = &it->second; return MG_OK; }
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                 const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                 CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        MG_CK(this, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn || q != cudaDriverEntryPointSuccess) return fail(MG_ECUDA, "cuTensorMapEncodeTiled not available");
        encode = (EncodeFn)fn;
    }
    CUtensorMap m;
    cuuint64_t gdim[3] = {(cuuint64_t)L, (cuuint64_t)L, (cuuint64_t)nplanes};
    cuuint64_t gstr[2] = {(cuuint64_t)L * elem, (cuuint64_t)L * L * elem};
    cuuint32_t box[3] = {(cuuint32_t)box_x, (cuuint32_t)box_y, 1};
    cuuint32_t est[3] = {1, 1, 1};
    CUresult r = encode(&m, elem == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                        const_cast<void *>(base), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE,
                        tma_promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                       : (tma_promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                                         : (tma_promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                                           : CU_TENSOR_MAP_L2_PROMOTION_L2_256B)),
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r);
        return MG_ECUDA;
    }
    auto ins = tmaps.emplace(key, m);
    *out = &ins.first->second;
    return MG_OK;
}

// Host <-> device for one buffer. Single GPU: the whole field. NCCL slab: the planes this rank
// owns (nzl * L^2 elements). LOCAL slab group: the GLOBAL field, scattered / gathered over the
// member slabs (replicated levels are written to every member and read from rank 0).
static inline int mg_copy_impl(mg_ctx *c, int which, int level, void *host, size_t bytes, bool in)
{
    size_t cap = 0;
    void *d0 = c->buffer(which, level, &cap);
    if (!d0 || !host) return c->fail(MG_EINVAL, "no such buffer (or not materialised in this mode)");
    int lv = c->nlevels - 1;
    if (which >= MG_BUF_r) { lv = 0; while ((1 << lv) < level) ++lv; }
    auto cp = [&](void *dev, void *h, size_t nb) -> cudaError_t {
        return in ? cudaMemcpyAsync(dev, h, nb, cudaMemcpyHostToDevice, c->stream)
                  : cudaMemcpyAsync(h, dev, nb, cudaMemcpyDeviceToHost, c->stream);
    };
    if (!c->group) {
        if (bytes > cap) return c->fail(MG_EINVAL, "too many bytes for this buffer");
        MG_CK(c, cp(d0, host, bytes));
    } else if (c->group->nccl) {
        const size_t own = c->own_elems(lv) * c->elem;
        if (bytes > own) return c->fail(MG_EINVAL, "too many bytes for this rank's slab");
        MG_CK(c, cp((char *)d0 + c->own_off_elems(lv) * c->elem, host, bytes));
    } else {
        const size_t full = c->plane_elems(lv) * c->elem * ((size_t)1 << lv);
        if (bytes != full) return c->fail(MG_EINVAL, "slab group: pass the whole global field");
        const size_t off = c->arena_off(d0);
        for (size_t r = 0; r < c->group->m.size(); ++r) {
            mg_ctx *m = c->group->m[r];
            char *dev = (char *)m->arena + off;
            if (int rc = m->activate()) return rc;
            auto cpm = [&](void *d, void *h, size_t nb) -> cudaError_t {
                return in ? cudaMemcpyAsync(d, h, nb, cudaMemcpyHostToDevice, m->stream)
                          : cudaMemcpyAsync(h, d, nb, cudaMemcpyDeviceToHost, m->stream);
            };
            if (c->dist[lv]) {
                const size_t own = c->own_elems(lv) * c->elem;
                MG_CK(c, cpm(dev + c->own_off_elems(lv) * c->elem, (char *)host + r * own, own));
            } else if (in || r == 0) {
                MG_CK(c, cpm(dev, host, full));
            }
        }
    }
    if (in)
        for (mg_ctx *m : (c->group ? c->group->m : std::vector<mg_ctx *>{c})) {
            if (which == MG_BUF_F) m->f_ghost_dirty = true;
            if (which == MG_BUF_PSI) m->u_ghost_dirty = true;
        }
    return c->sync();
}
inline int mg_ctx::copy_in(int which, int lv, const void *host, size_t bytes)
{
    return mg_copy_impl(this, which, lv, const_cast<void *>(host), bytes, true);
}
inline int mg_ctx::copy_out(int which, int lv, void *host, size_t bytes)
{
    return mg_copy_impl(this, which, lv, host, bytes, false);
}

inline void mg_ctx::pipe_release()
{
    if (!pipe) return;
    if (pipe->up) { cudaStreamSynchronize(pipe->up); cudaStreamDestroy(pipe->up); }
    if (pipe->down) { cudaStreamSynchronize(pipe->down); cudaStreamDestroy(pipe->down); }
    for (int s = 0; s < 2; ++s) {
        if (pipe->in_f[s]) cudaFree(pipe->in_f[s]);
        if (pipe->in_u[s]) cudaFree(pipe->in_u[s]);
        if (pipe->out_u[s]) cudaFree(pipe->out_u[s]);
        for (cudaEvent_t e : {pipe->up_done[s], pipe->in_free[s], pipe->comp_done[s], pipe->out_free[s]})
            if (e) cudaEventDestroy(e);
    }
    if (pipe->h_sum) cudaFreeHost(pipe->h_sum);
    delete pipe;
    pipe = nullptr;
}

// staging slots, copy streams and events of mg_step_host_batch, created on first use (3 x 2 fields of nb bytes)
inline int mg_ctx::pipe_ensure(size_t nb, int n)
{
    if (pipe && pipe->nb != nb) pipe_release();
    if (!pipe) {
        pipe = new (std::nothrow) HostPipe();
        if (!pipe) return fail(MG_ENOMEM, "out of host memory");
        pipe->nb = nb;
        MG_CK(this, cudaStreamCreateWithFlags(&pipe->up, cudaStreamNonBlocking));
        MG_CK(this, cudaStreamCreateWithFlags(&pipe->down, cudaStreamNonBlocking));
        for (int s = 0; s < 2; ++s) {
            MG_CK(this, cudaMalloc(&pipe->in_f[s], nb));
            MG_CK(this, cudaMalloc(&pipe->in_u[s], nb));
            MG_CK(this, cudaMalloc(&pipe->out_u[s], nb));
            for (cudaEvent_t *e : {&pipe->up_done[s], &pipe->in_free[s], &pipe->comp_done[s], &pipe->out_free[s]})
                MG_CK(this, cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        }
    }
    if (pipe->cap < n) {
        if (pipe->h_sum) cudaFreeHost(pipe->h_sum);
        pipe->h_sum = nullptr; pipe->cap = 0;
        MG_CK(this, cudaMallocHost((void **)&pipe->h_sum, sizeof(double) * (size_t)n));
        pipe->cap = n;
    }
    return MG_OK;
}

inline void mg_ctx::drop_graph()
{
    if (gexec) cudaGraphExecDestroy(gexec);
    if (graph) cudaGraphDestroy(graph);
    if (rep_gexec) cudaGraphExecDestroy(rep_gexec);
    if (rep_graph) cudaGraphDestroy(rep_graph);
    gexec = rep_gexec = nullptr; graph = rep_graph = nullptr; graph_nodes = rep_nodes = 0;
}

inline int mg_ctx::sync()
{
    if (group && group->multi) {
        for (mg_ctx *m : group->m) {
            if (m->activate() != MG_OK) return fail(MG_ECUDA, "cudaSetDevice failed");
            MG_CK(this, cudaStreamSynchronize(m->stream));
        }
    } else {
        MG_CK(this, cudaStreamSynchronize(stream));
    }
    return check_peers();
}

// A slab kernel that waited HS_TIMEOUT_NS for a neighbour gave up and raised the time-out word of its arena header
// (mg_stream3d.cuh, s3_wait_counter): report it instead of handing back garbage silently.
inline int mg_ctx::check_peers()
{
    if (!group || !group->concurrent()) return MG_OK;
    for (mg_ctx *m : group->m) {
        unsigned long long w = 0;
        if (m->activate() != MG_OK) return fail(MG_ECUDA, "cudaSetDevice failed");
        MG_CK(this, cudaMemcpy(&w, (const char *)m->arena + 8 * mg::HS_TIMEOUT, sizeof(w), cudaMemcpyDeviceToHost));
        if (w != 0) return fail(MG_ESTATE, "slab handshake timed out: a neighbouring rank did not answer within 20 s");
    }
    return MG_OK;
}

// twoGrid(1/size, psi, f, size) (cpu-raw.lua:247)
inline int mg_ctx::vcycle()
{
    const double h = 1.0 / size;  // cpu-raw.lua:242
    const int top = nlevels - 1;
    if (mode == MG_MODE_REFSEQ) {
        int rc = ensure_debug_arena();
        if (rc) return rc;
        return eng->twogrid_refseq(this, h, psi, f, top);
    }
    // Slabs: the first cycle after initCells / an upload refreshes ghosts with explicit exchanges and
    // runs eagerly; afterwards (fused halo exchange, one process per GPU) the whole cycle --
    // smoother passes, handshake kernels, the all-gather of the replicated level -- is one graph.
    // (Off by default: "slab_graph" option. With libnccl 2.28 the captured cycle dead-locked on a
    // 2-GPU box -- the all-gather inside the capture is the suspect -- so slabs launch eagerly.)
    const bool slab_graph = group && group->nccl && p2p && use_graph && slab_graph_opt && !f_ghost_dirty && !u_ghost_dirty;
    if (group && !slab_graph) return eng->slab_vcycle(group);
    if (!group && !use_graph) return eng->twogrid_fused(this, h, psi, f, top);
    if (!gexec) {
        cudaStream_t saved = stream;
        MG_CK(this, cudaStreamSynchronize(saved));
        MG_CK(this, cudaStreamBeginCapture(cap_stream, cudaStreamCaptureModeThreadLocal));
        stream = cap_stream; capturing = true; cap_nvl_bytes = 0;
        int rc = group ? eng->slab_vcycle(group) : eng->twogrid_fused(this, h, psi, f, top);
        stream = saved; capturing = false; graph_nvl_bytes = cap_nvl_bytes;
        cudaError_t e = cudaStreamEndCapture(cap_stream, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); graph = nullptr; return rc; }
        if (e != cudaSuccess) { graph = nullptr; return fail_cuda(e, "cudaStreamEndCapture"); }
        MG_CK(this, cudaGraphInstantiate(&gexec, graph, 0));
        size_t nn = 0;
        MG_CK(this, cudaGraphGetNodes(graph, nullptr, &nn));
        graph_nodes = nn;
    }
    MG_CK(this, cudaGraphLaunch(gexec, stream));
    launches += graph_nodes;
    nvl_bytes += graph_nvl_bytes;
    return MG_OK;
}

// loop body of run() (cpu-raw.lua:246-254)
inline int mg_ctx::step(double *err_out)
{
    if (group) {
        for (mg_ctx *m : group->m) {
            if (int rc = m->activate()) return rc;
            MG_CK(this, cudaMemcpyAsync(m->psiOld, m->psi, Ntop * elem, cudaMemcpyDeviceToDevice, m->stream));
        }
    } else {
        MG_CK(this, cudaMemcpyAsync(psiOld, psi, Ntop * elem, cudaMemcpyDeviceToDevice, stream));
    }
    int rc = vcycle();
    if (rc) return rc;
    double e;
    rc = eng->frob_err(this, &e, mode == MG_MODE_REFSEQ);
    if (rc) return rc;
    if (err_out) *err_out = e;
    return MG_OK;
}

inline int mg_ctx::trace_rec(char name, int L, const void *dev, size_t bytes)
{
    if (!trace_on) return MG_OK;
    mg::TraceRec rec;
    rec.name = name; rec.L = L;
    rec.data.resize(bytes);
    MG_CK(this, cudaMemcpyAsync(rec.data.data(), dev, bytes, cudaMemcpyDeviceToHost, stream));
    MG_CK(this, cudaStreamSynchronize(stream));
    trace.push_back(std::move(rec));
    return MG_OK;
}

// mg_api.cu -- C ABI (include/mgpoisson.h) of libmgpoisson.so: context, grid-hierarchy arena
// (K-f), V-cycle scheduling (reference sequence and fused), CUDA-graph replay, reductions,
// trace, measurement helpers. No torch types, no exceptions across the boundary.
//
// Replaces, in the reference: MultigridCPURaw/MultigridGPU :init, :run, :twoGrid,
// :inPlaceIterativeSolver (cpu-raw.lua:142-258, gpu.lua:26-373). See the header for the
// per-function citations.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/mgpoisson.h"
#include "mg_engine.cuh"

using namespace mg;

static thread_local std::string g_create_error;
int g_slab_min_planes = 32;   // "slab_min_planes" (mg_set_global_option), read by mg_ctx::init

extern "C" {

const char *mg_version(void) { return "mgpoisson-b200 0.1 (sm_100a)"; }

const char *mg_last_error(mg_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

static int create_common(int dim, int size, int real_kind, int smooth, int device, int rank,
                         int nranks, mg_ctx **out)
{
    if (!out) return MG_EINVAL;
    *out = nullptr;
    if ((dim != 2 && dim != 3) || size < 1 || (size & (size - 1)) != 0 || size > (1 << 20)) {
        g_create_error = "mg_create: dim must be 2 or 3 and size a power of two";
        return MG_EINVAL;
    }
    if (real_kind < 0 || real_kind > 2) {
        g_create_error = "mg_create: bad real_kind";
        return MG_EINVAL;
    }
    if (smooth < 0 || smooth > 4096) {
        g_create_error = "mg_create: smooth must be 0 (= the reference's 7) .. 4096";
        return MG_EINVAL;
    }
    mg_ctx *c = new (std::nothrow) mg_ctx();
    if (!c) return MG_ENOMEM;
    int rc = c->init(dim, size, real_kind, smooth, device, rank, nranks);
    if (rc != MG_OK) {
        g_create_error = c->err;
        c->release();
        delete c;
        return rc;
    }
    *out = c;
    return MG_OK;
}

int mg_set_global_option(const char *name, int value)
{
    if (!name) return MG_EINVAL;
    if (std::string(name) == "slab_min_planes") {
        if (value < 8) { g_create_error = "slab_min_planes must be >= 8 (two ghost depths)"; return MG_EINVAL; }
        g_slab_min_planes = value;
        return MG_OK;
    }
    g_create_error = "mg_set_global_option: unknown option";
    return MG_EINVAL;
}

int mg_create(int dim, int size, int real_kind, int smooth, int device, mg_ctx **out)
{
    return create_common(dim, size, real_kind, smooth, device, 0, 1, out);
}

int mg_destroy(mg_ctx *ctx)
{
    if (!ctx) return MG_OK;
    ctx->release();
    delete ctx;
    return MG_OK;
}

#define CTX_OR_FAIL(ctx) \
    if (!(ctx)) return MG_EINVAL; \
    if ((ctx)->activate() != MG_OK) return MG_ECUDA

int mg_set_mode(mg_ctx *ctx, int mode)
{
    CTX_OR_FAIL(ctx);
    if (mode != MG_MODE_FUSED && mode != MG_MODE_REFSEQ) return ctx->fail(MG_EINVAL, "bad mode");
    ctx->mode = mode;
    if (mode == MG_MODE_REFSEQ) return ctx->ensure_debug_arena();
    return MG_OK;
}

int mg_set_stream(mg_ctx *ctx, void *cuda_stream)
{
    CTX_OR_FAIL(ctx);
    if (ctx->group && !ctx->group->nccl && cuda_stream)
        return ctx->fail(MG_EUNSUPPORTED, "mg_set_stream: a slab group in one process runs on its own stream(s)");
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    ctx->borrowed_stream = cuda_stream != nullptr;
    return MG_OK;
}

int mg_set_tuning(mg_ctx *ctx, int tb, int small_L, int use_graph)
{
    CTX_OR_FAIL(ctx);
    if (tb > 4) return ctx->fail(MG_EINVAL, "tb must be 0..4");
    if (tb == 0 && ctx->group) return ctx->fail(MG_EINVAL, "tb must be 1..4 on a slab handle (the slab schedule needs the streaming smoother)");
    if (small_L > (1 << (SMALL_MAX_LEVELS - 1))) return ctx->fail(MG_EINVAL, "small_L too large");
    if (tb >= 0) ctx->tb = tb;
    if (small_L >= 0) ctx->small_L = small_L;
    if (use_graph >= 0) ctx->use_graph = use_graph;
    ctx->drop_graph();
    return MG_OK;
}

int mg_set_option(mg_ctx *ctx, const char *name, int value)
{
    CTX_OR_FAIL(ctx);
    if (!name) return MG_EINVAL;
    std::string n(name);
    if (n == "tb") {
        if (value < 0 || value > 4) return ctx->fail(MG_EINVAL, "tb must be 0..4");
        if (value == 0 && ctx->group) return ctx->fail(MG_EINVAL, "tb must be 1..4 on a slab handle (the slab schedule needs the streaming smoother)");
        ctx->tb = value;
    }
    else if (n == "small_L") { if (value < 1 || value > (1 << (SMALL_MAX_LEVELS - 1))) return ctx->fail(MG_EINVAL, "bad small_L"); ctx->small_L = value; }
    else if (n == "graph") ctx->use_graph = value != 0;
    else if (n == "stream_min_L") ctx->stream_min_L = value < 64 ? 64 : value;
    else if (n == "tz") ctx->tz_override = value;
    else if (n == "ncta") ctx->ncta_override = value < 0 ? 0 : value;
    else if (n == "cluster_L") {   // widest level of the one-cluster kernel (0 = off); it covers at most 256
        if (value < 0 || value > (1 << (SMALL_MAX_LEVELS - 1))) return ctx->fail(MG_EINVAL, "cluster_L must be 0..256");
        ctx->cluster_L = value;
    }
    else if (n == "cluster_ctas") {
        if (value != 0 && value != 1 && value != 2 && value != 4 && value != 8 && value != 16) return ctx->fail(MG_EINVAL, "cluster_ctas must be 0 (auto), 1, 2, 4, 8 or 16");
        ctx->cluster_ctas = value;
    }
    else if (n == "stream_flags") ctx->stream_flags = value;
    else if (n == "slab_p2p") {   // 0: halo planes by ncclSend/ncclRecv (or memcpy), 1: fused peer stores
        for (mg_ctx *m : (ctx->group ? ctx->group->m : std::vector<mg_ctx *>{ctx})) {
            if (value && !(m->peer_lo || m->peer_hi)) return ctx->fail(MG_ESTATE, "slab_p2p: no peer mapped");
            m->p2p = value != 0;
            m->u_ghost_dirty = m->f_ghost_dirty = true;
        }
    }
    else if (n == "lockstep") ctx->lockstep_opt = value;
    else if (n == "block_max_L") { for (mg_ctx *m : (ctx->group ? ctx->group->m : std::vector<mg_ctx *>{ctx})) { m->block_max_L = value; m->drop_graph(); } }
    else if (n == "slab_trace") {   // timeline of the slab passes (mg_slab_trace): 1 = record (and reset), 0 = off
        const size_t bytes = sizeof(unsigned long long) * (8 + 4 * (size_t)S3_TRACE_CAP);
        if (value && !ctx->slab_trace) MG_CK(ctx, cudaMalloc(&ctx->slab_trace, bytes));
        if (value) MG_CK(ctx, cudaMemset(ctx->slab_trace, 0, bytes));
        if (!value && ctx->slab_trace) { MG_CK(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->slab_trace); ctx->slab_trace = nullptr; }
    }
    else if (n == "small_smem") { for (mg_ctx *m : (ctx->group ? ctx->group->m : std::vector<mg_ctx *>{ctx})) { m->small_smem_opt = value; m->drop_graph(); } }
    else if (n == "pdl") { for (mg_ctx *m : (ctx->group ? ctx->group->m : std::vector<mg_ctx *>{ctx})) { m->pdl_opt = value; m->drop_graph(); } }
    else if (n == "colparts") { for (mg_ctx *m : (ctx->group ? ctx->group->m : std::vector<mg_ctx *>{ctx})) m->colparts_opt = value; }
    else if (n == "fastdiv") ctx->fastdiv_opt = value;
    else if (n == "fast_min_L") ctx->fast_min_L = value;
    else if (n == "slab_graph") ctx->slab_graph_opt = value != 0;
    else if (n == "tma_promo") { ctx->tma_promo = value; ctx->tmaps.clear(); }
    else if (n == "tb2") { if (value < 0 || value > 7) return ctx->fail(MG_EINVAL, "tb2 must be 0..7"); ctx->tb2 = value; }
    else if (n == "warp2d_min_L") ctx->warp2d_min_L = value < 32 ? 32 : value;
    else if (n == "ty") {   // rows per work item of the 2-D smoother: 0 = cost model, else even and >= 2
        if (value != 0 && (value < 2 || (value & 1))) return ctx->fail(MG_EINVAL, "ty must be 0 (automatic) or an even number >= 2");
        ctx->ty_override = value;
    }
    else return ctx->fail(MG_EINVAL, "mg_set_option: unknown option");
    ctx->drop_graph();
    return MG_OK;
}

// Weighted Jacobi u + omega * (J(u) - u). NOT in the reference, whose smoother is omega = 1 (cpu-raw.lua:34-44,176-184)
// and does not converge for it; omega = 1 (the default) keeps every result bit-identical to the
// reference path. omega != 1 is a labelled extension for time-to-solution reports: the level visits then run the
// one-sweep-per-launch kernels (the temporally blocked kernels implement the reference smoother only).
int mg_set_omega(mg_ctx *ctx, double omega)
{
    CTX_OR_FAIL(ctx);
    if (!(omega > 0.0 && omega < 2.0)) return ctx->fail(MG_EINVAL, "mg_set_omega: 0 < omega < 2");
    if (ctx->group && omega != 1.0) return ctx->fail(MG_EUNSUPPORTED, "mg_set_omega: slabs run the reference smoother (omega = 1) only");
    ctx->omega = omega;
    ctx->drop_graph();
    return MG_OK;
}

int mg_get_info(mg_ctx *ctx, int *dim, int *size, int *real_kind, int *smooth, int *nlevels,
                uint64_t *arena_bytes)
{
    if (!ctx) return MG_EINVAL;
    if (dim) *dim = ctx->dim;
    if (size) *size = ctx->size;
    if (real_kind) *real_kind = ctx->real_kind;
    if (smooth) *smooth = ctx->smooth;
    if (nlevels) *nlevels = ctx->nlevels;
    if (arena_bytes) *arena_bytes = ctx->arena_bytes + ctx->debug_arena_bytes;
    return MG_OK;
}

int mg_init_cells(mg_ctx *ctx)
{
    CTX_OR_FAIL(ctx);
    int rc = MG_OK;
    if (ctx->group && !ctx->group->nccl) {
        for (mg_ctx *m : ctx->group->m) {
            if ((rc = m->activate())) return rc;
            if ((rc = m->eng->init_cells(m))) return rc;
        }
    } else if ((rc = ctx->eng->init_cells(ctx))) {
        return rc;
    }
    return ctx->sync();
}

int mg_zero_corrections(mg_ctx *ctx)
{
    CTX_OR_FAIL(ctx);
    for (mg_ctx *m : (ctx->group ? ctx->group->m : std::vector<mg_ctx *>{ctx})) {
        if (int rc = m->activate()) return rc;
        for (int lv = 0; lv < m->nlevels; ++lv)
            if (m->V[lv]) MG_CK(ctx, cudaMemsetAsync(m->V[lv], 0, m->level_bytes(lv), m->stream));
    }
    return ctx->sync();
}

void *mg_device_ptr(mg_ctx *ctx, int which, int level)
{
    if (!ctx) return nullptr;
    return ctx->buffer(which, level, nullptr);
}

int mg_upload(mg_ctx *ctx, int which, int level, const void *host, size_t bytes)
{
    CTX_OR_FAIL(ctx);
    if (!ctx->group && (which == MG_BUF_ERRORBUF || which == MG_BUF_TMPU || which == MG_BUF_r || which == MG_BUF_v) &&
        ctx->ensure_debug_arena() != MG_OK)
        return MG_ENOMEM;
    return ctx->copy_in(which, level, host, bytes);
}

int mg_download(mg_ctx *ctx, int which, int level, void *host, size_t bytes)
{
    CTX_OR_FAIL(ctx);
    return ctx->copy_out(which, level, host, bytes);
}

void *mg_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
    return p;
}

int mg_host_free(void *p) { return cudaFreeHost(p) == cudaSuccess ? MG_OK : MG_ECUDA; }

// ------------------------------------------------------------------ hot path
int mg_vcycle_async(mg_ctx *ctx)
{
    CTX_OR_FAIL(ctx);
    return ctx->vcycle();
}

int mg_vcycle(mg_ctx *ctx)
{
    CTX_OR_FAIL(ctx);
    int rc = ctx->vcycle();
    if (rc) return rc;
    return ctx->sync();
}

int mg_synchronize(mg_ctx *ctx)
{
    CTX_OR_FAIL(ctx);
    return ctx->sync();
}

int mg_step(mg_ctx *ctx, double *err)
{
    CTX_OR_FAIL(ctx);
    return ctx->step(err);
}

int mg_run(mg_ctx *ctx, int max_cycles, double accuracy, double *errs, int *n_done)
{
    CTX_OR_FAIL(ctx);
    int it = 0;
    while (it < max_cycles) {
        double err;
        int rc = ctx->step(&err);
        if (rc) return rc;
        if (errs) errs[it] = err;
        ++it;
        if (err < accuracy || !std::isfinite(err)) break;  // cpu-raw.lua:256
    }
    if (n_done) *n_done = it;
    return MG_OK;
}

int mg_step_host(mg_ctx *ctx, const void *f_host, void *psi_host, double *err)
{
    CTX_OR_FAIL(ctx);
    if (!f_host || !psi_host) return ctx->fail(MG_EINVAL, "mg_step_host: null host pointer");
    const int top = ctx->nlevels - 1;
    // bytes the caller's buffers hold: the whole field, or this rank's planes on an NCCL slab
    size_t nb = (ctx->group && ctx->group->nccl ? ctx->own_elems(top) : ctx->N) * ctx->elem;
    int rc;
    if (!ctx->group) {
        MG_CK(ctx, cudaMemcpyAsync(ctx->f, f_host, nb, cudaMemcpyHostToDevice, ctx->stream));
        MG_CK(ctx, cudaMemcpyAsync(ctx->psi, psi_host, nb, cudaMemcpyHostToDevice, ctx->stream));
    } else {
        if ((rc = ctx->copy_in(MG_BUF_F, ctx->size, f_host, nb))) return rc;
        if ((rc = ctx->copy_in(MG_BUF_PSI, ctx->size, psi_host, nb))) return rc;
    }
    double e;
    if ((rc = ctx->step(&e))) return rc;
    if (!ctx->group) MG_CK(ctx, cudaMemcpyAsync(psi_host, ctx->psi, nb, cudaMemcpyDeviceToHost, ctx->stream));
    else if ((rc = ctx->copy_out(MG_BUF_PSI, ctx->size, psi_host, nb))) return rc;
    if (err) *err = e;
    return ctx->sync();
}

// n independent problems held in host memory, each taken through mg_step_host's sequence (upload f and psi; psiOld <- psi;
// one V-cycle; err; download psi) -- but pipelined: problem i+1 is uploaded by one copy engine and problem i-1 downloaded
// by the other while problem i's cycle runs, through two device staging slots per direction. PCIe is full duplex, so a
// batch costs its uploads (2 fields per problem) instead of uploads + cycle + downloads. Results are bit-identical to n
// calls of mg_step_host (same kernels, same order). Handles this covers: single GPU and one-process-per-GPU slabs (every
// rank passes its own planes, all ranks the same n); other groups and the reference-sequence mode run the calls one by one.
int mg_step_host_batch(mg_ctx *ctx, int n, const void **f_hosts, void **psi_hosts, double *errs)
{
    CTX_OR_FAIL(ctx);
    if (n < 0 || (n > 0 && (!f_hosts || !psi_hosts))) return ctx->fail(MG_EINVAL, "mg_step_host_batch: bad argument");
    for (int i = 0; i < n; ++i) {
        if (!f_hosts[i] || !psi_hosts[i]) return ctx->fail(MG_EINVAL, "mg_step_host_batch: null host pointer");
        for (int j = 0; j < i; ++j)   // an output buffer is written while later problems are still being read
            if (psi_hosts[j] == psi_hosts[i]) return ctx->fail(MG_EINVAL, "mg_step_host_batch: psi buffers must be distinct");
    }
    const bool nccl = ctx->group && ctx->group->nccl;
    if ((ctx->group && !nccl) || ctx->mode == MG_MODE_REFSEQ || n < 2) {
        for (int i = 0; i < n; ++i) {
            double e;
            if (int rc = mg_step_host(ctx, f_hosts[i], psi_hosts[i], &e)) return rc;
            if (errs) errs[i] = e;
        }
        return MG_OK;
    }
    const int top = ctx->nlevels - 1;
    const size_t nb = (nccl ? ctx->own_elems(top) : ctx->N) * ctx->elem;
    const size_t off = ctx->own_off_elems(top) * ctx->elem;
    int rc = ctx->pipe_ensure(nb, n);
    if (rc) return rc;
    mg_ctx::HostPipe *p = ctx->pipe;
    cudaStream_t st = ctx->stream;
    auto body = [&]() -> int {
        for (int i = 0; i < n; ++i) {
            const int s = i & 1;
            // copy engine 1: this problem's inputs into staging slot s (free once problem i-2 has left it)
            if (i >= 2) MG_CK(ctx, cudaStreamWaitEvent(p->up, p->in_free[s], 0));
            MG_CK(ctx, cudaMemcpyAsync(p->in_f[s], f_hosts[i], nb, cudaMemcpyHostToDevice, p->up));
            MG_CK(ctx, cudaMemcpyAsync(p->in_u[s], psi_hosts[i], nb, cudaMemcpyHostToDevice, p->up));
            MG_CK(ctx, cudaEventRecord(p->up_done[s], p->up));
            // the cycle's stream: staging -> the solver's fields, one step, result -> staging
            MG_CK(ctx, cudaStreamWaitEvent(st, p->up_done[s], 0));
            MG_CK(ctx, cudaMemcpyAsync((char *)ctx->f + off, p->in_f[s], nb, cudaMemcpyDeviceToDevice, st));
            MG_CK(ctx, cudaMemcpyAsync((char *)ctx->psi + off, p->in_u[s], nb, cudaMemcpyDeviceToDevice, st));
            MG_CK(ctx, cudaEventRecord(p->in_free[s], st));
            if (nccl) ctx->f_ghost_dirty = ctx->u_ghost_dirty = true;
            MG_CK(ctx, cudaMemcpyAsync(ctx->psiOld, ctx->psi, ctx->Ntop * ctx->elem, cudaMemcpyDeviceToDevice, st));
            if (int r = ctx->vcycle()) return r;
            if (int r = ctx->eng->frob_partial_sum(ctx, nullptr)) return r;   // cpu-raw.lua:249-254, sum left on the device
            if (nccl) {
                mg::SlabGroup *g = ctx->group;
                int e = g->api->AllReduce(ctx->d_scalar, ctx->d_scalar, 1, NcclApi::kFloat64, NcclApi::kSum, g->comm, st);
                if (e) return ctx->fail(MG_ECUDA, g->api->GetErrorString(e));
            }
            MG_CK(ctx, cudaMemcpyAsync(p->h_sum + i, ctx->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, st));
            if (i >= 2) MG_CK(ctx, cudaStreamWaitEvent(st, p->out_free[s], 0));
            MG_CK(ctx, cudaMemcpyAsync(p->out_u[s], (char *)ctx->psi + off, nb, cudaMemcpyDeviceToDevice, st));
            MG_CK(ctx, cudaEventRecord(p->comp_done[s], st));
            // copy engine 2: the result back to the caller
            MG_CK(ctx, cudaStreamWaitEvent(p->down, p->comp_done[s], 0));
            MG_CK(ctx, cudaMemcpyAsync(psi_hosts[i], p->out_u[s], nb, cudaMemcpyDeviceToHost, p->down));
            MG_CK(ctx, cudaEventRecord(p->out_free[s], p->down));
        }
        return MG_OK;
    };
    rc = body();
    // drain all three streams whatever happened, so that no copy is in flight when the caller gets its buffers back
    cudaError_t e1 = cudaStreamSynchronize(p->up), e2 = cudaStreamSynchronize(st), e3 = cudaStreamSynchronize(p->down);
    if (rc) return rc;
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)
        return ctx->fail_cuda(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3), "mg_step_host_batch");
    if (errs)
        for (int i = 0; i < n; ++i) errs[i] = std::sqrt(p->h_sum[i] / (double)ctx->N);
    return ctx->sync();
}

int mg_residual_norm(mg_ctx *ctx, double *rms)
{
    CTX_OR_FAIL(ctx);
    if (!rms) return MG_EINVAL;
    return ctx->eng->residual_norm(ctx, rms);
}

// ------------------------------------------------------------------ per-operator entry points
static inline int lv_of(mg_ctx *ctx, int L)
{
    if (L < 1 || (L & (L - 1)) || L > ctx->size) return -1;
    int k = 0;
    while ((1 << k) < L) ++k;
    return k;
}

int mg_twogrid(mg_ctx *ctx, double h, void *u, const void *f, int L)
{
    CTX_OR_FAIL(ctx);
    int lv = lv_of(ctx, L);
    if (lv < 0 || !u || !f) return ctx->fail(MG_EINVAL, "mg_twogrid: bad level or pointer");
    if (ctx->mode == MG_MODE_REFSEQ && ctx->ensure_debug_arena() != MG_OK) return MG_ENOMEM;
    int rc = ctx->mode == MG_MODE_REFSEQ ? ctx->eng->twogrid_refseq(ctx, h, u, f, lv)
                                         : ctx->eng->twogrid_fused(ctx, h, u, f, lv);
    if (rc) return rc;
    return ctx->sync();
}

int mg_smooth(mg_ctx *ctx, int L, void *u, const void *f, double h, int n)
{
    CTX_OR_FAIL(ctx);
    int lv = lv_of(ctx, L);
    if (lv < 0 || !u || !f || n < 0) return ctx->fail(MG_EINVAL, "mg_smooth: bad argument");
    int rc = ctx->eng->smooth(ctx, lv, u, f, h, n);
    if (rc) return rc;
    return ctx->sync();
}

int mg_jacobi(mg_ctx *ctx, int L, void *dest, const void *u, const void *f, double h)
{
    CTX_OR_FAIL(ctx);
    if (lv_of(ctx, L) < 0 || !dest || !u || !f) return ctx->fail(MG_EINVAL, "mg_jacobi: bad argument");
    int rc = ctx->eng->jacobi(ctx, L, dest, u, f, h);
    if (rc) return rc;
    return ctx->sync();
}

int mg_residual(mg_ctx *ctx, int L, void *r, const void *f, const void *u, double h)
{
    CTX_OR_FAIL(ctx);
    if (lv_of(ctx, L) < 0 || !r || !u || !f) return ctx->fail(MG_EINVAL, "mg_residual: bad argument");
    int rc = ctx->eng->residual(ctx, L, r, f, u, h);
    if (rc) return rc;
    return ctx->sync();
}

int mg_restrict(mg_ctx *ctx, int L2, void *R, const void *r)
{
    CTX_OR_FAIL(ctx);
    if (lv_of(ctx, L2) < 0 || 2 * L2 > ctx->size || !R || !r)
        return ctx->fail(MG_EINVAL, "mg_restrict: bad argument");
    int rc = ctx->eng->restrict_(ctx, L2, R, r);
    if (rc) return rc;
    return ctx->sync();
}

int mg_prolong(mg_ctx *ctx, int L2, void *v, const void *V)
{
    CTX_OR_FAIL(ctx);
    if (lv_of(ctx, L2) < 0 || 2 * L2 > ctx->size || !v || !V)
        return ctx->fail(MG_EINVAL, "mg_prolong: bad argument");
    int rc = ctx->eng->prolong(ctx, L2, v, V);
    if (rc) return rc;
    return ctx->sync();
}

int mg_add_to(mg_ctx *ctx, size_t n, void *u, const void *v)
{
    CTX_OR_FAIL(ctx);
    if (!u || !v) return ctx->fail(MG_EINVAL, "mg_add_to: null pointer");
    int rc = ctx->eng->add_to(ctx, n, u, v);
    if (rc) return rc;
    return ctx->sync();
}

int mg_frob_err(mg_ctx *ctx, double *err)
{
    CTX_OR_FAIL(ctx);
    if (!err) return MG_EINVAL;
    return ctx->eng->frob_err(ctx, err, ctx->mode == MG_MODE_REFSEQ);
}

int mg_smooth_residual_restrict(mg_ctx *ctx, int L, void *u, const void *f, double h, int n, void *R)
{
    CTX_OR_FAIL(ctx);
    int lv = lv_of(ctx, L);
    if (lv < 1 || !u || !f || !R || n < 0)
        return ctx->fail(MG_EINVAL, "mg_smooth_residual_restrict: bad argument");
    int rc = ctx->eng->pre_fused(ctx, lv, u, f, h, n, R);
    if (rc) return rc;
    return ctx->sync();
}

int mg_prolong_add_smooth(mg_ctx *ctx, int L, void *u, const void *f, double h, int n, const void *V)
{
    CTX_OR_FAIL(ctx);
    int lv = lv_of(ctx, L);
    if (lv < 1 || !u || !f || !V || n < 0)
        return ctx->fail(MG_EINVAL, "mg_prolong_add_smooth: bad argument");
    int rc = ctx->eng->post_fused(ctx, lv, u, f, h, n, V);
    if (rc) return rc;
    return ctx->sync();
}

// ------------------------------------------------------------------ Krylov comparator
int mg_cg(mg_ctx *ctx, int max_iter, double epsilon, double *err_hist, double *linf_hist, int *n_done)
{
    CTX_OR_FAIL(ctx);
    if (max_iter < 0) return ctx->fail(MG_EINVAL, "mg_cg: max_iter < 0");
    return ctx->eng->cg(ctx, max_iter, epsilon, err_hist, linf_hist, n_done);
}

int mg_linf_norm(mg_ctx *ctx, int which, int level, double *out)
{
    CTX_OR_FAIL(ctx);
    if (!out) return MG_EINVAL;
    if (ctx->group) return ctx->fail(MG_EUNSUPPORTED, "mg_linf_norm: single-GPU only");
    size_t cap = 0;
    void *d = ctx->buffer(which, level, &cap);
    if (!d) return ctx->fail(MG_EINVAL, "mg_linf_norm: no such buffer");
    return ctx->eng->linf_norm(ctx, d, cap / ctx->elem, out);
}

// ------------------------------------------------------------------ trace
int mg_trace_enable(mg_ctx *ctx, int on)
{
    if (!ctx) return MG_EINVAL;
    ctx->trace_on = on != 0;
    return MG_OK;
}
int mg_trace_clear(mg_ctx *ctx)
{
    if (!ctx) return MG_EINVAL;
    ctx->trace.clear();
    return MG_OK;
}
size_t mg_trace_count(mg_ctx *ctx) { return ctx ? ctx->trace.size() : 0; }
int mg_trace_get(mg_ctx *ctx, size_t i, char *name, int *L, const void **host_data, size_t *bytes)
{
    if (!ctx || i >= ctx->trace.size()) return MG_EINVAL;
    const TraceRec &r = ctx->trace[i];
    if (name) *name = r.name;
    if (L) *L = r.L;
    if (host_data) *host_data = r.data.data();
    if (bytes) *bytes = r.data.size();
    return MG_OK;
}

// ------------------------------------------------------------------ measurement
int mg_time_vcycles(mg_ctx *ctx, int n, float *ms_total)
{
    CTX_OR_FAIL(ctx);
    if (n < 1 || !ms_total) return MG_EINVAL;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = MG_OK;
    cudaError_t e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaEventRecord(e0, ctx->stream);
    for (int i = 0; e == cudaSuccess && i < n && rc == MG_OK; ++i) rc = ctx->vcycle();
    if (e == cudaSuccess) e = cudaEventRecord(e1, ctx->stream);
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    if (e == cudaSuccess) e = cudaEventElapsedTime(ms_total, e0, e1);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (e != cudaSuccess) return ctx->fail_cuda(e, "mg_time_vcycles");
    return rc;
}

uint64_t mg_launch_count(mg_ctx *ctx) { return ctx ? ctx->launches : 0; }

// One fused V-cycle, launched kernel by kernel (no graph) with a CUDA-event pair around every
// launch. Fills up to `cap` records {kind, L, sweeps, ms}; *n = number of launches.
int mg_profile_vcycle(mg_ctx *ctx, int cap, int *kind, int *L, int *sweeps, float *ms, int *n)
{
    CTX_OR_FAIL(ctx);
    if (ctx->mode != MG_MODE_FUSED) return ctx->fail(MG_ESTATE, "mg_profile_vcycle: fused mode only");
    ctx->prof.clear();
    ctx->prof_on = true;
    int rc = ctx->group ? ctx->eng->slab_vcycle(ctx->group)
                        : ctx->eng->twogrid_fused(ctx, 1.0 / ctx->size, ctx->psi, ctx->f, ctx->nlevels - 1);
    ctx->prof_on = false;
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    int cnt = 0;
    for (auto &r : ctx->prof) {
        if (e == cudaSuccess) cudaEventElapsedTime(&r.ms, r.e0, r.e1);
        if (cnt < cap) {
            if (kind) kind[cnt] = r.kind;
            if (L) L[cnt] = r.L;
            if (sweeps) sweeps[cnt] = r.sweeps;
            if (ms) ms[cnt] = r.ms;
        }
        ++cnt;
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    ctx->prof.clear();
    if (n) *n = cnt;
    if (rc) return rc;
    if (e != cudaSuccess) return ctx->fail_cuda(e, "mg_profile_vcycle");
    return MG_OK;
}

// ------------------------------------------------------------------ multi-GPU slabs (mg_slab.cuh)
int mg_nccl_unique_id(void *id, size_t bytes)
{
    if (!id || bytes < sizeof(NcclUniqueId)) return MG_EINVAL;
    NcclApi *api = NcclApi::get(g_create_error);
    if (!api) return MG_EUNSUPPORTED;
    NcclUniqueId u;
    int e = api->GetUniqueId(&u);
    if (e) { g_create_error = api->GetErrorString(e); return MG_ECUDA; }
    memcpy(id, &u, sizeof(u));
    return MG_OK;
}

int mg_create_slab(int dim, int size, int real_kind, int smooth, int device, int rank, int nranks,
                   const void *nccl_id, size_t id_bytes, mg_ctx **out)
{
    if (nranks == 1) return create_common(dim, size, real_kind, smooth, device, 0, 1, out);
    if (!out || !nccl_id || id_bytes < sizeof(NcclUniqueId) || rank < 0 || rank >= nranks) return MG_EINVAL;
    NcclApi *api = NcclApi::get(g_create_error);
    if (!api) return MG_EUNSUPPORTED;
    int rc = create_common(dim, size, real_kind, smooth, device, rank, nranks, out);
    if (rc) return rc;
    mg_ctx *c = *out;
    SlabGroup *g = new SlabGroup();
    g->nranks = nranks; g->nccl = true; g->api = api; g->m.push_back(c);
    NcclUniqueId u;
    memcpy(&u, nccl_id, sizeof(u));
    int e = api->CommInitRank(&g->comm, nranks, u, rank);
    if (e) {
        g_create_error = std::string("ncclCommInitRank: ") + api->GetErrorString(e);
        delete g; c->release(); delete c; *out = nullptr;
        return MG_ECUDA;
    }
    c->group = g; c->owns_group = true;
    return MG_OK;
}

// every slab in this process, on one device and one stream: the slab schedule on a single GPU
int mg_create_slab_local(int dim, int size, int real_kind, int smooth, int device, int nslabs, mg_ctx **out)
{
    if (nslabs == 1) return create_common(dim, size, real_kind, smooth, device, 0, 1, out);
    if (!out) return MG_EINVAL;
    SlabGroup *g = new SlabGroup();
    g->nranks = nslabs;
    for (int r = 0; r < nslabs; ++r) {
        mg_ctx *c = nullptr;
        int rc = create_common(dim, size, real_kind, smooth, device, r, nslabs, &c);
        if (rc) {
            for (size_t k = 0; k < g->m.size(); ++k) {
                mg_ctx *m = g->m[k];
                m->group = nullptr;
                if (k > 0) m->own_stream = nullptr;   // shared with member 0, which destroys it
                m->release();
                delete m;
            }
            delete g; *out = nullptr;
            return rc;
        }
        if (r > 0) {  // one stream for the whole group
            cudaStreamDestroy(c->own_stream);
            c->own_stream = g->m[0]->own_stream;
            c->stream = g->m[0]->stream;
        }
        c->group = g;
        g->m.push_back(c);
    }
    g->m[0]->owns_group = true;
    for (int r = 0; r < nslabs; ++r) {  // same process: the neighbours' arenas are directly addressable
        mg_ctx *c = g->m[r];
        c->peer_lo = r > 0 ? (char *)g->m[r - 1]->arena : nullptr;
        c->peer_hi = r < nslabs - 1 ? (char *)g->m[r + 1]->arena : nullptr;
        for (int o = 0; o < nslabs; ++o) c->peer[o] = (char *)g->m[o]->arena;
        c->p2p = true;
    }
    *out = g->m[0];
    return MG_OK;
}

// One process, one GPU per slab (the reference's host is a single LuaJIT process, test/test.lua:53-56): the slabs of
// the group live on `ndev` different devices, each with its own stream; kernels, fused halo stores and handshakes are
// those of the one-process-per-GPU transport, peer pointers come from cudaDeviceEnablePeerAccess.
int mg_create_slab_multi(int dim, int size, int real_kind, int smooth, int ndev, const int *devices, mg_ctx **out)
{
    if (!out) return MG_EINVAL;
    *out = nullptr;
    if (ndev == 1) return create_common(dim, size, real_kind, smooth, devices ? devices[0] : 0, 0, 1, out);
    if (ndev < 1 || ndev > S3_MAX_RANKS) { g_create_error = "mg_create_slab_multi: 1, 2, 4 or 8 devices"; return MG_EINVAL; }
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have < 1) { g_create_error = "no CUDA device: libmgpoisson has no CPU fallback"; return MG_ECUDA; }
    std::vector<int> dev(ndev);
    for (int r = 0; r < ndev; ++r) {
        dev[r] = devices ? devices[r] : r;
        if (dev[r] < 0 || dev[r] >= have) { g_create_error = "mg_create_slab_multi: no such device"; return MG_EINVAL; }
        for (int o = 0; o < r; ++o)
            if (dev[o] == dev[r]) { g_create_error = "mg_create_slab_multi: one slab per device (mg_create_slab_local puts several on one)"; return MG_EINVAL; }
    }
    for (int r = 0; r < ndev; ++r)
        for (int o = 0; o < ndev; ++o) {
            if (o == r) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, dev[r], dev[o]) != cudaSuccess || !can) {
                g_create_error = "mg_create_slab_multi: the devices cannot access each other's memory (no NVLink / P2P)";
                return MG_EUNSUPPORTED;
            }
            cudaSetDevice(dev[r]);
            cudaError_t e = cudaDeviceEnablePeerAccess(dev[o], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                g_create_error = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e);
                return MG_ECUDA;
            }
            cudaGetLastError();
        }
    SlabGroup *g = new SlabGroup();
    g->nranks = ndev; g->multi = true;
    for (int r = 0; r < ndev; ++r) {
        mg_ctx *c = nullptr;
        int rc = create_common(dim, size, real_kind, smooth, dev[r], r, ndev, &c);
        if (rc) {
            for (mg_ctx *m : g->m) { m->group = nullptr; m->release(); delete m; }
            delete g;
            return rc;
        }
        c->group = g;
        g->m.push_back(c);
    }
    g->m[0]->owns_group = true;
    for (int r = 0; r < ndev; ++r) {
        mg_ctx *c = g->m[r];
        c->peer_lo = r > 0 ? (char *)g->m[r - 1]->arena : nullptr;
        c->peer_hi = r < ndev - 1 ? (char *)g->m[r + 1]->arena : nullptr;
        for (int o = 0; o < ndev; ++o) c->peer[o] = (char *)g->m[o]->arena;
        c->p2p = true;
    }
    *out = g->m[0];
    return MG_OK;
}

// CUDA IPC handle of this rank's arena (64 bytes), to be shipped to its two neighbours
int mg_slab_ipc_export(mg_ctx *ctx, void *handle, size_t bytes)
{
    CTX_OR_FAIL(ctx);
    if (!handle || bytes < sizeof(cudaIpcMemHandle_t)) return ctx->fail(MG_EINVAL, "mg_slab_ipc_export: need 64 bytes");
    cudaIpcMemHandle_t h;
    MG_CK(ctx, cudaIpcGetMemHandle(&h, ctx->arena));
    memcpy(handle, &h, sizeof(h));
    return MG_OK;
}

// handles = nranks x 64 bytes, indexed by rank (e.g. the result of an all-gather). Maps the two
// neighbours' arenas and switches the handle to the fused halo exchange. Every rank must have
// attached (barrier) before the next V-cycle.
int mg_slab_ipc_attach(mg_ctx *ctx, const void *handles, size_t bytes)
{
    CTX_OR_FAIL(ctx);
    if (!ctx->group || !ctx->group->nccl) return ctx->fail(MG_ESTATE, "mg_slab_ipc_attach: not a multi-process slab");
    if (!handles || bytes < (size_t)ctx->nranks * sizeof(cudaIpcMemHandle_t)) return ctx->fail(MG_EINVAL, "mg_slab_ipc_attach: short buffer");
    const cudaIpcMemHandle_t *h = (const cudaIpcMemHandle_t *)handles;
    // every rank's arena, not only the neighbours': the RES pass of the last distributed level stores the restricted
    // residual into every rank's copy of the first replicated level, and the all-gather epochs go to every header
    for (int r = 0; r < ctx->nranks; ++r) {
        if (r == ctx->rank) { ctx->peer[r] = (char *)ctx->arena; continue; }
        cudaIpcMemHandle_t hh; memcpy(&hh, &h[r], sizeof(hh));
        MG_CK(ctx, cudaIpcOpenMemHandle((void **)&ctx->peer[r], hh, cudaIpcMemLazyEnablePeerAccess));
    }
    ctx->peer_lo = ctx->rank > 0 ? ctx->peer[ctx->rank - 1] : nullptr;
    ctx->peer_hi = ctx->rank < ctx->nranks - 1 ? ctx->peer[ctx->rank + 1] : nullptr;
    ctx->peer_ipc = true;
    ctx->p2p = true;
    ctx->u_ghost_dirty = ctx->f_ghost_dirty = true;
    return MG_OK;
}

int mg_slab_info(mg_ctx *ctx, int *rank, int *nranks, int *own_planes, int *ghost, uint64_t *exchanges,
                 uint64_t *exchanged_bytes)
{
    if (!ctx) return MG_EINVAL;
    const int top = ctx->nlevels - 1;
    if (rank) *rank = ctx->rank;
    if (nranks) *nranks = ctx->nranks;
    if (own_planes) *own_planes = ctx->dist[top] ? ctx->nzl[top] : ctx->size;
    if (ghost) *ghost = ctx->G;
    if (exchanges) *exchanges = ctx->group ? ctx->group->exchanges : 0;
    if (exchanged_bytes) *exchanged_bytes = ctx->group ? ctx->group->exchanged_bytes : 0;
    return MG_OK;
}

// NVLink traffic so far: bytes this handle's kernels stored straight into other GPUs' memory (fused halo exchange and
// fused all-gather; counted per launch from the planes a pass sends), and bytes moved by explicit exchanges
// (ncclSend/ncclRecv or peer copies: the ghost refresh after initCells / an upload, or every pass with slab_p2p = 0).
int mg_slab_traffic(mg_ctx *ctx, uint64_t *peer_store_bytes, uint64_t *exchange_bytes)
{
    if (!ctx) return MG_EINVAL;
    uint64_t b = 0;
    for (mg_ctx *m : (ctx->group ? ctx->group->m : std::vector<mg_ctx *>{ctx})) b += m->nvl_bytes;
    if (peer_store_bytes) *peer_store_bytes = b;
    if (exchange_bytes) *exchange_bytes = ctx->group ? ctx->group->exchanged_bytes : 0;
    return MG_OK;
}

// Timeline of this rank's slab passes since option "slab_trace" = 1: up to cap records of 4 words each
// {ns on entry, ns after the wait for the lower neighbour, ns CTA 0 waited for the upper neighbour, ns when the last CTA
// finished}, device globaltimer of THIS GPU (differences are meaningful, absolute values are not comparable across GPUs).
int mg_slab_trace(mg_ctx *ctx, uint64_t *records, size_t cap, size_t *n)
{
    CTX_OR_FAIL(ctx);
    if (!ctx->slab_trace) return ctx->fail(MG_ESTATE, "mg_slab_trace: set option slab_trace = 1 first");
    MG_CK(ctx, cudaStreamSynchronize(ctx->stream));
    unsigned long long cnt = 0;
    MG_CK(ctx, cudaMemcpy(&cnt, ctx->slab_trace, sizeof(cnt), cudaMemcpyDeviceToHost));
    size_t m = cnt < S3_TRACE_CAP ? (size_t)cnt : (size_t)S3_TRACE_CAP;
    if (m > cap) m = cap;
    if (records && m) MG_CK(ctx, cudaMemcpy(records, ctx->slab_trace + 8, m * 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (n) *n = m;
    return MG_OK;
}

}  // extern "C"

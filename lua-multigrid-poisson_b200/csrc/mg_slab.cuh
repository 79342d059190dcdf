// mg_slab.cuh -- slab decomposition of the 3-D hierarchy across GPUs.
//
// The reference has no multi-device path (gpu.lua:27-30 picks one device); this is new work
// behind the same V-cycle semantics. The grid is cut along z (the slowest axis): rank r owns
// planes [r*L/P, (r+1)*L/P) of every DISTRIBUTED level (L >= 64 and L/P >= 32 by default). Distributed
// fields carry G = 4 ghost planes on each side; before a smoother pass with NST pipeline stages
// the source field's ghosts are refreshed to depth NST from the two neighbours. Restriction and
// prolongation are communication-free (children 2K, 2K+1 live on the same rank); only ghosts of
// Rs[L/2] (as the next level's right-hand side) and Vs[L/2] (for the fused prolongation) are
// exchanged. Levels below the threshold are REPLICATED: the restricted residual is all-gathered
// and every rank finishes the coarse V-cycle redundantly, so nothing is scattered back
// (cpu-gpu.lua:17-52 hands coarse levels to another executor; here that executor is "everyone").
// Jacobi is order independent, so the sharded result is bit-identical to the single-GPU one.
//
// Three transports behind one schedule:
//   FUSED : (default, one process per GPU) the neighbours' arenas are mapped with CUDA IPC. The
//           smoother kernel that produces one of a rank's 4 boundary planes (or boundary coarse
//           residual planes) stores it straight into the neighbour's ghost planes over NVLink,
//           so there is no separate exchange step at all. The per-pass handshake is inside the
//           kernel too: every CTA acquires the neighbours' "passes done" counters before its
//           first TMA load, the last CTA to finish publishes ours (mg_stream3d.cuh, Stream3DArgs::hs).
//           The kernel that restricts into the first replicated level stores into EVERY rank's copy of it (the
//           all-gather, fused), framed by the epoch kernels below; NCCL is left with the error / residual
//           all-reduce, the ghost refresh after initCells / an upload and the CG comparator: no NCCL call inside
//           a V-cycle, which is replayed from one CUDA graph.
//   NCCL  : (option slab_p2p = 0) ncclSend/ncclRecv of the halo planes before every pass.
//           libnccl is dlopen()ed, so single-GPU users (LuaJIT) do not need it.
//   MULTI : one process, one GPU per slab (mg_create_slab_multi): the same fused kernels and handshakes as FUSED, peer
//           pointers through cudaDeviceEnablePeerAccess instead of CUDA IPC, cudaMemcpyPeerAsync for the ghost refresh
//           and host-side sums instead of NCCL. This is how a single LuaJIT process drives 2, 4 or 8 GPUs.
//   LOCAL : all slabs in one process on one device and one stream (peer pointers are plain
//           pointers; slab_p2p = 0 uses cudaMemcpyAsync). This is how the slab index arithmetic
//           and the fused stores are tested on a single GPU.
// Measured on B200s, 1024^3 fp32 (profiles/r2_bench_n*): 897 / 1663 / 2795 units/s on 2 / 4 / 8 GPUs vs 405-407 on one GPU
// of the same box = efficiency 1.10 / 1.02 / 0.863 (round 1, slower single-GPU kernel: 1733 vs 261 at 8;
// NCCL send/recv then: 1324; fused + separate handshake kernels: 1400). DESIGN.md section 7 has the breakdown.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <string>
#include <vector>

#include "mg_stream3d.cuh"

struct mg_ctx;

namespace mg {

struct NcclUniqueId {
    char internal[128];
};

struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(void **, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;

    static constexpr int kInt8 = 0, kFloat64 = 8, kSum = 0, kMax = 2;  // ncclInt8, ncclFloat64, ncclSum, ncclMax

    static NcclApi *get(std::string &err)
    {
        static NcclApi api;
        if (api.lib) return &api;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names)
            if ((api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
        if (!api.lib) {
            err = "libnccl.so.2 not found (needed only for multi-process slabs)";
            return nullptr;
        }
#define MG_SYM(field, name)                                             \
    *(void **)(&api.field) = dlsym(api.lib, name);                      \
    if (!api.field) {                                                   \
        err = std::string("symbol missing in libnccl: ") + name;        \
        api.lib = nullptr;                                              \
        return nullptr;                                                 \
    }
        MG_SYM(GetUniqueId, "ncclGetUniqueId")
        MG_SYM(CommInitRank, "ncclCommInitRank")
        MG_SYM(CommDestroy, "ncclCommDestroy")
        MG_SYM(GroupStart, "ncclGroupStart")
        MG_SYM(GroupEnd, "ncclGroupEnd")
        MG_SYM(Send, "ncclSend")
        MG_SYM(Recv, "ncclRecv")
        MG_SYM(AllGather, "ncclAllGather")
        MG_SYM(AllReduce, "ncclAllReduce")
        MG_SYM(GetErrorString, "ncclGetErrorString")
#undef MG_SYM
        return &api;
    }
};

// ---- all-gather epochs of the first replicated level (fused transport). The RES pass of the last distributed level
// stores the restricted residual into every rank's copy of the coarse cube (Stream3DArgs::rall). Behind it, on the same
// stream, k_slab_signal_all bumps this rank's epoch and publishes it in every peer's header; k_slab_wait_all, ahead of
// the replicated sub-cycle, waits until every rank has published the current epoch. No NCCL call is left inside a
// V-cycle, so the whole cycle can be one CUDA graph.
struct SlabPeers {
    unsigned long long *hs[S3_MAX_RANKS];   // every rank's arena header (this rank's own included)
    int nranks, rank;
};
__global__ void k_slab_signal_all(SlabPeers p)
{
    unsigned long long *own = p.hs[p.rank];
    const int r = (int)threadIdx.x;
    __threadfence_system();
    const unsigned long long e = own[HS_AG_EPOCH] + 1;   // every thread reads before thread 0 writes (same warp, lock step)
    __syncwarp();
    if (r < p.nranks) s3_st_release_sys(p.hs[r] + HS_AG_FROM + p.rank, e);
    __syncwarp();
    if (r == 0) s3_st_release_sys(own + HS_AG_EPOCH, e);
}
__global__ void k_slab_wait_all(SlabPeers p)
{
    unsigned long long *own = p.hs[p.rank];
    const int r = (int)threadIdx.x;
    const unsigned long long e = s3_ld_acquire_sys(own + HS_AG_EPOCH);
    if (r < p.nranks) s3_wait_counter(own + HS_AG_FROM + r, e, own);
}

// ahead of anything that READS this rank's ghost planes outside a smoother pass (the residual norm): both neighbours
// must have completed as many passes as this rank (their last pass is what fills our ghosts)
__global__ void k_slab_wait_neighbours(unsigned long long *hs, int has_lo, int has_hi)
{
    const unsigned long long n = s3_ld_acquire_sys(hs + HS_DONE);
    if (threadIdx.x == 0 && has_lo) s3_wait_counter(hs + HS_FROM_LO, n, hs);
    if (threadIdx.x == 1 && has_hi) s3_wait_counter(hs + HS_FROM_HI, n, hs);
}

struct SlabGroup {
    std::vector<mg_ctx *> m;  // LOCAL / MULTI: every rank's context; NCCL: this rank's only
    int nranks = 1;
    bool nccl = false;        // one process per GPU: NCCL communicator for the ghost refresh after an upload and the reductions
    bool multi = false;       // one process, one GPU per slab: peer copies and host-side sums instead
    bool concurrent() const { return nccl || multi; }   // slabs run at the same time on different GPUs: in-kernel handshakes
    void *comm = nullptr;
    NcclApi *api = nullptr;
    uint64_t exchanges = 0, exchanged_bytes = 0;
};

}  // namespace mg

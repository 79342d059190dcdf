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
//           NCCL is left with the all-gather of the first replicated level, the error all-reduce
//           and the ghost refresh after initCells / an upload.
//   NCCL  : (option slab_p2p = 0) ncclSend/ncclRecv of the halo planes before every pass.
//           libnccl is dlopen()ed, so single-GPU users (LuaJIT) do not need it.
//   LOCAL : all slabs in one process on one device and one stream (peer pointers are plain
//           pointers; slab_p2p = 0 uses cudaMemcpyAsync). This is how the slab index arithmetic
//           and the fused stores are tested on a single GPU.
// Measured on 8 x B200, 1024^3 fp32 (profiles/): 1733 units/s vs 261 on one GPU of the same box
// = 83 % parallel efficiency (NCCL send/recv: 1324 = 60 %; fused + separate handshake kernels: 68 %).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <string>
#include <vector>

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

    static constexpr int kInt8 = 0, kFloat64 = 8, kSum = 0;  // ncclInt8, ncclFloat64, ncclSum

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

struct SlabGroup {
    std::vector<mg_ctx *> m;  // LOCAL: every rank's context; NCCL: this rank's only
    int nranks = 1;
    bool nccl = false;
    void *comm = nullptr;
    NcclApi *api = nullptr;
    uint64_t exchanges = 0, exchanged_bytes = 0;
};

}  // namespace mg

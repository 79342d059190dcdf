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
// Two transports behind one schedule:
//   NCCL  : one process per GPU (torchrun), ncclSend/ncclRecv halo planes over NVLink,
//           ncclAllGather for the replicated level, ncclAllReduce for the error sum.
//           libnccl is dlopen()ed, so single-GPU users (LuaJIT) do not need it.
//   LOCAL : all slabs in one process on one device, exchanged with cudaMemcpyAsync on one
//           stream. This is how the slab index arithmetic is tested on a single GPU.
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

// s3d_bench.cu -- stand-alone harness for the streaming 3-D smoother (mg_stream3d.cuh), fp32.
// Builds in ~30 s (one TU, four kernel instantiations) instead of the 2 min of the whole
// library, so that kernel variants (-D flags) can be compared on one GPU box in one call:
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr \
//        -I lua-multigrid-poisson_b200/csrc [-DMG_F32_VX=2 ...] -o gpurun_out/s3d_NAME tools/s3d_bench.cu -lcuda
//   build/bin/s3d_NAME [L=512] [reps=10] [zsplit: -1 auto, 0 balanced, n] [promo 0..3] [check 0/1] [flags] [mode]
//
// For each of the four passes of a level visit ([4], [3+RES], [PRO+4], [3]) it prints the median
// CUDA-event time and checks the result BIT FOR BIT against the one-sweep kernels
// (mg_fused_simple.cuh), which share mg_math.cuh with everything else.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "mg_fused_simple.cuh"
#include "mg_stream3d.cuh"

#ifndef MG_TILE_X
#define MG_TILE_X 56
#endif
#ifndef MG_TILE_Y
#define MG_TILE_Y 40
#endif

using namespace mg;

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); exit(2); } \
    } while (0)

static CUtensorMap make_map(const float *base, int L, int box_x, int box_y, int promo)
{
    CUtensorMap m;
    cuuint64_t gdim[3] = {(cuuint64_t)L, (cuuint64_t)L, (cuuint64_t)L};
    cuuint64_t gstr[2] = {(cuuint64_t)L * 4, (cuuint64_t)L * L * 4};
    cuuint32_t box[3] = {(cuuint32_t)box_x, (cuuint32_t)box_y, 1};
    cuuint32_t est[3] = {1, 1, 1};
    CUtensorMapL2promotion pr = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                : promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(base), gdim, gstr, box, est,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, pr,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "cuTensorMapEncodeTiled: %d\n", (int)r); exit(2); }
    return m;
}

__global__ void k_fill(float *p, size_t n, unsigned seed, float scale)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned x = (unsigned)i * 2654435761u + seed;
        x ^= x >> 15; x *= 2246822519u; x ^= x >> 13; x *= 3266489917u; x ^= x >> 16;
        p[i] = ((float)(x >> 8) * (1.0f / 8388608.0f) - 1.0f) * scale;
    }
}
__global__ void k_diff(const unsigned *a, const unsigned *b, size_t n, unsigned long long *cnt)
{
    unsigned long long c = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) c += a[i] != b[i];
    if (c) atomicAdd(cnt, c);
}

static int g_zsplit = -1, g_promo = 3, g_nsm = 148, g_flags = 0, g_mode = 1;
static unsigned int *g_redo = nullptr;

// balance of the lock-step partition: columns below the split cost zs + OV steps, the helper CTAs
// share the rest (each chunk pays OV again)
static int pick_zsplit(int L, int ntiles, int ncta, int OV)
{
    if (ntiles >= ncta || ntiles * 10 < ncta * 7) return 0;
    const int nh = ncta - ntiles;
    int best = 0; double bestc = 1e30;
    for (int zs = L; zs >= L / 2; zs -= 2) {
        const double cl = zs + OV;
        const double planes = (double)ntiles * (L - zs) / nh;
        const double chunks = zs == L ? 0 : std::max(1.0, (double)ntiles / nh + 1.0);
        const double ch = planes + chunks * OV;
        const double c = std::max(cl, ch);
        if (c < bestc) { bestc = c; best = zs; }
    }
    return best == L ? 0 : best;
}

template <int S, bool PRO, bool RES>
static float run_pass(int L, float *dst, const float *src, const float *f, const float *Vp, float *Rout, int reps, int *zs_used)
{
    constexpr int TX = MG_TILE_X, TY = MG_TILE_Y;
    typedef Stream3DCfg<float, S, RES, TX, TY> C;
    CUtensorMap map = make_map(src, L, C::WX, C::WY, g_promo), fmap = make_map(f, L, C::WX, C::WY, g_promo);
    auto kern = g_mode ? k_stream3d<float, float, S, PRO, RES, TX, TY, S3_FAST> : k_stream3d<float, float, S, PRO, RES, TX, TY, S3_GUARDED>;
    auto kern2 = k_stream3d<float, float, S, PRO, RES, TX, TY, S3_RERUN>;
    for (auto k : {kern, kern2}) {
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        CK(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    int occ = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::NTHREADS, C::SMEM_BYTES));
    if (occ < 1) { fprintf(stderr, "kernel does not fit an SM (threads %d smem %d)\n", C::NTHREADS, C::SMEM_BYTES); exit(2); }
    const int ntiles = ((L + TX - 1) / TX) * ((L + TY - 1) / TY);
    long ncta = (long)g_nsm * occ;
    const long work = (long)ntiles * L;
    if (ncta > (work + 15) / 16) ncta = (work + 15) / 16;
    int zs = g_zsplit < 0 ? pick_zsplit(L, ntiles, (int)ncta, 3 * C::H - 1) : g_zsplit;
    if (zs > 0 && (ntiles >= ncta || zs >= L)) zs = 0;
    zs &= ~1;
    *zs_used = zs;
    int ncol = 0, zcol = 0, rem_cta0 = 0, rem_tile0 = 0, rem_z0 = 0;
    if (zs > 0) { ncol = ntiles; zcol = zs; rem_cta0 = ntiles; rem_z0 = zs; }
    else if (g_zsplit != 0 && ntiles >= ncta) { ncol = (int)(ntiles / ncta * ncta); zcol = L; rem_tile0 = ncol; *zs_used = -ncol; }
    Stream3DArgs<float> a{dst, Vp, Rout, L, g_flags, 0, L, 0, L, 0, 0, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr, ncol, zcol, rem_cta0, rem_tile0, rem_z0, g_redo, f};
    const Coef<float> cf = make_coef<float>(3, 1.0 / L);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    std::vector<float> ms;
    for (int r = 0; r < reps + 2; ++r) {
        CK(cudaEventRecord(e0));
        kern<<<(unsigned)ncta, C::NTHREADS, C::SMEM_BYTES>>>(map, fmap, a, cf);
        if (g_mode) kern2<<<(unsigned)ncta, C::NTHREADS, C::SMEM_BYTES>>>(map, fmap, a, cf);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float t; CK(cudaEventElapsedTime(&t, e0, e1));
        if (r >= 2) ms.push_back(t);
    }
    std::sort(ms.begin(), ms.end());
    return ms[ms.size() / 2];
}

static unsigned long long diff(const float *a, const float *b, size_t n, unsigned long long *d_cnt)
{
    CK(cudaMemset(d_cnt, 0, 8));
    k_diff<<<1184, 256>>>((const unsigned *)a, (const unsigned *)b, n, d_cnt);
    unsigned long long h;
    CK(cudaMemcpy(&h, d_cnt, 8, cudaMemcpyDeviceToHost));
    return h;
}

int main(int argc, char **argv)
{
    const int L = argc > 1 ? atoi(argv[1]) : 512;
    const int reps = argc > 2 ? atoi(argv[2]) : 10;
    g_zsplit = argc > 3 ? atoi(argv[3]) : -1;
    g_promo = argc > 4 ? atoi(argv[4]) : 3;
    const int check = argc > 5 ? atoi(argv[5]) : 1;
    g_flags = argc > 6 ? atoi(argv[6]) : 0;   // Stream3DArgs::flags (4 = force the guarded re-run)
    g_mode = argc > 7 ? atoi(argv[7]) : 1;    // 1 = fast + re-run kernels, 0 = the guarded kernel alone
    CK(cudaSetDevice(0));
    CK(cudaFree(0));
    CK(cudaDeviceGetAttribute(&g_nsm, cudaDevAttrMultiProcessorCount, 0));
    const size_t n = (size_t)L * L * L, n2 = n / 8;
    float *u, *f, *w, *ref, *ref2, *V, *Rc, *Rref;
    unsigned long long *d_cnt;
    CK(cudaMalloc(&u, n * 4)); CK(cudaMalloc(&f, n * 4)); CK(cudaMalloc(&w, n * 4));
    CK(cudaMalloc(&ref, n * 4)); CK(cudaMalloc(&ref2, n * 4));
    CK(cudaMalloc(&V, n2 * 4)); CK(cudaMalloc(&Rc, n2 * 4)); CK(cudaMalloc(&Rref, n2 * 4));
    CK(cudaMalloc(&d_cnt, 8));
    CK(cudaMalloc(&g_redo, 8)); CK(cudaMemset(g_redo, 0, 8));
    k_fill<<<1184, 256>>>(u, n, 1u, 1.0f);
    k_fill<<<1184, 256>>>(f, n, 2u, (float)L * L);
    k_fill<<<1184, 256>>>(V, n2, 3u, 1.0f);
    CK(cudaDeviceSynchronize());
    const Coef<float> cf = make_coef<float>(3, 1.0 / L);
    dim3 b(128, 2, 1), g((L + 127) / 128, (L + 1) / 2, L), g2((L / 2 + 127) / 128, (L / 2 + 1) / 2, L / 2);
    // reference: S sweeps with the one-sweep kernel, ping-pong ref <-> ref2; result pointer returned
    auto ref_sweeps = [&](int S, bool pro) -> float * {
        const float *src = u; float *dst = ref;
        for (int s = 0; s < S; ++s) {
            if (s == 0 && pro) k_sweep_pp<float, float, 3, true><<<g, b>>>(dst, src, f, V, L, cf);
            else k_sweep_pp<float, float, 3, false><<<g, b>>>(dst, src, f, nullptr, L, cf);
            src = dst; dst = dst == ref ? ref2 : ref;
        }
        CK(cudaDeviceSynchronize());
        return const_cast<float *>(src);
    };
    int zs = 0;
    double total = 0;
    unsigned long long bad = 0;
    printf("{\"L\": %d, \"tile\": [%d, %d], \"vx\": %d, \"threads\": %d, \"promo\": %d, \"flags\": %d, \"mode\": %d", L, MG_TILE_X, MG_TILE_Y, MG_F32_VX,
           Stream3DCfg<float, 4, false, MG_TILE_X, MG_TILE_Y>::NTHREADS, g_promo, g_flags, g_mode);
    {
        float t = run_pass<4, false, false>(L, w, u, f, nullptr, nullptr, reps, &zs);
        total += t;
        printf(", \"zsplit\": %d, \"s4_ms\": %.4f", zs, t);
        if (check) { float *r = ref_sweeps(4, false); unsigned long long d = diff(w, r, n, d_cnt); bad += d; printf(", \"s4_bad\": %llu", d); }
    }
    {
        CK(cudaMemset(Rc, 0, n2 * 4));
        float t = run_pass<3, false, true>(L, w, u, f, nullptr, Rc, reps, &zs);
        total += t;
        printf(", \"s3res_ms\": %.4f", t);
        if (check) {
            float *r = ref_sweeps(3, false);
            unsigned long long d = diff(w, r, n, d_cnt);
            k_residual_restrict<float, float, 3><<<g2, b>>>(Rref, f, r, L, cf);
            CK(cudaDeviceSynchronize());
            unsigned long long d2 = diff(Rc, Rref, n2, d_cnt);
            bad += d + d2;
            printf(", \"s3res_bad\": [%llu, %llu]", d, d2);
        }
    }
    {
        float t = run_pass<4, true, false>(L, w, u, f, V, nullptr, reps, &zs);
        total += t;
        printf(", \"pro4_ms\": %.4f", t);
        if (check) { float *r = ref_sweeps(4, true); unsigned long long d = diff(w, r, n, d_cnt); bad += d; printf(", \"pro4_bad\": %llu", d); }
    }
    {
        float t = run_pass<3, false, false>(L, w, u, f, nullptr, nullptr, reps, &zs);
        total += t;
        printf(", \"s3_ms\": %.4f", t);
        if (check) { float *r = ref_sweeps(3, false); unsigned long long d = diff(w, r, n, d_cnt); bad += d; printf(", \"s3_bad\": %llu", d); }
    }
    {   // alternative post-smoothing plan [PRO+3][4] (not part of level_visit_ms)
        float t = run_pass<3, true, false>(L, w, u, f, V, nullptr, reps, &zs);
        printf(", \"pro3_ms\": %.4f", t);
        if (check) { float *r = ref_sweeps(3, true); unsigned long long d = diff(w, r, n, d_cnt); bad += d; printf(", \"pro3_bad\": %llu", d); }
    }
    printf(", \"level_visit_ms\": %.4f, \"bad\": %llu}\n", total, bad);
    return bad ? 1 : 0;
}

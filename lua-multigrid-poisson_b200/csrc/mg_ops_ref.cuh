// mg_ops_ref.cuh -- one kernel per reference operator (MG_MODE_REFSEQ and the per-operator
// C-ABI entry points). One thread per cell, x fastest => coalesced rows; neighbour reuse is
// left to L1/L2. These are the parity anchors, not the fast path (see mg_stream3d.cuh,
// mg_warp2d.cuh, mg_small.cuh for that).
#pragma once
#include "mg_math.cuh"

namespace mg {

// cpu-raw.lua:8-20 initCells / gpu.lua:41-59 init
template <typename R, typename A, int DIM>
__global__ void k_init_cells(R *__restrict__ f, R *__restrict__ psi, int L, int plane0, int kglobal0)
{
    // slab view: local plane plane0 + blockIdx.z holds global plane kglobal0 + blockIdx.z
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    int kl = DIM == 3 ? plane0 + (int)blockIdx.z : 0;
    int k = DIM == 3 ? kglobal0 + (int)blockIdx.z : 0;
    if (i >= L || j >= L) return;
    size_t idx = (size_t)i + (size_t)L * ((size_t)j + (size_t)L * kl);
    int center = L / 2;
    A value = (A)0;
    if (i == center && j == center && (DIM == 2 || k == center)) value = -(A)1e+6 / (A)1;
    R fv = (R)value;
    f[idx] = fv;
    psi[idx] = (R)(-(A)fv);
}

// cpu-raw.lua:34-44 Jacobi / gpu.lua:83-102
template <typename R, typename A, int DIM>
__global__ void k_jacobi(R *__restrict__ dest, const R *__restrict__ u, const R *__restrict__ f,
                         int L, Coef<A> c)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    int k = DIM == 3 ? blockIdx.z : 0;
    if (i >= L || j >= L) return;
    size_t idx = (size_t)i + (size_t)L * ((size_t)j + (size_t)L * k);
    A S = stencil_sum<DIM, R, A>(u, i, j, k, L, idx);
    dest[idx] = (R)relax<A>(jacobi_point<DIM, A>(S, (A)f[idx], c), (A)u[idx], c);
}

// cpu-raw.lua:46-57 calcResidual / gpu.lua:104-124
template <typename R, typename A, int DIM>
__global__ void k_residual(R *__restrict__ r, const R *__restrict__ f, const R *__restrict__ u,
                           int L, Coef<A> c)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    int k = DIM == 3 ? blockIdx.z : 0;
    if (i >= L || j >= L) return;
    size_t idx = (size_t)i + (size_t)L * ((size_t)j + (size_t)L * k);
    A S = stencil_sum<DIM, R, A>(u, i, j, k, L, idx);
    r[idx] = (R)residual_point<A>(S, (A)f[idx], (A)u[idx], c);
}

// calcResidual on the planes [plane0, plane0 + gridDim.z) of a z-slab (3-D): the z-neighbours are read from the
// planes above and below unconditionally -- ghost planes hold the neighbouring rank's values, or +0 outside the grid,
// which is the reference's "a neighbour outside the grid reads 0" (cpu-raw.lua:36-39)
template <typename R, typename A>
__global__ void k_residual_slab(R *__restrict__ r, const R *__restrict__ f, const R *__restrict__ u, int L, int plane0, Coef<A> c)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= L || j >= L) return;
    const size_t sL = (size_t)L, sLL = sL * sL;
    const size_t idx = (size_t)i + sL * j + sLL * (size_t)(plane0 + (int)blockIdx.z);
    A xl = i > 0 ? (A)u[idx - 1] : (A)0, xr = i < L - 1 ? (A)u[idx + 1] : (A)0;
    A yl = j > 0 ? (A)u[idx - sL] : (A)0, yr = j < L - 1 ? (A)u[idx + sL] : (A)0;
    A S = Ar<A>::add(Ar<A>::add(Ar<A>::add(xl, xr), yl), yr);
    S = Ar<A>::add(Ar<A>::add(S, (A)u[idx - sLL]), (A)u[idx + sLL]);
    r[idx] = (R)residual_point<A>(S, (A)f[idx], (A)u[idx], c);
}

// sum of the 2^DIM children in the reference's order, times 2^-DIM (cpu-raw.lua:62)
template <typename R, typename A, int DIM>
__device__ __forceinline__ A restrict_children(const R *r, size_t srci, size_t sL, size_t sLL)
{
    A s = Ar<A>::add((A)r[srci], (A)r[srci + 1]);
    s = Ar<A>::add(s, (A)r[srci + sL]);
    s = Ar<A>::add(s, (A)r[srci + sL + 1]);
    if (DIM == 3) {
        s = Ar<A>::add(s, (A)r[srci + sLL]);
        s = Ar<A>::add(s, (A)r[srci + sLL + 1]);
        s = Ar<A>::add(s, (A)r[srci + sLL + sL]);
        s = Ar<A>::add(s, (A)r[srci + sLL + sL + 1]);
        return Ar<A>::mul((A).125, s);
    }
    return Ar<A>::mul((A).25, s);
}

// cpu-raw.lua:59-63 reduceResidual / gpu.lua:126-137
template <typename R, typename A, int DIM>
__global__ void k_restrict(R *__restrict__ Rc, const R *__restrict__ r, int L2)
{
    int I = blockIdx.x * blockDim.x + threadIdx.x;
    int J = blockIdx.y * blockDim.y + threadIdx.y;
    int K = DIM == 3 ? blockIdx.z : 0;
    if (I >= L2 || J >= L2) return;
    const size_t sL = (size_t)L2 * 2, sLL = sL * sL;
    size_t srci = ((size_t)I << 1) + sL * ((size_t)J << 1) + sLL * ((size_t)K << 1);
    Rc[(size_t)I + (size_t)L2 * ((size_t)J + (size_t)L2 * K)] =
        (R)restrict_children<R, A, DIM>(r, srci, sL, sLL);
}

// cpu-raw.lua:65-73 expandResidual / gpu.lua:139-161 (fine-grid sized launch: the commented
// "L-sized kernel" variant cpu-raw.lua:75-80 computes the same field)
template <typename R, int DIM>
__global__ void k_prolong(R *__restrict__ v, const R *__restrict__ V, int L)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    int k = DIM == 3 ? blockIdx.z : 0;
    if (i >= L || j >= L) return;
    const int L2 = L >> 1;
    v[(size_t)i + (size_t)L * ((size_t)j + (size_t)L * k)] =
        V[(size_t)(i >> 1) + (size_t)L2 * ((size_t)(j >> 1) + (size_t)L2 * (k >> 1))];
}

// cpu-raw.lua:83-85 addTo / gpu.lua:163-171
template <typename R, typename A>
__global__ void k_add_to(R *__restrict__ u, const R *__restrict__ v, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) u[i] = (R)Ar<A>::add((A)u[i], (A)v[i]);
}

// ---------------------------------------------------------------------------------------
// Reductions (K-e). Deterministic: every block writes one partial, a single-block second
// pass sums the partials in index order in double. Warp-shuffle tree inside the block.
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double block_sum(double v)
{
    __shared__ double ws[32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) ws[w] = v;
    __syncthreads();
    if (w == 0) {
        int nw = (blockDim.x + 31) >> 5;
        v = lane < nw ? ws[lane] : 0.0;
        v = warp_sum(v);
    }
    return v;  // valid in thread 0
}

// cpu-raw.lua:96-100 calcFrobErr (+ optional materialisation of errorBuf) and the partial
// sums of cpu-raw.lua:250-253. (psi-psiOld)^2 is rounded to the storage type before it is
// accumulated in double, as the reference stores it into errorBuf first.
template <typename R, typename A>
__global__ void k_frob_partial(const R *__restrict__ psi, const R *__restrict__ psiOld,
                               R *__restrict__ errorBuf, size_t n, double *__restrict__ partial)
{
    double acc = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        A d = Ar<A>::sub((A)psi[i], (A)psiOld[i]);
        R e = (R)Ar<A>::mul(d, d);
        if (errorBuf) errorBuf[i] = e;
        acc += (double)e;
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// sum of squares of a field (true residual norm)
template <typename R>
__global__ void k_sumsq_partial(const R *__restrict__ r, size_t n, double *__restrict__ partial)
{
    double acc = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        double d = (double)r[i];
        acc += d * d;
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

__global__ void k_final_sum(const double *__restrict__ partial, int np, double *__restrict__ out)
{
    double acc = 0.0;
    for (int i = threadIdx.x; i < np; i += blockDim.x) acc += partial[i];
    acc = block_sum(acc);
    if (threadIdx.x == 0) *out = acc;
}

}  // namespace mg

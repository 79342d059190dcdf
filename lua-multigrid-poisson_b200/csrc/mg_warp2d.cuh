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
// the y+1 neighbour completes the pending sum `acc` one step later. No shared memory, no
// __syncthreads(), no intermediate field ever touches L2/HBM. The summation order
// ((xl+xr)+yl)+yr of the reference is preserved => bit-identical to one sweep per launch.
//
// Columns/rows outside the grid stay exactly 0 at every stage (Dirichlet rule,
// cpu-raw.lua:36-39). Lanes near the strip edge compute garbage that never reaches the
// interior TXU = 128 - 2*HX columns the warp stores.
#pragma once
#include <cuda_runtime.h>

#include "mg_math.cuh"

namespace mg {

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
};

template <typename R, typename A, int S, bool PRO, bool RES>
__global__ void __launch_bounds__(128)
k_warp2d(R *__restrict__ dst, const R *__restrict__ src, const R *__restrict__ f, const R *__restrict__ Vp,
         R *__restrict__ Rout, int L, int TY, int nstrips, int nitems, Coef<A> cf)
{
    typedef Warp2DCfg<S, RES> C;
    constexpr int NST = C::NST, H = C::H;
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

    A acc[NST][4], prev[NST][4];
#pragma unroll
    for (int s = 0; s < NST; ++s)
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc[s][i] = (A)0; prev[s][i] = (A)0; }
    A rpart[2] = {(A)0, (A)0};

    for (int t = 0; t < nin; ++t) {
        const int q = yb + t;
        R row[4] = {(R)0, (R)0, (R)0, (R)0};
        if (xin && q >= 0 && q < L) {
            load4<R>(src + (size_t)gx0 + sL * (size_t)q, row);
            if (PRO) {
                const R *vp = Vp + (size_t)(gx0 >> 1) + (size_t)L2 * (size_t)(q >> 1);
                const R v0 = vp[0], v1 = vp[1];
                row[0] = (R)Ar<A>::add((A)row[0], (A)v0);
                row[1] = (R)Ar<A>::add((A)row[1], (A)v0);
                row[2] = (R)Ar<A>::add((A)row[2], (A)v1);
                row[3] = (R)Ar<A>::add((A)row[3], (A)v1);
            }
        }
#pragma unroll
        for (int sidx = 0; sidx < NST; ++sidx) {
            const int s = sidx + 1;
            if (t < 2 * sidx) break;           // stage not fed yet (warp-uniform)
            const bool emit = t >= 2 * s;
            const int p = q - s;                 // row this stage completes now
            const bool pin = p >= 0 && p < L;
            const bool is_res = RES && s == NST;
            R fv[4] = {(R)0, (R)0, (R)0, (R)0};
            if (emit && pin && xin) load4<R>(f + (size_t)gx0 + sL * (size_t)p, fv);
            const R lft = shfl_up1(row[3]), rgt = shfl_dn1(row[0]);
            R outv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const A xl = (A)(i == 0 ? lft : row[i - 1]), xr = (A)(i == 3 ? rgt : row[i + 1]);
                const A c = (A)row[i];
                const A tot = Ar<A>::add(acc[sidx][i], c);                       // pending row gets its y+1
                const A o = is_res ? residual_point<A>(tot, (A)fv[i], prev[sidx][i], cf)
                                   : jacobi_point<2, A>(tot, (A)fv[i], cf);
                outv[i] = (xin && pin) ? (R)o : (R)0;
                acc[sidx][i] = Ar<A>::add(Ar<A>::add(xl, xr), prev[sidx][i]);    // (xl+xr)+yl of this row
                prev[sidx][i] = c;
            }
            if (!emit) break;
            if (!is_res) {
                if (s == S && p >= y0 && p < y1 && xst) store4<R>(dst + (size_t)gx0 + sL * (size_t)p, outv);
#pragma unroll
                for (int i = 0; i < 4; ++i) row[i] = outv[i];
            } else if (p >= y0 && p < y1) {
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
    }
}

}  // namespace mg

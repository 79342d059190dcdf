// mg_block3d.cuh -- (K-e) temporally blocked Jacobi passes for the MID levels of a 3-D cycle (32^3, 64^3, optionally
// 128^3): a CTA loads a block of the level plus a halo of S cells into shared memory, performs S sweeps there (the
// valid region shrinks by one cell per sweep) and stores the block. One launch instead of S.
//
// Why: these levels hold < 2 % of the points, are L2 resident, and were 30 one-sweep launches per V-cycle
// (k_sweep_pp): a kernel boundary inside the cycle's CUDA graph costs ~2.7 us (measured with the kernels' own
// globaltimer stamps, bench.py slab_timeline) against ~1.5 us of work per sweep. The streaming kernel does not pay
// here (its z pipeline needs 3 S - 1 fill/drain steps per chunk and a 56 x 40 tile wastes half of a 64^2 plane).
// Redundant work in the halo is (1 + 2S/B)^3 -- irrelevant at this size.
//
// Semantics: cpu-raw.lua:34-44 (Jacobi, a neighbour outside the grid reads 0), :65-73,83-85 (PRO: the source is
// u + prolong(V), rounded to storage like addTo). Cells outside the grid are stored as +0 and never updated. All
// arithmetic from mg_math.cuh, summation order ((((xl+xr)+yl)+yr)+zl)+zr: bit-identical to one sweep per launch.
#pragma once
#include "mg_fused_simple.cuh"
#include "mg_math.cuh"

namespace mg {

template <int S, int BX, int BY, int BZ> struct Block3DCfg {
    static constexpr int EX = BX + 2 * S, EY = BY + 2 * S, EZ = BZ + 2 * S;
    static constexpr int NE = EX * EY * EZ, NEP = (NE + 3) / 4 * 4;
    static constexpr int NB = BX * BY * BZ;
    static constexpr int NTHREADS = NB >= 2048 ? 512 : 256;
    template <typename R> static constexpr int smem_bytes() { return 3 * NEP * (int)sizeof(R); }
};

template <typename R, typename A, int S, bool PRO, int BX, int BY, int BZ>
__global__ void __launch_bounds__((Block3DCfg<S, BX, BY, BZ>::NTHREADS))
k_block3d(R *__restrict__ dst, const R *__restrict__ src, const R *__restrict__ f, const R *__restrict__ Vp, int L, Coef<A> c)
{
    pdl_enter();
    typedef Block3DCfg<S, BX, BY, BZ> C;
    constexpr int EX = C::EX, EY = C::EY, EZ = C::EZ, NE = C::NE, NT = C::NTHREADS;
    extern __shared__ __align__(16) unsigned char block3d_smem_raw[];
    R *bufA = reinterpret_cast<R *>(block3d_smem_raw), *bufB = bufA + C::NEP, *const bufF = bufB + C::NEP;
    const int tid = (int)threadIdx.x;
    const int nbx = L / BX, nby = L / BY;
    const int b = (int)blockIdx.x, bx = b % nbx, by = (b / nbx) % nby, bz = b / (nbx * nby);
    const int gx0 = bx * BX - S, gy0 = by * BY - S, gz0 = bz * BZ - S;   // grid coordinates of extended cell (0, 0, 0)
    const size_t sL = (size_t)L, sLL = sL * sL;
    // ---- in: the block and its halo; outside the grid +0 in both buffers (those cells are never written again).
    // A thread requests CH cells' worth of global loads before it stores the first of them (a store that waits for its
    // load would otherwise serialise the round trips: 18 cells per thread).
    constexpr int CH = sizeof(R) == 4 ? 6 : 3;
    for (int e0 = tid; e0 < NE; e0 += CH * NT) {
        R u[CH], fv[CH];
#pragma unroll
        for (int q = 0; q < CH; ++q) {
            const int e = e0 + q * NT;
            const int i = e % EX, j = (e / EX) % EY, k = e / (EX * EY);
            const int gx = gx0 + i, gy = gy0 + j, gz = gz0 + k;
            const bool in = e < NE && (unsigned)gx < (unsigned)L && (unsigned)gy < (unsigned)L && (unsigned)gz < (unsigned)L;
            u[q] = (R)0; fv[q] = (R)0;
            if (in) {
                const size_t idx = (size_t)gx + sL * (size_t)gy + sLL * (size_t)gz;
                u[q] = PRO ? (R)corrected<R, A, 3>(src, Vp, gx, gy, gz, L, idx) : src[idx];
                fv[q] = f[idx];
            }
        }
#pragma unroll
        for (int q = 0; q < CH; ++q) {
            const int e = e0 + q * NT;
            if (e < NE) { bufA[e] = u[q]; bufB[e] = (R)0; bufF[e] = fv[q]; }
        }
    }
    __syncthreads();
    // ---- S sweeps, ping-pong; sweep s is valid on the extended cells [s, E - s) in every direction. Four cells per
    // thread at a time: their loads first (always inside the arrays), one division guard for the group, then the stores.
    auto sweep = [&](auto s_tag) {
        constexpr int s = decltype(s_tag)::value;
        constexpr int RX = EX - 2 * s, RY = EY - 2 * s, RZ = EZ - 2 * s, NR = RX * RY * RZ, G = 4;
        const R *const sa = bufA;
        R *const sb = bufB;
        for (int r0 = tid; r0 < NR; r0 += G * NT) {
            A num[G], out[G];
            int ee[G];
            bool ok[G];
#pragma unroll
            for (int q = 0; q < G; ++q) {
                const int r = r0 + q * NT < NR ? r0 + q * NT : r0;
                const int i = r % RX + s, j = (r / RX) % RY + s, k = r / (RX * RY) + s;
                const int gx = gx0 + i, gy = gy0 + j, gz = gz0 + k;
                ok[q] = r0 + q * NT < NR && (unsigned)gx < (unsigned)L && (unsigned)gy < (unsigned)L && (unsigned)gz < (unsigned)L;
                const int e = i + EX * (j + EY * k);
                ee[q] = e;
                A Ssum = Ar<A>::add(Ar<A>::add(Ar<A>::add((A)sa[e - 1], (A)sa[e + 1]), (A)sa[e - EX]), (A)sa[e + EX]);
                Ssum = Ar<A>::add(Ar<A>::add(Ssum, (A)sa[e - EX * EY]), (A)sa[e + EX * EY]);
                num[q] = jacobi_num<A>(Ssum, (A)bufF[e], c);
            }
            div_adiag_group<3, A, G>(num, out, c);
#pragma unroll
            for (int q = 0; q < G; ++q)
                if (ok[q]) sb[ee[q]] = (R)out[q];
        }
        __syncthreads();
        R *t = bufA; bufA = bufB; bufB = t;
    };
    sweep(std::integral_constant<int, 1>{});
    if constexpr (S >= 2) sweep(std::integral_constant<int, 2>{});
    if constexpr (S >= 3) sweep(std::integral_constant<int, 3>{});
    if constexpr (S >= 4) sweep(std::integral_constant<int, 4>{});
    static_assert(S >= 1 && S <= 4, "1..4 sweeps per launch");
    // ---- out: the block itself (always inside the grid: L is a multiple of the block)
    for (int r = tid; r < C::NB; r += NT) {
        const int i = r % BX, j = (r / BX) % BY, k = r / (BX * BY);
        const int e = (i + S) + EX * ((j + S) + EY * (k + S));
        dst[(size_t)(gx0 + S + i) + sL * (size_t)(gy0 + S + j) + sLL * (size_t)(gz0 + S + k)] = bufA[e];
    }
}

}  // namespace mg

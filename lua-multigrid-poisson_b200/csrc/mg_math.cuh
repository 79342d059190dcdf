// mg_math.cuh -- per-point arithmetic shared by EVERY kernel of the library.
//
// All kernels (one-operator-per-launch reference sequence, fused passes, temporally blocked
// streaming smoother, persistent small-level kernel) call the same __forceinline__ functions
// below, built only from explicitly rounded intrinsics (__fadd_rn, __fmaf_rn, ...), which
// nvcc never contracts or reassociates. That is what makes "per-point arithmetic identical
// across kernels" a structural property instead of a compiler accident, and what lets the
// fused path be compared bit-for-bit with the reference sequence and with the CPU oracle.
//
// Reference expressions (cpu-raw.lua:34-57, OpenCL twins gpu.lua:83-124), with h = 2^-k:
//   S       = ((u_xl + u_xr) + u_yl) + u_yr            [+ u_zl) + u_zr in 3-D]
//   Jacobi  : dest = (f - S/h^2) / adiag               adiag = -4/h^2 (2-D), -6/h^2 (3-D)
//   residual: r    = f - (S/h^2 + adiag*u)
// Because h^2 is a power of two, S/h^2 == S*inv_h2 exactly, so
//   numerator n = RN(f - S*inv_h2) = fma(-S, inv_h2, f)            (one rounding, as the reference)
//   2-D: n/adiag == n * (-h^2/4) exactly
//   3-D: n/adiag is one true rounding; computed as a Markstein-corrected reciprocal multiply
//        (exhaustively verified equal to IEEE division for every normal fp32 numerator and
//        every divisor 6*2^m, m>=3; numerators below the guard fall back to IEEE division).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mg {

// Programmatic dependent launch (the "pdl" option, mg_engine.cuh launch_k): the kernels of a V-cycle are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so the next kernel's CTAs are scheduled while this one still runs
// (PREEXIT) and block here (ACQBULK) until every prerequisite grid has completed and its writes are visible. First
// statement of every such kernel: nothing above it may touch global memory. Without the launch attribute both
// instructions do nothing.
__device__ __forceinline__ void pdl_enter()
{
    asm volatile("griddepcontrol.launch_dependents;\n\tgriddepcontrol.wait;" ::: "memory");
}

template <typename A> struct Ar;
template <> struct Ar<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float abs(float a) { return fabsf(a); }
    static __host__ __device__ __forceinline__ float tiny() { return 1e-20f; }
    // 2*|bits| - 1 (mod 2^32): +-0 -> 0xffffffff, otherwise monotone in |n|
    typedef unsigned int key_t;
    static __device__ __forceinline__ key_t guard_key(float n) { return 2u * __float_as_uint(n) - 1u; }
    static __device__ __forceinline__ key_t guard_threshold() { return 2u * __float_as_uint(1e-20f) - 1u; }
};
template <> struct Ar<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double abs(double a) { return fabs(a); }
    static __host__ __device__ __forceinline__ double tiny() { return 1e-200; }
    typedef unsigned long long key_t;
    static __device__ __forceinline__ key_t guard_key(double n) { return 2ull * (unsigned long long)__double_as_longlong(n) - 1ull; }
    static __device__ __forceinline__ key_t guard_threshold() { return 2ull * (unsigned long long)__double_as_longlong(1e-200) - 1ull; }
};

// Per-level coefficients, computed on the host in double and rounded to the arithmetic type
// (all exact: powers of two times 4 or 6, except yneg which is the rounded reciprocal).
template <typename A> struct Coef {
    A inv_h2;  // 1/h^2
    A adiag;   // -4/h^2 (2-D) or -6/h^2 (3-D)
    A cneg;    // 2-D: -h^2/4 (= 1/adiag, exact)
    A yneg;    // 3-D: RN(1/adiag)
    A nadiag;  // -adiag
    A omega;   // relaxation weight; the reference is omega = 1 (cpu-raw.lua:34-44), anything else is a labelled extension
    int weighted;   // omega != 1: the un-fused kernels apply u + omega * (Jacobi(u) - u); 0 keeps the reference arithmetic
};

template <typename A> static inline Coef<A> make_coef(int dim, double h, double omega = 1.0)
{
    Coef<A> c;
    double h2 = h * h;
    A ah = (A)h;                 // gpu.lua:291 passes h as real[1]; h is a power of two => exact
    A ah2 = ah * ah;
    (void)h2;
    c.inv_h2 = (A)1 / ah2;
    c.adiag = (dim == 2 ? (A)-4 : (A)-6) / ah2;
    c.cneg = (A)1 / c.adiag;
    c.yneg = (A)1 / c.adiag;
    c.nadiag = -c.adiag;
    c.omega = (A)omega;
    c.weighted = omega != 1.0;
    return c;
}

// weighted Jacobi (NOT in the reference, which is omega = 1): u + omega * (J(u) - u), two roundings on top of J
template <typename A> __device__ __forceinline__ A relax(A jac, A u, const Coef<A> &c)
{
    return c.weighted ? Ar<A>::fma(c.omega, Ar<A>::sub(jac, u), u) : jac;
}

// numerator of the Jacobi update: RN(f - S/h^2)
template <typename A> __device__ __forceinline__ A jacobi_num(A S, A f, const Coef<A> &c)
{
    return Ar<A>::fma(-S, c.inv_h2, f);
}

// n / adiag, correctly rounded.
template <int DIM, typename A> __device__ __forceinline__ A div_adiag(A n, const Coef<A> &c)
{
    if (DIM == 2) {
        return Ar<A>::mul(n, c.cneg);
    } else {
        A q = Ar<A>::mul(n, c.yneg);
        A r = Ar<A>::fma(c.nadiag, q, n);       // n - adiag*q, exact
        A q2 = Ar<A>::fma(r, c.yneg, q);
        if (Ar<A>::abs(n) < Ar<A>::tiny() && n != (A)0) q2 = Ar<A>::div(n, c.adiag);
        return q2;
    }
}

template <int DIM, typename A> __device__ __forceinline__ A jacobi_point(A S, A f, const Coef<A> &c)
{
    return div_adiag<DIM, A>(jacobi_num<A>(S, f, c), c);
}

// The same division for a group of N numerators with ONE branch for the whole group: the
// Markstein sequence for everybody unless some numerator is in the guarded tiny range, in
// which case everybody takes IEEE division (identical results outside that range).
template <int DIM, typename A, int N> __device__ __forceinline__ void div_adiag_group(const A *n, A *q, const Coef<A> &c)
{
    if (DIM == 2) {
#pragma unroll
        for (int i = 0; i < N; ++i) q[i] = Ar<A>::mul(n[i], c.cneg);
    } else {
        // Hot path: one key per numerator and a min chain for the group. The key orders
        // "tiny but non-zero" below everything else (zero maps to the largest key), so exact
        // zeros -- ubiquitous outside the grid and before the first correction -- stay on the
        // fast path. Fallback: IEEE division for the whole group (exact for every input).
        typename Ar<A>::key_t m = Ar<A>::guard_key(n[0]);
#pragma unroll
        for (int i = 1; i < N; ++i) {
            const typename Ar<A>::key_t k = Ar<A>::guard_key(n[i]);
            m = k < m ? k : m;
        }
        const bool slow = m < Ar<A>::guard_threshold();
        if (!slow) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                A q1 = Ar<A>::mul(n[i], c.yneg);
                A r = Ar<A>::fma(c.nadiag, q1, n[i]);
                q[i] = Ar<A>::fma(r, c.yneg, q1);
            }
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) q[i] = Ar<A>::div(n[i], c.adiag);  // unrolled: keeps n/q in registers
        }
    }
}

// r = f - (S/h^2 + adiag*u): product and sum separately rounded (cpu-raw.lua:55-56)
template <typename A> __device__ __forceinline__ A residual_point(A S, A f, A u, const Coef<A> &c)
{
    A askew = Ar<A>::mul(S, c.inv_h2);
    A au = Ar<A>::add(askew, Ar<A>::mul(c.adiag, u));
    return Ar<A>::sub(f, au);
}

// neighbour sum with out-of-range neighbours reading 0 (cpu-raw.lua:36-39)
template <int DIM, typename R, typename A>
__device__ __forceinline__ A stencil_sum(const R *u, int i, int j, int k, int L, size_t idx)
{
    const size_t sL = (size_t)L;
    A xl = i > 0 ? (A)u[idx - 1] : (A)0;
    A xr = i < L - 1 ? (A)u[idx + 1] : (A)0;
    A yl = j > 0 ? (A)u[idx - sL] : (A)0;
    A yr = j < L - 1 ? (A)u[idx + sL] : (A)0;
    A S = Ar<A>::add(Ar<A>::add(Ar<A>::add(xl, xr), yl), yr);
    if (DIM == 3) {
        const size_t sLL = sL * sL;
        A zl = k > 0 ? (A)u[idx - sLL] : (A)0;
        A zr = k < L - 1 ? (A)u[idx + sLL] : (A)0;
        S = Ar<A>::add(Ar<A>::add(S, zl), zr);
    }
    return S;
}

}  // namespace mg

/*
 * mg_oracle_impl.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Type-generic body of the CPU oracle. Included three times by mg_oracle.c with
 *   REAL  = storage type of every field        (cpu-raw.lua:143 `self.real`)
 *   ACC   = type every expression is evaluated in
 *   SFX   = symbol suffix
 *
 *   (double,double,_f64)    cpu-raw.lua default (`real or 'double'`, cpu-raw.lua:143)
 *   (float ,double,_f32a64) cpu-raw.lua with real='float': LuaJIT numbers are doubles, so
 *                           every expression is evaluated in double and rounded once when
 *                           stored into the float image
 *   (float ,float ,_f32)    gpu.lua on an fp32-only device (gpu.lua:32,39 `typedef float real`;
 *                           h is passed as real[1], gpu.lua:291)
 *
 * Every function restates one reference function; the citation is on the function.
 * Nothing here is copied: the reference is Lua, this is C, and the 3-D branches are the
 * dimensional extension defined in SURVEY.md section 8(a').
 *
 * Compile with -ffp-contract=off: LuaJIT on x86-64 never fuses a*b+c.
 */

#define ORC_CAT2(a, b) a##b
#define ORC_CAT(a, b) ORC_CAT2(a, b)
#define FN(name) ORC_CAT(name, SFX)

/* cpu-raw.lua:8-20 initCells (3-D: SURVEY 8(a'): source at (L/2,L/2,L/2)) */
static void FN(orc_init_cells)(int dim, int L, REAL *f, REAL *psi)
{
    int Lz = dim == 3 ? L : 1;
    int center = L / 2; /* math.floor(L / 2) */
    for (int k = 0; k < Lz; ++k)
        for (int j = 0; j < L; ++j)
            for (int i = 0; i < L; ++i) {
                size_t index = (size_t)i + (size_t)L * ((size_t)j + (size_t)L * k);
                ACC value = 0;
                if (i == center && j == center && (dim == 2 || k == center)) {
                    ACC charge = (ACC)1e+6;
                    ACC epsilon0 = 1;
                    value = -charge / epsilon0;
                }
                f[index] = (REAL)value;
                psi[index] = (REAL)(-(ACC)f[index]);
            }
}

/* cpu-raw.lua:34-44 Jacobi, driven by call2D (cpu-raw.lua:108-114): j outer, i inner.
 * OpenCL twin gpu.lua:83-102. Neighbour outside [0,L) reads 0 (cpu-raw.lua:36-39).
 * Sum is left-associated exactly as the Lua expression `u_xl + u_xr + u_yl + u_yr`. */
static void FN(orc_jacobi)(int dim, int L, REAL *destU, const REAL *u, const REAL *f, double h_in,
                           int nthreads)
{
    const ACC h = (ACC)h_in;
    const int Lz = dim == 3 ? L : 1;
    const size_t sL = (size_t)L, sLL = (size_t)L * L;
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for collapse(2) schedule(static) num_threads(nthreads) if (nthreads > 1)
#endif
    for (int k = 0; k < Lz; ++k)
        for (int j = 0; j < L; ++j)
            for (int i = 0; i < L; ++i) {
                size_t index = (size_t)i + sL * j + sLL * k;
                ACC u_xl = i > 0 ? (ACC)u[index - 1] : (ACC)0;
                ACC u_xr = i < L - 1 ? (ACC)u[index + 1] : (ACC)0;
                ACC u_yl = j > 0 ? (ACC)u[index - sL] : (ACC)0;
                ACC u_yr = j < L - 1 ? (ACC)u[index + sL] : (ACC)0;
                ACC hSq = h * h;
                ACC askew_u, adiag;
                if (dim == 2) {
                    askew_u = (u_xl + u_xr + u_yl + u_yr) / hSq;
                    adiag = -4 / hSq;
                } else {
                    ACC u_zl = k > 0 ? (ACC)u[index - sLL] : (ACC)0;
                    ACC u_zr = k < L - 1 ? (ACC)u[index + sLL] : (ACC)0;
                    askew_u = (u_xl + u_xr + u_yl + u_yr + u_zl + u_zr) / hSq;
                    adiag = -6 / hSq;
                }
                destU[index] = (REAL)(((ACC)f[index] - askew_u) / adiag);
            }
}

/* cpu-raw.lua:22-32 GaussSeidel: in place, lexicographic (i fastest), dead code in the
 * reference (cpu-raw.lua:177-179 is commented out). Kept only so the ordering is on record. */
static void FN(orc_gauss_seidel)(int dim, int L, REAL *u, const REAL *f, double h_in)
{
    const ACC h = (ACC)h_in;
    const int Lz = dim == 3 ? L : 1;
    const size_t sL = (size_t)L, sLL = (size_t)L * L;
    for (int k = 0; k < Lz; ++k)
        for (int j = 0; j < L; ++j)
            for (int i = 0; i < L; ++i) {
                size_t index = (size_t)i + sL * j + sLL * k;
                ACC u_xl = i > 0 ? (ACC)u[index - 1] : (ACC)0;
                ACC u_xr = i < L - 1 ? (ACC)u[index + 1] : (ACC)0;
                ACC u_yl = j > 0 ? (ACC)u[index - sL] : (ACC)0;
                ACC u_yr = j < L - 1 ? (ACC)u[index + sL] : (ACC)0;
                ACC hSq = h * h;
                ACC askew_u, adiag;
                if (dim == 2) {
                    askew_u = (u_xl + u_xr + u_yl + u_yr) / hSq;
                    adiag = -4 / hSq;
                } else {
                    ACC u_zl = k > 0 ? (ACC)u[index - sLL] : (ACC)0;
                    ACC u_zr = k < L - 1 ? (ACC)u[index + sLL] : (ACC)0;
                    askew_u = (u_xl + u_xr + u_yl + u_yr + u_zl + u_zr) / hSq;
                    adiag = -6 / hSq;
                }
                u[index] = (REAL)(((ACC)f[index] - askew_u) / adiag);
            }
}

/* cpu-raw.lua:46-57 calcResidual (OpenCL twin gpu.lua:104-124):
 * r = f - (askew_u + adiag*u); two separately rounded products/sums, no FMA. */
static void FN(orc_residual)(int dim, int L, REAL *r, const REAL *f, const REAL *u, double h_in,
                             int nthreads)
{
    const ACC h = (ACC)h_in;
    const int Lz = dim == 3 ? L : 1;
    const size_t sL = (size_t)L, sLL = (size_t)L * L;
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for collapse(2) schedule(static) num_threads(nthreads) if (nthreads > 1)
#endif
    for (int k = 0; k < Lz; ++k)
        for (int j = 0; j < L; ++j)
            for (int i = 0; i < L; ++i) {
                size_t index = (size_t)i + sL * j + sLL * k;
                ACC u_xl = i > 0 ? (ACC)u[index - 1] : (ACC)0;
                ACC u_xr = i < L - 1 ? (ACC)u[index + 1] : (ACC)0;
                ACC u_yl = j > 0 ? (ACC)u[index - sL] : (ACC)0;
                ACC u_yr = j < L - 1 ? (ACC)u[index + sL] : (ACC)0;
                ACC hSq = h * h;
                ACC askew_u, adiag;
                if (dim == 2) {
                    askew_u = (u_xl + u_xr + u_yl + u_yr) / hSq;
                    adiag = (ACC)-4. / hSq;
                } else {
                    ACC u_zl = k > 0 ? (ACC)u[index - sLL] : (ACC)0;
                    ACC u_zr = k < L - 1 ? (ACC)u[index + sLL] : (ACC)0;
                    askew_u = (u_xl + u_xr + u_yl + u_yr + u_zl + u_zr) / hSq;
                    adiag = (ACC)-6. / hSq;
                }
                ACC a_u = askew_u + adiag * (ACC)u[index];
                r[index] = (REAL)((ACC)f[index] - a_u);
            }
}

/* cpu-raw.lua:59-63 reduceResidual (gpu.lua:126-137): plain 2x2 average, left to right.
 * 3-D (SURVEY 8(a')): .125 * sum of the 8 children, i fastest, then j, then k. */
static void FN(orc_restrict)(int dim, int L2, REAL *R, const REAL *r, int nthreads)
{
    const int L = L2 << 1;
    const int L2z = dim == 3 ? L2 : 1;
    const size_t sL = (size_t)L, sLL = (size_t)L * L;
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for collapse(2) schedule(static) num_threads(nthreads) if (nthreads > 1)
#endif
    for (int K = 0; K < L2z; ++K)
        for (int J = 0; J < L2; ++J)
            for (int I = 0; I < L2; ++I) {
                size_t srci = ((size_t)I << 1) + sL * ((size_t)J << 1) + sLL * ((size_t)K << 1);
                size_t dst = (size_t)I + (size_t)L2 * ((size_t)J + (size_t)L2 * K);
                if (dim == 2) {
                    R[dst] = (REAL)((ACC).25 * ((ACC)r[srci] + (ACC)r[srci + 1] + (ACC)r[srci + sL] +
                                                (ACC)r[srci + sL + 1]));
                } else {
                    R[dst] = (REAL)((ACC).125 *
                                    ((ACC)r[srci] + (ACC)r[srci + 1] + (ACC)r[srci + sL] +
                                     (ACC)r[srci + sL + 1] + (ACC)r[srci + sLL] +
                                     (ACC)r[srci + sLL + 1] + (ACC)r[srci + sLL + sL] +
                                     (ACC)r[srci + sLL + sL + 1]));
                }
            }
}

/* cpu-raw.lua:65-73 expandResidual (gpu.lua:139-161): piecewise-constant injection. */
static void FN(orc_prolong)(int dim, int L2, REAL *v, const REAL *V, int nthreads)
{
    const int L = L2 << 1;
    const int L2z = dim == 3 ? L2 : 1;
    const size_t sL = (size_t)L, sLL = (size_t)L * L;
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for collapse(2) schedule(static) num_threads(nthreads) if (nthreads > 1)
#endif
    for (int K = 0; K < L2z; ++K)
        for (int J = 0; J < L2; ++J)
            for (int I = 0; I < L2; ++I) {
                size_t dsti = ((size_t)I << 1) + sL * ((size_t)J << 1) + sLL * ((size_t)K << 1);
                REAL src = V[(size_t)I + (size_t)L2 * ((size_t)J + (size_t)L2 * K)];
                v[dsti] = src;
                v[dsti + 1] = src;
                v[dsti + sL] = src;
                v[dsti + sL + 1] = src;
                if (dim == 3) {
                    v[dsti + sLL] = src;
                    v[dsti + sLL + 1] = src;
                    v[dsti + sLL + sL] = src;
                    v[dsti + sLL + sL + 1] = src;
                }
            }
}

/* cpu-raw.lua:83-85 addTo driven by call1D (cpu-raw.lua:102-106, 230) */
static void FN(orc_add_to)(size_t n, REAL *u, const REAL *v, int nthreads)
{
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads) if (nthreads > 1)
#endif
    for (size_t i = 0; i < n; ++i) u[i] = (REAL)((ACC)u[i] + (ACC)v[i]);
}

/* cpu-raw.lua:96-100 calcFrobErr, then cpu-raw.lua:250-254: sequential sum held in a Lua
 * number (double) whatever `real` is; err = sqrt(sum / size^dim).
 * For the (float,float) flavour this follows gpu.lua:189-200 (kernel in real) and
 * gpu.lua:365-369 (host sum in double). */
static double FN(orc_frob_err)(int dim, int size, REAL *errorBuf, const REAL *psi, const REAL *psiOld)
{
    size_t n = (size_t)size * size * (dim == 3 ? (size_t)size : 1);
    for (size_t index = 0; index < n; ++index) {
        ACC d = (ACC)psi[index] - (ACC)psiOld[index];
        errorBuf[index] = (REAL)(d * d);
    }
    double err = 0;
    for (size_t j = 0; j < n; ++j) err = err + (double)errorBuf[j];
    return sqrt(err / (double)n);
}

/* cpu-raw.lua:87-94 calcRelErr -- unused by run(); restated for completeness. */
static void FN(orc_rel_err)(size_t n, REAL *errorBuf, const REAL *psi, const REAL *psiOld)
{
    for (size_t index = 0; index < n; ++index) {
        if (psiOld[index] != 0 && psiOld[index] != psi[index])
            errorBuf[index] = (REAL)fabs((double)((ACC)1. - (ACC)psi[index] / (ACC)psiOld[index]));
        else
            errorBuf[index] = 0;
    }
}

/* cpu-raw.lua:176-184 inPlaceIterativeSolver: Jacobi into tmpU then copy back. */
static void FN(orc_in_place_solver)(orc_ctx *c, int L, REAL *u, const REAL *f, double h)
{
    size_t n = (size_t)L * L * (c->dim == 3 ? (size_t)L : 1);
    FN(orc_jacobi)(c->dim, L, (REAL *)c->tmpU, u, f, h, c->nthreads);
    memcpy(u, c->tmpU, n * sizeof(REAL));
}

/* cpu-raw.lua:186-237 twoGrid. The `show` sites (cpu-raw.lua:192,194,199-205,209-212,219,
 * 223,228,231,235) become trace records when tracing is on. Vs[L2] is NOT re-zeroed
 * (cpu-raw.lua:221-222). */
static void FN(orc_two_grid)(orc_ctx *c, double h, REAL *u, const REAL *f, int L)
{
    const size_t n = (size_t)L * L * (c->dim == 3 ? (size_t)L : 1);
    if (L == 1) {
        orc_trace(c, 'f', L, f, n * sizeof(REAL));
        FN(orc_in_place_solver)(c, L, u, f, h);
        orc_trace(c, 'u', L, u, n * sizeof(REAL));
        return;
    }
    for (int i = 1; i <= c->smooth; ++i) {
        if (L == c->size) orc_trace(c, 'f', L, f, n * sizeof(REAL));
        FN(orc_in_place_solver)(c, L, u, f, h);
        orc_trace(c, 'u', L, u, n * sizeof(REAL));
    }
    int lv = orc_log2(L);
    REAL *r = (REAL *)c->rs[lv];
    orc_trace(c, 'f', L, f, n * sizeof(REAL));
    orc_trace(c, 'u', L, u, n * sizeof(REAL));
    FN(orc_residual)(c->dim, L, r, f, u, h, c->nthreads);
    orc_trace(c, 'r', L, r, n * sizeof(REAL));

    int L2 = L / 2;
    const size_t n2 = (size_t)L2 * L2 * (c->dim == 3 ? (size_t)L2 : 1);
    REAL *R = (REAL *)c->Rs[lv - 1];
    FN(orc_restrict)(c->dim, L2, R, r, c->nthreads);
    orc_trace(c, 'R', L2, R, n2 * sizeof(REAL));

    REAL *V = (REAL *)c->Vs[lv - 1];
    FN(orc_two_grid)(c, 2 * h, V, R, L2);
    orc_trace(c, 'V', L2, V, n2 * sizeof(REAL));

    REAL *v = (REAL *)c->vs[lv];
    FN(orc_prolong)(c->dim, L2, v, V, c->nthreads);
    orc_trace(c, 'v', L, v, n * sizeof(REAL));

    FN(orc_add_to)(n, u, v, c->nthreads);
    orc_trace(c, 'u', L, u, n * sizeof(REAL));

    for (int i = 1; i <= c->smooth; ++i) {
        FN(orc_in_place_solver)(c, L, u, f, h);
        orc_trace(c, 'u', L, u, n * sizeof(REAL));
    }
}

/* one iteration of the loop body of cpu-raw.lua:245-256 */
static double FN(orc_step)(orc_ctx *c)
{
    size_t n = c->N;
    double h = 1.0 / c->size; /* cpu-raw.lua:242 */
    memcpy(c->psiOld, c->psi, n * sizeof(REAL));
    FN(orc_two_grid)(c, h, (REAL *)c->psi, (const REAL *)c->f, c->size);
    return FN(orc_frob_err)(c->dim, c->size, (REAL *)c->errorBuf, (const REAL *)c->psi,
                            (const REAL *)c->psiOld);
}

/* Krylov comparator (test infrastructure for mg_cg): textbook conjugate gradient on the operator of
 * test/converge-multigrid-vs-krylov.lua:48-58, A(u) = (u_xl + u_xr + u_yl + u_yr - 4 u) / h^2, h = 1/width
 * (3-D: six neighbours, -6 u), b = f, x given (the experiment passes -f, :45-46); err = ||r||_2/||b||_2;
 * per-iteration ||x||_inf as the experiment records it (:62-64). `solver.conjgrad` itself is an
 * un-vendored dependency of the reference (thenumbernine/lua-solver, unpinned): PARITY UNPINNED. */
static void FN(orc_apply_A)(int dim, int L, REAL *out, const REAL *u)
{
    const ACC h = (ACC)1 / (ACC)L;
    const int Lz = dim == 3 ? L : 1;
    const size_t sL = (size_t)L, sLL = (size_t)L * L;
    for (int k = 0; k < Lz; ++k)
        for (int j = 0; j < L; ++j)
            for (int i = 0; i < L; ++i) {
                size_t index = (size_t)i + sL * j + sLL * k;
                ACC u_xl = i > 0 ? (ACC)u[index - 1] : (ACC)0;
                ACC u_xr = i < L - 1 ? (ACC)u[index + 1] : (ACC)0;
                ACC u_yl = j > 0 ? (ACC)u[index - sL] : (ACC)0;
                ACC u_yr = j < L - 1 ? (ACC)u[index + sL] : (ACC)0;
                ACC v;
                if (dim == 2) {
                    v = (u_xl + u_xr + u_yl + u_yr - 4 * (ACC)u[index]) / (h * h);
                } else {
                    ACC u_zl = k > 0 ? (ACC)u[index - sLL] : (ACC)0;
                    ACC u_zr = k < L - 1 ? (ACC)u[index + sLL] : (ACC)0;
                    v = (u_xl + u_xr + u_yl + u_yr + u_zl + u_zr - 6 * (ACC)u[index]) / (h * h);
                }
                out[index] = (REAL)v;
            }
}

static int FN(orc_cg)(int dim, int L, REAL *x, const REAL *b, int max_iter, double epsilon, double *err_hist,
                      double *linf_hist, int *n_done)
{
    const size_t n = (size_t)L * L * (dim == 3 ? (size_t)L : 1);
    REAL *r = (REAL *)malloc(n * sizeof(REAL)), *p = (REAL *)malloc(n * sizeof(REAL)), *Ap = (REAL *)malloc(n * sizeof(REAL));
    if (!r || !p || !Ap) return -1;
    FN(orc_apply_A)(dim, L, Ap, x);
    double rr = 0, bb = 0;
    for (size_t i = 0; i < n; ++i) {
        r[i] = (REAL)((ACC)b[i] - (ACC)Ap[i]);
        p[i] = r[i];
        rr += (double)r[i] * (double)r[i];
        bb += (double)b[i] * (double)b[i];
    }
    if (bb == 0) bb = 1;
    int it = 0;
    double err = sqrt(rr / bb);
    if (!(err < epsilon)) {
        for (it = 1; it <= max_iter; ++it) {
            FN(orc_apply_A)(dim, L, Ap, p);
            double pAp = 0;
            for (size_t i = 0; i < n; ++i) pAp += (double)p[i] * (double)Ap[i];
            const ACC alpha = (ACC)(rr / pAp);
            double rrn = 0, mx = 0;
            for (size_t i = 0; i < n; ++i) {
                x[i] = (REAL)fma((double)alpha, (double)p[i], (double)x[i]);
                r[i] = (REAL)fma(-(double)alpha, (double)Ap[i], (double)r[i]);
                rrn += (double)r[i] * (double)r[i];
                if (fabs((double)x[i]) > mx) mx = fabs((double)x[i]);
            }
            err = sqrt(rrn / bb);
            if (err_hist) err_hist[it - 1] = err;
            if (linf_hist) linf_hist[it - 1] = mx;
            if (err < epsilon || !isfinite(err)) break;
            const ACC beta = (ACC)(rrn / rr);
            for (size_t i = 0; i < n; ++i) p[i] = (REAL)fma((double)beta, (double)p[i], (double)r[i]);
            rr = rrn;
        }
        if (it > max_iter) it = max_iter;
    }
    free(r); free(p); free(Ap);
    if (n_done) *n_done = it;
    return 0;
}

#undef FN
#undef REAL
#undef ACC
#undef SFX

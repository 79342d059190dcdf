/*
 * mg_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the reference's V-cycle path, `MultigridCPURaw` in
 * /root/reference/cpu-raw.lua, used ONLY as the checker for the CUDA path:
 *   - tests/            (parity tests)
 *   - __graft_entry__.smoke()
 *   - bench.py          (the `cpu_baseline` leg and `--impl reference`)
 * Nothing in the product (lua-multigrid-poisson_b200/) may import, link or call it.
 *
 * PARITY PINNING. The reference ships no golden vectors, no assertions and no fixtures
 * (SURVEY.md section 4), no Lua runtime exists in this image, and the reference is not
 * C/C++/Python, so neither `oracle/_ref` nor imported-Python fixtures are possible. Instead the
 * reference's OWN SOURCE TEXT (cpu-raw.lua, unmodified) is executed by a purpose-built Lua
 * interpreter (oracle/minilua.py, driver oracle/run_reference.py) and what it computes in run()
 * -- both err values, f, psi, psiOld, rs/Rs/vs/Vs of every level and the complete sequence of
 * `show` dumps of twoGrid -- is committed as tests/golden/ref_2d_*.npz (sizes 2..64, real =
 * double and float). The same is done for gpu.lua (oracle/run_reference_gpu.py): its host code
 * runs under minilua on a fake in-memory OpenCL device whose kernels are gpu.lua's own OpenCL C
 * source compiled by gcc (-ffp-contract=off) into oracle/_ref/ -> tests/golden/refgpu_2d_*.npz
 * (real = float there is fp32 ARITHMETIC), and for cpu.lua (oracle/run_reference_cpu.py, the
 * variant that re-zeroes the coarse corrections every cycle) -> tests/golden/refcpu_2d_*.npz.
 * tests/test_reference_source.py requires this oracle
 * to reproduce all those fixtures BIT FOR BIT; tests/test_gpu_vcycle.py requires the same of the
 * CUDA path. Status: the three 2-D modes (f64, f32 storage + f64 arithmetic, f32) are PINNED TO
 * THE REFERENCE SOURCE AS EXECUTED BY minilua (+ gcc for the kernels) -- not by LuaJIT / an
 * OpenCL device: see the fidelity arguments in minilua.py and run_reference_gpu.py. The 3-D rules
 * are this project's extension (SURVEY.md section 8(a')) and have no reference run to be pinned
 * to. The hand-derived known answers KA1-KA5 of SURVEY.md section 8(c)
 * (tests/test_oracle_known_answers.py) and the oracle-generated fixtures
 * (tests/golden/make_golden.py) remain as independent checks.
 *
 * Reference functions restated (file:line in /root/reference):
 *   initCells        cpu-raw.lua:8-20        -> orc_init_cells
 *   GaussSeidel      cpu-raw.lua:22-32       -> orc_gauss_seidel (dead code in the reference)
 *   Jacobi           cpu-raw.lua:34-44       -> orc_jacobi
 *   calcResidual     cpu-raw.lua:46-57       -> orc_residual
 *   reduceResidual   cpu-raw.lua:59-63       -> orc_restrict
 *   expandResidual   cpu-raw.lua:65-73       -> orc_prolong
 *   addTo            cpu-raw.lua:83-85       -> orc_add_to
 *   calcRelErr       cpu-raw.lua:87-94       -> orc_rel_err
 *   calcFrobErr+sum  cpu-raw.lua:96-100,249-254 -> orc_frob_err
 *   call1D/call2D    cpu-raw.lua:102-114     -> loop nests (j outer, i inner)
 *   init             cpu-raw.lua:142-174     -> orc_create
 *   inPlaceIterativeSolver cpu-raw.lua:176-184 -> orc_in_place_solver
 *   twoGrid          cpu-raw.lua:186-237     -> orc_two_grid
 *   run              cpu-raw.lua:239-258     -> orc_run
 *   A + conjgrad call test/converge-multigrid-vs-krylov.lua:38-69 -> orc_apply_A, orc_cg (solver lib un-vendored)
 *
 * Threads: `nthreads` > 1 parallelises the order-independent loops (Jacobi, residual,
 * restriction, prolongation, add) with OpenMP; results are bit-identical to 1 thread
 * because every cell is computed independently. The error sum stays sequential.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAX_LEVELS 32

enum { ORC_REAL_F64 = 0, ORC_REAL_F32 = 1, ORC_REAL_F32_ACC64 = 2 };
enum {
    ORC_BUF_F = 0,
    ORC_BUF_PSI = 1,
    ORC_BUF_PSIOLD = 2,
    ORC_BUF_ERRORBUF = 3,
    ORC_BUF_TMPU = 4,
    ORC_BUF_r = 5,
    ORC_BUF_R = 6,
    ORC_BUF_v = 7,
    ORC_BUF_V = 8
};

typedef struct orc_trace_rec {
    char name;
    int L;
    size_t offset, bytes;
} orc_trace_rec;

typedef struct orc_ctx {
    int dim, size, real_kind, smooth, nlevels, nthreads;
    size_t N, elem;
    void *f, *psi, *psiOld, *errorBuf, *tmpU;
    void *rs[ORC_MAX_LEVELS], *Rs[ORC_MAX_LEVELS], *vs[ORC_MAX_LEVELS], *Vs[ORC_MAX_LEVELS];
    /* trace (the reference's `debugging` dumps, cpu-raw.lua:121,126-140) */
    int trace_on;
    orc_trace_rec *recs;
    size_t nrecs, caprecs;
    unsigned char *tdata;
    size_t tbytes, tcap;
} orc_ctx;

static int orc_log2(int L)
{
    int k = 0;
    while ((1 << k) < L) ++k;
    return k;
}

static void orc_trace(orc_ctx *c, char name, int L, const void *data, size_t bytes)
{
    if (!c->trace_on) return;
    if (c->nrecs == c->caprecs) {
        c->caprecs = c->caprecs ? 2 * c->caprecs : 256;
        c->recs = (orc_trace_rec *)realloc(c->recs, c->caprecs * sizeof(orc_trace_rec));
    }
    if (c->tbytes + bytes > c->tcap) {
        while (c->tbytes + bytes > c->tcap) c->tcap = c->tcap ? 2 * c->tcap : (1u << 20);
        c->tdata = (unsigned char *)realloc(c->tdata, c->tcap);
    }
    memcpy(c->tdata + c->tbytes, data, bytes);
    c->recs[c->nrecs].name = name;
    c->recs[c->nrecs].L = L;
    c->recs[c->nrecs].offset = c->tbytes;
    c->recs[c->nrecs].bytes = bytes;
    c->nrecs++;
    c->tbytes += bytes;
}

#define REAL double
#define ACC double
#define SFX _f64
#include "mg_oracle_impl.h"

#define REAL float
#define ACC double
#define SFX _f32a64
#include "mg_oracle_impl.h"

#define REAL float
#define ACC float
#define SFX _f32
#include "mg_oracle_impl.h"

#define ORC_DISPATCH(kind, call)                      \
    switch (kind) {                                   \
    case ORC_REAL_F64: call(_f64, double); break;     \
    case ORC_REAL_F32: call(_f32, float); break;      \
    case ORC_REAL_F32_ACC64: call(_f32a64, float); break; \
    default: return -1;                               \
    }

/* ---------------------------------------------------------------- stateless operators */

int orc_op_init_cells(int dim, int real_kind, int L, void *f, void *psi)
{
#define CALL(sfx, T) orc_init_cells##sfx(dim, L, (T *)f, (T *)psi)
    ORC_DISPATCH(real_kind, CALL)
#undef CALL
    return 0;
}
int orc_op_jacobi(int dim, int real_kind, int L, void *dest, const void *u, const void *f, double h,
                  int nthreads)
{
#define CALL(sfx, T) orc_jacobi##sfx(dim, L, (T *)dest, (const T *)u, (const T *)f, h, nthreads)
    ORC_DISPATCH(real_kind, CALL)
#undef CALL
    return 0;
}
int orc_op_gauss_seidel(int dim, int real_kind, int L, void *u, const void *f, double h)
{
#define CALL(sfx, T) orc_gauss_seidel##sfx(dim, L, (T *)u, (const T *)f, h)
    ORC_DISPATCH(real_kind, CALL)
#undef CALL
    return 0;
}
int orc_op_residual(int dim, int real_kind, int L, void *r, const void *f, const void *u, double h,
                    int nthreads)
{
#define CALL(sfx, T) orc_residual##sfx(dim, L, (T *)r, (const T *)f, (const T *)u, h, nthreads)
    ORC_DISPATCH(real_kind, CALL)
#undef CALL
    return 0;
}
int orc_op_restrict(int dim, int real_kind, int L2, void *R, const void *r, int nthreads)
{
#define CALL(sfx, T) orc_restrict##sfx(dim, L2, (T *)R, (const T *)r, nthreads)
    ORC_DISPATCH(real_kind, CALL)
#undef CALL
    return 0;
}
int orc_op_prolong(int dim, int real_kind, int L2, void *v, const void *V, int nthreads)
{
#define CALL(sfx, T) orc_prolong##sfx(dim, L2, (T *)v, (const T *)V, nthreads)
    ORC_DISPATCH(real_kind, CALL)
#undef CALL
    return 0;
}
int orc_op_add_to(int real_kind, size_t n, void *u, const void *v, int nthreads)
{
#define CALL(sfx, T) orc_add_to##sfx(n, (T *)u, (const T *)v, nthreads)
    ORC_DISPATCH(real_kind, CALL)
#undef CALL
    return 0;
}
int orc_op_rel_err(int real_kind, size_t n, void *errorBuf, const void *psi, const void *psiOld)
{
#define CALL(sfx, T) orc_rel_err##sfx(n, (T *)errorBuf, (const T *)psi, (const T *)psiOld)
    ORC_DISPATCH(real_kind, CALL)
#undef CALL
    return 0;
}
int orc_op_frob_err(int dim, int real_kind, int size, void *errorBuf, const void *psi,
                    const void *psiOld, double *err)
{
#define CALL(sfx, T) \
    *err = orc_frob_err##sfx(dim, size, (T *)errorBuf, (const T *)psi, (const T *)psiOld)
    ORC_DISPATCH(real_kind, CALL)
#undef CALL
    return 0;
}

int orc_op_apply_A(int dim, int real_kind, int L, void *out, const void *u)
{
#define CALL(sfx, T) orc_apply_A##sfx(dim, L, (T *)out, (const T *)u)
    ORC_DISPATCH(real_kind, CALL)
#undef CALL
    return 0;
}
int orc_op_cg(int dim, int real_kind, int L, void *x, const void *b, int max_iter, double epsilon,
              double *err_hist, double *linf_hist, int *n_done)
{
    int rc = -1;
#define CALL(sfx, T) rc = orc_cg##sfx(dim, L, (T *)x, (const T *)b, max_iter, epsilon, err_hist, linf_hist, n_done)
    ORC_DISPATCH(real_kind, CALL)
#undef CALL
    return rc;
}

/* ---------------------------------------------------------------- solver object */

/* cpu-raw.lua:142-174 init: five full-size images, four images per level L = 1,2,4..size,
 * zero-filled once, then initCells. */
orc_ctx *orc_create(int dim, int size, int real_kind, int smooth, int nthreads)
{
    if ((dim != 2 && dim != 3) || size < 1 || (size & (size - 1)) != 0) return NULL;
    if (real_kind < 0 || real_kind > 2) return NULL;
    orc_ctx *c = (orc_ctx *)calloc(1, sizeof(orc_ctx));
    if (!c) return NULL;
    c->dim = dim;
    c->size = size;
    c->real_kind = real_kind;
    c->smooth = smooth > 0 ? smooth : 7; /* cpu-raw.lua:123 */
    c->nthreads = nthreads > 0 ? nthreads : 1;
    c->elem = real_kind == ORC_REAL_F64 ? 8 : 4;
    c->N = (size_t)size * size * (dim == 3 ? (size_t)size : 1);
    c->nlevels = orc_log2(size) + 1;
    c->f = calloc(c->N, c->elem);
    c->psi = calloc(c->N, c->elem);
    c->psiOld = calloc(c->N, c->elem);
    c->errorBuf = calloc(c->N, c->elem);
    c->tmpU = calloc(c->N, c->elem);
    for (int i = 0; i < c->nlevels; ++i) {
        size_t L = (size_t)1 << i;
        size_t n = L * L * (dim == 3 ? L : 1);
        c->rs[i] = calloc(n, c->elem);
        c->Rs[i] = calloc(n, c->elem);
        c->vs[i] = calloc(n, c->elem);
        c->Vs[i] = calloc(n, c->elem);
    }
    orc_op_init_cells(dim, real_kind, size, c->f, c->psi);
    return c;
}

void orc_destroy(orc_ctx *c)
{
    if (!c) return;
    free(c->f);
    free(c->psi);
    free(c->psiOld);
    free(c->errorBuf);
    free(c->tmpU);
    for (int i = 0; i < c->nlevels; ++i) {
        free(c->rs[i]);
        free(c->Rs[i]);
        free(c->vs[i]);
        free(c->Vs[i]);
    }
    free(c->recs);
    free(c->tdata);
    free(c);
}

void orc_set_threads(orc_ctx *c, int nthreads) { c->nthreads = nthreads > 0 ? nthreads : 1; }

void *orc_buffer(orc_ctx *c, int which, int L)
{
    int lv = orc_log2(L > 0 ? L : c->size);
    switch (which) {
    case ORC_BUF_F: return c->f;
    case ORC_BUF_PSI: return c->psi;
    case ORC_BUF_PSIOLD: return c->psiOld;
    case ORC_BUF_ERRORBUF: return c->errorBuf;
    case ORC_BUF_TMPU: return c->tmpU;
    case ORC_BUF_r: return lv < c->nlevels ? c->rs[lv] : NULL;
    case ORC_BUF_R: return lv < c->nlevels ? c->Rs[lv] : NULL;
    case ORC_BUF_v: return lv < c->nlevels ? c->vs[lv] : NULL;
    case ORC_BUF_V: return lv < c->nlevels ? c->Vs[lv] : NULL;
    }
    return NULL;
}

/* twoGrid(h, u, f, L) on caller-chosen buffers (cpu-raw.lua:186) */
int orc_two_grid(orc_ctx *c, double h, void *u, const void *f, int L)
{
#define CALL(sfx, T) orc_two_grid##sfx(c, h, (T *)u, (const T *)f, L)
    ORC_DISPATCH(c->real_kind, CALL)
#undef CALL
    return 0;
}

/* twoGrid(1/size, psi, f, size) (cpu-raw.lua:247) */
int orc_vcycle(orc_ctx *c) { return orc_two_grid(c, 1.0 / c->size, c->psi, c->f, c->size); }

/* loop body of run() (cpu-raw.lua:246-254) */
int orc_step(orc_ctx *c, double *err)
{
#define CALL(sfx, T) *err = orc_step##sfx(c)
    ORC_DISPATCH(c->real_kind, CALL)
#undef CALL
    return 0;
}

/* cpu-raw.lua:239-258 run(): the reference hard-wires max_cycles = 2 (cpu-raw.lua:245) and
 * accuracy = 1e-10 (cpu-raw.lua:124); both are parameters here. */
int orc_run(orc_ctx *c, int max_cycles, double accuracy, double *errs, int *n_done)
{
    int it = 0;
    for (it = 0; it < max_cycles;) {
        double err;
        if (orc_step(c, &err)) return -1;
        if (errs) errs[it] = err;
        ++it;
        if (err < accuracy || !isfinite(err)) break;
    }
    if (n_done) *n_done = it;
    return 0;
}

/* true residual RMS ||f - A psi|| / sqrt(N): NOT in the reference (its `err` is the update
 * RMS); evaluated with the reference's own calcResidual, summed in double. */
int orc_residual_rms(orc_ctx *c, double *rms)
{
    void *r = c->rs[c->nlevels - 1];
    double h = 1.0 / c->size;
    if (orc_op_residual(c->dim, c->real_kind, c->size, r, c->f, c->psi, h, c->nthreads)) return -1;
    double s = 0;
    if (c->real_kind == ORC_REAL_F64) {
        const double *p = (const double *)r;
        for (size_t i = 0; i < c->N; ++i) s += p[i] * p[i];
    } else {
        const float *p = (const float *)r;
        for (size_t i = 0; i < c->N; ++i) s += (double)p[i] * (double)p[i];
    }
    *rms = sqrt(s / (double)c->N);
    return 0;
}

/* ---------------------------------------------------------------- trace access */
void orc_trace_enable(orc_ctx *c, int on) { c->trace_on = on; }
void orc_trace_clear(orc_ctx *c) { c->nrecs = 0; c->tbytes = 0; }
size_t orc_trace_count(orc_ctx *c) { return c->nrecs; }
int orc_trace_get(orc_ctx *c, size_t i, char *name, int *L, const void **data, size_t *bytes)
{
    if (i >= c->nrecs) return -1;
    *name = c->recs[i].name;
    *L = c->recs[i].L;
    *data = c->tdata + c->recs[i].offset;
    *bytes = c->recs[i].bytes;
    return 0;
}

/* the reference's debug text layout (cpu-raw.lua:126-134): name, then L rows of
 * ' '..value with the FIRST index as the row -- note `im[j+L*i]`, i.e. row i, column j. */
int orc_show_text(FILE *out, const char *name, const double *im, int L)
{
    fprintf(out, "%s\n", name);
    for (int i = 0; i < L; ++i) {
        for (int j = 0; j < L; ++j) fprintf(out, " %.14g", im[j + L * i]);
        fprintf(out, "\n");
    }
    return 0;
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}

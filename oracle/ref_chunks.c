/*
 * oracle/ref_chunks.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).  Never linked into or
 * called from the product library.
 *
 * Restates, in plain C with 0-based indices, the part of the reference that the reference
 * itself implements (everything except the UMFPACK call): the dense column-chunk repack of
 * the sparse L and U factors and the chunked triangular solves.
 *
 *   chunk ranges      reference src/SharedMemSparseLU.jl:101-149  (get_chunking_parameters)
 *   chunk shapes      reference src:151-178                       (allocate_chunks)
 *   chunk fill        reference src:180-243                       (fill_chunks!, rectangular part negated)
 *   lsolve!           reference src:349-367   trsv('L','N','U') + gemm(N=1) accumulate
 *   rsolve!           reference src:374-392   trsv('U','N','N') + gemm(N=1) accumulate
 *   ldiv!             reference src:286-342   wrk = (Rs.*b)[p]; lsolve; rsolve; x[q] = wrk
 *
 * The BLAS calls are replaced by the textbook column-oriented loops of the reference BLAS
 * (dtrsv / dgemv, no-transpose), so this file has no dependencies.
 *
 * PARITY UNPINNED for the factors fed into it (see ref_lu.c); for the solve path itself
 * the reference's tests pin only identities (L y = b, U z = y, A x = b at 1e-12 / 1e-10,
 * test/runtests.jl:51,70,86,104,120,163) and hold no golden vectors; tests/ check this
 * restatement against those identities and against SciPy SuperLU's own solve.
 */
#include <stdlib.h>
#include <string.h>

typedef long long i64;

typedef struct {
    i64 m, chunk_size, total_chunks;
    /* per chunk: column range [c0,c1), rectangular row range [r0,r1) (may be empty) */
    i64 *lc0, *lc1, *lr0, *lr1;
    i64 *uc0, *uc1, *ur0, *ur1;
    double **Ltri, **Lrect, **Utri, **Urect;   /* dense column-major blocks */
    double bytes;                               /* dense storage actually allocated */
} ref_chunks_t;

void ref_chunks_free(ref_chunks_t *C) {
    if (!C) return;
    for (i64 c = 0; c < C->total_chunks; ++c) {
        if (C->Ltri) free(C->Ltri[c]);
        if (C->Lrect) free(C->Lrect[c]);
        if (C->Utri) free(C->Utri[c]);
        if (C->Urect) free(C->Urect[c]);
    }
    free(C->Ltri); free(C->Lrect); free(C->Utri); free(C->Urect);
    free(C->lc0); free(C->lc1); free(C->lr0); free(C->lr1);
    free(C->uc0); free(C->uc1); free(C->ur0); free(C->ur1);
    free(C);
}

static i64 imin(i64 a, i64 b) { return a < b ? a : b; }

/* Dense bytes the reference would allocate for these factors, without allocating. */
double ref_chunks_predict_bytes(i64 m, i64 cs, const i64 *Lp, const i64 *Li,
                                const i64 *Up, const i64 *Ui) {
    if (cs > m) cs = m;
    if (cs < 1) return 0.0;
    i64 C = (m + cs - 1) / cs;
    double tot = 0.0;
    for (i64 c = 0; c < C; ++c) {
        i64 c0 = c * cs, c1 = imin(m, (c + 1) * cs), w = c1 - c0;
        i64 rmax = -1;
        for (i64 j = c0; j < c1; ++j) if (Li[Lp[j + 1] - 1] > rmax) rmax = Li[Lp[j + 1] - 1];
        i64 rows = rmax + 1 - c1; if (rows < 0) rows = 0;
        tot += (double)w * (double)w + (double)rows * (double)w;
        i64 u0 = (C - 1 - c) * cs, u1 = imin(m, (C - c) * cs), uw = u1 - u0;
        i64 rmin = m;
        for (i64 j = u0; j < u1; ++j) if (Ui[Up[j]] < rmin) rmin = Ui[Up[j]];
        i64 urows = u0 - rmin; if (urows < 0) urows = 0;
        tot += (double)uw * (double)uw + (double)urows * (double)uw;
    }
    return 8.0 * tot;
}

/*
 * Build the chunked representation from CSC factors (0-based, rows sorted, L with its
 * unit diagonal stored explicitly as UMFPACK returns it).  chunk_size <= 0 selects the
 * reference default 8 (src:67-70); it is clamped to the matrix size (src:72).
 */
ref_chunks_t *ref_chunks_build(i64 m, i64 chunk_size,
                               const i64 *Lp, const i64 *Li, const double *Lx,
                               const i64 *Up, const i64 *Ui, const double *Ux) {
    ref_chunks_t *C = (ref_chunks_t *)calloc(1, sizeof(ref_chunks_t));
    if (!C) return NULL;
    if (chunk_size <= 0) chunk_size = 8;
    if (chunk_size > m) chunk_size = m;
    C->m = m; C->chunk_size = chunk_size;
    i64 T = (m + chunk_size - 1) / chunk_size;
    C->total_chunks = T;
    size_t sz = (size_t)(T > 0 ? T : 1);
    C->lc0 = calloc(sz, sizeof(i64)); C->lc1 = calloc(sz, sizeof(i64));
    C->lr0 = calloc(sz, sizeof(i64)); C->lr1 = calloc(sz, sizeof(i64));
    C->uc0 = calloc(sz, sizeof(i64)); C->uc1 = calloc(sz, sizeof(i64));
    C->ur0 = calloc(sz, sizeof(i64)); C->ur1 = calloc(sz, sizeof(i64));
    C->Ltri = calloc(sz, sizeof(double *)); C->Lrect = calloc(sz, sizeof(double *));
    C->Utri = calloc(sz, sizeof(double *)); C->Urect = calloc(sz, sizeof(double *));

    /* ---- ranges (src:111-123 for L, src:132-144 for U; U chunk 0 holds the LAST columns) ---- */
    for (i64 c = 0; c < T; ++c) {
        i64 c0 = c * chunk_size, c1 = imin(m, (c + 1) * chunk_size);
        i64 last = -1;
        for (i64 j = c0; j < c1; ++j) { i64 r = Li[Lp[j + 1] - 1]; if (r > last) last = r; }
        C->lc0[c] = c0; C->lc1[c] = c1;
        C->lr0[c] = c1; C->lr1[c] = (last + 1 > c1) ? last + 1 : c1;   /* empty if nothing below */

        i64 u0 = (T - 1 - c) * chunk_size, u1 = imin(m, (T - c) * chunk_size);
        i64 first = m;
        for (i64 j = u0; j < u1; ++j) { i64 r = Ui[Up[j]]; if (r < first) first = r; }
        C->uc0[c] = u0; C->uc1[c] = u1;
        C->ur0[c] = (first < u0) ? first : u0; C->ur1[c] = u0;
    }
    /* ---- allocate zeroed dense blocks (src:156-175) and fill (src:186-242) ---- */
    for (i64 c = 0; c < T; ++c) {
        i64 w = C->lc1[c] - C->lc0[c], rows = C->lr1[c] - C->lr0[c];
        C->Ltri[c] = calloc((size_t)(w * w > 0 ? w * w : 1), sizeof(double));
        C->Lrect[c] = calloc((size_t)(rows * w > 0 ? rows * w : 1), sizeof(double));
        C->bytes += 8.0 * ((double)w * w + (double)rows * w);
        for (i64 j = C->lc0[c]; j < C->lc1[c]; ++j)
            for (i64 t = Lp[j]; t < Lp[j + 1]; ++t) {
                i64 r = Li[t];
                if (r < C->lc1[c]) C->Ltri[c][(r - C->lc0[c]) + (j - C->lc0[c]) * w] = Lx[t];
                else C->Lrect[c][(r - C->lr0[c]) + (j - C->lc0[c]) * rows] = -Lx[t];
            }
        i64 uw = C->uc1[c] - C->uc0[c], urows = C->ur1[c] - C->ur0[c];
        C->Utri[c] = calloc((size_t)(uw * uw > 0 ? uw * uw : 1), sizeof(double));
        C->Urect[c] = calloc((size_t)(urows * uw > 0 ? urows * uw : 1), sizeof(double));
        C->bytes += 8.0 * ((double)uw * uw + (double)urows * uw);
        for (i64 j = C->uc0[c]; j < C->uc1[c]; ++j)
            for (i64 t = Up[j]; t < Up[j + 1]; ++t) {
                i64 r = Ui[t];
                if (r >= C->uc0[c]) C->Utri[c][(r - C->uc0[c]) + (j - C->uc0[c]) * uw] = Ux[t];
                else C->Urect[c][(r - C->ur0[c]) + (j - C->uc0[c]) * urows] = -Ux[t];
            }
    }
    return C;
}

double ref_chunks_bytes(const ref_chunks_t *C) { return C->bytes; }
i64 ref_chunks_total(const ref_chunks_t *C) { return C->total_chunks; }
void ref_chunks_ranges(const ref_chunks_t *C, i64 *lc0, i64 *lc1, i64 *lr0, i64 *lr1,
                       i64 *uc0, i64 *uc1, i64 *ur0, i64 *ur1) {
    size_t b = (size_t)C->total_chunks * sizeof(i64);
    memcpy(lc0, C->lc0, b); memcpy(lc1, C->lc1, b); memcpy(lr0, C->lr0, b); memcpy(lr1, C->lr1, b);
    memcpy(uc0, C->uc0, b); memcpy(uc1, C->uc1, b); memcpy(ur0, C->ur0, b); memcpy(ur1, C->ur1, b);
}

/* y += A x, A rows-by-w column-major: the N=1 gemm of src:362-363 / 387-388. */
static void gemv_acc(i64 rows, i64 w, const double *A, const double *x, double *y) {
    for (i64 j = 0; j < w; ++j) {
        double t = x[j];
        const double *a = A + j * rows;
        for (i64 i = 0; i < rows; ++i) y[i] += t * a[i];
    }
}

/* In-place x <- L^{-1} x  (src:355-364). */
void ref_chunks_lsolve(const ref_chunks_t *C, double *x) {
    for (i64 c = 0; c < C->total_chunks; ++c) {
        i64 c0 = C->lc0[c], w = C->lc1[c] - c0;
        const double *T = C->Ltri[c];
        double *xc = x + c0;
        for (i64 j = 0; j < w; ++j) {            /* lower, no-transpose, UNIT diagonal */
            double t = xc[j];
            if (t != 0.0) for (i64 i = j + 1; i < w; ++i) xc[i] -= t * T[i + j * w];
        }
        gemv_acc(C->lr1[c] - C->lr0[c], w, C->Lrect[c], xc, x + C->lr0[c]);
    }
}

/* In-place x <- U^{-1} x  (src:380-389); chunk 0 is the last block of columns. */
void ref_chunks_rsolve(const ref_chunks_t *C, double *x) {
    for (i64 c = 0; c < C->total_chunks; ++c) {
        i64 c0 = C->uc0[c], w = C->uc1[c] - c0;
        const double *T = C->Utri[c];
        double *xc = x + c0;
        for (i64 j = w - 1; j >= 0; --j) {       /* upper, no-transpose, NON-unit diagonal */
            if (xc[j] != 0.0) {
                xc[j] /= T[j + j * w];
                double t = xc[j];
                for (i64 i = 0; i < j; ++i) xc[i] -= t * T[i + j * w];
            }
        }
        gemv_acc(C->ur1[c] - C->ur0[c], w, C->Urect[c], xc, x + C->ur0[c]);
    }
}

/* x = A \ b  (src:318-341).  Returns -1 on the reference's DimensionMismatch conditions. */
int ref_chunks_ldiv(const ref_chunks_t *C, i64 nx, i64 nb, const i64 *p, const i64 *q,
                    const double *Rs, const double *b, double *x, double *wrk) {
    i64 n = C->m;
    if (nx != n || nb != n) return -1;
    for (i64 i = 0; i < n; ++i) wrk[i] = Rs[p[i]] * b[p[i]];
    ref_chunks_lsolve(C, wrk);
    ref_chunks_rsolve(C, wrk);
    for (i64 i = 0; i < n; ++i) x[q[i]] = wrk[i];
    return 0;
}

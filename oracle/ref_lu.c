/*
 * oracle/ref_lu.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).  Never linked into or called
 * from the product library; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it.
 *
 * PARITY UNPINNED: the reference (johnomotani/SharedMemSparseLU.jl) delegates its numeric
 * factorization to SuiteSparse UMFPACK through Julia's SparseArrays stdlib
 * (reference src/SharedMemSparseLU.jl:74 `lu(A)`, :247 `lu!(F.lu_object, A)`); UMFPACK's
 * source is not under /root/reference, no version is pinned (no Manifest, no [compat]),
 * neither Julia nor SuiteSparse exists in this image, and the reference's tests hold no
 * golden vectors for L, U, p, q or Rs.  This file therefore restates the *contract* the
 * reference documents at src:305-316,
 *
 *        L * U == (Rs .* A)[p, q]       L unit lower, U upper, p/q permutations,
 *
 * with the published left-looking sparse LU of Gilbert & Peierls (the column algorithm
 * UMFPACK's results are defined by when the pivot sequence is fixed): column j of the
 * permuted, row-scaled matrix is solved against the already computed columns of L, the
 * pivot is either prescribed (static mode: same (p,q) => same L,U as any other
 * elimination with those pivots, up to summation order) or chosen by threshold partial
 * pivoting with a preference for the diagonal (the rule UMFPACK's symmetric strategy and
 * SuperLU's diag_pivot_thresh use).  Row scaling follows UMFPACK's default "SUM" rule:
 * Rs[i] = 1 / sum_j |a_ij|  (Rs are multipliers, as the reference applies them, src:326).
 *
 * Updates into column j are applied in ASCENDING pivot order so the floating point
 * summation order is well defined.
 *
 * All indices are 0-based int64 at this interface; CSC with sorted rows on output.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef long long i64;

typedef struct {
    i64 n;
    i64 *Lp, *Li; double *Lx; i64 lnz, lcap;
    i64 *Up, *Ui; double *Ux; i64 unz, ucap;
    i64 *p;          /* p[k] = original row chosen as k-th pivot          */
    i64 *q;          /* q[k] = original column eliminated at step k        */
    double *Rs;      /* row multipliers, indexed by ORIGINAL row           */
    double flops;    /* 2*mults + divisions actually performed             */
    i64 bad_col;     /* first column with a zero/NaN pivot, or -1          */
} oracle_lu_t;

static int cmp_i64(const void *a, const void *b) {
    i64 x = *(const i64 *)a, y = *(const i64 *)b;
    return (x > y) - (x < y);
}

void oracle_lu_free(oracle_lu_t *F) {
    if (!F) return;
    free(F->Lp); free(F->Li); free(F->Lx);
    free(F->Up); free(F->Ui); free(F->Ux);
    free(F->p); free(F->q); free(F->Rs);
    free(F);
}

/* UMFPACK default scaling: divide each row by the sum of absolute values of its entries. */
void oracle_row_scale_sum(i64 n, const i64 *Ap, const i64 *Ai, const double *Ax, double *Rs) {
    for (i64 i = 0; i < n; ++i) Rs[i] = 0.0;
    for (i64 j = 0; j < n; ++j)
        for (i64 t = Ap[j]; t < Ap[j + 1]; ++t) Rs[Ai[t]] += fabs(Ax[t]);
    for (i64 i = 0; i < n; ++i) Rs[i] = (Rs[i] > 0.0) ? 1.0 / Rs[i] : 1.0;
}

static int grow(i64 **I, double **X, i64 *cap, i64 need) {
    if (need <= *cap) return 0;
    i64 nc = *cap * 2; if (nc < need) nc = need + 1024;
    i64 *ni = (i64 *)realloc(*I, (size_t)nc * sizeof(i64));
    if (!ni) return -1; *I = ni;
    double *nx = (double *)realloc(*X, (size_t)nc * sizeof(double));
    if (!nx) return -1; *X = nx;
    *cap = nc; return 0;
}

/*
 * Factorize.  q (length n) must be given (column order).  Rs may be NULL (=> ones).
 * pivot_mode 0: static -- p (length n) is given and obeyed exactly.
 * pivot_mode 1: threshold partial pivoting; p_in ignored.  The diagonal candidate
 *               (original row == q[k]) is kept when |x_diag| >= diag_tol * max|x|;
 *               otherwise the largest entry wins (ties: smallest original row index).
 * During elimination rows keep ORIGINAL numbering; pinv maps them to pivot positions.
 */
oracle_lu_t *oracle_lu_factor(i64 n, const i64 *Ap, const i64 *Ai, const double *Ax,
                              const i64 *p_in, const i64 *q_in, const double *Rs_in,
                              int pivot_mode, double diag_tol) {
    oracle_lu_t *F = (oracle_lu_t *)calloc(1, sizeof(oracle_lu_t));
    if (!F) return NULL;
    F->n = n; F->bad_col = -1;
    i64 annz = Ap[n];
    F->lcap = 4 * annz + n + 16; F->ucap = 4 * annz + n + 16;
    F->Lp = (i64 *)calloc((size_t)n + 1, sizeof(i64));
    F->Up = (i64 *)calloc((size_t)n + 1, sizeof(i64));
    F->Li = (i64 *)malloc((size_t)F->lcap * sizeof(i64));
    F->Ui = (i64 *)malloc((size_t)F->ucap * sizeof(i64));
    F->Lx = (double *)malloc((size_t)F->lcap * sizeof(double));
    F->Ux = (double *)malloc((size_t)F->ucap * sizeof(double));
    F->p = (i64 *)malloc((size_t)(n + 1) * sizeof(i64));
    F->q = (i64 *)malloc((size_t)(n + 1) * sizeof(i64));
    F->Rs = (double *)malloc((size_t)(n + 1) * sizeof(double));
    i64 *pinv = (i64 *)malloc((size_t)(n + 1) * sizeof(i64));    /* orig row -> pivot step or -1 */
    double *x = (double *)calloc((size_t)n + 1, sizeof(double)); /* dense accumulator, orig rows */
    char *mark = (char *)calloc((size_t)n + 1, 1);
    i64 *patt = (i64 *)malloc((size_t)(n + 1) * sizeof(i64));    /* all touched orig rows        */
    i64 *piv_list = (i64 *)malloc((size_t)(n + 1) * sizeof(i64)); /* pivot steps reached          */
    i64 *stack = (i64 *)malloc((size_t)(n + 1) * sizeof(i64));
    i64 *spos = (i64 *)malloc((size_t)(n + 1) * sizeof(i64));
    if (!F->Lp || !F->Up || !F->Li || !F->Ui || !F->Lx || !F->Ux || !F->p || !F->q || !F->Rs ||
        !pinv || !x || !mark || !patt || !piv_list || !stack || !spos) {
        oracle_lu_free(F); F = NULL; goto done;
    }
    for (i64 i = 0; i < n; ++i) {
        pinv[i] = -1;
        F->p[i] = -1;
        F->q[i] = q_in[i];
        F->Rs[i] = Rs_in ? Rs_in[i] : 1.0;
    }

    for (i64 k = 0; k < n; ++k) {
        i64 col = F->q[k];
        i64 np = 0, nk = 0;
        /* ---- symbolic: rows reachable from the entries of A(:,col) through columns of L ---- */
        for (i64 t = Ap[col]; t < Ap[col + 1]; ++t) {
            i64 r0 = Ai[t];
            if (mark[r0]) continue;
            /* iterative DFS; a row that is already pivotal expands into its L column */
            i64 sp = 0; stack[0] = r0; mark[r0] = 1; patt[np++] = r0;
            spos[0] = (pinv[r0] >= 0) ? F->Lp[pinv[r0]] : -1;
            if (pinv[r0] >= 0) piv_list[nk++] = pinv[r0];
            while (sp >= 0) {
                i64 r = stack[sp];
                i64 kk = pinv[r];
                int descended = 0;
                if (kk >= 0) {
                    i64 end = F->Lp[kk + 1];
                    while (spos[sp] < end) {
                        i64 c = F->Li[spos[sp]++];
                        if (mark[c]) continue;
                        mark[c] = 1; patt[np++] = c;
                        if (pinv[c] >= 0) piv_list[nk++] = pinv[c];
                        ++sp; stack[sp] = c;
                        spos[sp] = (pinv[c] >= 0) ? F->Lp[pinv[c]] : -1;
                        descended = 1;
                        break;
                    }
                }
                if (!descended) --sp;
            }
        }
        /* ---- numeric: x = scaled column, then eliminate with earlier pivots, ascending ---- */
        for (i64 t = Ap[col]; t < Ap[col + 1]; ++t) x[Ai[t]] = F->Rs[Ai[t]] * Ax[t];
        qsort(piv_list, (size_t)nk, sizeof(i64), cmp_i64);
        for (i64 a = 0; a < nk; ++a) {
            i64 kk = piv_list[a];
            double ukj = x[F->p[kk]];
            /* L column kk holds (orig row, multiplier) for the rows below the pivot only;
               the unit diagonal is added when the factors are exported.                  */
            for (i64 t = F->Lp[kk]; t < F->Lp[kk + 1]; ++t) x[F->Li[t]] -= F->Lx[t] * ukj;
            F->flops += 2.0 * (double)(F->Lp[kk + 1] - F->Lp[kk]);
        }
        /* ---- pivot choice ---- */
        i64 piv = -1;
        if (pivot_mode == 0) {
            piv = p_in[k];
            if (pinv[piv] >= 0) { F->bad_col = k; break; }   /* p is not a permutation */
            if (!mark[piv]) { mark[piv] = 1; patt[np++] = piv; x[piv] = 0.0; }
        } else {
            double amax = -1.0; i64 imax = -1;
            for (i64 a = 0; a < np; ++a) {
                i64 r = patt[a];
                if (pinv[r] >= 0) continue;
                double v = fabs(x[r]);
                if (v > amax || (v == amax && r < imax)) { amax = v; imax = r; }
            }
            piv = imax;
            if (piv >= 0 && pinv[col] < 0 && mark[col] && fabs(x[col]) >= diag_tol * amax &&
                fabs(x[col]) > 0.0)
                piv = col;
            if (piv < 0) { F->bad_col = k; break; }
        }
        double pv = x[piv];
        if (!(fabs(pv) > 0.0) || pv != pv) { F->bad_col = k; }   /* keep going: inf/nan propagate */
        F->p[k] = piv;
        /* ---- store U(:,k) (pivotal rows incl. the new pivot) and L(:,k) (the rest) ---- */
        if (grow(&F->Ui, &F->Ux, &F->ucap, F->unz + nk + 1) ||
            grow(&F->Li, &F->Lx, &F->lcap, F->lnz + np + 1)) { F->bad_col = -2; break; }
        for (i64 a = 0; a < nk; ++a) {          /* piv_list is sorted => U rows sorted */
            F->Ui[F->unz] = piv_list[a];
            F->Ux[F->unz++] = x[F->p[piv_list[a]]];
        }
        F->Ui[F->unz] = k; F->Ux[F->unz++] = pv;
        F->Up[k + 1] = F->unz;
        for (i64 a = 0; a < np; ++a) {
            i64 r = patt[a];
            if (pinv[r] >= 0 || r == piv) continue;
            F->Li[F->lnz] = r;                   /* original row for now */
            F->Lx[F->lnz++] = x[r] / pv;
            F->flops += 1.0;
        }
        F->Lp[k + 1] = F->lnz;
        pinv[piv] = k;
        for (i64 a = 0; a < np; ++a) { x[patt[a]] = 0.0; mark[patt[a]] = 0; }
        if (F->bad_col >= 0 && pivot_mode == 0 && !(fabs(pv) > 0.0)) break;
    }
    /* stopped early (singular / bad permutation): keep the column pointers well formed */
    for (i64 k = 0; k < n; ++k) {
        if (F->Lp[k + 1] < F->Lp[k]) F->Lp[k + 1] = F->Lp[k];
        if (F->Up[k + 1] < F->Up[k]) F->Up[k + 1] = F->Up[k];
    }
done:
    free(pinv); free(x); free(mark); free(patt); free(piv_list); free(stack); free(spos);
    return F;
}

i64 oracle_lu_n(const oracle_lu_t *F) { return F->n; }
i64 oracle_lu_bad_col(const oracle_lu_t *F) { return F->bad_col; }
double oracle_lu_flops(const oracle_lu_t *F) { return F->flops; }
/* nnz of the exported factors; L includes its explicit unit diagonal (reference src:47-48:
   F.L as returned by UMFPACK carries the unit diagonal).                                  */
i64 oracle_lu_nnzL(const oracle_lu_t *F) { return F->lnz + F->n; }
i64 oracle_lu_nnzU(const oracle_lu_t *F) { return F->unz; }

/* Export in permuted numbering, CSC, rows sorted, L with explicit unit diagonal. */
void oracle_lu_export(const oracle_lu_t *F, i64 *Lp, i64 *Li, double *Lx,
                      i64 *Up, i64 *Ui, double *Ux, i64 *p, i64 *q, double *Rs) {
    i64 n = F->n;
    i64 *pinv = (i64 *)malloc((size_t)(n + 1) * sizeof(i64));
    for (i64 k = 0; k < n; ++k) pinv[F->p[k]] = k;
    typedef struct { i64 r; double v; } ent;
    i64 maxc = 1;
    for (i64 k = 0; k < n; ++k) if (F->Lp[k + 1] - F->Lp[k] + 1 > maxc) maxc = F->Lp[k + 1] - F->Lp[k] + 1;
    i64 *ord = (i64 *)malloc((size_t)maxc * 2 * sizeof(i64));
    i64 w = 0;
    Lp[0] = 0;
    for (i64 k = 0; k < n; ++k) {
        i64 cnt = F->Lp[k + 1] - F->Lp[k];
        /* sort (permuted row, slot) pairs encoded in one i64 array */
        for (i64 a = 0; a < cnt; ++a) { ord[2 * a] = pinv[F->Li[F->Lp[k] + a]]; ord[2 * a + 1] = F->Lp[k] + a; }
        qsort(ord, (size_t)cnt, 2 * sizeof(i64), cmp_i64);
        Li[w] = k; Lx[w++] = 1.0;
        for (i64 a = 0; a < cnt; ++a) { Li[w] = ord[2 * a]; Lx[w++] = F->Lx[ord[2 * a + 1]]; }
        Lp[k + 1] = w;
    }
    for (i64 k = 0; k <= n; ++k) Up[k] = F->Up[k];
    for (i64 t = 0; t < F->unz; ++t) { Ui[t] = F->Ui[t]; Ux[t] = F->Ux[t]; }
    for (i64 k = 0; k < n; ++k) { p[k] = F->p[k]; q[k] = F->q[k]; Rs[k] = F->Rs[k]; }
    free(ord); free(pinv);
}

/* ---------- plain sparse triangular solves on exported CSC factors (the reference's
 * test oracle `F.L \ b`, `F.U \ b`, test/runtests.jl:51,70,86,104) ---------------------- */
void oracle_csc_lsolve(i64 n, const i64 *Lp, const i64 *Li, const double *Lx, double *x) {
    for (i64 j = 0; j < n; ++j) {
        /* diagonal is first entry (sorted rows) and is 1 */
        double xj = x[j] / Lx[Lp[j]];
        x[j] = xj;
        for (i64 t = Lp[j] + 1; t < Lp[j + 1]; ++t) x[Li[t]] -= Lx[t] * xj;
    }
}
void oracle_csc_usolve(i64 n, const i64 *Up, const i64 *Ui, const double *Ux, double *x) {
    for (i64 j = n - 1; j >= 0; --j) {
        double xj = x[j] / Ux[Up[j + 1] - 1];   /* diagonal is last entry */
        x[j] = xj;
        for (i64 t = Up[j]; t < Up[j + 1] - 1; ++t) x[Ui[t]] -= Ux[t] * xj;
    }
}
/* x = A \ b through the factors, in the reference's order of operations (src:322-339). */
void oracle_lu_solve(i64 n, const i64 *Lp, const i64 *Li, const double *Lx,
                     const i64 *Up, const i64 *Ui, const double *Ux,
                     const i64 *p, const i64 *q, const double *Rs,
                     const double *b, double *x, double *wrk) {
    for (i64 i = 0; i < n; ++i) wrk[i] = Rs[p[i]] * b[p[i]];
    oracle_csc_lsolve(n, Lp, Li, Lx, wrk);
    oracle_csc_usolve(n, Up, Ui, Ux, wrk);
    for (i64 i = 0; i < n; ++i) x[q[i]] = wrk[i];
}

// oracle/ref_mf.cpp -- TEST / BASELINE INFRASTRUCTURE ONLY (never linked into libsmslu.so).
//
// The CPU side of the comparison: a multifrontal sparse LU with static pivots on the host cores,
// doing on the CPU what UMFPACK does for the reference inside `lu!` (reference
// src/SharedMemSparseLU.jl:247: numeric refactorization with the symbolic analysis reused) and
// what the reference's lsolve!/rsolve! do (src:349-392), with
//   * the SAME ordering / supernodes as the GPU path (csrc/symbolic.cpp is compiled into this
//     library too: host-only analysis code, no kernels, libsmslu.so is not loaded),
//   * BLAS-3 in the fronts (dgemm / dtrsm of the OpenBLAS that SciPy ships, handed in as function
//     pointers), like UMFPACK's frontal kernels,
//   * all host cores: OpenMP over the fronts of a level for the small fronts, threaded BLAS for
//     the big ones.
// PARITY UNPINNED (as oracle/ref_lu.c): UMFPACK itself is not installed here.  This port is pinned
// against oracle/ref_lu.c entry by entry in tests/test_oracle.py.
#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>

#include "../sharedmemsparselu.jl_b200/csrc/symbolic.hpp"

using namespace smslu;

typedef void (*dgemm_t)(char*, char*, int*, int*, int*, double*, double*, int*, double*, int*, double*, double*, int*);
typedef void (*dtrsm_t)(char*, char*, char*, char*, int*, int*, double*, double*, int*, double*, int*);
typedef void (*setthr_t)(int);

namespace {

struct MF {
    Symbolic S;
    std::vector<int64_t> Ap, Ai;
    std::vector<double> lu, Rs, upd;
    std::vector<double*> cb;          // contribution block of every front: a slot of cbpool (arena planned by the analysis), null once consumed
    std::vector<double> cbpool;
    dgemm_t dgemm = nullptr;
    dtrsm_t dtrsm = nullptr;
    setthr_t blas_threads = nullptr;   // openblas_set_num_threads of the same library
    int nthreads = 1;
    int64_t bad = -1;
    double t_factor = 0, t_solve = 0;
    std::string err;
};

constexpr int BIG_F = 1536;   // fronts with at least this many rows: one front at a time, threaded BLAS, OpenMP assembly
constexpr int BLAS_F = 96;    // smaller fronts (one per thread): BLAS-3 on the calling thread from this size on, plain loops below
constexpr int PB = 128;       // panel width of the blocked pivot-block factorization

double now_s() {
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

// extend-add of child c's contribution block into the parent's P / T / C
void extend_add(const Symbolic& S, int c, const double* Cc, int64_t k, int64_t r, double* P, double* T, double* C, bool par) {
    const int64_t rc = S.rows_ptr[c + 1] - S.rows_ptr[c], f = k + r;
    const int* rel = S.rel.data() + S.rows_ptr[c];
#pragma omp parallel for schedule(static) if (par)
    for (int64_t b = 0; b < rc; ++b) {
        const int64_t rb = rel[b];
        const double* src = Cc + b * rc;
        if (rb < k) {
            double* d = P + rb * f;
            for (int64_t a = 0; a < rc; ++a) d[rel[a]] += src[a];
        } else {
            double* dc = C + (rb - k) * r - k;
            for (int64_t a = 0; a < rc; ++a) {
                const int64_t ra = rel[a];
                if (ra < k) T[(rb - k) + ra * r] += src[a];
                else dc[ra] += src[a];
            }
        }
    }
}

// unblocked right-looking elimination of columns [j0, j1) of the f x k panel P, rows [j0, f)
// (rows [j0, rend): rend = f for the whole panel, rend = j1 for the diagonal block only)
void panel_unblocked(double* P, int64_t f, int64_t j0, int64_t j1, int c0, int64_t& bad, double& growth, int64_t rend = -1) {
    const int64_t fe = rend < 0 ? f : rend;
    for (int64_t j = j0; j < j1; ++j) {
        const double piv = P[j + j * f];
        if (!(std::fabs(piv) > 0.0) || !std::isfinite(piv)) { if (bad < 0 || c0 + j < bad) bad = c0 + j; }
        const double rinv = 1.0 / piv;
        double* cj = P + j * f;
        double g = 0.0;
        for (int64_t i = j + 1; i < fe; ++i) { cj[i] *= rinv; g = std::max(g, std::fabs(cj[i])); }
        growth = std::max(growth, g);
        for (int64_t c = j + 1; c < j1; ++c) {
            const double u = P[j + c * f];
            double* cc = P + c * f;
            for (int64_t i = j + 1; i < fe; ++i) cc[i] -= cj[i] * u;
        }
    }
}

void factor_front(MF& h, int s, int64_t& bad, double& growth, bool big) {
    const Symbolic& S = h.S;
    const int c0 = S.sn_start[s];
    const int64_t k = S.sn_start[s + 1] - c0, r = S.rows_ptr[s + 1] - S.rows_ptr[s], f = k + r;
    double* P = h.lu.data() + S.Loff[s];
    double* T = h.lu.data() + S.Uoff[s];
    double* C = r > 0 ? h.cbpool.data() + S.CBoff[s] : nullptr;
    h.cb[s] = C;
    if (big) {
#pragma omp parallel for schedule(static)
        for (int64_t e = 0; e < r * r; ++e) C[e] = 0.0;
    } else if (r > 0) memset(C, 0, sizeof(double) * r * r);
    for (int ci = S.child_ptr[s]; ci < S.child_ptr[s + 1]; ++ci) {
        const int c = S.child_idx[ci];
        if (!h.cb[c]) continue;
        extend_add(S, c, h.cb[c], k, r, P, T, C, big);
        h.cb[c] = nullptr;
    }
    if (f >= BLAS_F && h.dgemm && h.dtrsm) {
        char N = 'N', Tr = 'T', L = 'L', R = 'R', U = 'U';
        double one = 1.0, mone = -1.0;
        for (int64_t j0 = 0; j0 < k; j0 += PB) {
            const int64_t j1 = std::min(k, j0 + PB);
            panel_unblocked(P, f, j0, j1, c0, bad, growth, j1);          // the diagonal block ...
            if (j1 < f) {                                                 // ... then the rows below it: L = A U_bb^{-1}
                int mr = (int)(f - j1), nb = (int)(j1 - j0), ld = (int)f;
                h.dtrsm(&R, &U, &N, &N, &mr, &nb, &one, P + j0 + j0 * f, &ld, P + j1 + j0 * f, &ld);
            }
            if (j1 < k) {
                int m = (int)(j1 - j0), n = (int)(k - j1), ld = (int)f, mr = (int)(f - j1);
                // U block row: P[j0:j1, j1:k] <- L_bb^{-1} P[j0:j1, j1:k]
                h.dtrsm(&L, &L, &N, &U, &m, &n, &one, P + j0 + j0 * f, &ld, P + j0 + j1 * f, &ld);
                // trailing part of the panel: P[j1:f, j1:k] -= P[j1:f, j0:j1] P[j0:j1, j1:k]
                h.dgemm(&N, &N, &mr, &n, &m, &mone, P + j1 + j0 * f, &ld, P + j0 + j1 * f, &ld, &one, P + j1 + j1 * f, &ld);
            }
        }
        if (r > 0) {
            int m = (int)r, n = (int)k, ldp = (int)f, ldt = (int)r;
            // U12' = A12' L11^{-T}
            h.dtrsm(&R, &L, &Tr, &U, &m, &n, &one, P, &ldp, T, &ldt);
            // C -= L21 U12
            h.dgemm(&N, &Tr, &m, &m, &n, &mone, P + k, &ldp, T, &ldt, &one, C, &ldt);
        }
        double g = 0.0;
#pragma omp parallel for schedule(static) reduction(max : g) if (big)
        for (int64_t j = 0; j < k; ++j)
            for (int64_t i = j + 1; i < f; ++i) g = std::max(g, std::fabs(P[i + j * f]));
        growth = std::max(growth, g);
    } else {
        panel_unblocked(P, f, 0, k, c0, bad, growth);
        for (int64_t j = 0; j < k; ++j)                      // U12 (stored transposed)
            for (int64_t i = j + 1; i < k; ++i) {
                const double l = P[i + j * f];
                const double* tj = T + j * r;
                double* ti = T + i * r;
                for (int64_t a = 0; a < r; ++a) ti[a] -= l * tj[a];
            }
        for (int64_t p = 0; p < k; ++p) {
            const double* lp = P + k + p * f;
            const double* tp = T + p * r;
            for (int64_t b = 0; b < r; ++b) {
                const double u = tp[b];
                double* cc = C + b * r;
                for (int64_t a = 0; a < r; ++a) cc[a] -= lp[a] * u;
            }
        }
    }
}

}  // namespace

extern "C" {

void* mf_create(int64_t n, const int64_t* Ap, const int64_t* Ai, int ordering, const int* grid, int nthreads) {
    MF* h = new MF();
    h->Ap.assign(Ap, Ap + n + 1);
    h->Ai.assign(Ai, Ai + Ap[n]);
    SymOptions o;
    o.ordering = ordering;
    if (grid) for (int d = 0; d < 3; ++d) o.grid[d] = grid[d];
    // same ordering as the GPU path, but a wide separator stays ONE front (the GPU layout chains it into
    // 128-column fronts whose trailing matrix is handed on in the GEMM epilogue; on the CPU that hand-over
    // would be a copy of the whole trailing matrix per 128 columns)
    o.max_width = 1 << 30;
    int rc = analyze((int)n, Ap, Ai, nullptr, nullptr, o, h->S, h->err);
    if (rc != 0) { fprintf(stderr, "mf_create: %s\n", h->err.c_str()); delete h; return nullptr; }
    h->nthreads = nthreads > 0 ? nthreads : omp_get_max_threads();
    h->cb.assign(h->S.nsn, nullptr);
    return h;
}

void mf_set_blas(void* hv, void* dgemm, void* dtrsm, void* set_threads) {
    MF* h = (MF*)hv;
    h->dgemm = (dgemm_t)dgemm;
    h->dtrsm = (dtrsm_t)dtrsm;
    h->blas_threads = (setthr_t)set_threads;
}

void mf_free(void* hv) {
    MF* h = (MF*)hv;
    delete h;
}

// info: n, nsn, nlevels, nnzL_exact, lu_size, threads, big fronts, max_front ; flops_exact
void mf_info(void* hv, int64_t* out, double* flops) {
    MF* h = (MF*)hv;
    const Symbolic& S = h->S;
    int64_t nbig = 0;
    for (int s = 0; s < S.nsn; ++s) nbig += (S.sn_start[s + 1] - S.sn_start[s]) + (S.rows_ptr[s + 1] - S.rows_ptr[s]) >= BIG_F;
    int64_t v[] = {S.n, S.nsn, S.nlevels, S.nnzL_exact, S.lu_size, h->nthreads, nbig, S.max_front};
    memcpy(out, v, sizeof v);
    flops[0] = S.flops_exact;
}

void mf_perm(void* hv, int64_t* p, int64_t* q) {
    const Symbolic& S = ((MF*)hv)->S;
    for (int k = 0; k < S.n; ++k) { p[k] = S.p[k]; q[k] = S.q[k]; }
}

// Numeric factorization (`lu!`): Rs given, or NULL for UMFPACK's default row scaling Rs[i] = 1 / sum_j |a_ij|.
// Returns the permuted column of the first zero / non-finite pivot, or -1; *growth = max |l_ij|.
int64_t mf_factor(void* hv, const double* Ax, const double* Rs, double* growth_out) {
    MF* h = (MF*)hv;
    const Symbolic& S = h->S;
    const int n = S.n;
    const double t0 = now_s();
    omp_set_num_threads(h->nthreads);
    h->Rs.assign(n, 0.0);
    if (Rs) h->Rs.assign(Rs, Rs + n);
    else {
        for (int c = 0; c < n; ++c)                               // ascending column per row: same order as the oracle
            for (int64_t t = h->Ap[c]; t < h->Ap[c + 1]; ++t) h->Rs[h->Ai[t]] += std::fabs(Ax[t]);
        for (int i = 0; i < n; ++i) h->Rs[i] = h->Rs[i] > 0.0 ? 1.0 / h->Rs[i] : 1.0;
    }
    if ((int64_t)h->lu.size() != S.lu_size) h->lu.resize(S.lu_size);
    {
        double* lu = h->lu.data();
        const int64_t tot = S.lu_size;
#pragma omp parallel for schedule(static)
        for (int64_t e = 0; e < tot; ++e) lu[e] = 0.0;
#pragma omp parallel for schedule(static)
        for (int c = 0; c < n; ++c)
            for (int64_t t = h->Ap[c]; t < h->Ap[c + 1]; ++t) lu[S.a_dst[t]] = h->Rs[h->Ai[t]] * Ax[t];
    }
    std::fill(h->cb.begin(), h->cb.end(), nullptr);
    if ((int64_t)h->cbpool.size() != S.cb_size) h->cbpool.resize(S.cb_size);
    int64_t bad = -1;
    double growth = 0.0;
    auto F = [&](int s) { return (int64_t)(S.sn_start[s + 1] - S.sn_start[s]) + (S.rows_ptr[s + 1] - S.rows_ptr[s]); };
    for (int l = 0; l < S.nlevels; ++l) {
        const int* sn = S.level_sn.data() + S.level_ptr[l];
        const int cnt = S.level_ptr[l + 1] - S.level_ptr[l];
        // small and medium fronts of the level: one front per thread (BLAS calls stay on the calling thread)
        bool any_small = false, any_big = false;
        for (int t = 0; t < cnt; ++t) { if (F(sn[t]) < BIG_F) any_small = true; else any_big = true; }
        if (any_small && h->blas_threads) h->blas_threads(1);
        if (any_small)
#pragma omp parallel
        {
            int64_t mybad = -1;
            double myg = 0.0;
#pragma omp for schedule(dynamic, 4) nowait
            for (int t = 0; t < cnt; ++t)
                if (F(sn[t]) < BIG_F) factor_front(*h, sn[t], mybad, myg, false);
#pragma omp critical
            {
                if (mybad >= 0 && (bad < 0 || mybad < bad)) bad = mybad;
                growth = std::max(growth, myg);
            }
        }
        // big fronts: one at a time, the cores are inside BLAS
        if (any_big && h->blas_threads) h->blas_threads(h->nthreads);
        for (int t = 0; t < cnt; ++t)
            if (F(sn[t]) >= BIG_F) factor_front(*h, sn[t], bad, growth, true);
    }
    h->bad = bad;
    h->t_factor = now_s() - t0;
    if (growth_out) *growth_out = growth;
    return bad;
}

// `ldiv!(x, F, b)` (reference src:286-342): w = (Rs .* b)[p]; L w; U w; x[q] = w.
void mf_solve(void* hv, const double* b, double* x) {
    MF* h = (MF*)hv;
    const Symbolic& S = h->S;
    const int n = S.n;
    const double t0 = now_s();
    omp_set_num_threads(h->nthreads);
    std::vector<double> w(n);
    for (int i = 0; i < n; ++i) w[i] = h->Rs[S.p[i]] * b[S.p[i]];
    h->upd.assign(S.sum_r, 0.0);
    const double* lu = h->lu.data();
    double* upd = h->upd.data();
    for (int l = 0; l < S.nlevels; ++l) {                        // forward: pull the children's update vectors
        const int lo = S.level_ptr[l], hi = S.level_ptr[l + 1];
#pragma omp parallel for schedule(dynamic, 4)
        for (int u = lo; u < hi; ++u) {
            const int s = S.level_sn[u];
            const int c0 = S.sn_start[s];
            const int64_t k = S.sn_start[s + 1] - c0, r = S.rows_ptr[s + 1] - S.rows_ptr[s], f = k + r;
            const double* P = lu + S.Loff[s];
            double* us = upd + S.rows_ptr[s];
            for (int ci = S.child_ptr[s]; ci < S.child_ptr[s + 1]; ++ci) {
                const int c = S.child_idx[ci];
                const int64_t rc = S.rows_ptr[c + 1] - S.rows_ptr[c];
                const int* rel = S.rel.data() + S.rows_ptr[c];
                const double* uc = upd + S.rows_ptr[c];
                for (int64_t a = 0; a < rc; ++a) {
                    if (rel[a] < k) w[c0 + rel[a]] += uc[a];
                    else us[rel[a] - k] += uc[a];
                }
            }
            for (int64_t j = 0; j < k; ++j) {
                const double xj = w[c0 + j];
                const double* col = P + j * f;
                for (int64_t i = j + 1; i < k; ++i) w[c0 + i] -= col[i] * xj;
                for (int64_t a = 0; a < r; ++a) us[a] -= col[k + a] * xj;
            }
        }
    }
    for (int l = S.nlevels - 1; l >= 0; --l) {                   // backward
        const int lo = S.level_ptr[l], hi = S.level_ptr[l + 1];
#pragma omp parallel for schedule(dynamic, 4)
        for (int u = lo; u < hi; ++u) {
            const int s = S.level_sn[u];
            const int c0 = S.sn_start[s];
            const int64_t k = S.sn_start[s + 1] - c0, r = S.rows_ptr[s + 1] - S.rows_ptr[s], f = k + r;
            const double* P = lu + S.Loff[s];
            const double* T = lu + S.Uoff[s];
            const int* rows = S.rows.data() + S.rows_ptr[s];
            for (int64_t i = k - 1; i >= 0; --i) {
                double acc = w[c0 + i];
                const double* ti = T + i * r;
                for (int64_t a = 0; a < r; ++a) acc -= ti[a] * w[rows[a]];
                for (int64_t j = i + 1; j < k; ++j) acc -= P[i + j * f] * w[c0 + j];
                w[c0 + i] = acc / P[i + i * f];
            }
        }
    }
    for (int i = 0; i < n; ++i) x[S.q[i]] = w[i];
    h->t_solve = now_s() - t0;
}

void mf_times(void* hv, double* out) {
    MF* h = (MF*)hv;
    out[0] = h->t_factor;
    out[1] = h->t_solve;
}

void mf_get_factors(void* hv, int64_t* Lp, int64_t* Li, double* Lx, int64_t* Up, int64_t* Ui, double* Ux, double* Rs) {
    MF* h = (MF*)hv;
    std::vector<int64_t> ptr;
    std::vector<int> idx;
    exact_structure(h->S, h->Ap.data(), h->Ai.data(), ptr, idx);
    export_factors(h->S, ptr, idx, h->lu.data(), 0, Lp, Li, Lx, Up, Ui, Ux);
    if (Rs) memcpy(Rs, h->Rs.data(), sizeof(double) * h->S.n);
}

}  // extern "C"

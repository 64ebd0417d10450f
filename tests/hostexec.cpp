// tests/hostexec.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A sequential CPU walk over exactly the data structures the GPU kernels consume (the Symbolic
// layout of csrc/symbolic.cpp: panels, contribution blocks, rel maps, A scatter map, levels).
// It lets the CPU test-suite validate the host-side symbolic analysis end to end (against the
// oracle) without a GPU.  It is never linked into libsmslu.so and the product never calls it.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../sharedmemsparselu.jl_b200/csrc/symbolic.hpp"

using namespace smslu;

struct HX {
    Symbolic S;
    std::vector<int64_t> Ap, Ai;
    std::vector<double> lu, cb, Rs, upd;
    std::string err;
};

extern "C" {

void* hx_create(int64_t n, const int64_t* Ap, const int64_t* Ai, int ordering, const int* grid,
                int nd_leaf, int relax, int max_width, const int64_t* p, const int64_t* q) {
    HX* h = new HX();
    h->Ap.assign(Ap, Ap + n + 1);
    h->Ai.assign(Ai, Ai + Ap[n]);
    SymOptions o;
    o.ordering = ordering;
    if (grid) for (int d = 0; d < 3; ++d) o.grid[d] = grid[d];
    if (nd_leaf > 0) o.nd_leaf = nd_leaf;
    o.relax = relax;
    if (max_width > 0) o.max_width = max_width;
    std::vector<int> pp, qq;
    if (p && q) { pp.assign(p, p + n); qq.assign(q, q + n); }
    int rc = analyze((int)n, Ap, Ai, pp.empty() ? nullptr : pp.data(), qq.empty() ? nullptr : qq.data(), o, h->S, h->err);
    if (rc != 0) { fprintf(stderr, "hx_create: %s\n", h->err.c_str()); delete h; return nullptr; }
    return h;
}

void hx_free(void* hv) { delete (HX*)hv; }

// info: n, nsn, nlevels, lu_size, cb_size, nnzL_exact, nnzL_stored, sum_r, max_front, max_k, max_children
void hx_info(void* hv, int64_t* out, double* flops) {
    const Symbolic& S = ((HX*)hv)->S;
    int64_t v[] = {S.n, S.nsn, S.nlevels, S.lu_size, S.cb_size, S.nnzL_exact, S.nnzL_stored,
                   S.sum_r, S.max_front, S.max_k, S.max_children};
    memcpy(out, v, sizeof(v));
    flops[0] = S.flops_exact;
    flops[1] = S.flops_stored;
}

void hx_perm(void* hv, int64_t* p, int64_t* q) {
    const Symbolic& S = ((HX*)hv)->S;
    for (int k = 0; k < S.n; ++k) { p[k] = S.p[k]; q[k] = S.q[k]; }
}

// Numeric multifrontal factorization, level by level, children in ascending order.
// Returns the permuted column of the first zero/non-finite pivot, or -1.
int64_t hx_factor(void* hv, const double* Ax, const double* Rs) {
    HX* h = (HX*)hv;
    const Symbolic& S = h->S;
    const int n = S.n;
    h->Rs.assign(n, 1.0);
    if (Rs) h->Rs.assign(Rs, Rs + n);
    h->lu.assign(S.lu_size, 0.0);
    h->cb.assign(S.cb_size, 0.0);
    std::vector<char> zeroed(S.nsn, 0);
    for (int c = 0; c < n; ++c)
        for (int64_t t = h->Ap[c]; t < h->Ap[c + 1]; ++t) h->lu[S.a_dst[t]] += h->Rs[h->Ai[t]] * Ax[t];
    int64_t bad = -1;
    for (int l = 0; l < S.nlevels; ++l)
        for (int u = S.level_ptr[l]; u < S.level_ptr[l + 1]; ++u) {
            const int s = S.level_sn[u];
            const int c0 = S.sn_start[s];
            const int64_t k = S.sn_start[s + 1] - c0, r = S.rows_ptr[s + 1] - S.rows_ptr[s], f = k + r;
            double* P = h->lu.data() + S.Loff[s];
            double* T = h->lu.data() + S.Uoff[s];
            double* C = h->cb.data() + S.CBoff[s];
            const bool has_children = S.child_ptr[s + 1] > S.child_ptr[s];
            // blocks that receive '+=' contributions are zeroed before the first one arrives: at
            // this level for extend-add children, one level earlier for a direct child (done there)
            if (has_children && !S.cb_assigned[s] && !zeroed[s]) { for (int64_t e = 0; e < r * r; ++e) C[e] = 0.0; zeroed[s] = 1; }
            // extend-add of the children that did not write directly
            for (int ci = S.child_ptr[s]; ci < S.child_ptr[s + 1]; ++ci) {
                const int c = S.child_idx[ci];
                if (S.direct[c]) continue;
                const int64_t rc = S.rows_ptr[c + 1] - S.rows_ptr[c];
                const int* rel = S.rel.data() + S.rows_ptr[c];
                const double* Cc = h->cb.data() + S.CBoff[c];
                for (int64_t b = 0; b < rc; ++b)
                    for (int64_t a = 0; a < rc; ++a) {
                        const double v = Cc[a + b * rc];
                        const int64_t ra = rel[a], rb = rel[b];
                        if (rb < k) P[ra + rb * f] += v;
                        else if (ra < k) T[(rb - k) + ra * r] += v;
                        else C[(ra - k) + (rb - k) * r] += v;
                    }
            }
            // dense partial factorization of the front
            for (int64_t j = 0; j < k; ++j) {
                const double piv = P[j + j * f];
                if (!(std::fabs(piv) > 0.0) || !std::isfinite(piv)) { if (bad < 0) bad = c0 + j; }
                for (int64_t i = j + 1; i < f; ++i) P[i + j * f] /= piv;
                for (int64_t c = j + 1; c < k; ++c) {          // columns inside the pivot block
                    const double ujc = P[j + c * f];
                    for (int64_t i = j + 1; i < f; ++i) P[i + c * f] -= P[i + j * f] * ujc;
                }
                for (int64_t a = 0; a < r; ++a) {              // U12 (stored transposed)
                    const double uja = T[a + j * r];
                    for (int64_t i = j + 1; i < k; ++i) T[a + i * r] -= P[i + j * f] * uja;
                }
            }
            // Schur update.  beta: the block holds assembled contributions; leaves start from 0.
            if (!has_children) for (int64_t e = 0; e < r * r; ++e) C[e] = 0.0;
            for (int64_t p = 0; p < k; ++p)
                for (int64_t b = 0; b < r; ++b) {
                    const double upb = T[b + p * r];
                    for (int64_t a = 0; a < r; ++a) C[a + b * r] -= P[(k + a) + p * f] * upb;
                }
            if (S.direct[s]) {   // hand the finished block straight to the parent
                const int ps = S.sn_parent[s];
                if (!S.cb_assigned[ps] && !zeroed[ps]) {
                    const int64_t pr0 = S.rows_ptr[ps + 1] - S.rows_ptr[ps];
                    double* PC0 = h->cb.data() + S.CBoff[ps];
                    for (int64_t e = 0; e < pr0 * pr0; ++e) PC0[e] = 0.0;
                    zeroed[ps] = 1;
                }
                const int64_t pk = S.sn_start[ps + 1] - S.sn_start[ps], pr = S.rows_ptr[ps + 1] - S.rows_ptr[ps], pf = pk + pr;
                double* PP = h->lu.data() + S.Loff[ps];
                double* PT = h->lu.data() + S.Uoff[ps];
                double* PC = h->cb.data() + S.CBoff[ps];
                const int* rel = S.rel.data() + S.rows_ptr[s];
                for (int64_t b = 0; b < r; ++b)
                    for (int64_t a = 0; a < r; ++a) {
                        const double v = C[a + b * r];
                        const int64_t ra = rel[a], rb = rel[b];
                        if (rb < pk) PP[ra + rb * pf] += v;
                        else if (ra < pk) PT[(rb - pk) + ra * pr] += v;
                        else if (S.cb_assigned[ps]) PC[(ra - pk) + (rb - pk) * pr] = v;
                        else PC[(ra - pk) + (rb - pk) * pr] += v;
                    }
            }
        }
    return bad;
}

void hx_lsolve(void* hv, double* x) {
    HX* h = (HX*)hv;
    const Symbolic& S = h->S;
    h->upd.assign(S.sum_r, 0.0);
    for (int l = 0; l < S.nlevels; ++l)
        for (int u = S.level_ptr[l]; u < S.level_ptr[l + 1]; ++u) {
            const int s = S.level_sn[u];
            const int c0 = S.sn_start[s];
            const int64_t k = S.sn_start[s + 1] - c0, r = S.rows_ptr[s + 1] - S.rows_ptr[s], f = k + r;
            const double* P = h->lu.data() + S.Loff[s];
            double* us = h->upd.data() + S.rows_ptr[s];
            for (int ci = S.child_ptr[s]; ci < S.child_ptr[s + 1]; ++ci) {
                const int c = S.child_idx[ci];
                const int64_t rc = S.rows_ptr[c + 1] - S.rows_ptr[c];
                const int* rel = S.rel.data() + S.rows_ptr[c];
                const double* uc = h->upd.data() + S.rows_ptr[c];
                for (int64_t a = 0; a < rc; ++a) {
                    if (rel[a] < k) x[c0 + rel[a]] += uc[a];
                    else us[rel[a] - k] += uc[a];
                }
            }
            for (int64_t j = 0; j < k; ++j) {
                const double xj = x[c0 + j];
                for (int64_t i = j + 1; i < k; ++i) x[c0 + i] -= P[i + j * f] * xj;
                for (int64_t a = 0; a < r; ++a) us[a] -= P[(k + a) + j * f] * xj;
            }
        }
}

void hx_rsolve(void* hv, double* x) {
    HX* h = (HX*)hv;
    const Symbolic& S = h->S;
    for (int l = S.nlevels - 1; l >= 0; --l)
        for (int u = S.level_ptr[l]; u < S.level_ptr[l + 1]; ++u) {
            const int s = S.level_sn[u];
            const int c0 = S.sn_start[s];
            const int64_t k = S.sn_start[s + 1] - c0, r = S.rows_ptr[s + 1] - S.rows_ptr[s], f = k + r;
            const double* P = h->lu.data() + S.Loff[s];
            const double* T = h->lu.data() + S.Uoff[s];
            const int* rows = S.rows.data() + S.rows_ptr[s];
            for (int64_t i = k - 1; i >= 0; --i) {
                double acc = x[c0 + i];
                for (int64_t a = 0; a < r; ++a) acc -= T[a + i * r] * x[rows[a]];
                for (int64_t j = i + 1; j < k; ++j) acc -= P[i + j * f] * x[c0 + j];
                x[c0 + i] = acc / P[i + i * f];
            }
        }
}

void hx_solve(void* hv, const double* b, double* x) {
    HX* h = (HX*)hv;
    const Symbolic& S = h->S;
    std::vector<double> w(S.n);
    for (int i = 0; i < S.n; ++i) w[i] = h->Rs[S.p[i]] * b[S.p[i]];
    hx_lsolve(hv, w.data());
    hx_rsolve(hv, w.data());
    for (int i = 0; i < S.n; ++i) x[S.q[i]] = w[i];
}

void hx_get_factors(void* hv, int64_t* Lp, int64_t* Li, double* Lx, int64_t* Up, int64_t* Ui, double* Ux) {
    HX* h = (HX*)hv;
    std::vector<int64_t> ptr;
    std::vector<int> idx;
    exact_structure(h->S, h->Ap.data(), h->Ai.data(), ptr, idx);
    export_factors(h->S, ptr, idx, h->lu.data(), 0, Lp, Li, Lx, Up, Ui, Ux);
}

}  // extern "C"

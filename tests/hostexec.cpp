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
    std::vector<double> lu, cb, Rs, upd, vbuf;
    std::vector<char> zeroed;
    std::vector<int64_t> voff;      // per supernode: offset of its slice of vbuf (interface fronts), else -1
    int64_t bad = -1;
    std::string err;
};

extern "C" {

void* hx_create2(int64_t n, const int64_t* Ap, const int64_t* Ai, int ordering, const int* grid,
                 int nd_leaf, int relax, int max_width, const int64_t* p, const int64_t* q, int nranks) {
    HX* h = new HX();
    h->Ap.assign(Ap, Ap + n + 1);
    h->Ai.assign(Ai, Ai + Ap[n]);
    SymOptions o;
    o.ordering = ordering;
    if (grid) for (int d = 0; d < 3; ++d) o.grid[d] = grid[d];
    if (nd_leaf > 0) o.nd_leaf = nd_leaf;
    o.relax = relax;
    if (max_width > 0) o.max_width = max_width;
    o.nranks = nranks > 0 ? nranks : 1;
    std::vector<int> pp, qq;
    if (p && q) { pp.assign(p, p + n); qq.assign(q, q + n); }
    int rc = analyze((int)n, Ap, Ai, pp.empty() ? nullptr : pp.data(), qq.empty() ? nullptr : qq.data(), o, h->S, h->err);
    if (rc != 0) { fprintf(stderr, "hx_create: %s\n", h->err.c_str()); delete h; return nullptr; }
    const Symbolic& S = h->S;
    h->voff.assign(S.nsn, -1);
    int64_t v = 0;
    for (int s = 0; s < S.nsn; ++s)
        if (S.iface[s]) { h->voff[s] = v; v += (S.sn_start[s + 1] - S.sn_start[s]) + (S.rows_ptr[s + 1] - S.rows_ptr[s]); }
    h->vbuf.assign(v, 0.0);
    return h;
}

void* hx_create(int64_t n, const int64_t* Ap, const int64_t* Ai, int ordering, const int* grid,
                int nd_leaf, int relax, int max_width, const int64_t* p, const int64_t* q) {
    return hx_create2(n, Ap, Ai, ordering, grid, nd_leaf, relax, max_width, p, q, 1);
}

void hx_free(void* hv) { delete (HX*)hv; }

// info: n, nsn, nlevels, lu_size, cb_size, nnzL_exact, nnzL_stored, sum_r, max_front, max_k, max_children,
//       nranks, lu_top_size, cb_xchg_size (exchange slots of the subtree roots), vbuf length, number of top supernodes
void hx_info(void* hv, int64_t* out, double* flops) {
    HX* h = (HX*)hv;
    const Symbolic& S = h->S;
    int64_t ntop = 0;
    for (int s = 0; s < S.nsn; ++s) ntop += S.owner[s] == -1;
    int64_t v[] = {S.n, S.nsn, S.nlevels, S.lu_size, S.cb_size, S.nnzL_exact, S.nnzL_stored,
                   S.sum_r, S.max_front, S.max_k, S.max_children,
                   S.nranks, S.lu_top_size, S.cb_xchg_size, (int64_t)h->vbuf.size(), ntop};
    memcpy(out, v, sizeof(v));
    flops[0] = S.flops_exact;
    flops[1] = S.flops_stored;
}

void hx_perm(void* hv, int64_t* p, int64_t* q) {
    const Symbolic& S = ((HX*)hv)->S;       // ORD_GIVEN: the caller's own p, q (see Symbolic::post)
    const bool given = !S.post.empty();
    for (int k = 0; k < S.n; ++k) { p[k] = given ? S.p_given[k] : S.p[k]; q[k] = given ? S.q_given[k] : S.q[k]; }
}

void hx_owner(void* hv, int64_t* owner) {
    const Symbolic& S = ((HX*)hv)->S;
    for (int s = 0; s < S.nsn; ++s) owner[s] = S.owner[s];
}

// raw views for the exchange steps of the partitioned walk: 0 = factor pool, 1 = contribution pool,
// 2 = forward-solve interface vectors
double* hx_buffer(void* hv, int which, int64_t* len) {
    HX* h = (HX*)hv;
    std::vector<double>& v = which == 0 ? h->lu : which == 1 ? h->cb : h->vbuf;
    *len = (int64_t)v.size();
    return v.data();
}

// One phase of the numeric multifrontal factorization as rank `rank` of the partition runs it:
// phase 0 = the supernodes the rank owns (plus their contributions into top fronts), phase 1 = the
// top of the tree (after the caller has summed lu[0:lu_top_size] and cb[0:cb_iface_size] over ranks).
// With nranks == 1 phase 0 is the whole factorization.  Children in ascending order.
// Returns the permuted column of the first zero/non-finite pivot met so far, or -1.
int64_t hx_factor_phase(void* hv, const double* Ax, const double* Rs, int rank, int phase) {
    HX* h = (HX*)hv;
    const Symbolic& S = h->S;
    const int n = S.n;
    const int mine = phase == 0 ? rank : -1;
    if (phase == 0) {
        h->Rs.assign(n, 1.0);
        if (Rs) h->Rs.assign(Rs, Rs + n);
        h->lu.assign(S.lu_size, 0.0);
        h->cb.assign(S.cb_size, 0.0);
        h->zeroed.assign(S.nsn, 0);
        h->bad = -1;
        std::vector<int> cinv(n);
        for (int k2 = 0; k2 < n; ++k2) cinv[S.q[k2]] = k2;
        for (int c = 0; c < n; ++c)
            for (int64_t t = h->Ap[c]; t < h->Ap[c + 1]; ++t) {
                const int s = S.a_sn[t], o = S.owner[s];
                bool mine = o == rank;
                if (o == -1) {   // top front: pivot columns belong to the panel owner, an entry of U12 to its column's owner
                    const int jj = cinv[c];
                    mine = (jj < S.sn_start[s + 1] ? S.top_owner[s] : S.col_owner[jj]) == rank;
                }
                if (mine) h->lu[S.a_dst[t]] += h->Rs[h->Ai[t]] * Ax[t];
            }
    }
    if (phase == 1) return h->bad;      // the distributed top is walked level by level: hx_top_panels / hx_top_update
    int64_t& bad = h->bad;
    std::vector<char>& zeroed = h->zeroed;
    for (int l = 0; l < S.nlevels; ++l)
        for (int u = S.level_ptr[l]; u < S.level_ptr[l + 1]; ++u) {
            const int s = S.level_sn[u];
            const bool own = S.owner[s] == mine;
            if (!own) continue;        // (a top parent pulls the subtree roots' blocks in hx_top_panels)
            const int c0 = S.sn_start[s];
            const int64_t k = S.sn_start[s + 1] - c0, r = S.rows_ptr[s + 1] - S.rows_ptr[s], f = k + r;
            double* P = h->lu.data() + S.Loff[s];
            double* T = h->lu.data() + S.Uoff[s];
            double* C = h->cb.data() + S.CBoff[s];
            const bool has_children = S.child_ptr[s + 1] > S.child_ptr[s];
            // blocks that receive '+=' contributions are zeroed before the first one arrives: at
            // this level for extend-add children, one level earlier for a direct child (done there);
            // interface blocks were zeroed with the pool and arrive summed over the ranks
            if (own && has_children && !S.cb_assigned[s] && !zeroed[s] && !S.iface[s]) { for (int64_t e = 0; e < r * r; ++e) C[e] = 0.0; zeroed[s] = 1; }
            // extend-add of the children that did not write directly
            for (int ci = S.child_ptr[s]; ci < S.child_ptr[s + 1]; ++ci) {
                const int c = S.child_idx[ci];
                if (S.direct[c] || S.owner[c] != mine) continue;
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
            if (!own) continue;
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

// ---- the distributed top of the tree (nranks > 1), mirroring the GPU schedule (api.cu: build_top_fac) -------------
// levels that hold top fronts
int hx_top_levels(void* hv, int64_t* out) {
    const Symbolic& S = ((HX*)hv)->S;
    int n = 0;
    for (int l = 0; l < S.nlevels; ++l) {
        bool any = false;
        for (int u = S.level_ptr[l]; u < S.level_ptr[l + 1]; ++u) any |= S.owner[S.level_sn[u]] == -1;
        if (any) out[n++] = l;
    }
    return n;
}
// (offset, length) of the panels P_s of the level's top fronts (kind 0) or of their U12' blocks T_s (kind 1)
int hx_top_segments(void* hv, int level, int kind, int64_t* out) {
    const Symbolic& S = ((HX*)hv)->S;
    int n = 0;
    for (int u = S.level_ptr[level]; u < S.level_ptr[level + 1]; ++u) {
        const int s = S.level_sn[u];
        if (S.owner[s] != -1) continue;
        const int64_t k = S.sn_start[s + 1] - S.sn_start[s], r = S.rows_ptr[s + 1] - S.rows_ptr[s];
        out[2 * n] = kind == 0 ? S.Loff[s] : S.Uoff[s];
        out[2 * n + 1] = kind == 0 ? (k + r) * k : r * k;
        ++n;
    }
    return n;
}
// Step 1 of a level: assembly of the destination columns this rank owns (children: top fronts and subtree roots, whose
// blocks the caller has delivered), then the panel owner factors the pivot block and L21 of its fronts.
void hx_top_panels(void* hv, int rank, int level) {
    HX* h = (HX*)hv;
    const Symbolic& S = h->S;
    for (int u = S.level_ptr[level]; u < S.level_ptr[level + 1]; ++u) {
        const int s = S.level_sn[u];
        if (S.owner[s] != -1) continue;
        const int c0 = S.sn_start[s];
        const int64_t k = S.sn_start[s + 1] - c0, r = S.rows_ptr[s + 1] - S.rows_ptr[s], f = k + r;
        double* P = h->lu.data() + S.Loff[s];
        double* T = h->lu.data() + S.Uoff[s];
        double* C = h->cb.data() + S.CBoff[s];
        const int* rows = S.rows.data() + S.rows_ptr[s];
        auto dstown = [&](int64_t pb) { return pb < k ? S.top_owner[s] : S.col_owner[rows[pb - k]]; };
        const bool has_children = S.child_ptr[s + 1] > S.child_ptr[s];
        if (has_children && !S.cb_assigned[s] && !h->zeroed[s]) { for (int64_t e = 0; e < r * r; ++e) C[e] = 0.0; h->zeroed[s] = 1; }
        for (int ci = S.child_ptr[s]; ci < S.child_ptr[s + 1]; ++ci) {
            const int c = S.child_idx[ci];
            if (S.direct[c]) continue;
            const int64_t rc = S.rows_ptr[c + 1] - S.rows_ptr[c];
            const int* rel = S.rel.data() + S.rows_ptr[c];
            const double* Cc = h->cb.data() + S.CBoff[c];
            for (int64_t b = 0; b < rc; ++b) {
                const int64_t rb = rel[b];
                if (dstown(rb) != rank) continue;
                for (int64_t a = 0; a < rc; ++a) {
                    const double v = Cc[a + b * rc];
                    const int64_t ra = rel[a];
                    if (rb < k) P[ra + rb * f] += v;
                    else if (ra < k) T[(rb - k) + ra * r] += v;
                    else C[(ra - k) + (rb - k) * r] += v;
                }
            }
        }
        if (S.top_owner[s] != rank) continue;
        for (int64_t j = 0; j < k; ++j) {
            const double piv = P[j + j * f];
            if (!(std::fabs(piv) > 0.0) || !std::isfinite(piv)) { if (h->bad < 0) h->bad = c0 + j; }
            for (int64_t i = j + 1; i < f; ++i) P[i + j * f] /= piv;
            for (int64_t c = j + 1; c < k; ++c) {
                const double ujc = P[j + c * f];
                for (int64_t i = j + 1; i < f; ++i) P[i + c * f] -= P[i + j * f] * ujc;
            }
        }
    }
}
// Step 2 (after the caller has made the level's panels known to every rank): the rows of U12' and the columns of the
// Schur update this rank owns; a direct child hands its columns straight to the parent.
int64_t hx_top_update(void* hv, int rank, int level) {
    HX* h = (HX*)hv;
    const Symbolic& S = h->S;
    for (int u = S.level_ptr[level]; u < S.level_ptr[level + 1]; ++u) {
        const int s = S.level_sn[u];
        if (S.owner[s] != -1) continue;
        const int64_t k = S.sn_start[s + 1] - S.sn_start[s], r = S.rows_ptr[s + 1] - S.rows_ptr[s], f = k + r;
        const double* P = h->lu.data() + S.Loff[s];
        double* T = h->lu.data() + S.Uoff[s];
        double* C = h->cb.data() + S.CBoff[s];
        const int* rows = S.rows.data() + S.rows_ptr[s];
        const bool has_children = S.child_ptr[s + 1] > S.child_ptr[s];
        std::vector<char> mine(r);
        for (int64_t a = 0; a < r; ++a) mine[a] = S.col_owner[rows[a]] == rank;
        for (int64_t j = 0; j < k; ++j)
            for (int64_t a = 0; a < r; ++a) {
                if (!mine[a]) continue;
                const double uja = T[a + j * r];
                for (int64_t i = j + 1; i < k; ++i) T[a + i * r] -= P[i + j * f] * uja;
            }
        if (!has_children) for (int64_t e = 0; e < r * r; ++e) C[e] = 0.0;
        for (int64_t p = 0; p < k; ++p)
            for (int64_t b = 0; b < r; ++b) {
                if (!mine[b]) continue;
                const double upb = T[b + p * r];
                for (int64_t a = 0; a < r; ++a) C[a + b * r] -= P[(k + a) + p * f] * upb;
            }
        if (S.direct[s]) {
            const int ps = S.sn_parent[s];
            const int64_t pk = S.sn_start[ps + 1] - S.sn_start[ps], pr = S.rows_ptr[ps + 1] - S.rows_ptr[ps], pf = pk + pr;
            double* PP = h->lu.data() + S.Loff[ps];
            double* PT = h->lu.data() + S.Uoff[ps];
            double* PC = h->cb.data() + S.CBoff[ps];
            if (!S.cb_assigned[ps] && !h->zeroed[ps]) { for (int64_t e = 0; e < pr * pr; ++e) PC[e] = 0.0; h->zeroed[ps] = 1; }
            const int* rel = S.rel.data() + S.rows_ptr[s];
            for (int64_t b = 0; b < r; ++b) {
                if (!mine[b]) continue;
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
    }
    return h->bad;
}

int64_t hx_factor(void* hv, const double* Ax, const double* Rs) { return hx_factor_phase(hv, Ax, Rs, 0, 0); }

// Forward substitution, one phase (see hx_factor_phase).  Phase 0 also sums the update vectors of the
// rank's subtree roots into vbuf (one dense slice per interface front); the caller sums vbuf over the
// ranks before phase 1, where an interface front takes that slice instead of its children below the cut.
void hx_lsolve_phase(void* hv, double* x, int rank, int phase) {
    HX* h = (HX*)hv;
    const Symbolic& S = h->S;
    const int mine = phase == 0 ? rank : -1;
    if (phase == 0) { h->upd.assign(S.sum_r, 0.0); std::fill(h->vbuf.begin(), h->vbuf.end(), 0.0); }
    for (int l = 0; l < S.nlevels; ++l)
        for (int u = S.level_ptr[l]; u < S.level_ptr[l + 1]; ++u) {
            const int s = S.level_sn[u];
            if (S.owner[s] != mine) continue;
            const int c0 = S.sn_start[s];
            const int64_t k = S.sn_start[s + 1] - c0, r = S.rows_ptr[s + 1] - S.rows_ptr[s], f = k + r;
            const double* P = h->lu.data() + S.Loff[s];
            double* us = h->upd.data() + S.rows_ptr[s];
            if (S.iface[s]) {
                const double* vb = h->vbuf.data() + h->voff[s];
                for (int64_t i = 0; i < k; ++i) x[c0 + i] += vb[i];
                for (int64_t a = 0; a < r; ++a) us[a] += vb[k + a];
            }
            for (int ci = S.child_ptr[s]; ci < S.child_ptr[s + 1]; ++ci) {
                const int c = S.child_idx[ci];
                if (S.owner[c] != mine) continue;
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
            const int ps = S.sn_parent[s];
            if (phase == 0 && ps != -1 && S.owner[ps] == -1) {      // subtree root: hand over across the cut
                double* vb = h->vbuf.data() + h->voff[ps];
                const int* rel = S.rel.data() + S.rows_ptr[s];
                for (int64_t a = 0; a < r; ++a) vb[rel[a]] += us[a];
            }
        }
}

// lsolve! / rsolve! at the boundary work in the caller's labelling: permute through post when the analysis relabelled
static void with_post(void* hv, double* x, void (*fn)(void*, double*, int, int)) {
    const Symbolic& S = ((HX*)hv)->S;
    if (S.post.empty()) { fn(hv, x, 0, 0); return; }
    std::vector<double> w(S.n);
    for (int k = 0; k < S.n; ++k) w[k] = x[S.post[k]];
    fn(hv, w.data(), 0, 0);
    for (int k = 0; k < S.n; ++k) x[S.post[k]] = w[k];
}
void hx_lsolve(void* hv, double* x) { with_post(hv, x, hx_lsolve_phase); }

// Backward substitution, one phase: phase 1 (top) runs first, then phase 0 (the rank's subtrees).
void hx_rsolve_phase(void* hv, double* x, int rank, int phase) {
    HX* h = (HX*)hv;
    const Symbolic& S = h->S;
    const int mine = phase == 0 ? rank : -1;
    for (int l = S.nlevels - 1; l >= 0; --l)
        for (int u = S.level_ptr[l]; u < S.level_ptr[l + 1]; ++u) {
            const int s = S.level_sn[u];
            if (S.owner[s] != mine) continue;
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

void hx_rsolve(void* hv, double* x) { with_post(hv, x, hx_rsolve_phase); }

// zero the entries of a permuted-space vector this rank is not responsible for
void hx_mask_owned(void* hv, double* z, int rank) {
    const Symbolic& S = ((HX*)hv)->S;
    for (int j = 0; j < S.n; ++j) {
        const int o = S.owner[S.col2sn[j]];
        if (!(o == rank || (o == -1 && rank == 0))) z[j] = 0.0;
    }
}

void hx_permute_scale(void* hv, const double* b, double* w) {
    HX* h = (HX*)hv;
    for (int i = 0; i < h->S.n; ++i) w[i] = h->Rs[h->S.p[i]] * b[h->S.p[i]];
}
void hx_unpermute(void* hv, const double* w, double* x) {
    HX* h = (HX*)hv;
    for (int i = 0; i < h->S.n; ++i) x[h->S.q[i]] = w[i];
}

void hx_solve(void* hv, const double* b, double* x) {
    HX* h = (HX*)hv;
    const Symbolic& S = h->S;
    std::vector<double> w(S.n);
    for (int i = 0; i < S.n; ++i) w[i] = h->Rs[S.p[i]] * b[S.p[i]];
    hx_lsolve_phase(hv, w.data(), 0, 0);
    hx_rsolve_phase(hv, w.data(), 0, 0);
    for (int i = 0; i < S.n; ++i) x[S.q[i]] = w[i];
}

void hx_get_factors(void* hv, int64_t* Lp, int64_t* Li, double* Lx, int64_t* Up, int64_t* Ui, double* Ux) {
    HX* h = (HX*)hv;
    std::vector<int64_t> ptr;
    std::vector<int> idx;
    exact_structure(h->S, h->Ap.data(), h->Ai.data(), ptr, idx);
    export_factors(h->S, ptr, idx, h->lu.data(), 0, Lp, Li, Lx, Up, Ui, Ux);
    if (!h->S.post.empty()) {
        relabel_csc(h->S.n, h->S.post.data(), 0, Lp, Li, Lx);
        relabel_csc(h->S.n, h->S.post.data(), 0, Up, Ui, Ux);
    }
}

}  // extern "C"

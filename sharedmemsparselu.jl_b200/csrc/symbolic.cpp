// Host-side symbolic analysis (see symbolic.hpp).  Plain C++17, no device code: testable on CPU.
#include "symbolic.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <mutex>
#include <system_error>
#include <thread>
#include <cmath>
#include <cstring>
#include <map>
#include <numeric>

#include "../../include/smslu.h"

namespace smslu {
namespace {

struct Graph {
    int n = 0;
    std::vector<int64_t> xadj;
    std::vector<int> adj;   // sorted, unique, no self loops
};

// Run fn(i0, i1) over [0, n) cut into ranges of about equal weight (prefix[i] = weight of the rows before i), on a few
// host threads when the total weight makes that worthwhile.  The ranges are disjoint, so any fn that only writes what
// belongs to its own rows gives the sequential result.
// Host threads of the analysis: min(16, cores), or SMSLU_HOST_THREADS (the results do not depend on it).
int host_threads() {
    if (const char* e = getenv("SMSLU_HOST_THREADS")) return std::max(1, std::min(64, atoi(e)));
    return (int)std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency()));
}

template <class F>
void parallel_rows(int n, const std::vector<int64_t>& prefix, F fn) {
    const int64_t total = prefix[n];
    int nthreads = host_threads();
    if (total < (int64_t)2000000 || n < 1024) nthreads = 1;
    if (nthreads == 1) { fn(0, n); return; }
    std::vector<int> cut(nthreads + 1, n);
    cut[0] = 0;
    for (int t = 1; t < nthreads; ++t)
        cut[t] = (int)(std::lower_bound(prefix.begin(), prefix.begin() + n, total * t / nthreads) - prefix.begin());
    std::vector<std::thread> pool;
    int started = 1;                       // ranges [1, started) run on their own threads
    try {
        for (; started < nthreads; ++started)
            if (cut[started + 1] > cut[started]) pool.emplace_back(fn, cut[started], cut[started + 1]);
    } catch (const std::system_error&) {}  // no more threads to be had: the rest runs here
    if (cut[1] > cut[0]) fn(cut[0], cut[1]);
    for (int t = started; t < nthreads; ++t)
        if (cut[t + 1] > cut[t]) fn(cut[t], cut[t + 1]);
    for (auto& th : pool) th.join();
}

// Graph of pattern(B + B') for B = A[p,q]; rinv/cinv map original row/col -> permuted index.
Graph build_sym_graph(int n, const int64_t* Ap, const int64_t* Ai, const int* rinv, const int* cinv) {
    Graph G;
    G.n = n;
    std::vector<int64_t> cnt(n + 1, 0);
    for (int c = 0; c < n; ++c)
        for (int64_t t = Ap[c]; t < Ap[c + 1]; ++t) {
            int i = rinv[Ai[t]], j = cinv[c];
            if (i == j) continue;
            ++cnt[i + 1];
            ++cnt[j + 1];
        }
    for (int i = 0; i < n; ++i) cnt[i + 1] += cnt[i];
    std::vector<int> tmp(cnt[n]);
    std::vector<int64_t> w(cnt.begin(), cnt.end() - 1);
    for (int c = 0; c < n; ++c)
        for (int64_t t = Ap[c]; t < Ap[c + 1]; ++t) {
            int i = rinv[Ai[t]], j = cinv[c];
            if (i == j) continue;
            tmp[w[i]++] = j;
            tmp[w[j]++] = i;
        }
    // sort + unique every row inside its own slot (rows are independent), then compact
    G.xadj.assign(n + 1, 0);
    parallel_rows(n, cnt, [&](int i0, int i1) {
        for (int i = i0; i < i1; ++i) {
            std::sort(tmp.begin() + cnt[i], tmp.begin() + cnt[i + 1]);
            G.xadj[i + 1] = std::unique(tmp.begin() + cnt[i], tmp.begin() + cnt[i + 1]) - (tmp.begin() + cnt[i]);
        }
    });
    for (int i = 0; i < n; ++i) G.xadj[i + 1] += G.xadj[i];
    G.adj.resize(G.xadj[n]);
    parallel_rows(n, cnt, [&](int i0, int i1) {
        for (int i = i0; i < i1; ++i)
            std::copy(tmp.begin() + cnt[i], tmp.begin() + cnt[i] + (G.xadj[i + 1] - G.xadj[i]), G.adj.begin() + G.xadj[i]);
    });
    return G;
}

// new vertex k is old vertex perm[k]
Graph relabel(const Graph& G, const std::vector<int>& perm) {
    int n = G.n;
    std::vector<int> inv(n);
    for (int k = 0; k < n; ++k) inv[perm[k]] = k;
    Graph H;
    H.n = n;
    H.xadj.assign(n + 1, 0);
    for (int k = 0; k < n; ++k) H.xadj[k + 1] = H.xadj[k] + (G.xadj[perm[k] + 1] - G.xadj[perm[k]]);
    H.adj.resize(G.adj.size());
    parallel_rows(n, H.xadj, [&](int k0, int k1) {
        for (int k = k0; k < k1; ++k) {
            int v = perm[k];
            int64_t o = H.xadj[k];
            for (int64_t t = G.xadj[v]; t < G.xadj[v + 1]; ++t) H.adj[o++] = inv[G.adj[t]];
            std::sort(H.adj.begin() + H.xadj[k], H.adj.begin() + H.xadj[k + 1]);
        }
    });
    return H;
}

// ------------------------------------------------------------------ geometric nested dissection
void nd_grid(const int dims[3], int leaf, std::vector<int>& order) {
    const int nx = dims[0], ny = std::max(dims[1], 1), nz = std::max(dims[2], 1);
    order.resize((size_t)nx * ny * nz);
    struct Box { int lo[3], hi[3]; int64_t pos; };
    std::vector<Box> st;
    st.push_back(Box{{0, 0, 0}, {nx, ny, nz}, 0});
    auto emit = [&](const int lo[3], const int hi[3], int64_t pos) {
        for (int k = lo[2]; k < hi[2]; ++k)
            for (int j = lo[1]; j < hi[1]; ++j)
                for (int i = lo[0]; i < hi[0]; ++i) order[pos++] = i + nx * (j + ny * k);
        return pos;
    };
    while (!st.empty()) {
        Box b = st.back();
        st.pop_back();
        int64_t ext[3] = {b.hi[0] - b.lo[0], b.hi[1] - b.lo[1], b.hi[2] - b.lo[2]};
        int64_t vol = ext[0] * ext[1] * ext[2];
        int d = 0;
        for (int a = 1; a < 3; ++a) if (ext[a] > ext[d]) d = a;
        if (vol <= leaf || ext[d] < 3) { emit(b.lo, b.hi, b.pos); continue; }
        int mid = b.lo[d] + (int)(ext[d] / 2);
        Box L = b, R = b, Sp = b;
        L.hi[d] = mid;
        R.lo[d] = mid + 1;
        Sp.lo[d] = mid; Sp.hi[d] = mid + 1;
        int64_t vl = vol / ext[d] * (mid - b.lo[d]), vr = vol / ext[d] * (b.hi[d] - mid - 1);
        L.pos = b.pos;
        R.pos = b.pos + vl;
        emit(Sp.lo, Sp.hi, b.pos + vl + vr);
        st.push_back(L);
        st.push_back(R);
    }
}

// ------------------------------------------------------------------ graph nested dissection
// George-style automatic nested dissection: the separator is a (trimmed) level set of a
// rooted level structure started at a pseudo-peripheral vertex, chosen near the median.
class GraphND {
  public:
    GraphND(const Graph& G, const SymOptions& o) : G_(G), opt_(o), region_(G.n), level_(G.n, -1) {
        for (int v = 0; v < G.n; ++v) region_[v].store(-1, std::memory_order_relaxed);
    }

    struct Item { std::vector<int> verts; int64_t pos; };

    // The two sides of a separator are independent subproblems (disjoint vertices, disjoint ranges of `order`),
    // and an item's result does not depend on when it is processed: a few worker threads share the open
    // subdomains.  The ordering is bit-for-bit the sequential one.
    void run(std::vector<int>& order) {
        const int n = G_.n;
        order.assign(n, -1);
        // dense vertices go last (they would wreck every level structure)
        double thr = std::max(16.0, opt_.dense_factor * std::sqrt((double)n));
        std::vector<int> normal, dense;
        for (int v = 0; v < n; ++v)
            ((double)(G_.xadj[v + 1] - G_.xadj[v]) > thr ? dense : normal).push_back(v);
        if (normal.empty()) { std::iota(order.begin(), order.end(), 0); return; }
        int64_t pos_dense = (int64_t)normal.size();
        for (int v : dense) order[pos_dense++] = v;   // region_ stays -1 => invisible to BFS
        std::vector<Item> st;
        st.push_back(Item{std::move(normal), 0});
        int nthreads = host_threads();
        if (n < 100000) nthreads = 1;
        if (nthreads > 1) {
            // shared stack of large open subdomains; a worker keeps one side of every split for itself and hands
            // the other one over while it is still large
            constexpr size_t SHARE_MIN = 8192;
            std::mutex mu;
            std::condition_variable cv;
            int active = 0;
            auto worker = [&]() {
                std::vector<Item> mine;
                for (;;) {
                    {
                        std::unique_lock<std::mutex> lk(mu);
                        cv.wait(lk, [&] { return !st.empty() || active == 0; });
                        if (st.empty()) { cv.notify_all(); return; }
                        mine.push_back(std::move(st.back()));
                        st.pop_back();
                        ++active;
                    }
                    while (!mine.empty()) {
                        Item it = std::move(mine.back());
                        mine.pop_back();
                        process(std::move(it), mine, order);
                        while (mine.size() > 1 && mine.back().verts.size() >= SHARE_MIN) {
                            std::lock_guard<std::mutex> lk(mu);
                            st.push_back(std::move(mine.back()));
                            mine.pop_back();
                            cv.notify_one();
                        }
                    }
                    {
                        std::lock_guard<std::mutex> lk(mu);
                        --active;
                        if (active == 0 && st.empty()) cv.notify_all();
                    }
                }
            };
            std::vector<std::thread> pool;
            try {
                for (int t = 1; t < nthreads; ++t) pool.emplace_back(worker);
            } catch (const std::system_error&) {}      // fewer workers, same result
            worker();
            for (auto& th : pool) th.join();
            return;
        }
        while (!st.empty()) {
            Item it = std::move(st.back());
            st.pop_back();
            process(std::move(it), st, order);
        }
    }

  private:
    // One step of the recursion: order a leaf, or split `it` and push the two sides onto st.
    void process(Item it, std::vector<Item>& st, std::vector<int>& order) {
        const int rid = next_region_.fetch_add(1);
        for (int v : it.verts) region_[v].store(rid, std::memory_order_relaxed);
        // ---- connected components
        std::vector<std::vector<int>> comps;
        for (int v : it.verts) level_[v] = -1;
        for (int v : it.verts) {
            if (level_[v] != -1) continue;
            comps.emplace_back();
            bfs(v, rid, comps.back());
        }
        if (comps.size() > 1) {
            int64_t pos = it.pos;
            for (auto& c : comps) {
                int64_t sz = (int64_t)c.size();
                st.push_back(Item{std::move(c), pos});
                pos += sz;
            }
            return;
        }
        std::vector<int>& comp = comps[0];
        const int m = (int)comp.size();
        if (m <= opt_.nd_leaf) {   // leaf: Cuthill-McKee style order from a peripheral vertex
            int root = pseudo_peripheral(comp, rid);
            std::vector<int> ord;
            for (int v : comp) level_[v] = -1;
            bfs(root, rid, ord);
            for (int k = 0; k < m; ++k) order[it.pos + k] = ord[k];
            return;
        }
        int root = pseudo_peripheral(comp, rid);
        std::vector<int> ord;
        for (int v : comp) level_[v] = -1;
        bfs(root, rid, ord);
        int nlev = level_[ord.back()] + 1;
        if (nlev < 3) {   // clique-like: no useful separator
            for (int k = 0; k < m; ++k) order[it.pos + k] = ord[k];
            return;
        }
        std::vector<int> lcount(nlev, 0);
        for (int v : ord) ++lcount[level_[v]];
        // candidate separator levels: both sides keep >= 25% of the vertices; among those the
        // narrowest level wins (penalised by imbalance).  Fallback: the median level.
        int best = -1, median = -1; double best_score = 1e300; int64_t cum = 0;
        for (int l = 0; l < nlev; ++l) {
            int64_t before = cum; cum += lcount[l];
            if (median < 0 && 2 * cum >= m) median = l;
            if (l == 0 || l == nlev - 1) continue;
            double fb = (double)before / m, fa = (double)(m - cum) / m;
            if (fb < 0.25 || fa < 0.25) continue;
            double score = (double)lcount[l] * (1.0 + std::fabs(fb - fa));
            if (score < best_score) { best_score = score; best = l; }
        }
        if (best < 0) best = std::min(std::max(median, 1), nlev - 2);
        std::vector<int> A, B, Sv;
        for (int v : ord) {
            int l = level_[v];
            if (l < best) A.push_back(v);
            else if (l > best) B.push_back(v);
            else {
                bool touches_after = false;
                for (int64_t t = G_.xadj[v]; t < G_.xadj[v + 1] && !touches_after; ++t) {
                    int u = G_.adj[t];
                    if (region_[u].load(std::memory_order_relaxed) == rid && level_[u] == best + 1) touches_after = true;
                }
                (touches_after ? Sv : A).push_back(v);
            }
        }
        if (Sv.empty() || A.empty() || B.empty()) {
            for (int k = 0; k < m; ++k) order[it.pos + k] = ord[k];
            return;
        }
        int64_t pa = it.pos, pb = pa + (int64_t)A.size(), ps = pb + (int64_t)B.size();
        for (size_t k = 0; k < Sv.size(); ++k) { order[ps + (int64_t)k] = Sv[k]; region_[Sv[k]].store(-2, std::memory_order_relaxed); }
        st.push_back(Item{std::move(A), pa});
        st.push_back(Item{std::move(B), pb});
    }

    // BFS inside region rid from root over vertices with level_ == -1; appends visit order.
    void bfs(int root, int rid, std::vector<int>& out) {
        size_t head = out.size();
        level_[root] = 0;
        out.push_back(root);
        while (head < out.size()) {
            int v = out[head++];
            for (int64_t t = G_.xadj[v]; t < G_.xadj[v + 1]; ++t) {
                int u = G_.adj[t];
                if (region_[u].load(std::memory_order_relaxed) != rid || level_[u] != -1) continue;
                level_[u] = level_[v] + 1;
                out.push_back(u);
            }
        }
    }
    int pseudo_peripheral(const std::vector<int>& comp, int rid) {
        int root = comp[0];
        int64_t bestdeg = INT64_MAX;
        for (int v : comp) {
            int64_t d = G_.xadj[v + 1] - G_.xadj[v];
            if (d < bestdeg) { bestdeg = d; root = v; }
        }
        int depth = -1;
        std::vector<int> ord;
        for (int iter = 0; iter < 4; ++iter) {
            for (int v : comp) level_[v] = -1;
            ord.clear();
            bfs(root, rid, ord);
            int d = level_[ord.back()];
            if (d <= depth) break;
            depth = d;
            // smallest-degree vertex of the last level
            int cand = ord.back(); int64_t cd = INT64_MAX;
            for (size_t k = ord.size(); k-- > 0 && level_[ord[k]] == d;) {
                int64_t dg = G_.xadj[ord[k] + 1] - G_.xadj[ord[k]];
                if (dg < cd) { cd = dg; cand = ord[k]; }
            }
            root = cand;
        }
        return root;
    }
    const Graph& G_;
    const SymOptions& opt_;
    std::vector<std::atomic<int>> region_;   // other threads' regions are read (never matched) while they change
    std::vector<int> level_;
    std::atomic<int> next_region_{0};
};

// ------------------------------------------------------------------ elimination tree etc.
void etree(const Graph& G, std::vector<int>& parent) {
    int n = G.n;
    parent.assign(n, -1);
    std::vector<int> anc(n, -1);
    for (int j = 0; j < n; ++j)
        for (int64_t t = G.xadj[j]; t < G.xadj[j + 1]; ++t) {
            int r = G.adj[t];
            if (r >= j) break;
            while (anc[r] != -1 && anc[r] != j) { int nx = anc[r]; anc[r] = j; r = nx; }
            if (anc[r] == -1) { anc[r] = j; parent[r] = j; }
        }
}

void postorder(const std::vector<int>& parent, std::vector<int>& post) {
    int n = (int)parent.size();
    std::vector<int> head(n, -1), next(n, -1);
    for (int j = n - 1; j >= 0; --j)
        if (parent[j] != -1) { next[j] = head[parent[j]]; head[parent[j]] = j; }
    post.clear();
    post.reserve(n);
    std::vector<int> st;
    for (int r = 0; r < n; ++r) {
        if (parent[r] != -1) continue;
        st.push_back(r);
        while (!st.empty()) {
            int v = st.back();
            int c = head[v];
            if (c == -1) { post.push_back(v); st.pop_back(); }
            else { head[v] = next[c]; st.push_back(c); }
        }
    }
}

// Column counts of the Cholesky-like factor of a postordered symmetric pattern
// (skeleton-matrix / least-common-ancestor method of Gilbert, Ng and Peyton).
void column_counts(const Graph& G, const std::vector<int>& parent, std::vector<int>& cc) {
    int n = G.n;
    std::vector<int> size(n, 1), first(n), maxfirst(n, -1), prevleaf(n, -1), anc(n);
    std::vector<int64_t> delta(n);
    for (int j = 0; j < n; ++j) if (parent[j] != -1) size[parent[j]] += size[j];
    for (int j = 0; j < n; ++j) { first[j] = j - size[j] + 1; delta[j] = (size[j] == 1) ? 1 : 0; anc[j] = j; }
    for (int j = 0; j < n; ++j) {
        if (parent[j] != -1) --delta[parent[j]];
        for (int64_t t = G.xadj[j]; t < G.xadj[j + 1]; ++t) {
            int i = G.adj[t];
            if (i <= j || first[j] <= maxfirst[i]) continue;
            maxfirst[i] = first[j];
            int jprev = prevleaf[i];
            prevleaf[i] = j;
            ++delta[j];
            if (jprev != -1) {
                int qn = jprev;
                while (qn != anc[qn]) qn = anc[qn];
                for (int s = jprev; s != qn;) { int sp = anc[s]; anc[s] = qn; s = sp; }
                --delta[qn];
            }
        }
        if (parent[j] != -1) anc[j] = parent[j];
    }
    cc.resize(n);
    std::vector<int64_t> acc(delta);
    for (int j = 0; j < n; ++j) if (parent[j] != -1) acc[parent[j]] += acc[j];
    for (int j = 0; j < n; ++j) cc[j] = (int)acc[j];
}

// best-fit allocator over a growing arena, used to plan contribution-block lifetimes
class Arena {
  public:
    int64_t alloc(int64_t len) {
        if (len == 0) return 0;
        auto it = by_size_.lower_bound(len);
        if (it != by_size_.end()) {
            int64_t sz = it->first, off = it->second;
            by_size_.erase(it);
            by_off_.erase(off);
            if (sz > len) insert_free(off + len, sz - len);
            return off;
        }
        // extend the arena; reuse a trailing free block if there is one
        if (!by_off_.empty()) {
            auto last = std::prev(by_off_.end());
            if (last->first + last->second == top_) {
                int64_t off = last->first, sz = last->second;
                erase_free(off, sz);
                top_ = off + len;
                return off;
            }
        }
        int64_t off = top_;
        top_ += len;
        return off;
    }
    void release(int64_t off, int64_t len) {
        if (len == 0) return;
        auto nx = by_off_.lower_bound(off);
        if (nx != by_off_.end() && off + len == nx->first) {
            int64_t nsz = nx->second, noff = nx->first;
            erase_free(noff, nsz);
            len += nsz;
        }
        auto pv = by_off_.lower_bound(off);
        if (pv != by_off_.begin()) {
            --pv;
            if (pv->first + pv->second == off) {
                int64_t poff = pv->first, psz = pv->second;
                erase_free(poff, psz);
                off = poff;
                len += psz;
            }
        }
        insert_free(off, len);
    }
    int64_t top() const { return top_; }

  private:
    void insert_free(int64_t off, int64_t len) { by_off_[off] = len; by_size_.emplace(len, off); }
    void erase_free(int64_t off, int64_t len) {
        by_off_.erase(off);
        auto rng = by_size_.equal_range(len);
        for (auto it = rng.first; it != rng.second; ++it)
            if (it->second == off) { by_size_.erase(it); break; }
    }
    std::map<int64_t, int64_t> by_off_;
    std::multimap<int64_t, int64_t> by_size_;
    int64_t top_ = 0;
};

inline int64_t align2(int64_t x) { return (x + 1) & ~(int64_t)1; }

}  // namespace

int analyze(int n, const int64_t* Ap, const int64_t* Ai, const int* p_in, const int* q_in,
            const SymOptions& opt, Symbolic& S, std::string& err) {
    // SMSLU_ANALYZE_TIMES=1 (debug): wall time of every numbered step below on stderr
    const bool step_times = getenv("SMSLU_ANALYZE_TIMES") != nullptr;
    const char* step_name = nullptr;
    auto step_clock = std::chrono::steady_clock::now();
    auto step_mark = [&](const char* nm) {
        if (!step_times) return;
        const auto t = std::chrono::steady_clock::now();
        if (step_name) fprintf(stderr, "[smslu analyze] %-36s %9.1f ms\n", step_name, std::chrono::duration<double, std::milli>(t - step_clock).count());
        step_name = nm; step_clock = t;
    };

    S = Symbolic();
    S.n = n;
    if (n <= 0) { err = "matrix must have at least one row"; return SMSLU_E_DIM; }
    S.annz = Ap[n];
    {
        std::vector<int> seen(n, -1);     // a duplicate (row, column) would silently lose a value in the scatter
        for (int c = 0; c < n; ++c) {
            if (Ap[c + 1] < Ap[c]) { err = "colptr is not monotone"; return SMSLU_E_PATTERN; }
            for (int64_t t = Ap[c]; t < Ap[c + 1]; ++t) {
                if (Ai[t] < 0 || Ai[t] >= n) { err = "row index out of range"; return SMSLU_E_PATTERN; }
                if (seen[Ai[t]] == c) { err = "duplicate (row, column) entry in the pattern: sum duplicates first"; return SMSLU_E_PATTERN; }
                seen[Ai[t]] = c;
            }
        }
    }
    step_mark("1 initial ordering");
    // ---------------------------------------------------------------- 1. initial ordering
    std::vector<int> p0(n), q0(n);
    int ordering = opt.ordering;
    if (ordering == ORD_AUTO) {
        int64_t g = (int64_t)opt.grid[0] * std::max(opt.grid[1], 1) * std::max(opt.grid[2], 1);
        ordering = (opt.grid[0] > 0 && g == n) ? ORD_ND_GRID : ORD_ND_GRAPH;
    }
    if (ordering == ORD_GIVEN) {
        if (!p_in || !q_in) { err = "ordering GIVEN needs p and q"; return SMSLU_E_ARG; }
        std::vector<char> seenp(n, 0), seenq(n, 0);
        for (int k = 0; k < n; ++k) {
            if (p_in[k] < 0 || p_in[k] >= n || q_in[k] < 0 || q_in[k] >= n || seenp[p_in[k]] || seenq[q_in[k]]) {
                err = "p/q is not a permutation";
                return SMSLU_E_ARG;
            }
            seenp[p_in[k]] = seenq[q_in[k]] = 1;
            p0[k] = p_in[k];
            q0[k] = q_in[k];
        }
    } else if (ordering == ORD_NATURAL) {
        std::iota(p0.begin(), p0.end(), 0);
        q0 = p0;
    } else if (ordering == ORD_ND_GRID) {
        int64_t g = (int64_t)opt.grid[0] * std::max(opt.grid[1], 1) * std::max(opt.grid[2], 1);
        if (opt.grid[0] <= 0 || g != n) { err = "grid hint does not match n"; return SMSLU_E_ARG; }
        nd_grid(opt.grid, std::max(opt.nd_leaf, 1), p0);
        q0 = p0;
    } else if (ordering == ORD_ND_GRAPH) {
        std::vector<int> id(n);
        std::iota(id.begin(), id.end(), 0);
        Graph G0 = build_sym_graph(n, Ap, Ai, id.data(), id.data());
        GraphND nd(G0, opt);
        nd.run(p0);
        q0 = p0;
    } else { err = "unknown ordering"; return SMSLU_E_ARG; }

    step_mark("2 etree + postorder");
    // ---------------------------------------------------------------- 2. etree + postorder
    std::vector<int> rinv(n), cinv(n);
    for (int k = 0; k < n; ++k) { rinv[p0[k]] = k; cinv[q0[k]] = k; }
    Graph G1 = build_sym_graph(n, Ap, Ai, rinv.data(), cinv.data());
    std::vector<int> par1, post;
    etree(G1, par1);
    postorder(par1, post);
    S.p.resize(n);
    S.q.resize(n);
    std::vector<int> ipost(n);
    for (int k = 0; k < n; ++k) { S.p[k] = p0[post[k]]; S.q[k] = q0[post[k]]; ipost[post[k]] = k; }
    if (ordering == ORD_GIVEN) {
        bool ident = true;
        for (int k = 0; k < n && ident; ++k) ident = post[k] == k;
        if (!ident) { S.p_given = p0; S.q_given = q0; S.post = post; }
    }
    Graph G = relabel(G1, post);
    G1 = Graph();
    S.parent.resize(n);
    for (int k = 0; k < n; ++k) S.parent[k] = par1[post[k]] == -1 ? -1 : ipost[par1[post[k]]];
    for (int k = 0; k < n; ++k) { rinv[S.p[k]] = k; cinv[S.q[k]] = k; }
    const std::vector<int>& parent = S.parent;

    step_mark("3 column counts");
    // ---------------------------------------------------------------- 3. column counts
    column_counts(G, parent, S.colcount);
    const std::vector<int>& cc = S.colcount;
    S.nnzL_exact = 0;
    S.flops_exact = 0;
    for (int j = 0; j < n; ++j) {
        S.nnzL_exact += cc[j];
        double c = cc[j] - 1;
        S.flops_exact += c * (2.0 * c + 1.0);
    }

    step_mark("4 supernodes");
    // ---------------------------------------------------------------- 4. supernodes
    std::vector<int> fs_start;   // fundamental (maximal) supernodes
    const int W = std::max(1, opt.max_width);
    for (int j = 0; j < n; ++j)   // wide supernodes become chains of <= W-column fronts
        if (j == 0 || !(parent[j - 1] == j && cc[j] == cc[j - 1] - 1) || j - fs_start.back() >= W)
            fs_start.push_back(j);
    int nfs = (int)fs_start.size();
    fs_start.push_back(n);
    std::vector<int> gfirst(nfs);
    std::vector<double> tru(nfs);
    std::vector<char> dead(nfs, 0);
    std::vector<int> col2fs(n);
    for (int s = 0; s < nfs; ++s)
        for (int j = fs_start[s]; j < fs_start[s + 1]; ++j) col2fs[j] = s;
    for (int s = 0; s < nfs; ++s) {
        gfirst[s] = fs_start[s];
        double t = 0;
        for (int j = fs_start[s]; j < fs_start[s + 1]; ++j) t += cc[j];
        tru[s] = t;
        if (!opt.relax) continue;
        const int last = fs_start[s + 1] - 1;
        const double rs = cc[last] - 1;
        for (;;) {
            int c = gfirst[s] - 1;
            if (c < 0) break;
            int pc = parent[c];
            if (pc < gfirst[s] || pc > last) break;   // the group ending at c is not a child
            int tgrp = col2fs[c];
            double kt = c - gfirst[tgrp] + 1, ks = last - gfirst[s] + 1;
            double kn = kt + ks;
            if (kn > std::min(W, opt.relax_width)) break;
            double stored = kn * (kn + 1) / 2 + kn * rs;
            double frac = 1.0 - (tru[tgrp] + tru[s]) / stored;
            double lim = kn <= opt.relax_k1 ? opt.relax_f1 : (kn <= opt.relax_k2 ? opt.relax_f2 : opt.relax_f3);
            if (!(kn <= opt.relax_always || frac <= lim)) break;
            gfirst[s] = gfirst[tgrp];
            tru[s] += tru[tgrp];
            dead[tgrp] = 1;
        }
    }
    S.sn_start.clear();
    for (int s = 0; s < nfs; ++s) if (!dead[s]) S.sn_start.push_back(gfirst[s]);
    S.nsn = (int)S.sn_start.size();
    S.sn_start.push_back(n);
    const int nsn = S.nsn;
    S.col2sn.resize(n);
    for (int s = 0; s < nsn; ++s)
        for (int j = S.sn_start[s]; j < S.sn_start[s + 1]; ++j) S.col2sn[j] = s;
    S.sn_parent.assign(nsn, -1);
    for (int s = 0; s < nsn; ++s) {
        int pl = parent[S.sn_start[s + 1] - 1];
        S.sn_parent[s] = pl == -1 ? -1 : S.col2sn[pl];
    }
    S.child_ptr.assign(nsn + 1, 0);
    for (int s = 0; s < nsn; ++s) if (S.sn_parent[s] != -1) ++S.child_ptr[S.sn_parent[s] + 1];
    for (int s = 0; s < nsn; ++s) S.child_ptr[s + 1] += S.child_ptr[s];
    S.child_idx.resize(S.child_ptr[nsn]);
    {
        std::vector<int> w(S.child_ptr.begin(), S.child_ptr.end() - 1);
        for (int s = 0; s < nsn; ++s) if (S.sn_parent[s] != -1) S.child_idx[w[S.sn_parent[s]]++] = s;
    }
    S.max_children = 0;
    for (int s = 0; s < nsn; ++s) S.max_children = std::max(S.max_children, S.child_ptr[s + 1] - S.child_ptr[s]);

    step_mark("5 supernodal row structure");
    // ---------------------------------------------------------------- 5. supernodal row structure
    S.rows_ptr.assign(nsn + 1, 0);
    S.rows.clear();
    {
        std::vector<int> marker(n, -1), list;
        for (int s = 0; s < nsn; ++s) {
            const int c0 = S.sn_start[s], last = S.sn_start[s + 1] - 1;
            list.clear();
            for (int j = c0; j <= last; ++j)
                for (int64_t t = G.xadj[j + 1]; t-- > G.xadj[j];) {
                    int i = G.adj[t];
                    if (i <= last) break;
                    if (marker[i] != s) { marker[i] = s; list.push_back(i); }
                }
            for (int t = S.child_ptr[s]; t < S.child_ptr[s + 1]; ++t) {
                int c = S.child_idx[t];
                for (int64_t u = S.rows_ptr[c + 1]; u-- > S.rows_ptr[c];) {
                    int i = S.rows[u];
                    if (i <= last) break;
                    if (marker[i] != s) { marker[i] = s; list.push_back(i); }
                }
            }
            std::sort(list.begin(), list.end());
            if ((int)list.size() != cc[last] - 1) {
                err = "internal: supernode row count disagrees with column count";
                return SMSLU_E_INTERNAL;
            }
            S.rows.insert(S.rows.end(), list.begin(), list.end());
            S.rows_ptr[s + 1] = (int64_t)S.rows.size();
        }
    }
    step_mark("6 child -> parent index maps");
    // ---------------------------------------------------------------- 6. child -> parent index maps
    S.rel.assign(S.rows.size(), -1);
    for (int c = 0; c < nsn; ++c) {
        int s = S.sn_parent[c];
        if (s == -1) {
            if (S.rows_ptr[c + 1] != S.rows_ptr[c]) { err = "internal: root supernode with rows"; return SMSLU_E_INTERNAL; }
            continue;
        }
        const int pc0 = S.sn_start[s], plast = S.sn_start[s + 1] - 1, pk = plast - pc0 + 1;
        int64_t u = S.rows_ptr[s];
        for (int64_t t = S.rows_ptr[c]; t < S.rows_ptr[c + 1]; ++t) {
            int i = S.rows[t];
            if (i <= plast) { S.rel[t] = i - pc0; continue; }
            while (u < S.rows_ptr[s + 1] && S.rows[u] < i) ++u;
            if (u == S.rows_ptr[s + 1] || S.rows[u] != i) { err = "internal: child row missing in parent"; return SMSLU_E_INTERNAL; }
            S.rel[t] = pk + (int)(u - S.rows_ptr[s]);
        }
    }
    step_mark("7 levels");
    // ---------------------------------------------------------------- 7. levels
    S.sn_level.assign(nsn, 0);
    S.nlevels = 0;
    for (int s = 0; s < nsn; ++s) {
        if (S.sn_parent[s] != -1) S.sn_level[S.sn_parent[s]] = std::max(S.sn_level[S.sn_parent[s]], S.sn_level[s] + 1);
        S.nlevels = std::max(S.nlevels, S.sn_level[s] + 1);
    }
    S.level_ptr.assign(S.nlevels + 1, 0);
    for (int s = 0; s < nsn; ++s) ++S.level_ptr[S.sn_level[s] + 1];
    for (int l = 0; l < S.nlevels; ++l) S.level_ptr[l + 1] += S.level_ptr[l];
    S.level_sn.resize(nsn);
    {
        std::vector<int> w(S.level_ptr.begin(), S.level_ptr.end() - 1);
        for (int s = 0; s < nsn; ++s) S.level_sn[w[S.sn_level[s]]++] = s;
    }
    step_mark("7b partition over GPUs");
    // ---------------------------------------------------------------- 7b. partition over GPUs
    // owner[s] = rank that factors supernode s, or -1 for the "top" of the tree, which every rank
    // factors redundantly after the contributions of the subtrees have been summed (all-reduce).
    // Greedy: start from the roots, repeatedly move the heaviest remaining subtree root into the top
    // set and replace it by its children; keep the cut with the smallest  top work + heaviest rank.
    const int NR = std::max(1, opt.nranks);
    S.nranks = NR;
    S.owner.assign(nsn, 0);
    S.iface.assign(nsn, 0);
    if (NR > 1) {
        std::vector<double> wsub(nsn, 0.0), wself(nsn);
        for (int s = 0; s < nsn; ++s) {
            const double kd = S.sn_start[s + 1] - S.sn_start[s], rd = (double)(S.rows_ptr[s + 1] - S.rows_ptr[s]);
            wself[s] = 2.0 * kd * kd * kd / 3.0 + 2.0 * kd * kd * rd + 2.0 * kd * rd * rd + 1.0e5;
        }
        for (int s = 0; s < nsn; ++s) {          // children precede parents (postorder)
            wsub[s] += wself[s];
            if (S.sn_parent[s] != -1) wsub[S.sn_parent[s]] += wsub[s];
        }
        auto lpt = [&](const std::vector<int>& q, std::vector<int>* assign) {
            std::vector<int> ord(q.size());
            std::iota(ord.begin(), ord.end(), 0);
            std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return wsub[q[a]] > wsub[q[b]]; });
            std::vector<double> load(NR, 0.0);
            if (assign) assign->assign(q.size(), 0);
            for (int t : ord) {
                int best = 0;
                for (int g = 1; g < NR; ++g) if (load[g] < load[best]) best = g;
                load[best] += wsub[q[t]];
                if (assign) (*assign)[t] = best;
            }
            return *std::max_element(load.begin(), load.end());
        };
        std::vector<int> q, best_q;
        std::vector<char> is_top(nsn, 0), best_top;
        for (int s = 0; s < nsn; ++s) if (S.sn_parent[s] == -1) q.push_back(s);
        double top_w = 0.0, best_cost = 1e300;      // at least the roots go to the top
        best_q = q; best_top = is_top;
        for (int it = 0; it < 96 * NR && !q.empty(); ++it) {
            int hi = -1;
            for (size_t t = 0; t < q.size(); ++t)
                if (S.child_ptr[q[t] + 1] > S.child_ptr[q[t]] && (hi < 0 || wsub[q[t]] > wsub[q[hi]])) hi = (int)t;
            if (hi < 0) break;
            const int s = q[hi];
            q.erase(q.begin() + hi);
            is_top[s] = 1;
            top_w += wself[s];
            for (int u = S.child_ptr[s]; u < S.child_ptr[s + 1]; ++u) q.push_back(S.child_idx[u]);
            // the top is shared by the ranks (column-distributed Schur updates; panels and the chain of dependent
            // steps are not): count it at 60 % parallel efficiency
            const double cost = top_w / (0.6 * NR) + lpt(q, nullptr);
            if (cost < best_cost) { best_cost = cost; best_q = q; best_top = is_top; }
        }
        std::vector<int> assign;
        lpt(best_q, &assign);
        for (int s = 0; s < nsn; ++s) S.owner[s] = best_top[s] ? -1 : -2;
        for (size_t t = 0; t < best_q.size(); ++t) S.owner[best_q[t]] = assign[t];
        for (int s = nsn - 1; s >= 0; --s)       // parents precede children in this sweep
            if (S.owner[s] == -2) S.owner[s] = S.owner[S.sn_parent[s]];
        for (int s = 0; s < nsn; ++s)
            if (S.owner[s] >= 0 && S.sn_parent[s] != -1 && S.owner[S.sn_parent[s]] == -1) S.iface[S.sn_parent[s]] = 1;
    }
    S.top_owner.assign(nsn, -1);
    S.col_owner.assign(n, -1);
    S.xroot.assign(nsn, 0);
    if (NR > 1) {
        int t = 0;                                // consecutive fronts of a chain go to consecutive ranks
        for (int s = 0; s < nsn; ++s) {
            if (S.owner[s] != -1) continue;
            S.top_owner[s] = t++ % NR;
            for (int j = S.sn_start[s]; j < S.sn_start[s + 1]; ++j) S.col_owner[j] = S.top_owner[s];
        }
        for (int s = 0; s < nsn; ++s)
            S.xroot[s] = S.owner[s] >= 0 && S.sn_parent[s] != -1 && S.owner[S.sn_parent[s]] == -1;
    }
    step_mark("8 storage plan");
    // ---------------------------------------------------------------- 8. storage plan
    S.Loff.resize(nsn);
    S.Uoff.resize(nsn);
    S.CBoff.assign(nsn, 0);
    S.nnzL_stored = 0;
    S.flops_stored = 0;
    S.sum_r = (int64_t)S.rows.size();
    S.small.assign(nsn, 0);
    for (int s = 0; s < nsn; ++s) {
        const int64_t k = S.sn_start[s + 1] - S.sn_start[s], r = S.rows_ptr[s + 1] - S.rows_ptr[s];
        // (a subtree root hands its contribution block to other ranks from the Schur-update kernel: never "small")
        S.small[s] = S.owner[s] >= 0 && !S.xroot[s] && k <= opt.small_k_max && k + r <= opt.small_front_max;
    }
    auto is_small = [&](int s) { return S.small[s] != 0; };
    // Pool order: top fronts (summed across ranks in one all-reduce), then per rank its big fronts
    // (zero-filled and scattered into in HBM at the start of every refactorization), then per rank
    // its small fronts (assembled in shared memory, written exactly once, never zero-filled).
    int64_t off = 0;
    S.lu_big_begin.assign(NR, 0); S.lu_big_end.assign(NR, 0);
    for (int pass = 0; pass < 1 + 2 * NR; ++pass) {
        const int want_owner = pass == 0 ? -1 : (pass - 1) % NR;
        const bool want_small = pass > NR;
        if (pass >= 1 && pass <= NR) S.lu_big_begin[want_owner] = off;
        for (int s = 0; s < nsn; ++s) {
            if (S.owner[s] != want_owner || (pass > 0 && is_small(s) != want_small)) continue;
            int64_t k = S.sn_start[s + 1] - S.sn_start[s], r = S.rows_ptr[s + 1] - S.rows_ptr[s], f = k + r;
            S.Loff[s] = off; off = align2(off + f * k);
            S.Uoff[s] = off; off = align2(off + r * k);
            S.nnzL_stored += k * (k + 1) / 2 + k * r;
            S.max_front = std::max<int>(S.max_front, (int)f);
            S.max_k = std::max<int>(S.max_k, (int)k);
            double kd = (double)k, rd = (double)r;
            S.flops_stored += 2.0 * kd * kd * kd / 3.0 + 2.0 * kd * kd * rd + 2.0 * kd * rd * rd;
        }
        if (pass == 0) S.lu_top_size = off;
        if (pass >= 1 && pass <= NR) S.lu_big_end[want_owner] = off;
        if (pass == NR) S.lu_big_size = off;
    }
    S.lu_size = off;
    // direct-write eligibility
    S.direct.assign(nsn, 0);
    S.cb_assigned.assign(nsn, 0);
    for (int c = 0; c < nsn; ++c) {
        const int s = S.sn_parent[c];
        if (s == -1) continue;
        const int64_t rc = S.rows_ptr[c + 1] - S.rows_ptr[c];
        if (is_small(c) || is_small(s)) continue;   // small fronts are assembled in shared memory
        if (S.owner[c] != S.owner[s]) continue;     // contributions that cross the partition are summed by all-reduce
        if (S.child_ptr[s + 1] - S.child_ptr[s] != 1) continue;
        S.direct[c] = 1;
        const int64_t ks = S.sn_start[s + 1] - S.sn_start[s], rs = S.rows_ptr[s + 1] - S.rows_ptr[s];
        int64_t inside = 0;   // rows of c that are pivot columns of s
        for (int64_t t = S.rows_ptr[c]; t < S.rows_ptr[c + 1] && S.rel[t] < ks; ++t) ++inside;
        if (rc - inside == rs) S.cb_assigned[s] = 1;
    }
    {
        // Exchange slots: the contribution block of every subtree root lives in a permanent region at the start of
        // the pool, at the same offset on every rank.  Its owner's Schur update writes column b of the block into
        // the slot of the rank that owns global column rows[b] (its own memory or a peer's, over NVLink); the top
        // fronts' assembly then reads its own columns locally.
        int64_t ioff = 0;
        std::vector<char> have(nsn, 0);
        for (int s = 0; s < nsn; ++s)
            if (S.xroot[s]) {
                int64_t r = S.rows_ptr[s + 1] - S.rows_ptr[s];
                S.CBoff[s] = ioff; ioff += align2(r * r);
                have[s] = 1;
            }
        S.cb_xchg_size = ioff;
        S.cb_iface_size = 0;
        Arena arena;
        auto need = [&](int s) {
            if (have[s]) return;
            int64_t r = S.rows_ptr[s + 1] - S.rows_ptr[s];
            S.CBoff[s] = ioff + arena.alloc(align2(r * r));
            have[s] = 1;
        };
        for (int l = 0; l < S.nlevels; ++l) {
            for (int t = S.level_ptr[l]; t < S.level_ptr[l + 1]; ++t) need(S.level_sn[t]);
            // a direct child writes into its parent's block while its own level runs
            for (int t = S.level_ptr[l]; t < S.level_ptr[l + 1]; ++t)
                if (S.direct[S.level_sn[t]]) need(S.sn_parent[S.level_sn[t]]);
            for (int t = S.level_ptr[l]; t < S.level_ptr[l + 1]; ++t) {
                int s = S.level_sn[t];
                for (int u = S.child_ptr[s]; u < S.child_ptr[s + 1]; ++u) {
                    int c = S.child_idx[u];
                    if (S.xroot[c]) continue;
                    int64_t r = S.rows_ptr[c + 1] - S.rows_ptr[c];
                    arena.release(S.CBoff[c] - ioff, align2(r * r));
                }
            }
        }
        S.cb_size = ioff + arena.top();
    }
    step_mark("9 A -> factor scatter map");
    // ---------------------------------------------------------------- 9. A -> factor scatter map
    S.a_dst.resize(S.annz);
    S.a_sn.resize(S.annz);
    S.a_loc.resize(S.annz);
    std::atomic<int> outside{0};          // 1: entry outside the L structure, 2: outside the U structure
    const std::vector<int64_t> col_prefix(Ap, Ap + n + 1);
    parallel_rows(n, col_prefix, [&](int cbeg, int cend) {     // every entry writes only its own slots
    for (int c = cbeg; c < cend; ++c) {
        const int j = cinv[c];
        for (int64_t t = Ap[c]; t < Ap[c + 1]; ++t) {
            const int i = rinv[Ai[t]];
            const int s = S.col2sn[std::min(i, j)];
            const int c0 = S.sn_start[s], last = S.sn_start[s + 1] - 1;
            const int64_t k = last - c0 + 1, r = S.rows_ptr[s + 1] - S.rows_ptr[s], f = k + r;
            const int* rb = S.rows.data() + S.rows_ptr[s];
            if (j <= last) {   // lands in the L panel (diagonal block included)
                int64_t lr;
                if (i <= last) lr = i - c0;
                else {
                    const int* it = std::lower_bound(rb, rb + r, i);
                    if (it == rb + r || *it != i) { outside.store(1); return; }
                    lr = k + (it - rb);
                }
                S.a_dst[t] = S.Loff[s] + (int64_t)(j - c0) * f + lr;
                S.a_sn[t] = s;
                S.a_loc[t] = f < 65536 ? (int)(lr | ((int64_t)(j - c0) << 16)) : 0;
            } else {           // row in the pivot block, column beyond: U panel (stored transposed)
                const int* it = std::lower_bound(rb, rb + r, j);
                if (it == rb + r || *it != j) { outside.store(2); return; }
                S.a_dst[t] = S.Uoff[s] + (int64_t)(i - c0) * r + (it - rb);
                S.a_sn[t] = s;
                S.a_loc[t] = f < 65536 ? (int)((i - c0) | ((k + (it - rb)) << 16)) : 0;
            }
        }
    }
    });
    if (outside.load()) {
        err = outside.load() == 1 ? "internal: A entry outside L structure" : "internal: A entry outside U structure";
        return SMSLU_E_INTERNAL;
    }
    step_mark("end");
    return 0;
}

void exact_structure(const Symbolic& S, const int64_t* Ap, const int64_t* Ai,
                     std::vector<int64_t>& ptr, std::vector<int>& idx) {
    const int n = S.n;
    std::vector<int> rinv(n), cinv(n);
    for (int k = 0; k < n; ++k) { rinv[S.p[k]] = k; cinv[S.q[k]] = k; }
    Graph G = build_sym_graph(n, Ap, Ai, rinv.data(), cinv.data());
    std::vector<int> head(n, -1), next(n, -1);
    for (int j = n - 1; j >= 0; --j)
        if (S.parent[j] != -1) { next[j] = head[S.parent[j]]; head[S.parent[j]] = j; }
    ptr.assign(n + 1, 0);
    for (int j = 0; j < n; ++j) ptr[j + 1] = ptr[j] + S.colcount[j] - 1;
    idx.resize(ptr[n]);
    std::vector<int> marker(n, -1);
    for (int j = 0; j < n; ++j) {
        int64_t o = ptr[j];
        for (int64_t t = G.xadj[j]; t < G.xadj[j + 1]; ++t) {
            int i = G.adj[t];
            if (i > j && marker[i] != j) { marker[i] = j; idx[o++] = i; }
        }
        for (int c = head[j]; c != -1; c = next[c])
            for (int64_t t = ptr[c]; t < ptr[c + 1]; ++t) {
                int i = idx[t];
                if (i > j && marker[i] != j) { marker[i] = j; idx[o++] = i; }
            }
        std::sort(idx.begin() + ptr[j], idx.begin() + o);
    }
}

}  // namespace smslu

namespace smslu {

void relabel_csc(int n, const int* post, int64_t base, int64_t* colptr, int64_t* idx, double* val) {
    const int64_t nnz = colptr[n] - base;
    std::vector<int64_t> np(n + 1, 0);
    for (int j = 0; j < n; ++j) np[post[j] + 1] = colptr[j + 1] - colptr[j];
    for (int c = 0; c < n; ++c) np[c + 1] += np[c];
    std::vector<int64_t> ti(idx ? nnz : 0);
    std::vector<double> tv(val ? nnz : 0);
    std::vector<std::pair<int64_t, double>> col;
    for (int j = 0; j < n; ++j) {
        const int64_t lo = colptr[j] - base, cnt = colptr[j + 1] - colptr[j], o = np[post[j]];
        if (!idx) {            // values only: the order inside the column still follows the relabelled rows, which
            continue;          // needs the row indices -- callers that want values pass idx too
        }
        col.resize(cnt);
        for (int64_t t = 0; t < cnt; ++t) col[t] = {(int64_t)post[idx[lo + t] - base], val ? val[lo + t] : 0.0};
        std::sort(col.begin(), col.end(), [](const std::pair<int64_t, double>& a, const std::pair<int64_t, double>& b) { return a.first < b.first; });
        for (int64_t t = 0; t < cnt; ++t) { ti[o + t] = col[t].first + base; if (val) tv[o + t] = col[t].second; }
    }
    if (idx) std::copy(ti.begin(), ti.end(), idx);
    if (val && idx) std::copy(tv.begin(), tv.end(), val);
    for (int c = 0; c <= n; ++c) colptr[c] = np[c] + base;
}

void export_factors(const Symbolic& S, const std::vector<int64_t>& ptr, const std::vector<int>& idx,
                    const double* lu, int64_t base, int64_t* Lp, int64_t* Li, double* Lx,
                    int64_t* Up, int64_t* Ui, double* Ux, const char* col_mine) {
    const int n = S.n;
    // position of row i inside the front of supernode s
    auto front_pos = [&](int s, int i) -> int64_t {
        const int c0 = S.sn_start[s], last = S.sn_start[s + 1] - 1;
        if (i <= last) return i - c0;
        const int* rb = S.rows.data() + S.rows_ptr[s];
        const int64_t r = S.rows_ptr[s + 1] - S.rows_ptr[s];
        return (last - c0 + 1) + (std::lower_bound(rb, rb + r, i) - rb);
    };
    // ---- L: column j = unit diagonal + exact structure below it
    if (Lp) {
        Lp[0] = base;
        for (int j = 0; j < n; ++j) Lp[j + 1] = Lp[j] + 1 + (ptr[j + 1] - ptr[j]);
    }
    if (Li || Lx) {
        int64_t w = 0;
        for (int j = 0; j < n; ++j) {
            const int s = S.col2sn[j];
            const int c0 = S.sn_start[s];
            const int64_t k = S.sn_start[s + 1] - c0, f = k + (S.rows_ptr[s + 1] - S.rows_ptr[s]);
            const double* col = lu + S.Loff[s] + (int64_t)(j - c0) * f;
            if (Li) Li[w] = j + base;
            if (Lx) Lx[w] = (!col_mine || col_mine[j]) ? 1.0 : 0.0;
            ++w;
            for (int64_t t = ptr[j]; t < ptr[j + 1]; ++t, ++w) {
                if (Li) Li[w] = idx[t] + base;
                if (Lx) Lx[w] = col[front_pos(s, idx[t])];
            }
        }
    }
    // ---- U: row i of U has the structure of column i of L; emit by columns with sorted rows
    if (Up || Ui || Ux) {
        std::vector<int64_t> cnt(n + 1, 0);
        for (int i = 0; i < n; ++i) {
            ++cnt[i + 1];   // diagonal
            for (int64_t t = ptr[i]; t < ptr[i + 1]; ++t) ++cnt[idx[t] + 1];
        }
        for (int j = 0; j < n; ++j) cnt[j + 1] += cnt[j];
        if (Up) for (int j = 0; j <= n; ++j) Up[j] = cnt[j] + base;
        if (Ui || Ux) {
            std::vector<int64_t> w(cnt.begin(), cnt.end() - 1);
            for (int i = 0; i < n; ++i) {   // ascending i => rows sorted inside every column
                const int s = S.col2sn[i];
                const int c0 = S.sn_start[s], last = S.sn_start[s + 1] - 1;
                const int64_t k = last - c0 + 1, r = S.rows_ptr[s + 1] - S.rows_ptr[s], f = k + r;
                const double* P = lu + S.Loff[s];
                const double* T = lu + S.Uoff[s];
                // strictly-upper entries of row i: columns idx[t] > i
                for (int64_t t = ptr[i]; t < ptr[i + 1]; ++t) {
                    const int j = idx[t];
                    double v;
                    if (j <= last) v = P[(int64_t)(j - c0) * f + (i - c0)];
                    else v = T[(int64_t)(i - c0) * r + (front_pos(s, j) - k)];
                    // column j receives rows in ascending order, but the diagonal of column j must
                    // come last: it is written when i == j below, after all smaller rows
                    int64_t o = w[j]++;
                    if (Ui) Ui[o] = i + base;
                    if (Ux) Ux[o] = v;
                }
                int64_t o = w[i]++;   // all rows < i of column i were emitted in earlier iterations
                if (Ui) Ui[o] = i + base;
                if (Ux) Ux[o] = P[(int64_t)(i - c0) * f + (i - c0)];
            }
        }
    }
}

}  // namespace smslu

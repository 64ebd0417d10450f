// C ABI of libsmslu.so (include/smslu.h): handle, upload of the symbolic layout, level schedules,
// and the numeric entry points.  No CPU fallback: numeric calls need a CUDA device.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>     // types and prototypes only: the library is loaded lazily (see nccl_api below)

#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/smslu.h"
#include "kernels.cuh"
#include "symbolic.hpp"

using namespace smslu;

namespace {

// NCCL is resolved with dlopen on first multi-GPU use instead of being a link-time dependency: a process
// that never partitions a matrix never loads it, and a host that already carries its own NCCL (PyTorch
// bundles one under the same soname) is not handed a second, possibly older, copy.
struct NcclApi {
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    bool ok = false;
};
NcclApi& nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (lib) {
            api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))dlsym(lib, "ncclCommInitRank");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(lib, "ncclCommDestroy");
            api.AllReduce = (decltype(api.AllReduce))dlsym(lib, "ncclAllReduce");
            api.AllGather = (decltype(api.AllGather))dlsym(lib, "ncclAllGather");
            api.GroupStart = (decltype(api.GroupStart))dlsym(lib, "ncclGroupStart");
            api.GroupEnd = (decltype(api.GroupEnd))dlsym(lib, "ncclGroupEnd");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(lib, "ncclGetErrorString");
            api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.AllGather && api.GroupStart &&
                     api.GroupEnd && api.GetErrorString;
        }
    }
    return api;
}

enum LaunchKind { L_ZERO = SMSLU_K_ZERO, L_EXTEND = SMSLU_K_EXTEND, L_SMALL = SMSLU_K_SMALL, L_PANEL = SMSLU_K_PANEL,
                  L_GEMM = SMSLU_K_GEMM, L_FWD = SMSLU_K_FWD, L_BWD = SMSLU_K_BWD,
                  L_FWD_SMALL = SMSLU_K_FWD_SMALL, L_BWD_SMALL = SMSLU_K_BWD_SMALL,
                  // partitioned top (timed as SMSLU_K_ALLREDUCE = "exchange"): publish panels to the peers, signal, wait
                  L_REPL = 100, L_SIGNAL = 101, L_WAIT = 102,
                  // persistent chain solves (single right-hand side)
                  L_FWD_CHAIN = 103, L_BWD_RECT = 104, L_BWD_CHAIN = 105,
                  // 32 right-hand sides per sweep on the FP64 tensor pipe (big fronts)
                  L_FWD32 = 106, L_BWD32 = 107 };

constexpr int NLANES = 4;
constexpr int GEMM_STRIP_DEFAULT = 1;
constexpr bool CHAINS_DEFAULT = false;
constexpr int CHAIN_MAXC_DEFAULT = 32;

constexpr int PANEL_GROUP_CTAS = 296;  // panel launches aim at about this many CTAs (2 per SM) ...
constexpr int PANEL_GROUP_MAX = 16;    // ... by giving one CTA up to this many 128-row tiles of its front
constexpr int PANEL_FIT_MAX = 40;      // ... or up to this many where that makes the launch fit one wave of resident CTAs
constexpr int FWD_WIDE_TILES = 148;   // 64-row forward tiles of a level beyond which 256-row tiles are used

struct Launch {
    int kind;
    int64_t off;
    int ntasks;
    int fmax;
    int level;     // launches of one level of one schedule are independent across the two lanes
    int lane;      // 0 = big-front kernels (main stream), 1 / 2 = small-front kernels (auxiliary streams)
};

double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

}  // namespace

struct smslu_handle_s {
    int n = 0;
    int index_base = 0;
    int64_t annz = 0;
    std::vector<int64_t> Ap, Ai;   // 0-based pattern as given by the caller
    smslu_options_t opt{};
    Symbolic S;
    bool analyzed = false, uploaded = false, factored = false;
    std::string err;
    smslu_stats_t st{};

    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t aux_stream[NLANES - 1] = {nullptr, nullptr, nullptr};   // lanes 1..3 of a level
    cudaEvent_t ev_fork = nullptr, ev_join[NLANES - 1] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_scatter = nullptr;        // big fronts' panels zero-filled and scattered into
    bool scatter_pending = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    std::vector<void*> dev_allocs;
    DevCtx cx{};
    int64_t* d_big_dst = nullptr;        // entries of A that land in big fronts: offset into lu,
    const int *d_big_row = nullptr, *d_big_src = nullptr;   // original row, index into nzval
    int64_t nnz_big = 0;
    const double* cur_av = nullptr;      // device nzval of the refactorization being enqueued
    int64_t *d_rowptr = nullptr, *d_rowidx = nullptr;
    int *d_p = nullptr, *d_q = nullptr;
    int* d_post = nullptr;               // ORD_GIVEN with a relabelling postorder: caller position of internal index (else null)
    double *d_Rs = nullptr, *d_aval = nullptr, *d_w = nullptr, *d_z = nullptr, *d_xb = nullptr;
    int4* d_tasks = nullptr;
    int2* d_inv_tasks = nullptr;        // (front, 32-column block) pairs of k_diag_inverse
    int n_inv_tasks = 0;
    std::vector<int> asm_meta;                       // (child, first column) pairs of the assembly tasks
    std::vector<Launch> fac, fwd, bwd;               // this rank's supernodes (everything when nranks == 1)
    std::vector<Launch> fac_top, fwd_top, bwd_top;   // top of the tree, replicated on every rank
    // single right-hand side: the same sweeps with every set of parallel chains of fronts in one persistent kernel
    std::vector<Launch> fwd1, bwd1, fwd_top1, bwd_top1;
    // 32-wide sweeps (one GPU): the small fronts' launches of fwd / bwd + k_fwd32 / k_bwd32 for every big front;
    // their work vectors are allocated by the first solve with that many right-hand sides (ensure_wide)
    std::vector<Launch> fwd32, bwd32;
    int64_t bpart32_slots = 0;
    double *d_w32 = nullptr, *d_z32 = nullptr, *d_xb32 = nullptr;
    bool have_wide = false, wide_failed = false;
    DevCtx cx32{};
    int* d_chain_desc = nullptr;
    int solve_epoch = 0;
    int64_t chain_part_vecs = 0;
    int rank = 0, nranks = 1;
    ncclComm_t comm = nullptr;
    int4* d_vtasks = nullptr;            // interface fronts: x = front, y = virtual child, z..w = range in d_vlist
    int* d_vlist = nullptr;              // this rank's subtree roots below each interface front
    int nvtasks = 0;
    int64_t vupd_off = 0, vupd_len = 0;  // region of cx.upd holding the virtual children's vectors
    int* d_colowner = nullptr;           // owner of every permuted column (-1 = top)
    // distributed top of the tree: peers' pools mapped through CUDA IPC, cross-GPU flags, per-level publish lists
    int* d_xflags = nullptr;             // [sync point * nranks + source rank] = epoch of the last signal
    int64_t* d_segs = nullptr;           // (offset, length) pairs of the panels this rank publishes
    int* d_wait_slots = nullptr;
    int4* d_trow_tasks = nullptr;        // runs of U12' rows this rank owns (published at the end of a refactorization)
    int n_trow_tasks = 0;
    int epoch = 0;
    int* d_barrier = nullptr;            // one int, all-reduced as a barrier between the phases
    std::vector<void*> ipc_opened;
    int64_t peer_bytes_refactor = 0;     // bytes this rank stores into peer memory per refactorization
    int64_t bpart_slots = 0, ncounters = 0;

    bool own_stream = true;
    cudaStream_t user_stream = nullptr;
    bool have_user_stream = false;
    bool profile = false;
    // SMSLU_LEVEL_TIMES=1 (debug): an event on the main stream at every level boundary, printed by smslu_sync
    bool level_times = getenv("SMSLU_LEVEL_TIMES") != nullptr;
    struct LevelMark { cudaEvent_t ev; int level; int nlaunch; int ntasks; const void* sched; };
    std::vector<LevelMark> lvl_marks;
    std::vector<cudaEvent_t> pev;          // event pool for per-launch profiling
    std::vector<int> pev_kind;             // kind of the launch between pev[2i], pev[2i+1]
    size_t pev_used = 0;
    bool pending_refactor = false;         // an async refactor has not been checked yet
    int* h_flag = nullptr;                 // pinned

    std::vector<int64_t> ex_ptr;   // exact structure, built lazily for get_factors
    std::vector<int> ex_idx;
    bool have_exact = false;
};

namespace {

int fail(smslu_handle_t h, int code, const std::string& msg) {
    if (h) h->err = msg;
    return code;
}

// Every entry point that touches the device selects the handle's GPU; the caller's current device is put back on return.
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); } }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(h, e_ == cudaErrorMemoryAllocation ? SMSLU_E_OOM : SMSLU_E_CUDA,          \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                      \
    } while (0)

#define NCCLCHK(call)                                                                                 \
    do {                                                                                          \
        ncclResult_t r_ = (call);                                                                 \
        if (r_ != ncclSuccess) return fail(h, SMSLU_E_NCCL, std::string(#call) + ": " + nccl_api().GetErrorString(r_)); \
    } while (0)

template <class T>
int dev_alloc(smslu_handle_t h, T** p, size_t count) {
    *p = nullptr;
    size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    cudaError_t e = cudaMalloc((void**)p, bytes);
    if (e != cudaSuccess) return fail(h, SMSLU_E_OOM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    h->dev_allocs.push_back(*p);
    return 0;
}
template <class T>
int dev_upload(smslu_handle_t h, T** p, const std::vector<T>& v) {
    int rc = dev_alloc(h, p, v.size());
    if (rc) return rc;
    if (!v.empty()) CU(cudaMemcpy(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

bool is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// ---------------------------------------------------------------- schedules
// Launch lists are built once per pattern.  With a partition (nranks > 1) a rank runs
//   phase A: the supernodes it owns, level by level (their contributions into top fronts included),
//   all-reduce of the top panels and interface contribution blocks,
//   phase B: the top of the tree, level by level, identically on every rank.
// phase 0 = this rank's supernodes, phase 1 = top supernodes (owner -1).
void build_schedules(smslu_handle_t h, std::vector<int4>& tasks) {
    const Symbolic& S = h->S;
    const int rank = h->rank;
    auto K = [&](int s) { return S.sn_start[s + 1] - S.sn_start[s]; };
    auto R = [&](int s) { return (int64_t)(S.rows_ptr[s + 1] - S.rows_ptr[s]); };
    auto NC = [&](int s) { return S.child_ptr[s + 1] - S.child_ptr[s]; };
    auto SMALL = [&](int s) { return S.small[s] != 0; };
    int64_t ncounters = 0, slots = 0;
    int cur_level = 0, big_lane = 0;
    // Schur-update CTAs walk strips of this many row tiles of one tile column (k_gemm_strip); 1 = one tile per CTA (k_gemm_cb)
    const bool panel_fit = !(getenv("SMSLU_PANEL_FIT") && atoi(getenv("SMSLU_PANEL_FIT")) == 0);   // debugging aid / A-B switch
    const int gemm_strip = getenv("SMSLU_GEMM_STRIP") ? std::max(1, atoi(getenv("SMSLU_GEMM_STRIP"))) : GEMM_STRIP_DEFAULT;
    auto push = [&](std::vector<Launch>& v, int kind, int64_t off, int fmax) {
        int nt = (int)((int64_t)tasks.size() - off);
        // small fronts: factor classes up to 40 rows on lane 1, the wider ones on lane 2; solves on lane 1;
        // big fronts of the factorization: lane 0 (wide pivot blocks) or lane 3 (at most 64 pivot columns)
        const int lane = kind == L_SMALL ? (fmax <= 40 ? 1 : 2) : ((kind == L_FWD_SMALL || kind == L_BWD_SMALL) ? 1 :
                         ((kind == L_ZERO || kind == L_EXTEND || kind == L_PANEL || kind == L_GEMM) ? big_lane :
                          ((kind == L_FWD || kind == L_BWD) && (fmax & 255) <= NB ? 3 : 0)));     // solves: narrow big fronts on lane 3
        if (nt > 0) v.push_back(Launch{kind, off, nt, fmax, cur_level, lane});
    };
    for (int ph = 0; ph < 2; ++ph) {
        std::vector<Launch>& fac = ph == 0 ? h->fac : h->fac_top;
        std::vector<Launch>& fwd = ph == 0 ? h->fwd : h->fwd_top;
        std::vector<Launch>& bwd = ph == 0 ? h->bwd : h->bwd_top;
        fac.clear(); fwd.clear(); bwd.clear();
        const int mine = ph == 0 ? rank : -1;
        auto IN = [&](int s) { return S.owner[s] == mine; };
        for (int l = 0; l < S.nlevels; ++l) {
            cur_level = l;
            const int* sn = S.level_sn.data() + S.level_ptr[l];
            const int cnt = S.level_ptr[l + 1] - S.level_ptr[l];
            int64_t off;
            if (ph == 0) {                          // (the top's factorization schedule: build_top_fac)
            // small fronts by shared-memory class
            const int classes[5] = {32, 40, 48, 64, front_small_limit()};
            int lo = 0;
            for (int ci = 0; ci < 5; ++ci) {
                off = (int64_t)tasks.size();
                for (int t = 0; t < cnt; ++t) {
                    int s = sn[t];
                    int64_t f = K(s) + R(s);
                    if (IN(s) && SMALL(s) && f > lo && f <= classes[ci]) tasks.push_back(make_int4(s, 0, 0, 0));
                }
                push(fac, L_SMALL, off, classes[ci]);
                lo = classes[ci];
            }
            // Big fronts in two independent groups, each on its own lane: pivot blocks of at most 64 columns (one or
            // two panel steps, then their Schur update) do not wait for the 3- and 4-step fronts of the level.
            // (Only where the level's big fronts do not fill the machine anyway: with thousands of Schur-update tiles the
            // two groups would just compete, and the split costs extra launches.)
            int64_t level_tiles = 0;
            for (int t = 0; t < cnt; ++t)
                if (IN(sn[t]) && !SMALL(sn[t])) { const int64_t nt = (R(sn[t]) + GEMM_TILE - 1) / GEMM_TILE; level_tiles += nt * nt; }
            const bool split = level_tiles <= 4096;
            auto NARROW = [&](int s) { return split && K(s) <= 2 * NB; };
            for (int grp = 0; grp < 2; ++grp) {
                big_lane = grp == 1 ? 3 : 0;
                // Zero the contribution blocks that receive '+=' contributions.  Extend-add children add
                // at this level; a direct child adds one level earlier, so its parent is zeroed there.
                // Interface blocks are zero-filled once, before phase A, and arrive here all-reduced.
                off = (int64_t)tasks.size();
                auto zero_tasks = [&](int s) {
                    int64_t tiles = (R(s) * R(s) + ZERO_TILE - 1) / ZERO_TILE;
                    for (int64_t i = 0; i < tiles; ++i) tasks.push_back(make_int4(s, (int)i, 0, 0));
                };
                for (int t = 0; t < cnt; ++t) {
                    int s = sn[t];
                    if (!IN(s) || NARROW(s) != (grp == 1)) continue;
                    if (S.direct[s] && !S.cb_assigned[S.sn_parent[s]] && R(S.sn_parent[s]) > 0) zero_tasks(S.sn_parent[s]);
                }
                push(fac, L_ZERO, off, 0);
                // assembly of the children's contribution blocks into big parents: one launch, the CTA that
                // owns a range of destination columns also zero-fills its part of the parent's block.
                // In phase A a top parent receives this rank's subtree contributions here as well.
                off = (int64_t)tasks.size();
                int64_t asm_fmax = 0;                  // rows of the largest parent assembled in this launch
                for (int t = 0; t < cnt; ++t) {
                    int s = sn[t];
                    if (SMALL(s) || NC(s) == 0 || NARROW(s) != (grp == 1)) continue;   // small parents pull their children
                    if (!IN(s)) continue;       // (a top parent pulls the subtree roots' blocks in the top phase)
                    std::vector<int> kids;
                    for (int u = S.child_ptr[s]; u < S.child_ptr[s + 1]; ++u) {
                        const int c = S.child_idx[u];
                        if (IN(c) && !S.direct[c] && R(c) > 0) kids.push_back(c);
                    }
                    if (kids.empty()) continue;
                    const bool zero = IN(s) && R(s) > 0 && !S.iface[s];
                    const int64_t f = K(s) + R(s);
                    asm_fmax = std::max<int64_t>(asm_fmax, f);
                    for (int64_t pb0 = 0; pb0 < f; pb0 += ASM_COLS) {
                        const int64_t moff = (int64_t)h->asm_meta.size();
                        int np = 0;
                        for (int c : kids) {            // first column of c whose parent position is >= pb0
                            const int* rb = S.rel.data() + S.rows_ptr[c];
                            const int* re = S.rel.data() + S.rows_ptr[c + 1];
                            const int* it = std::lower_bound(rb, re, (int)pb0);
                            if (it == re || *it >= pb0 + ASM_COLS) continue;       // nothing of c lands in this range
                            h->asm_meta.push_back(c); h->asm_meta.push_back((int)(it - rb));
                            ++np;
                        }
                        if (np == 0 && !zero) continue;
                        for (int ch = 0; ch * (int64_t)asm_chunk_rows(f) < f; ++ch)          // one task per row chunk of a tall parent
                            tasks.push_back(make_int4(s, (int)pb0 | (ch << 20), (int)moff, (zero ? 1 : 0) | (np << 8)));
                    }
                }
                push(fac, L_EXTEND, off, (int)std::min<int64_t>(asm_fmax, 1 << 30));
                // big fronts: left-looking panel steps (one launch per 32 pivot columns), then Schur update
                int max_blk = 0;
                for (int t = 0; t < cnt; ++t) if (IN(sn[t]) && !SMALL(sn[t]) && NARROW(sn[t]) == (grp == 1)) max_blk = std::max(max_blk, (K(sn[t]) + NB - 1) / NB);
                for (int g = 0; g < max_blk; ++g) {
                    // rows per tile: 128 when that already gives the machine enough CTAs, else 32.  On bulk levels a CTA
                    // takes up to `group` consecutive tiles of its front (one staging and one factorization of D_gg
                    // for all of them): group = tiles of the launch / (2 CTAs x 148 SMs).
                    int rows = PANEL_ROWS;
                    auto tiles_of = [&](int s, int rws, int& tl, int& tt, int& ti) {
                        const int k = K(s);
                        const int64_t r = R(s), f = k + r;
                        const int j1 = std::min(k, (g + 1) * NB);
                        tl = (int)((f - j1 + rws - 1) / rws);            // rows below the diagonal block
                        tt = (int)((r + rws - 1) / rws);                 // rows of U12'
                        ti = (k - j1 + rws - 1) / rws;                   // columns right of it
                    };
                    auto in_step = [&](int s) {
                        return IN(s) && !SMALL(s) && NARROW(s) == (grp == 1) && g < (K(s) + NB - 1) / NB;
                    };
                    int64_t tiles_total = 0;
                    for (int t = 0; t < cnt; ++t) {
                        if (!in_step(sn[t])) continue;
                        int tl, tt, ti; tiles_of(sn[t], rows, tl, tt, ti);
                        tiles_total += std::max(1, tl + tt + ti);
                    }
                    if (tiles_total > 0 && tiles_total < 120) rows = PANEL_ROWS_TOP;     // few tiles: small CTAs
                    int group = rows == PANEL_ROWS ? (int)std::min<int64_t>(PANEL_GROUP_MAX, std::max<int64_t>(1, tiles_total / PANEL_GROUP_CTAS)) : 1;
                    if (panel_fit && rows == PANEL_ROWS && tiles_total > PANEL_GROUP_CTAS) {
                        // the per-front rounding (ceil(tiles / group) CTAs each) pushes a launch past the 2 x 148 resident CTAs and
                        // the stragglers cost a second round: widen the groups until the launch fits one wave
                        auto total_ctas = [&](int gsz) {
                            int64_t tot = 0;
                            for (int t = 0; t < cnt; ++t) {
                                if (!in_step(sn[t])) continue;
                                int tl, tt, ti; tiles_of(sn[t], rows, tl, tt, ti);
                                tot += (std::max(1, tl + tt + ti) + gsz - 1) / gsz;
                            }
                            return tot;
                        };
                        while (group < PANEL_FIT_MAX && total_ctas(group) > PANEL_GROUP_CTAS) ++group;
                    }
                    off = (int64_t)tasks.size();
                    for (int t = 0; t < cnt; ++t) {
                        int s = sn[t];
                        if (!in_step(s)) continue;
                        int tl, tt, ti; tiles_of(s, rows, tl, tt, ti);
                        const int ntile = std::max(1, tl + tt + ti);         // someone has to factor D_gg
                        const int nctas = (ntile + group - 1) / group;
                        const int cidx = (int)ncounters++;
                        for (int c = 0; c < nctas; ++c) {                    // even split of the tiles over the CTAs
                            const int t0 = (int)((int64_t)ntile * c / nctas), t1 = (int)((int64_t)ntile * (c + 1) / nctas);
                            tasks.push_back(make_int4(s, g | ((t1 - t0) << 4) | (nctas << 16), t0, cidx));
                        }
                    }
                    push(fac, L_PANEL, off, g | (rows << 8));
                }
                off = (int64_t)tasks.size();
                for (int t = 0; t < cnt; ++t) {
                    int s = sn[t];
                    int64_t r = R(s);
                    if (!IN(s) || SMALL(s) || NARROW(s) != (grp == 1)) continue;
                    int nt = (int)((r + GEMM_TILE - 1) / GEMM_TILE);
                    const int flags = (NC(s) > 0 ? 1 : 0) | (S.direct[s] ? 2 : 0) |
                                      (S.direct[s] && S.cb_assigned[S.sn_parent[s]] ? 4 : 0) |
                                      (S.xroot[s] ? 8 : 0) |      // subtree root: columns go to their owners' pools
                                      (S.direct[s] && r == K(S.sn_parent[s]) + R(S.sn_parent[s]) ? 32 : 0);   // chain link: identity map
                    for (int j = 0; j < nt; ++j)
                        for (int i = 0; i < nt; i += gemm_strip) tasks.push_back(make_int4(s, gemm_strip > 1 ? (i | (std::min(gemm_strip, nt - i) << 16)) : i, j, flags));
                }
                push(fac, L_GEMM, off, gemm_strip > 1 ? 1 : 0);
            }
            big_lane = 0;
            }
            // forward solve level: warp-per-front kernel for the small fronts; narrow (k <= 32) and
            // wide big fronts go to separate launches because the kernel stages the whole pivot block
            // in shared memory (8 KB vs up to 129 KB)
            off = (int64_t)tasks.size();
            for (int t = 0; t < cnt; ++t) if (IN(sn[t]) && SMALL(sn[t])) tasks.push_back(make_int4(sn[t], 0, 0, 0));
            push(fwd, L_FWD_SMALL, off, 0);
            for (int cls = 0; cls < 2; ++cls) {
                // 64-row tiles where the level is a few fronts (latency); 256-row tiles where the 64-row tiles
                // would not be resident at once anyway (every tile repeats the pivot-block solve)
                int64_t tiles64 = 0, tiles128 = 0;
                for (int t = 0; t < cnt; ++t) {
                    int s = sn[t];
                    if (!IN(s) || SMALL(s) || (K(s) > NB) != (cls == 1)) continue;
                    tiles64 += std::max<int64_t>(1, (R(s) + FWD_ROWS - 1) / FWD_ROWS);
                    tiles128 += std::max<int64_t>(1, (R(s) + FWD_ROWS_MID - 1) / FWD_ROWS_MID);
                }
                const int rows = tiles64 <= FWD_WIDE_TILES ? FWD_ROWS : (tiles128 <= 2 * FWD_WIDE_TILES ? FWD_ROWS_MID : FWD_ROWS_WIDE);
                off = (int64_t)tasks.size();
                int kmax = 0;
                for (int t = 0; t < cnt; ++t) {
                    int s = sn[t];
                    if (!IN(s) || SMALL(s) || (K(s) > NB) != (cls == 1)) continue;
                    kmax = std::max(kmax, K(s));
                    int nt = (int)std::max<int64_t>(1, (R(s) + rows - 1) / rows);
                    for (int i = 0; i < nt; ++i) tasks.push_back(make_int4(s, i, 0, 0));
                }
                push(fwd, L_FWD, off, kmax | (rows << 8));
            }
        }
        for (int l = S.nlevels - 1; l >= 0; --l) {
            cur_level = l;
            int64_t off0 = (int64_t)tasks.size();
            for (int t = S.level_ptr[l]; t < S.level_ptr[l + 1]; ++t)
                if (IN(S.level_sn[t]) && SMALL(S.level_sn[t])) tasks.push_back(make_int4(S.level_sn[t], 0, 0, 0));
            push(bwd, L_BWD_SMALL, off0, 0);
            for (int cls = 0; cls < 2; ++cls) {
                int64_t off = (int64_t)tasks.size();
                int kmax = 0;
                // few row tiles on the level: split the pivot columns of every tile over 2 or 4 CTAs
                int64_t lvl_tiles = 0;
                for (int t = S.level_ptr[l]; t < S.level_ptr[l + 1]; ++t) {
                    int s = S.level_sn[t];
                    if (!IN(s) || SMALL(s) || (K(s) > NB) != (cls == 1)) continue;
                    lvl_tiles += std::max<int64_t>(1, (R(s) + BWD_ROWS - 1) / BWD_ROWS);
                }
                const int want = cls == 0 ? 1 : (lvl_tiles <= 37 ? 4 : (lvl_tiles <= 148 ? 2 : 1));
                for (int t = S.level_ptr[l]; t < S.level_ptr[l + 1]; ++t) {
                    int s = S.level_sn[t];
                    if (!IN(s) || SMALL(s) || (K(s) > NB) != (cls == 1)) continue;
                    kmax = std::max(kmax, K(s));
                    int nt = (int)std::max<int64_t>(1, (R(s) + BWD_ROWS - 1) / BWD_ROWS);
                    const int nsp = (R(s) > 0 && K(s) >= 16 * want) ? want : 1;
                    for (int i = 0; i < nt; ++i)
                        for (int cp = 0; cp < nsp; ++cp) tasks.push_back(make_int4(s, i | (cp << 24) | (nsp << 28), nt, (int)slots));
                    if (nt * nsp > 1) slots += nt;
                }
                push(bwd, L_BWD, off, kmax);
            }
        }
    }
    h->bpart_slots = slots;
    h->ncounters = ncounters;
    // 32-wide sweeps (one GPU only): small fronts as in fwd / bwd, every big front through the tensor-pipe kernels
    h->fwd32.clear(); h->bwd32.clear(); h->bpart32_slots = 0;
    if (h->nranks == 1) {
        for (int l = 0; l < S.nlevels; ++l) {
            cur_level = l;
            int64_t off = (int64_t)tasks.size();
            for (int t = S.level_ptr[l]; t < S.level_ptr[l + 1]; ++t) if (SMALL(S.level_sn[t])) tasks.push_back(make_int4(S.level_sn[t], 0, 0, 0));
            push(h->fwd32, L_FWD_SMALL, off, 0);
            off = (int64_t)tasks.size();
            int64_t wide_tiles = 0;
            for (int t = S.level_ptr[l]; t < S.level_ptr[l + 1]; ++t)
                if (!SMALL(S.level_sn[t])) wide_tiles += std::max<int64_t>(1, (R(S.level_sn[t]) + RB_WIDE_ROWS - 1) / RB_WIDE_ROWS);
            const int trows = wide_tiles < 74 ? RB_WIDE_ROWS_TOP : RB_WIDE_ROWS;      // few tiles: spread the level over more SMs (one wave of 64-row tiles)
            for (int t = S.level_ptr[l]; t < S.level_ptr[l + 1]; ++t) {
                const int s = S.level_sn[t];
                if (SMALL(s)) continue;
                const int nt = (int)std::max<int64_t>(1, (R(s) + trows - 1) / trows);
                for (int i = 0; i < nt; ++i) tasks.push_back(make_int4(s, i, 0, 0));
            }
            if ((int64_t)tasks.size() > off) h->fwd32.push_back(Launch{L_FWD32, off, (int)((int64_t)tasks.size() - off), trows, l, 0});
        }
        int64_t slots32 = 0;
        for (int l = S.nlevels - 1; l >= 0; --l) {
            cur_level = l;
            int64_t off = (int64_t)tasks.size();
            for (int t = S.level_ptr[l]; t < S.level_ptr[l + 1]; ++t) if (SMALL(S.level_sn[t])) tasks.push_back(make_int4(S.level_sn[t], 0, 0, 0));
            push(h->bwd32, L_BWD_SMALL, off, 0);
            off = (int64_t)tasks.size();
            for (int t = S.level_ptr[l]; t < S.level_ptr[l + 1]; ++t) {
                const int s = S.level_sn[t];
                if (SMALL(s)) continue;
                const int nt = (int)std::max<int64_t>(1, (R(s) + RB_WIDE_ROWS - 1) / RB_WIDE_ROWS);
                for (int i = 0; i < nt; ++i) tasks.push_back(make_int4(s, i, nt, (int)slots32));
                if (nt > 1) slots32 += nt;
            }
            if ((int64_t)tasks.size() > off) h->bwd32.push_back(Launch{L_BWD32, off, (int)((int64_t)tasks.size() - off), 0, l, 0});
        }
        h->bpart32_slots = slots32;
    }
}

// Chains of fronts (links of at most KW pivot columns cut out of one wide separator; each link the only child of the
// next, its row list exactly the next link's front) are solved by persistent kernels when there is one right-hand side:
// find the sets of level-aligned parallel chains in a sweep's schedule, build their descriptors, and derive the chain
// variants of the forward / backward schedules (the per-level launches of the chains' wide fronts are dropped).
struct ChainSet { int l0, l1; std::vector<std::vector<int>> chains; };

void build_chain_schedules(smslu_handle_t h, std::vector<int4>& tasks, const std::vector<int>& dchild_ptr,
                           const std::vector<int>& dchild_idx, std::vector<int>& desc, std::vector<int>& links,
                           std::vector<int>& boffs) {
    const Symbolic& S = h->S;
    auto K = [&](int s) { return S.sn_start[s + 1] - S.sn_start[s]; };
    auto R = [&](int s) { return (int64_t)(S.rows_ptr[s + 1] - S.rows_ptr[s]); };
    // default off until validated on the GPU in this round: SMSLU_CHAINS=1 switches the persistent chain kernels on
    const bool enabled = CHAINS_DEFAULT ? !(getenv("SMSLU_NO_CHAINS") && atoi(getenv("SMSLU_NO_CHAINS")) != 0)
                                        : (getenv("SMSLU_CHAINS") && atoi(getenv("SMSLU_CHAINS")) != 0);
    // only sets of at most this many parallel chains (1 = the separators that stand alone on their levels: the top of the tree)
    const int max_chains = getenv("SMSLU_CHAIN_MAXC") ? std::max(1, std::min(32, atoi(getenv("SMSLU_CHAIN_MAXC")))) : CHAIN_MAXC_DEFAULT;
    for (int ph = 0; ph < 2; ++ph) {
        const std::vector<Launch>& fwd = ph == 0 ? h->fwd : h->fwd_top;
        const std::vector<Launch>& bwd = ph == 0 ? h->bwd : h->bwd_top;
        std::vector<Launch>& fwd1 = ph == 0 ? h->fwd1 : h->fwd_top1;
        std::vector<Launch>& bwd1 = ph == 0 ? h->bwd1 : h->bwd_top1;
        const int mine = ph == 0 ? h->rank : -1;
        auto WIDE = [&](int s) { return S.owner[s] == mine && !S.small[s] && K(s) > NB; };
        // wide fronts per level
        std::vector<std::vector<int>> C(S.nlevels);
        for (int l = 0; l < S.nlevels; ++l)
            for (int t = S.level_ptr[l]; t < S.level_ptr[l + 1]; ++t) if (WIDE(S.level_sn[t])) C[l].push_back(S.level_sn[t]);
        // the wide front below s when s continues a chain: its only device child, whose row list is s's front
        auto below = [&](int s) -> int {
            if (dchild_ptr[s + 1] - dchild_ptr[s] != 1) return -1;
            const int c = dchild_idx[dchild_ptr[s]];
            if (c >= S.nsn || !WIDE(c) || R(c) != K(s) + R(s)) return -1;
            return c;
        };
        std::vector<ChainSet> sets;
        for (int l = 0; enabled && l < S.nlevels;) {
            if (C[l].empty()) { ++l; continue; }
            int l1 = l;
            while (l1 + 1 < S.nlevels && C[l1 + 1].size() == C[l].size()) {
                bool ok = true;
                std::vector<int> seen;
                for (int s : C[l1 + 1]) {
                    const int c = below(s);
                    if (c < 0 || S.sn_level[c] != l1 || std::find(seen.begin(), seen.end(), c) != seen.end()) { ok = false; break; }
                    seen.push_back(c);
                }
                if (!ok) break;
                ++l1;
            }
            if (l1 - l + 1 >= 4 && l1 - l + 1 <= KW && (int)C[l].size() <= max_chains) {        // (the kernels keep per-link geometry of <= KW links in shared memory)
                ChainSet cs; cs.l0 = l; cs.l1 = l1;
                for (int top : C[l1]) {                    // follow every chain down from its top link
                    std::vector<int> ch(1, top);
                    for (int lv = l1; lv > l; --lv) ch.push_back(below(ch.back()));
                    std::reverse(ch.begin(), ch.end());
                    cs.chains.push_back(ch);
                }
                // CTAs in proportion to the chains' blocks; every CTA may own at most CHAIN_MAXOWN of them
                std::vector<int> nb(cs.chains.size());
                int64_t tot = 0;
                for (size_t q = 0; q < cs.chains.size(); ++q) {
                    const int last = cs.chains[q].back();
                    nb[q] = (int)cs.chains[q].size() + (int)((R(last) + KW - 1) / KW);
                    tot += nb[q];
                }
                bool fits = true;
                int used = 0;
                std::vector<int> nct(cs.chains.size());
                for (size_t q = 0; q < cs.chains.size(); ++q) {
                    nct[q] = std::max<int>(1, (int)((int64_t)CHAIN_CTAS * nb[q] / tot));
                    used += nct[q];
                    if (nb[q] > CHAIN_MAXOWN * nct[q]) fits = false;
                }
                if (fits && used <= CHAIN_CTAS) sets.push_back(cs);
            }
            l = l1 + 1;
        }
        // descriptors + the backward sweep's tasks for the rows above the chains
        struct SetDev { int desc0, nch, nctas; int64_t rect_off; int rect_n; };
        std::vector<SetDev> dev(sets.size());
        for (size_t si = 0; si < sets.size(); ++si) {
            const ChainSet& cs = sets[si];
            std::vector<int> nb(cs.chains.size());
            int64_t tot = 0;
            for (size_t q = 0; q < cs.chains.size(); ++q) { nb[q] = (int)cs.chains[q].size() + (int)((R(cs.chains[q].back()) + KW - 1) / KW); tot += nb[q]; }
            dev[si].desc0 = (int)desc.size() / CHAIN_DESC; dev[si].nch = (int)cs.chains.size();
            dev[si].rect_off = (int64_t)tasks.size();
            int cta0 = 0;
            int64_t part0 = 0;
            for (size_t q = 0; q < cs.chains.size(); ++q) {
                const std::vector<int>& ch = cs.chains[q];
                const int m = (int)ch.size(), last = ch.back();
                const int64_t ra = R(last);
                const int ntile = (int)((ra + BWD_ROWS - 1) / BWD_ROWS);
                const int nct = std::max<int>(1, (int)((int64_t)CHAIN_CTAS * nb[q] / tot));
                const int d8[CHAIN_DESC] = {(int)links.size(), m, (int)boffs.size(), nb[q], cta0, nct, (int)part0, ntile};
                desc.insert(desc.end(), d8, d8 + CHAIN_DESC);
                int o = 0;
                for (int s : ch) { links.push_back(s); boffs.push_back(o); o += K(s); }
                for (int64_t a = 0; a < ra; a += KW) boffs.push_back(o + (int)a);
                boffs.push_back(o + (int)ra);
                for (int j = 0; j < m; ++j)                 // rows above the chain are the last `ra` rows of every link's row list
                    for (int tl = 0; tl < ntile; ++tl) {
                        const int64_t first = R(ch[j]) - ra + (int64_t)tl * BWD_ROWS;
                        tasks.push_back(make_int4(ch[j], (int)first, (int)std::min<int64_t>(BWD_ROWS, ra - (int64_t)tl * BWD_ROWS), (int)(part0 + (int64_t)j * ntile + tl)));
                    }
                part0 += (int64_t)m * ntile;
                cta0 += nct;
            }
            dev[si].nctas = cta0;
            dev[si].rect_n = (int)((int64_t)tasks.size() - dev[si].rect_off);
            h->chain_part_vecs = std::max(h->chain_part_vecs, part0);
        }
        auto in_set = [&](int level) -> int {
            for (size_t si = 0; si < sets.size(); ++si) if (level >= sets[si].l0 && level <= sets[si].l1) return (int)si;
            return -1;
        };
        // forward: ascending levels; the chain kernel of a set goes in front of the first launch at or above its bottom level
        fwd1.clear();
        {
            size_t next = 0;
            auto emit_upto = [&](int level) {
                while (next < sets.size() && sets[next].l0 <= level) {
                    fwd1.push_back(Launch{L_FWD_CHAIN, dev[next].desc0, dev[next].nch, dev[next].nctas, sets[next].l0, 0});
                    ++next;
                }
            };
            for (const Launch& L : fwd) {
                emit_upto(L.level);
                if (L.kind == L_FWD && (L.fmax & 255) > NB && in_set(L.level) >= 0) continue;
                fwd1.push_back(L);
            }
            emit_upto(S.nlevels);
        }
        // backward: descending levels; rows above the chains first (one launch), then the chain kernel
        bwd1.clear();
        {
            int next = (int)sets.size() - 1;
            auto emit_downto = [&](int level) {
                while (next >= 0 && sets[next].l1 >= level) {
                    if (dev[next].rect_n > 0) bwd1.push_back(Launch{L_BWD_RECT, dev[next].rect_off, dev[next].rect_n, 0, sets[next].l1, 0});
                    bwd1.push_back(Launch{L_BWD_CHAIN, dev[next].desc0, dev[next].nch, dev[next].nctas, sets[next].l1, 0});
                    --next;
                }
            };
            for (const Launch& L : bwd) {
                emit_downto(L.level);
                if (L.kind == L_BWD && L.fmax > NB && in_set(L.level) >= 0) continue;
                bwd1.push_back(L);
            }
            emit_downto(-1);
        }
    }
}

// Map every peer's factor pool, contribution pool and flag array into this process (CUDA IPC over NVLink): the handles
// travel through one ncclAllGather.  Offsets into the pools are the same on every rank (same deterministic analysis).
int map_peers(smslu_handle_t h) {
    if (h->nranks > MAX_RANKS) return fail(h, SMSLU_E_ARG, "at most 8 ranks (one NVSwitch box)");
    struct Handles { cudaIpcMemHandle_t lu, cb, fl; };
    Handles mine;
    CU(cudaIpcGetMemHandle(&mine.lu, h->cx.lu));
    CU(cudaIpcGetMemHandle(&mine.cb, h->cx.cb));
    CU(cudaIpcGetMemHandle(&mine.fl, h->d_xflags));
    char* d_all = nullptr;
    int rc;
    if ((rc = dev_alloc(h, &d_all, sizeof(Handles) * (size_t)h->nranks))) return rc;
    CU(cudaMemcpy(d_all + sizeof(Handles) * h->rank, &mine, sizeof(Handles), cudaMemcpyHostToDevice));
    NCCLCHK(nccl_api().AllGather(d_all + sizeof(Handles) * h->rank, d_all, sizeof(Handles), ncclChar, h->comm, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    std::vector<Handles> all(h->nranks);
    CU(cudaMemcpy(all.data(), d_all, sizeof(Handles) * (size_t)h->nranks, cudaMemcpyDeviceToHost));
    for (int g = 0; g < h->nranks; ++g) {
        if (g == h->rank) continue;
        void *plu = nullptr, *pcb = nullptr, *pfl = nullptr;
        CU(cudaIpcOpenMemHandle(&plu, all[g].lu, cudaIpcMemLazyEnablePeerAccess));
        CU(cudaIpcOpenMemHandle(&pcb, all[g].cb, cudaIpcMemLazyEnablePeerAccess));
        CU(cudaIpcOpenMemHandle(&pfl, all[g].fl, cudaIpcMemLazyEnablePeerAccess));
        h->ipc_opened.push_back(plu); h->ipc_opened.push_back(pcb); h->ipc_opened.push_back(pfl);
        h->cx.lu_peer[g] = (double*)plu; h->cx.cb_peer[g] = (double*)pcb; h->cx.xflag_peer[g] = (int*)pfl;
    }
    return 0;
}

// Factorization schedule of the distributed top (nranks > 1).  Per level l two groups of launches (run_schedule forks /
// joins the lanes of a group):
//   group 2l  : zero-fill / assembly (destination columns this rank owns); panel owner's pass over the fronts it did not
//               already factor ahead (k_panel mode 1: pivot block and L21), published to every peer's pool, signal;
//               wait for the other owners of the level; rows of U12' this rank owns (k_panel mode 2); the Schur-update
//               tiles that feed the pivot columns of a parent this rank owns (look-ahead part)
//   group 2l+1: lane 0: the rest of the Schur update (columns this rank owns);
//               lane 1 (high-priority stream): LOOK-AHEAD -- the panels of the next level's fronts this rank owns whose
//               only input is that look-ahead part (links of a chain), published and signalled while every rank is
//               still inside the current Schur update.
void build_top_fac(smslu_handle_t h, std::vector<int4>& tasks, std::vector<int64_t>& segs, std::vector<int>& wait_slots,
                   std::vector<int4>& trow_tasks, int64_t& ncounters) {
    const Symbolic& S = h->S;
    const int rank = h->rank, NR = h->nranks;
    auto K = [&](int s) { return S.sn_start[s + 1] - S.sn_start[s]; };
    auto R = [&](int s) { return (int64_t)(S.rows_ptr[s + 1] - S.rows_ptr[s]); };
    auto NC = [&](int s) { return S.child_ptr[s + 1] - S.child_ptr[s]; };
    auto ROWOWN = [&](int s, int64_t a) { return S.col_owner[S.rows[S.rows_ptr[s] + a]]; };
    // owner of destination column pb of front s (pivot columns: the panel owner)
    auto DSTOWN = [&](int s, int64_t pb) { return pb < K(s) ? S.top_owner[s] : ROWOWN(s, pb - K(s)); };
    std::vector<Launch>& fac = h->fac_top;
    fac.clear();
    std::vector<int> slot_of(S.nsn, -1);          // sync slot of a top front = its ordinal among the top fronts
    { int t = 0; for (int s = 0; s < S.nsn; ++s) if (S.owner[s] == -1) slot_of[s] = t++; }
    std::vector<char> ahead(S.nsn, 0);            // panel already factored and published by the look-ahead lane
    const bool lookahead = !(getenv("SMSLU_NO_LOOKAHEAD") && atoi(getenv("SMSLU_NO_LOOKAHEAD")) != 0);
    int64_t peer_doubles = 0;
    auto push = [&](int kind, int64_t off, int nt, int fmax, int group, int lane) {
        if (nt > 0) fac.push_back(Launch{kind, off, nt, fmax, group, lane});
    };
    // the panel owner's pass (pass 1: tiles of kinds 0 and 2 of the fronts in `list` this rank owns) or every rank's pass
    // over the rows of U12' it owns (pass 2), one launch per 32 pivot columns; then, for pass 1, publish + signal
    auto panel_pass = [&](const std::vector<int>& list, int pass, int group, int lane) {
        int max_blk = 0;
        for (int s : list) max_blk = std::max(max_blk, (K(s) + NB - 1) / NB);
        for (int g = 0; g < max_blk; ++g) {
            auto ntiles = [&](int s, int rws) -> int {
                const int k = K(s);
                const int64_t r = R(s), f = k + r;
                const int j1 = std::min(k, (g + 1) * NB);
                if (pass == 1) return (int)((f - j1 + rws - 1) / rws) + (k - j1 + rws - 1) / rws;
                return (int)((r + rws - 1) / rws);
            };
            int64_t tiles_total = 0;
            for (int s : list) if (g < (K(s) + NB - 1) / NB) tiles_total += std::max(1, ntiles(s, PANEL_ROWS));
            const int rows = (tiles_total > 0 && tiles_total < 120) ? PANEL_ROWS_TOP : PANEL_ROWS;
            const int64_t off = (int64_t)tasks.size();
            for (int s : list) {
                if (g >= (K(s) + NB - 1) / NB) continue;
                if (pass == 1) {
                    const int nt = std::max(1, ntiles(s, rows));         // someone has to factor D_gg
                    const int cidx = (int)ncounters++;
                    for (int t = 0; t < nt; ++t) tasks.push_back(make_int4(s, (int)(g | (1 << 4) | (nt << 16) | (1u << 30)), t, cidx));
                } else {
                    const int nt = ntiles(s, rows);
                    for (int t = 0; t < nt; ++t) {
                        bool mine = false;
                        for (int64_t a = (int64_t)t * rows; a < std::min<int64_t>(R(s), (int64_t)(t + 1) * rows) && !mine; ++a) mine = ROWOWN(s, a) == rank;
                        if (mine) tasks.push_back(make_int4(s, (int)(g | (1 << 4) | (1 << 16) | (2u << 30)), t, 0));
                    }
                }
            }
            push(L_PANEL, off, (int)((int64_t)tasks.size() - off), g | (rows << 8), group, lane);
        }
        if (pass == 1 && !list.empty()) {
            const int64_t s0 = (int64_t)segs.size() / 2;
            for (int s : list) {
                const int64_t len = ((K(s) + R(s)) * K(s) + 1) & ~(int64_t)1;
                segs.push_back(S.Loff[s]); segs.push_back(len);
                peer_doubles += len * (NR - 1);
            }
            push(L_REPL, s0, (int)list.size(), 0, group, lane);
            for (int s : list) fac.push_back(Launch{L_SIGNAL, (int64_t)slot_of[s], 1, 0, group, lane});
        }
    };
    for (int l = 0; l < S.nlevels; ++l) {
        std::vector<int> top;
        for (int t = S.level_ptr[l]; t < S.level_ptr[l + 1]; ++t) if (S.owner[S.level_sn[t]] == -1) top.push_back(S.level_sn[t]);
        if (top.empty()) continue;
        const int gA = 2 * l, gB = 2 * l + 1;
        // zero-fill of the parents that direct children of this level add into (whole blocks: local memory)
        int64_t off = (int64_t)tasks.size();
        for (int s : top)
            if (S.direct[s] && !S.cb_assigned[S.sn_parent[s]] && R(S.sn_parent[s]) > 0) {
                const int ps = S.sn_parent[s];
                const int64_t tiles = (R(ps) * R(ps) + ZERO_TILE - 1) / ZERO_TILE;
                for (int64_t i = 0; i < tiles; ++i) tasks.push_back(make_int4(ps, (int)i, 0, 0));
            }
        push(L_ZERO, off, (int)((int64_t)tasks.size() - off), 0, gA, 0);
        // assembly: every non-direct child (top fronts and the subtree roots, whose columns arrived in this rank's
        // exchange slots), destination columns this rank owns, tasks cut where the owner changes
        off = (int64_t)tasks.size();
        int64_t asm_fmax = 0;
        for (int s : top) {
            if (NC(s) == 0) continue;
            std::vector<int> kids;
            for (int u = S.child_ptr[s]; u < S.child_ptr[s + 1]; ++u) {
                const int c = S.child_idx[u];
                if (!S.direct[c] && R(c) > 0) kids.push_back(c);
            }
            if (kids.empty()) continue;
            const bool zero = R(s) > 0;
            const int64_t f = K(s) + R(s);
            for (int64_t pb0 = 0; pb0 < f;) {
                const int own = DSTOWN(s, pb0);
                int64_t pb1 = pb0 + 1;
                while (pb1 < f && pb1 < pb0 + ASM_COLS && DSTOWN(s, pb1) == own) ++pb1;
                if (own == rank) {
                    const int64_t moff = (int64_t)h->asm_meta.size();
                    int np = 0;
                    for (int c : kids) {
                        const int* rb = S.rel.data() + S.rows_ptr[c];
                        const int* re = S.rel.data() + S.rows_ptr[c + 1];
                        const int* it = std::lower_bound(rb, re, (int)pb0);
                        if (it == re || *it >= pb1) continue;
                        h->asm_meta.push_back(c); h->asm_meta.push_back((int)(it - rb));
                        ++np;
                    }
                    if (np > 0 || zero) {
                        for (int ch = 0; ch * (int64_t)asm_chunk_rows(f) < f; ++ch)
                            tasks.push_back(make_int4(s, (int)pb0 | (ch << 20), (int)moff, (zero ? 1 : 0) | ((int)(pb1 - pb0) << 4) | (np << 8)));
                        asm_fmax = std::max<int64_t>(asm_fmax, f);
                    }
                }
                pb0 = pb1;
            }
        }
        push(L_EXTEND, off, (int)((int64_t)tasks.size() - off), (int)std::min<int64_t>(asm_fmax, 1 << 30), gA, 0);
        // panels of the fronts this rank owns and has not factored ahead; then wait for everybody else's
        std::vector<int> mine_now;
        const int64_t w0 = (int64_t)wait_slots.size();
        for (int s : top) {
            if (S.top_owner[s] == rank) { if (!ahead[s]) mine_now.push_back(s); }
            else wait_slots.push_back(slot_of[s]);
        }
        panel_pass(mine_now, 1, gA, 0);
        push(L_WAIT, w0, (int)((int64_t)wait_slots.size() - w0), 0, gA, 0);
        panel_pass(top, 2, gA, 0);
        // Schur update of the columns this rank owns: first the tile columns that feed the pivot columns of a parent this
        // rank owns (then that parent's panel runs ahead on lane 1), the rest beside it on lane 0
        std::vector<int> next_mine;
        for (int part = 0; part < 2; ++part) {
            off = (int64_t)tasks.size();
            for (int s : top) {
                const int64_t r = R(s);
                const int nt = (int)((r + GEMM_TILE - 1) / GEMM_TILE);
                const int ps = S.sn_parent[s];
                int64_t nb = 0;                          // columns [0, nb) feed the pivot columns of a parent this rank owns
                if (lookahead && S.direct[s] && ps != -1 && S.owner[ps] == -1 && S.top_owner[ps] == rank) {
                    while (nb < r && S.rel[S.rows_ptr[s] + nb] < K(ps)) ++nb;
                    if (part == 0) { next_mine.push_back(ps); ahead[ps] = 1; }
                }
                const int flags = (NC(s) > 0 ? 1 : 0) | (S.direct[s] ? 2 : 0) | (S.direct[s] && S.cb_assigned[ps] ? 4 : 0) | 16 |
                                  (S.direct[s] && r == K(ps) + R(ps) ? 32 : 0) | 64;
                // column tiles restart at every change of owner (and at nb): runs of columns this rank owns, cut into <= 64
                for (int64_t b0 = part == 0 ? 0 : nb; b0 < (part == 0 ? nb : r);) {
                    const int own = ROWOWN(s, b0);
                    const int64_t lim = part == 0 ? nb : r;
                    int64_t b1 = b0 + 1;
                    while (b1 < lim && ROWOWN(s, b1) == own) ++b1;
                    if (own == rank)
                        for (int64_t c0 = b0; c0 < b1; c0 += GEMM_TILE) {
                            const int w = (int)std::min<int64_t>(GEMM_TILE, b1 - c0);
                            for (int i = 0; i < nt; ++i) tasks.push_back(make_int4(s, i, (int)c0, flags | (w << 8)));
                        }
                    b0 = b1;
                }
            }
            push(L_GEMM, off, (int)((int64_t)tasks.size() - off), 0, part == 0 ? gA : gB, 0);
        }
        panel_pass(next_mine, 1, gB, 1);
        // rows of U12' this rank owns, published to the peers at the end of the refactorization (the solves read all of T_s)
        for (int s : top) {
            const int64_t r = R(s);
            for (int64_t a = 0; a < r;) {
                int64_t b = a + 1;
                const int own = ROWOWN(s, a);
                while (b < r && b < a + 256 && ROWOWN(s, b) == own) ++b;
                if (own == rank) { trow_tasks.push_back(make_int4(s, (int)a, (int)(b - a), 0)); peer_doubles += (b - a) * K(s) * (NR - 1); }
                a = b;
            }
        }
    }
    h->peer_bytes_refactor = 8 * peer_doubles;
}

// Everything the handle owns on the device: streams, events, allocations, peer mappings.  Used by smslu_destroy and when an
// upload fails half way (so that a retry starts from a clean handle instead of leaking the first attempt's objects).
void release_device_state(smslu_handle_t h) {
    if (!(h->uploaded || h->stream || !h->dev_allocs.empty())) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (void* p : h->ipc_opened) cudaIpcCloseMemHandle(p);
    h->ipc_opened.clear();
    for (void* p : h->dev_allocs) cudaFree(p);
    h->dev_allocs.clear();
    cudaEvent_t* evs[] = {&h->ev0, &h->ev1, &h->ev2, &h->ev3, &h->ev_fork, &h->ev_scatter};
    for (cudaEvent_t* e : evs) if (*e) { cudaEventDestroy(*e); *e = nullptr; }
    for (cudaEvent_t e : h->pev) cudaEventDestroy(e);
    h->pev.clear(); h->pev_kind.clear(); h->pev_used = 0;
    for (auto& m : h->lvl_marks) cudaEventDestroy(m.ev);
    h->lvl_marks.clear();
    if (h->h_flag) { cudaFreeHost(h->h_flag); h->h_flag = nullptr; }
    if (h->stream && h->own_stream) cudaStreamDestroy(h->stream);
    h->stream = nullptr;
    for (int a = 0; a < NLANES - 1; ++a) {
        if (h->aux_stream[a]) { cudaStreamSynchronize(h->aux_stream[a]); cudaStreamDestroy(h->aux_stream[a]); h->aux_stream[a] = nullptr; }
        if (h->ev_join[a]) { cudaEventDestroy(h->ev_join[a]); h->ev_join[a] = nullptr; }
    }
    h->uploaded = false; h->factored = false; h->have_wide = false; h->wide_failed = false; h->scatter_pending = false;
    h->cx = DevCtx{}; h->cx32 = DevCtx{};
    cudaGetLastError();
}

int upload_impl(smslu_handle_t h);
int ensure_uploaded(smslu_handle_t h) {
    if (h->uploaded) { CU(cudaSetDevice(h->device)); return 0; }
    const int rc = upload_impl(h);
    if (rc) { const std::string msg = h->err; release_device_state(h); h->err = msg; }
    return rc;
}

int upload_impl(smslu_handle_t h) {
    if (!h->analyzed) return fail(h, SMSLU_E_ARG, "smslu_analyze has not been called");
    if (h->nranks > 1 && !h->comm) return fail(h, SMSLU_E_ARG, "partitioned handle: call smslu_comm_init on every rank first");
    double t0 = now_ms();
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(h, SMSLU_E_CUDA, std::string("no usable CUDA device (no CPU fallback exists): ") + cudaGetErrorString(e));
    if (h->opt.device >= 0) h->device = h->opt.device;
    else CU(cudaGetDevice(&h->device));
    CU(cudaSetDevice(h->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, h->device));
    if (prop.major < 10) return fail(h, SMSLU_E_CUDA, "device is not sm_100 class; this library is built for sm_100a only");
    CU(kernels_init());
    if (h->have_user_stream) { h->stream = h->user_stream; h->own_stream = false; }
    else CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    const bool lanes = !(getenv("SMSLU_NO_LANES") && atoi(getenv("SMSLU_NO_LANES")) != 0);   // debugging aid: one stream only
    for (int a = 0; a < NLANES - 1; ++a) {
        if (lanes) {       // lane 1 carries the look-ahead panels of the distributed top: ahead of the queued Schur-update CTAs
            int plo = 0, phi = 0;
            CU(cudaDeviceGetStreamPriorityRange(&plo, &phi));
            CU(cudaStreamCreateWithPriority(&h->aux_stream[a], cudaStreamNonBlocking, a == 0 ? phi : plo));
        }
        CU(cudaEventCreateWithFlags(&h->ev_join[a], cudaEventDisableTiming));
    }
    CU(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_scatter, cudaEventDisableTiming));
    CU(cudaMallocHost((void**)&h->h_flag, 2 * sizeof(int)));
    CU(cudaEventCreate(&h->ev0)); CU(cudaEventCreate(&h->ev1));
    CU(cudaEventCreate(&h->ev2)); CU(cudaEventCreate(&h->ev3));
    const Symbolic& S = h->S;
    const int n = S.n;
    int rc;
    int *d_sn_start, *d_rows, *d_rel, *d_sn_parent, *d_child_ptr, *d_child_idx;
    int64_t *d_rows_ptr, *d_Loff, *d_Uoff, *d_CBoff;
    // Device copies of the tree.  With a partition every interface front gets one extra "virtual
    // child" whose update vector is the all-reduced sum of the forward-solve contributions of the
    // subtrees below the cut (rel = identity over the parent's front); it replaces those subtree
    // roots in the parent's child list, so the solve kernels need no special case.
    std::vector<int64_t> rows_ptr_d(S.rows_ptr);
    std::vector<int> rows_d(S.rows), rel_d(S.rel), child_ptr_d(S.child_ptr), child_idx_d(S.child_idx);
    h->vupd_off = S.sum_r; h->vupd_len = 0; h->nvtasks = 0;
    if (S.nranks > 1) {
        std::vector<int4> vtasks;
        std::vector<int> vlist;
        child_ptr_d.assign(S.nsn + 1, 0);
        child_idx_d.clear();
        int nv = 0;
        for (int s2 = 0; s2 < S.nsn; ++s2) {
            if (S.iface[s2]) {
                const int v = S.nsn + nv++;
                const int64_t f = (S.sn_start[s2 + 1] - S.sn_start[s2]) + (S.rows_ptr[s2 + 1] - S.rows_ptr[s2]);
                for (int64_t i = 0; i < f; ++i) { rows_d.push_back(0); rel_d.push_back((int)i); }
                rows_ptr_d.push_back((int64_t)rows_d.size());
                child_idx_d.push_back(v);
                const int l0 = (int)vlist.size();
                for (int u = S.child_ptr[s2]; u < S.child_ptr[s2 + 1]; ++u)
                    if (S.owner[S.child_idx[u]] == h->rank) vlist.push_back(S.child_idx[u]);
                vtasks.push_back(make_int4(s2, v, l0, (int)vlist.size()));
            }
            for (int u = S.child_ptr[s2]; u < S.child_ptr[s2 + 1]; ++u)
                if (!S.iface[s2] || S.owner[S.child_idx[u]] == -1) child_idx_d.push_back(S.child_idx[u]);
            child_ptr_d[s2 + 1] = (int)child_idx_d.size();
        }
        h->vupd_len = (int64_t)rows_d.size() - S.sum_r;
        h->nvtasks = (int)vtasks.size();
        if ((rc = dev_upload(h, &h->d_vtasks, vtasks))) return rc;
        if ((rc = dev_upload(h, &h->d_vlist, vlist))) return rc;
        std::vector<int> colowner(n);
        for (int j = 0; j < n; ++j) colowner[j] = S.owner[S.col2sn[j]];
        if ((rc = dev_upload(h, &h->d_colowner, colowner))) return rc;
    }
    signed char* d_rowown;
    {   // owner of the global column behind every entry of `rows` (the virtual children's rows: none)
        std::vector<signed char> rowown(rows_d.size(), (signed char)-1);
        for (int64_t t = 0; t < S.sum_r; ++t) rowown[t] = (signed char)S.col_owner[S.rows[t]];
        if ((rc = dev_upload(h, &d_rowown, rowown))) return rc;
    }
    if ((rc = dev_upload(h, &d_sn_start, S.sn_start))) return rc;
    if ((rc = dev_upload(h, &d_rows_ptr, rows_ptr_d))) return rc;
    if ((rc = dev_upload(h, &d_rows, rows_d))) return rc;
    if ((rc = dev_upload(h, &d_rel, rel_d))) return rc;
    if ((rc = dev_upload(h, &d_Loff, S.Loff))) return rc;
    if ((rc = dev_upload(h, &d_Uoff, S.Uoff))) return rc;
    if ((rc = dev_upload(h, &d_CBoff, S.CBoff))) return rc;
    if ((rc = dev_upload(h, &d_sn_parent, S.sn_parent))) return rc;
    if ((rc = dev_upload(h, &d_child_ptr, child_ptr_d))) return rc;
    if ((rc = dev_upload(h, &d_child_idx, child_idx_d))) return rc;
    if ((rc = dev_upload(h, &h->d_p, S.p))) return rc;
    if ((rc = dev_upload(h, &h->d_q, S.q))) return rc;
    if (!S.post.empty() && (rc = dev_upload(h, &h->d_post, S.post))) return rc;
    {   // row index of every nonzero (for the scaling) and a row-major view of the pattern
        std::vector<int64_t> rowptr(n + 1, 0), rowidx(h->annz);
        for (int64_t t = 0; t < h->annz; ++t) ++rowptr[h->Ai[t] + 1];
        for (int i = 0; i < n; ++i) rowptr[i + 1] += rowptr[i];
        std::vector<int64_t> w(rowptr.begin(), rowptr.end() - 1);
        for (int c = 0; c < n; ++c)
            for (int64_t t = h->Ap[c]; t < h->Ap[c + 1]; ++t) rowidx[w[h->Ai[t]]++] = t;
        if ((rc = dev_upload(h, &h->d_rowptr, rowptr))) return rc;
        if ((rc = dev_upload(h, &h->d_rowidx, rowidx))) return rc;
    }
    int *d_a_ptr, *d_a_src, *d_a_row, *d_a_pos;
    {   // entries of A grouped by the small front that pulls them; the big fronts' entries follow
        const int nsn = S.nsn;
        if (h->annz > INT_MAX) return fail(h, SMSLU_E_ARG, "more than 2^31-1 nonzeros");
        // this rank assembles the entries of its own fronts and, in the top fronts, the entries of the columns it owns
        // (pivot columns of a top front: its panel owner; an entry of U12 in global column j: the owner of column j)
        auto small = [&](int s) { return S.small[s] != 0 && S.owner[s] == h->rank; };
        std::vector<int> cinv;
        if (S.nranks > 1) { cinv.resize(n); for (int k2 = 0; k2 < n; ++k2) cinv[S.q[k2]] = k2; }
        std::vector<char> top_mine;               // per nonzero, only filled for entries of top fronts
        if (S.nranks > 1) {
            top_mine.assign(h->annz, 0);
            for (int c = 0; c < n; ++c)
                for (int64_t t = h->Ap[c]; t < h->Ap[c + 1]; ++t) {
                    const int s = S.a_sn[t];
                    if (S.owner[s] != -1) continue;
                    const int jj = cinv[c];
                    top_mine[t] = (jj < S.sn_start[s + 1] ? S.top_owner[s] : S.col_owner[jj]) == h->rank;
                }
        }
        auto mine = [&](int s, int64_t t) { return S.owner[s] == h->rank || (S.owner[s] == -1 && top_mine[t]); };
        std::vector<int> a_ptr(nsn + 1, 0);
        for (int64_t t = 0; t < h->annz; ++t) if (small(S.a_sn[t])) ++a_ptr[S.a_sn[t] + 1];
        for (int s = 0; s < nsn; ++s) a_ptr[s + 1] += a_ptr[s];
        const int nsmall = a_ptr[nsn];
        std::vector<int> a_src(h->annz), a_row(h->annz), a_pos(nsmall), w(a_ptr.begin(), a_ptr.end() - 1);
        std::vector<int64_t> dst_big(h->annz - nsmall);     // upper bound; nb entries are used
        int64_t nb = 0;
        for (int64_t t = 0; t < h->annz; ++t) {
            const int s = S.a_sn[t];
            int64_t o;
            if (small(s)) { o = w[s]++; a_pos[o] = S.a_loc[t]; }
            else if (mine(s, t)) { o = nsmall + nb; dst_big[nb++] = S.a_dst[t]; }
            else continue;
            a_src[o] = (int)t;
            a_row[o] = (int)h->Ai[t];
        }
        if ((rc = dev_upload(h, &d_a_ptr, a_ptr))) return rc;
        if ((rc = dev_upload(h, &d_a_src, a_src))) return rc;
        if ((rc = dev_upload(h, &d_a_row, a_row))) return rc;
        if ((rc = dev_upload(h, &d_a_pos, a_pos))) return rc;
        if ((rc = dev_upload(h, &h->d_big_dst, dst_big))) return rc;
        h->d_big_src = d_a_src + nsmall;
        h->d_big_row = d_a_row + nsmall;
        h->nnz_big = nb;
    }
    double *d_lu, *d_cb, *d_upd, *d_dinv;
    int *d_counters, *d_flag;
    if ((rc = dev_alloc(h, &d_lu, (size_t)S.lu_size))) return rc;
    if ((rc = dev_alloc(h, &d_cb, (size_t)S.cb_size))) return rc;
    if ((rc = dev_alloc(h, &d_upd, (size_t)(S.sum_r + h->vupd_len) * RB_MAX))) return rc;
    if ((rc = dev_alloc(h, &d_flag, 2))) return rc;
    if ((rc = dev_alloc(h, &d_dinv, (size_t)n))) return rc;
    if ((rc = dev_alloc(h, &h->d_Rs, (size_t)n))) return rc;
    if ((rc = dev_alloc(h, &h->d_aval, (size_t)h->annz))) return rc;
    if ((rc = dev_alloc(h, &h->d_w, (size_t)n * RB_MAX))) return rc;
    if ((rc = dev_alloc(h, &h->d_z, (size_t)n * RB_MAX))) return rc;
    if ((rc = dev_alloc(h, &h->d_xb, (size_t)n * RB_MAX))) return rc;
    {
        std::vector<double> ones(n, 1.0);
        CU(cudaMemcpy(h->d_Rs, ones.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
    }
    std::vector<int4> tasks;
    build_schedules(h, tasks);
    if (S.nranks > 1) {
        std::vector<int64_t> segs;
        std::vector<int> wait_slots;
        std::vector<int4> trow;
        build_top_fac(h, tasks, segs, wait_slots, trow, h->ncounters);
        h->n_trow_tasks = (int)trow.size();
        if ((rc = dev_upload(h, &h->d_segs, segs))) return rc;
        if ((rc = dev_upload(h, &h->d_wait_slots, wait_slots))) return rc;
        if ((rc = dev_upload(h, &h->d_trow_tasks, trow))) return rc;
        if ((rc = dev_alloc(h, &h->d_xflags, (size_t)S.nsn + 1))) return rc;
        CU(cudaMemset(h->d_xflags, 0, sizeof(int) * ((size_t)S.nsn + 1)));
        if ((rc = dev_alloc(h, &h->d_barrier, 1))) return rc;
        CU(cudaMemset(h->d_barrier, 0, sizeof(int)));
        h->st.allreduce_doubles_refactor = h->peer_bytes_refactor / 8;
    }
    int *d_chain_links, *d_chain_boff, *d_chain_flags;
    double* d_chain_part;
    {
        std::vector<int> cdesc, clinks, cboff;
        build_chain_schedules(h, tasks, child_ptr_d, child_idx_d, cdesc, clinks, cboff);
        if ((rc = dev_upload(h, &h->d_chain_desc, cdesc))) return rc;
        if ((rc = dev_upload(h, &d_chain_links, clinks))) return rc;
        if ((rc = dev_upload(h, &d_chain_boff, cboff))) return rc;
        if ((rc = dev_alloc(h, &d_chain_flags, (size_t)2 * S.nsn))) return rc;
        CU(cudaMemset(d_chain_flags, 0, sizeof(int) * 2 * (size_t)std::max(S.nsn, 1)));
        if ((rc = dev_alloc(h, &d_chain_part, (size_t)h->chain_part_vecs * KW))) return rc;
    }
    if ((rc = dev_upload(h, &h->d_tasks, tasks))) return rc;
    int* d_asm_meta;
    if ((rc = dev_upload(h, &d_asm_meta, h->asm_meta))) return rc;
    h->cx.asm_meta = d_asm_meta;
    int* d_Doff; double* d_dblk;
    {   // inverted diagonal blocks of the big fronts this rank factors (its own and the replicated top)
        std::vector<int> Doff(S.nsn, 0);
        std::vector<int2> inv_tasks;
        for (int s2 = 0; s2 < S.nsn; ++s2) {
            if (S.small[s2] || !(S.owner[s2] == h->rank || S.owner[s2] == -1)) continue;
            Doff[s2] = (int)inv_tasks.size();
            const int k2 = S.sn_start[s2 + 1] - S.sn_start[s2];
            for (int g = 0; g * NB < k2; ++g) inv_tasks.push_back(make_int2(s2, g));
        }
        h->n_inv_tasks = (int)inv_tasks.size();
        if ((rc = dev_upload(h, &d_Doff, Doff))) return rc;
        if ((rc = dev_upload(h, &h->d_inv_tasks, inv_tasks))) return rc;
        if ((rc = dev_alloc(h, &d_dblk, (size_t)std::max<size_t>(inv_tasks.size(), 1) * NB * NB))) return rc;
    }
    double* d_bpart; int* d_counters2;
    if ((rc = dev_alloc(h, &d_counters, (size_t)h->ncounters))) return rc;
    if ((rc = dev_alloc(h, &d_bpart, (size_t)h->bpart_slots * KMAX * RB_MAX))) return rc;
    if ((rc = dev_alloc(h, &d_counters2, (size_t)S.nsn))) return rc;
    CU(cudaMemset(d_counters2, 0, sizeof(int) * std::max(S.nsn, 1)));
    DevCtx& cx = h->cx;
    cx.sn_start = d_sn_start; cx.rows_ptr = d_rows_ptr; cx.rows = d_rows; cx.rel = d_rel;
    cx.Loff = d_Loff; cx.Uoff = d_Uoff; cx.CBoff = d_CBoff; cx.sn_parent = d_sn_parent;
    cx.child_ptr = d_child_ptr; cx.child_idx = d_child_idx;
    cx.lu = d_lu; cx.cb = d_cb; cx.upd = d_upd; cx.counters = d_counters; cx.flag = d_flag;
    cx.bpart = d_bpart; cx.counters2 = d_counters2; cx.dinv = d_dinv;
    cx.Doff = d_Doff; cx.dblk = d_dblk;
    {
        const double tol = h->opt.pivot_tol == 0.0 ? 1.0e-3 : h->opt.pivot_tol;
        cx.lmax = tol > 0.0 ? (1.0 + 1.0e-6) / tol : HUGE_VAL;   // a hair above 1/tol: pivots chosen by a host threshold search pass
    }
    cx.a_ptr = d_a_ptr; cx.a_src = d_a_src; cx.a_row = d_a_row; cx.a_pos = d_a_pos;
    cx.chain_links = d_chain_links; cx.chain_boff = d_chain_boff; cx.chain_flags = d_chain_flags; cx.chain_nsn = S.nsn;
    cx.chain_part = d_chain_part;
    cx.rowown = d_rowown; cx.rank = h->rank; cx.nranks = h->nranks;
    for (int g = 0; g < MAX_RANKS; ++g) { cx.cb_peer[g] = nullptr; cx.lu_peer[g] = nullptr; cx.xflag_peer[g] = nullptr; }
    cx.cb_peer[h->rank] = d_cb; cx.lu_peer[h->rank] = d_lu; cx.xflag_peer[h->rank] = h->d_xflags;
    CU(cudaDeviceSynchronize());
    if (S.nranks > 1 && (rc = map_peers(h))) return rc;
    h->uploaded = true;
    h->st.ms_upload = now_ms() - t0;
    return 0;
}

// ---- optional per-launch timing (smslu_set_profile): events around every launch, summed by kind
int prof_begin(smslu_handle_t h, int kind) {
    if (!h->profile) return 0;
    if (h->pev_used + 2 > h->pev.size()) {
        for (int i = 0; i < 256; ++i) { cudaEvent_t e; CU(cudaEventCreate(&e)); h->pev.push_back(e); }
    }
    h->pev_kind.push_back(kind);
    CU(cudaEventRecord(h->pev[h->pev_used++], h->stream));
    return 0;
}
int prof_end(smslu_handle_t h) {
    if (!h->profile) return 0;
    CU(cudaEventRecord(h->pev[h->pev_used++], h->stream));
    return 0;
}
int prof_collect(smslu_handle_t h) {   // stream must be synchronized
    if (!h->lvl_marks.empty()) {
        for (size_t i = 0; i + 1 < h->lvl_marks.size(); ++i) {
            const auto& m = h->lvl_marks[i];
            if (!m.sched) continue;
            float ms = 0;
            cudaEventElapsedTime(&ms, m.ev, h->lvl_marks[i + 1].ev);
            const char* nm = m.sched == &h->fac ? "fac" : m.sched == &h->fwd ? "fwd" : m.sched == &h->bwd ? "bwd" : "top";
            fprintf(stderr, "[smslu level] %s level %3d launches %2d tasks %7d  %8.1f us\n", nm, m.level, m.nlaunch, m.ntasks, 1e3 * ms);
        }
        for (auto& m : h->lvl_marks) cudaEventDestroy(m.ev);
        h->lvl_marks.clear();
    }
    if (!h->profile) return 0;
    for (size_t i = 0; i + 1 < h->pev_used; i += 2) {
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, h->pev[i], h->pev[i + 1]));
        int k = h->pev_kind[i / 2];
        h->st.ms_kernel[k] += ms;
        h->st.launches_kernel[k] += 1;
    }
    h->pev_used = 0;
    h->pev_kind.clear();
    return 0;
}

int launch_one(smslu_handle_t h, cudaStream_t st, const Launch& L, const double* win, double* zx, int rb) {
    const int4* tk = h->d_tasks + ((L.kind < L_REPL || L.kind >= L_FWD32) ? L.off : 0);
    const DevCtx& cxv = rb == RB_WIDE ? h->cx32 : h->cx;      // the wide sweeps have their own update vectors / scratch
    switch (L.kind) {
        case L_ZERO: launch_zero_cb(st, h->cx, tk, L.ntasks); break;
        case L_EXTEND: launch_assemble(st, h->cx, tk, L.ntasks, L.fmax); break;
        case L_SMALL: launch_front_small(st, h->cx, tk, L.ntasks, L.fmax, h->cur_av, h->d_Rs); break;
        case L_FWD_SMALL: launch_small_fwd(st, cxv, tk, L.ntasks, win, zx, rb); break;
        case L_BWD_SMALL: launch_small_bwd(st, cxv, tk, L.ntasks, zx, rb); break;
        case L_FWD32: launch_fwd32(st, cxv, tk, L.ntasks, L.fmax, win, zx); break;
        case L_BWD32: launch_bwd32(st, cxv, tk, L.ntasks, zx); break;
        case L_PANEL: launch_panel(st, h->cx, tk, L.ntasks, L.fmax & 255, L.fmax >> 8); break;
        case L_GEMM: if (L.fmax == 1) launch_gemm_strip(st, h->cx, tk, L.ntasks); else launch_gemm_cb(st, h->cx, tk, L.ntasks); break;
        case L_FWD: launch_fwd(st, h->cx, tk, L.ntasks, L.fmax >> 8, win, zx, rb); break;
        case L_BWD: launch_bwd(st, h->cx, tk, L.ntasks, zx, rb); break;
        case L_REPL: launch_replicate(st, h->cx, h->d_segs + 2 * L.off, L.ntasks); break;
        case L_SIGNAL: launch_signal(st, h->cx, (int)L.off, h->epoch); break;
        case L_WAIT: launch_wait(st, h->cx, h->d_wait_slots + L.off, L.ntasks, h->epoch); break;
        case L_FWD_CHAIN:
            if (launch_fwd_chain(st, h->cx, h->d_chain_desc + L.off * CHAIN_DESC, L.ntasks, L.fmax, win, zx, ++h->solve_epoch) != cudaSuccess) return -1;
            break;
        case L_BWD_RECT: launch_bwd_rect(st, h->cx, h->d_tasks + L.off, L.ntasks, zx); break;
        case L_BWD_CHAIN:
            if (launch_bwd_chain(st, h->cx, h->d_chain_desc + L.off * CHAIN_DESC, L.ntasks, L.fmax, zx, ++h->solve_epoch) != cudaSuccess) return -1;
            break;
    }
    return 0;
}

// Inside one level the small-front kernels and the big-front kernels touch disjoint fronts; where a level has
// both, the small-front launches go to an auxiliary stream (fork / join with events) so that the two chains of
// launches overlap.  Per-launch profiling keeps everything on one stream.
int run_schedule(smslu_handle_t h, const std::vector<Launch>& sched, const double* win, double* zx, int rb = 1) {
    int rc;
    for (size_t i = 0; i < sched.size();) {
        size_t j = i;
        bool used[NLANES] = {false, false, false, false};
        while (j < sched.size() && sched[j].level == sched[i].level) { used[sched[j].lane] = true; ++j; }
        const int nused = (int)used[0] + (int)used[1] + (int)used[2] + (int)used[3];
        if (h->scatter_pending && (used[0] || used[NLANES - 1])) {      // first level with a big front: its panels must be ready
            CU(cudaStreamWaitEvent(h->stream, h->ev_scatter, 0));
            h->scatter_pending = false;
        }
        if (h->level_times) {
            smslu_handle_s::LevelMark m;
            CU(cudaEventCreate(&m.ev));
            CU(cudaEventRecord(m.ev, h->stream));
            m.level = sched[i].level; m.nlaunch = (int)(j - i); m.ntasks = 0; m.sched = &sched;
            for (size_t t = i; t < j; ++t) m.ntasks += sched[t].ntasks;
            h->lvl_marks.push_back(m);
        }
        const bool fork = nused > 1 && !h->profile && h->aux_stream[0];
        auto lane_stream = [&](int lane) { return (fork && lane > 0) ? h->aux_stream[lane - 1] : h->stream; };
        if (fork) {
            CU(cudaEventRecord(h->ev_fork, h->stream));
            for (int a = 1; a < NLANES; ++a) if (used[a]) CU(cudaStreamWaitEvent(h->aux_stream[a - 1], h->ev_fork, 0));
        }
        for (size_t t = i; t < j; ++t) {
            const Launch& L = sched[t];
            if ((rc = prof_begin(h, (L.kind == L_FWD_CHAIN || L.kind == L_FWD32) ? SMSLU_K_FWD : (L.kind >= L_BWD_RECT ? SMSLU_K_BWD : (L.kind >= L_REPL ? SMSLU_K_ALLREDUCE : L.kind))))) return rc;
            if (launch_one(h, lane_stream(L.lane), L, win, zx, rb) != 0) return fail(h, SMSLU_E_CUDA, std::string("cooperative launch of a chain kernel failed: ") + cudaGetErrorString(cudaGetLastError()));
            if ((rc = prof_end(h))) return rc;
        }
        if (fork) {
            for (int a = 1; a < NLANES; ++a)
                if (used[a]) {
                    CU(cudaEventRecord(h->ev_join[a - 1], h->aux_stream[a - 1]));
                    CU(cudaStreamWaitEvent(h->stream, h->ev_join[a - 1], 0));
                }
        }
        i = j;
    }
    if (h->level_times) {
        smslu_handle_s::LevelMark m;
        CU(cudaEventCreate(&m.ev));
        CU(cudaEventRecord(m.ev, h->stream));
        m.level = -1; m.nlaunch = 0; m.ntasks = 0; m.sched = nullptr;
        h->lvl_marks.push_back(m);
    }
    CU(cudaGetLastError());
    return 0;
}

// Enqueue one numeric refactorization on h->stream; av / Rs are DEVICE pointers (Rs may be null).
int enqueue_refactor(smslu_handle_t h, const double* av, bool rs_given) {
    const Symbolic& S = h->S;
    int rc;
    if (!rs_given && h->opt.scaling == SMSLU_SCALE_SUM) {
        if ((rc = prof_begin(h, SMSLU_K_ROWSCALE))) return rc;
        launch_rowscale(h->stream, h->n, h->d_rowptr, h->d_rowidx, av, h->d_Rs);
        if ((rc = prof_end(h))) return rc;
    }
    if ((rc = prof_begin(h, SMSLU_K_SCATTER))) return rc;
    CU(cudaMemsetAsync(h->cx.flag, 0x7f, 2 * sizeof(int), h->stream));   // 0x7f7f7f7f = clean
    CU(cudaMemsetAsync(h->cx.counters, 0, sizeof(int) * std::max<int64_t>(h->ncounters, 1), h->stream));
    // The zero-fill and scatter of the big fronts' panels only matter from the first level that has a big front:
    // they run on the big-front lane's stream while the main stream starts on the small-front levels.
    cudaStream_t zs = h->stream;
    if (!h->profile && h->aux_stream[NLANES - 2]) {
        zs = h->aux_stream[NLANES - 2];
        CU(cudaEventRecord(h->ev_fork, h->stream));
        CU(cudaStreamWaitEvent(zs, h->ev_fork, 0));
    }
    if (S.lu_top_size > 0) CU(cudaMemsetAsync(h->cx.lu, 0, sizeof(double) * S.lu_top_size, zs));
    if (S.lu_big_end[h->rank] > S.lu_big_begin[h->rank])
        CU(cudaMemsetAsync(h->cx.lu + S.lu_big_begin[h->rank], 0,
                           sizeof(double) * (S.lu_big_end[h->rank] - S.lu_big_begin[h->rank]), zs));
    launch_scatter(zs, h->nnz_big, h->d_big_dst, h->d_big_row, h->d_big_src, h->d_Rs, av, h->cx.lu);
    if (zs != h->stream) { CU(cudaEventRecord(h->ev_scatter, zs)); h->scatter_pending = true; }
    h->cur_av = av;
    if ((rc = prof_end(h))) return rc;
    if ((rc = run_schedule(h, h->fac, nullptr, nullptr))) return rc;
    if (h->scatter_pending) { CU(cudaStreamWaitEvent(h->stream, h->ev_scatter, 0)); h->scatter_pending = false; }
    if (h->nranks > 1) {
        // The subtree roots' Schur updates have stored their blocks column by column into the owners' pools.  Barrier
        // (every rank is past its subtrees, and past the solves that read the previous factors), then the distributed
        // top: per level the panel owners publish their panels into every pool and signal; Schur updates by column owner.
        if ((rc = prof_begin(h, SMSLU_K_ALLREDUCE))) return rc;
        NCCLCHK(nccl_api().AllReduce(h->d_barrier, h->d_barrier, 1, ncclInt, ncclMax, h->comm, h->stream));
        if ((rc = prof_end(h))) return rc;
        ++h->epoch;
        if ((rc = run_schedule(h, h->fac_top, nullptr, nullptr))) return rc;
        if ((rc = prof_begin(h, SMSLU_K_ALLREDUCE))) return rc;
        launch_replicate_rows(h->stream, h->cx, h->d_trow_tasks, h->n_trow_tasks);     // U12' rows to every pool (for the solves)
        // pivot flags of all ranks; also the barrier behind which every pool holds the complete top factors
        NCCLCHK(nccl_api().AllReduce(h->cx.flag, h->cx.flag, 2, ncclInt, ncclMin, h->comm, h->stream));
        if ((rc = prof_end(h))) return rc;
    }
    // the solves apply the 32 x 32 diagonal blocks of the big fronts through their inverses
    if ((rc = prof_begin(h, SMSLU_K_PANEL))) return rc;
    launch_diag_inverse(h->stream, h->cx, h->d_inv_tasks, h->n_inv_tasks);
    if ((rc = prof_end(h))) return rc;
    CU(cudaMemcpyAsync(h->h_flag, h->cx.flag, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    h->pending_refactor = true;
    return 0;
}

// kernels of one refactorization besides the optional row scaling: scatter, the schedules, the diagonal-block inverses
int64_t refactor_launches(smslu_handle_t h) {
    return 1 + (int64_t)h->fac.size() + (int64_t)h->fac_top.size() + (h->n_inv_tasks > 0 ? 1 : 0);
}

constexpr int FLAG_CLEAN = 0x7f7f7f7f;

// After the stream has been synchronized: turn the device pivot flag into a status.
int finish_refactor(smslu_handle_t h) {
    h->pending_refactor = false;
    h->st.n_refactor++;
    const int flag = *h->h_flag;
    if (flag == -3 || flag == -4) {          // a bounded device-side wait gave up (k_wait: a peer GPU; chain_wait: a CTA of the same kernel)
        h->factored = false;
        return fail(h, SMSLU_E_INTERNAL, flag == -3 ? "timed out waiting for a panel from a peer GPU (a rank died or fell out of step)"
                                                    : "a chain solve kernel timed out waiting for another CTA");
    }
    if (flag != FLAG_CLEAN) {
        h->factored = false;
        h->st.bad_pivot_col = flag;
        char buf[160];
        snprintf(buf, sizeof buf, "zero or non-finite pivot at permuted column %d under the static pivot order", flag);
        return fail(h, SMSLU_E_PIVOT, buf);
    }
    h->st.bad_pivot_col = -1;
    h->factored = true;
    const int tflag = h->h_flag[1];
    h->st.threshold_col = tflag != FLAG_CLEAN ? tflag : -1;
    if (tflag != FLAG_CLEAN) {
        // the factors are complete and usable, but the static pivots fail the threshold test for these values
        char buf[240];
        snprintf(buf, sizeof buf, "a multiplier exceeds 1/pivot_tol = %.3g in the front starting at permuted column %d: the static "
                 "pivot order fails the threshold test for these values; re-analyse with fresh pivots", h->cx.lmax, tflag);
        return fail(h, SMSLU_E_REPIVOT, buf);
    }
    return 0;
}

// Partitioned forward solve across the cut: sum this rank's subtree contributions per interface front
// into the virtual children's vectors, all-reduce them, then run the top of the tree.
int enqueue_top_forward(smslu_handle_t h, int rb) {
    int rc;
    if ((rc = prof_begin(h, SMSLU_K_ALLREDUCE))) return rc;
    launch_vgather(h->stream, h->cx, h->d_vtasks, h->nvtasks, h->d_vlist, rb);
    if (h->vupd_len > 0)
        NCCLCHK(nccl_api().AllReduce(h->cx.upd + h->vupd_off * rb, h->cx.upd + h->vupd_off * rb, (size_t)h->vupd_len * rb, ncclDouble, ncclSum, h->comm, h->stream));
    if ((rc = prof_end(h))) return rc;
    return run_schedule(h, rb == 1 ? h->fwd_top1 : h->fwd_top, h->d_w, h->d_z, rb);
}

// Every rank holds the solution on its own columns and on the top columns; zero the rest (rank 0
// keeps the top) and sum over ranks so that every rank ends with the full vector.
int enqueue_gather_solution(smslu_handle_t h, int rb) {
    int rc;
    if ((rc = prof_begin(h, SMSLU_K_ALLREDUCE))) return rc;
    launch_mask_owned(h->stream, h->n, h->d_colowner, h->rank, h->d_z, rb);
    NCCLCHK(nccl_api().AllReduce(h->d_z, h->d_z, (size_t)h->n * rb, ncclDouble, ncclSum, h->comm, h->stream));
    return prof_end(h);
}

// nv right-hand sides (columns bdev + q * ldb) are swept together in rb = 1, 4 or 8 slots (nv <= rb).
int enqueue_solve(smslu_handle_t h, double* xdev, int64_t ldx, const double* bdev, int64_t ldb, int rb, int nv) {
    int rc;
    if ((rc = prof_begin(h, SMSLU_K_PERMUTE))) return rc;
    if (rb == RB_WIDE) {                 // one GPU: forward and backward sweeps of 32 right-hand sides
        launch_permute_scale(h->stream, h->n, h->d_p, h->d_Rs, bdev, ldb, h->d_w32, rb, nv);
        if ((rc = prof_end(h))) return rc;
        if ((rc = run_schedule(h, h->fwd32, h->d_w32, h->d_z32, rb))) return rc;
        if ((rc = run_schedule(h, h->bwd32, nullptr, h->d_z32, rb))) return rc;
        if ((rc = prof_begin(h, SMSLU_K_UNPERMUTE))) return rc;
        launch_unpermute(h->stream, h->n, h->d_q, h->d_z32, xdev, ldx, rb, nv);
        return prof_end(h);
    }
    launch_permute_scale(h->stream, h->n, h->d_p, h->d_Rs, bdev, ldb, h->d_w, rb, nv);
    if ((rc = prof_end(h))) return rc;
    if ((rc = run_schedule(h, rb == 1 ? h->fwd1 : h->fwd, h->d_w, h->d_z, rb))) return rc;
    if (h->nranks > 1) {
        if ((rc = enqueue_top_forward(h, rb))) return rc;
        if ((rc = run_schedule(h, rb == 1 ? h->bwd_top1 : h->bwd_top, nullptr, h->d_z, rb))) return rc;
    }
    if ((rc = run_schedule(h, rb == 1 ? h->bwd1 : h->bwd, nullptr, h->d_z, rb))) return rc;
    if (h->nranks > 1 && (rc = enqueue_gather_solution(h, rb))) return rc;
    if ((rc = prof_begin(h, SMSLU_K_UNPERMUTE))) return rc;
    launch_unpermute(h->stream, h->n, h->d_q, h->d_z, xdev, ldx, rb, nv);
    if ((rc = prof_end(h))) return rc;
    return 0;
}

// Work vectors of the 32-wide sweeps, allocated by the first solve that has enough right-hand sides for one (one GPU only).
// If the device cannot hold them the solve quietly stays with the 8-wide sweeps.
bool ensure_wide(smslu_handle_t h) {
    if (h->have_wide) return true;
    if (h->wide_failed || h->nranks > 1 || h->fwd32.empty()) return false;
    if (getenv("SMSLU_NO_WIDE") && atoi(getenv("SMSLU_NO_WIDE")) != 0) { h->wide_failed = true; return false; }
    const Symbolic& S = h->S;
    const size_t n = (size_t)h->n;
    double *w = nullptr, *z = nullptr, *xb = nullptr, *upd = nullptr, *bp = nullptr;
    int* cnt = nullptr;
    auto grab = [&](void** p, size_t bytes) {
        if (cudaMalloc(p, std::max<size_t>(bytes, 8)) != cudaSuccess) { cudaGetLastError(); *p = nullptr; return false; }
        h->dev_allocs.push_back(*p);
        return true;
    };
    const bool ok = grab((void**)&w, sizeof(double) * n * RB_WIDE) && grab((void**)&z, sizeof(double) * n * RB_WIDE) &&
                    grab((void**)&xb, sizeof(double) * n * RB_WIDE) && grab((void**)&upd, sizeof(double) * (size_t)S.sum_r * RB_WIDE) &&
                    grab((void**)&bp, sizeof(double) * (size_t)h->bpart32_slots * KMAX * RB_WIDE) && grab((void**)&cnt, sizeof(int) * (size_t)std::max(S.nsn, 1));
    if (!ok || cudaMemset(cnt, 0, sizeof(int) * (size_t)std::max(S.nsn, 1)) != cudaSuccess) { cudaGetLastError(); h->wide_failed = true; return false; }
    h->d_w32 = w; h->d_z32 = z; h->d_xb32 = xb;
    h->cx32 = h->cx;
    h->cx32.upd = upd; h->cx32.bpart = bp; h->cx32.counters2 = cnt;
    h->have_wide = true;
    return true;
}

// slots of the next sweep: a partly filled sweep of 4 or 8 beats several narrower ones; from WIDE_MIN right-hand sides on a
// 32-wide sweep on the tensor pipe (one GPU) beats two or more 8-wide ones
constexpr int WIDE_MIN = 9;
inline int chunk_rb(int64_t left) { return left >= 5 ? 8 : (left >= 2 ? 4 : 1); }

}  // namespace

extern "C" {

int smslu_version(void) { return 100; }
int smslu_debug_trace(int64_t* out32) { return out32 ? debug_read_trace((long long*)out32) : SMSLU_E_ARG; }
int smslu_allocate_shared(void) { return 0; }

int smslu_host_alloc(void** ptr, int64_t bytes) {
    if (!ptr || bytes < 0) return SMSLU_E_ARG;
    cudaError_t e = cudaMallocHost(ptr, (size_t)std::max<int64_t>(bytes, 1));
    if (e != cudaSuccess) { *ptr = nullptr; cudaGetLastError(); return e == cudaErrorMemoryAllocation ? SMSLU_E_OOM : SMSLU_E_CUDA; }
    return 0;
}
int smslu_host_free(void* ptr) {
    if (!ptr) return 0;
    return cudaFreeHost(ptr) == cudaSuccess ? 0 : SMSLU_E_CUDA;
}

int smslu_options_default(smslu_options_t* o) {
    if (!o) return SMSLU_E_ARG;
    memset(o, 0, sizeof(*o));
    o->ordering = SMSLU_ORD_AUTO;
    o->nd_leaf = 48;
    o->relax = 1;
    o->max_width = KMAX;
    o->scaling = SMSLU_SCALE_SUM;
    o->device = -1;
    o->pivot_tol = 0.0;   // = 1e-3
    return 0;
}

int smslu_create(smslu_handle_t* hp, int64_t n, const int64_t* colptr, const int64_t* rowval,
                 int32_t index_base, const smslu_options_t* opts) {
    if (!hp) return SMSLU_E_ARG;
    *hp = nullptr;
    if (!colptr || !rowval || n <= 0 || n > INT_MAX - 1 || (index_base != 0 && index_base != 1)) return SMSLU_E_ARG;
    smslu_handle_t h = new (std::nothrow) smslu_handle_s();
    if (!h) return SMSLU_E_OOM;
    if (opts) h->opt = *opts; else smslu_options_default(&h->opt);
    if (h->opt.max_width <= 0 || h->opt.max_width > KMAX) h->opt.max_width = KMAX;
    h->nranks = std::max(1, (int)h->opt.nranks);
    h->rank = h->nranks > 1 ? (int)h->opt.rank : 0;
    if (h->rank < 0 || h->rank >= h->nranks) { delete h; return SMSLU_E_ARG; }
    h->n = (int)n;
    h->index_base = index_base;
    h->annz = colptr[n] - index_base;
    if (h->annz < 0) { delete h; return SMSLU_E_PATTERN; }
    h->Ap.resize(n + 1);
    h->Ai.resize(h->annz);
    for (int64_t c = 0; c <= n; ++c) h->Ap[c] = colptr[c] - index_base;
    for (int64_t t = 0; t < h->annz; ++t) h->Ai[t] = rowval[t] - index_base;
    h->st.n = n;
    h->st.nnz_a = h->annz;
    h->st.bad_pivot_col = -1;
    h->st.threshold_col = -1;
    *hp = h;
    return 0;
}

int smslu_analyze(smslu_handle_t h, const int64_t* p, const int64_t* q) {
    if (!h) return SMSLU_E_ARG;
    if (h->uploaded) return fail(h, SMSLU_E_ARG, "pattern already uploaded; create a new handle to re-analyse");
    double t0 = now_ms();
    SymOptions o;
    o.ordering = h->opt.ordering;
    for (int d = 0; d < 3; ++d) o.grid[d] = h->opt.grid[d];
    if (h->opt.nd_leaf > 0) o.nd_leaf = h->opt.nd_leaf;
    o.relax = h->opt.relax;
    o.max_width = h->opt.max_width;
    o.small_front_max = front_small_limit();
    o.small_k_max = NB;
    o.relax_width = NB;
    o.nranks = h->nranks;
    std::vector<int> pp, qq;
    if (o.ordering == ORD_GIVEN) {
        if (!p || !q) return fail(h, SMSLU_E_ARG, "ordering GIVEN needs p and q");
        pp.resize(h->n); qq.resize(h->n);
        const int64_t base = h->index_base;   // p, q use the index base of the pattern
        for (int i = 0; i < h->n; ++i) { pp[i] = (int)(p[i] - base); qq[i] = (int)(q[i] - base); }
    }
    int rc = analyze(h->n, h->Ap.data(), h->Ai.data(), pp.empty() ? nullptr : pp.data(),
                     qq.empty() ? nullptr : qq.data(), o, h->S, h->err);
    if (rc) return rc;
    h->analyzed = true;
    h->have_exact = false;
    const Symbolic& S = h->S;
    smslu_stats_t& st = h->st;
    st.nnz_l_exact = st.nnz_u_exact = S.nnzL_exact;
    st.nnz_l_stored = st.nnz_u_stored = S.nnzL_stored;
    st.n_supernodes = S.nsn; st.n_levels = S.nlevels; st.max_front = S.max_front;
    st.max_pivot_block = S.max_k; st.max_children = S.max_children; st.sum_rows = S.sum_r;
    st.lu_pool_doubles = S.lu_size; st.cb_pool_doubles = S.cb_size;
    st.flops_exact = S.flops_exact; st.flops_stored = S.flops_stored;
    st.n_top_supernodes = st.n_local_supernodes = 0;
    for (int s2 = 0; s2 < S.nsn; ++s2) {
        if (S.owner[s2] == -1) ++st.n_top_supernodes;
        if (S.owner[s2] == h->rank) ++st.n_local_supernodes;
    }
    int64_t vlen = 0;
    for (int s2 = 0; s2 < S.nsn; ++s2)
        if (S.iface[s2]) vlen += (S.sn_start[s2 + 1] - S.sn_start[s2]) + (S.rows_ptr[s2 + 1] - S.rows_ptr[s2]);
    st.allreduce_doubles_refactor = 0;     // set at upload: doubles this rank stores into peer pools per refactorization
    st.allreduce_doubles_solve = h->nranks > 1 ? vlen + S.n : 0;
    st.ms_analyze = now_ms() - t0;
    return 0;
}

int smslu_comm_unique_id(void* id, int64_t nbytes) {
    if (!id || nbytes < (int64_t)sizeof(ncclUniqueId)) return SMSLU_E_ARG;
    if (!nccl_api().ok) return SMSLU_E_NCCL;
    ncclUniqueId u;
    if (nccl_api().GetUniqueId(&u) != ncclSuccess) return SMSLU_E_NCCL;
    memcpy(id, &u, sizeof(u));
    return 0;
}

int smslu_comm_init(smslu_handle_t h, const void* id, int64_t nbytes) {
    DeviceGuard device_guard_;
    if (!h || !id || nbytes < (int64_t)sizeof(ncclUniqueId)) return SMSLU_E_ARG;
    if (h->nranks <= 1) return 0;
    if (h->comm) return fail(h, SMSLU_E_ARG, "communicator already initialised");
    if (!nccl_api().ok) return fail(h, SMSLU_E_NCCL, "libnccl.so.2 could not be loaded");
    if (h->opt.device >= 0) h->device = h->opt.device;
    else CU(cudaGetDevice(&h->device));
    CU(cudaSetDevice(h->device));
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    NCCLCHK(nccl_api().CommInitRank(&h->comm, h->nranks, u, h->rank));
    return 0;
}

int smslu_refactor(smslu_handle_t h, const double* nzval, const double* Rs) {
    DeviceGuard device_guard_;
    if (!h || !nzval) return SMSLU_E_ARG;
    int rc = ensure_uploaded(h);
    if (rc) return rc;
    h->factored = false;
    CU(cudaEventRecord(h->ev0, h->stream));
    const double* av = nzval;
    if (!is_device_ptr(nzval)) {
        CU(cudaMemcpyAsync(h->d_aval, nzval, sizeof(double) * h->annz, cudaMemcpyHostToDevice, h->stream));
        av = h->d_aval;
    }
    if (Rs) CU(cudaMemcpyAsync(h->d_Rs, Rs, sizeof(double) * h->n, cudaMemcpyDefault, h->stream));
    else if (h->opt.scaling != SMSLU_SCALE_SUM) {
        std::vector<double> ones(h->n, 1.0);
        CU(cudaMemcpyAsync(h->d_Rs, ones.data(), sizeof(double) * h->n, cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    CU(cudaEventRecord(h->ev1, h->stream));
    if ((rc = enqueue_refactor(h, av, Rs != nullptr))) return rc;
    CU(cudaEventRecord(h->ev2, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1)); h->st.ms_refactor_h2d = ms;
    CU(cudaEventElapsedTime(&ms, h->ev1, h->ev2)); h->st.ms_refactor = ms;
    h->st.launches_refactor = refactor_launches(h) + ((!Rs && h->opt.scaling == SMSLU_SCALE_SUM) ? 1 : 0);
    if ((rc = prof_collect(h))) return rc;
    return finish_refactor(h);
}

int smslu_refactor_async(smslu_handle_t h, const double* nzval_dev, const double* Rs_dev) {
    DeviceGuard device_guard_;
    if (!h || !nzval_dev) return SMSLU_E_ARG;
    int rc = ensure_uploaded(h);
    if (rc) return rc;
    if (!is_device_ptr(nzval_dev) || (Rs_dev && !is_device_ptr(Rs_dev)))
        return fail(h, SMSLU_E_ARG, "smslu_refactor_async needs device pointers");
    if (Rs_dev) CU(cudaMemcpyAsync(h->d_Rs, Rs_dev, sizeof(double) * h->n, cudaMemcpyDeviceToDevice, h->stream));
    else if (h->opt.scaling != SMSLU_SCALE_SUM) return fail(h, SMSLU_E_ARG, "async refactor needs Rs or SUM scaling");
    h->st.launches_refactor = refactor_launches(h) + ((!Rs_dev) ? 1 : 0);
    h->factored = false;                       // valid again once smslu_sync has seen the pivot flags
    return enqueue_refactor(h, nzval_dev, Rs_dev != nullptr);
}

int smslu_solve_async(smslu_handle_t h, double* x_dev, const double* b_dev) {
    DeviceGuard device_guard_;
    if (!h || !x_dev || !b_dev) return SMSLU_E_ARG;
    if (!h->uploaded) return fail(h, SMSLU_E_ARG, "no factorization has been enqueued");
    if (!h->factored && !h->pending_refactor) return fail(h, SMSLU_E_ARG, "no valid factorization (call smslu_refactor)");
    if (!is_device_ptr(x_dev) || !is_device_ptr(b_dev)) return fail(h, SMSLU_E_ARG, "smslu_solve_async needs device pointers");
    CU(cudaSetDevice(h->device));
    h->st.launches_solve = (int64_t)h->fwd1.size() + (int64_t)h->bwd1.size() + (int64_t)h->fwd_top1.size() + (int64_t)h->bwd_top1.size() + 2;
    return enqueue_solve(h, x_dev, h->n, b_dev, h->n, 1, 1);
}

int smslu_sync(smslu_handle_t h) {
    DeviceGuard device_guard_;
    if (!h) return SMSLU_E_ARG;
    if (!h->uploaded) return 0;
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    int rc;
    if ((rc = prof_collect(h))) return rc;
    if (h->pending_refactor) return finish_refactor(h);
    return 0;
}

int smslu_set_stream(smslu_handle_t h, void* stream) {
    if (!h) return SMSLU_E_ARG;
    if (h->uploaded) {
        CU(cudaSetDevice(h->device));
        CU(cudaStreamSynchronize(h->stream));
        if (h->own_stream) cudaStreamDestroy(h->stream);
        h->stream = (cudaStream_t)stream;
        h->own_stream = false;
    }
    h->user_stream = (cudaStream_t)stream;
    h->have_user_stream = true;
    return 0;
}

int smslu_set_profile(smslu_handle_t h, int32_t on) {
    if (!h) return SMSLU_E_ARG;
    h->profile = on != 0;
    if (on) { memset(h->st.ms_kernel, 0, sizeof h->st.ms_kernel); memset(h->st.launches_kernel, 0, sizeof h->st.launches_kernel); }
    return 0;
}

static int check_vec(smslu_handle_t h, int64_t len, int64_t nrhs, int64_t ld, const char* what) {
    if (len != h->n) return fail(h, SMSLU_E_DIM, std::string("`") + what + "` does not have same size as F");
    if (nrhs < 1 || (nrhs > 1 && ld < h->n)) return fail(h, SMSLU_E_ARG, "bad nrhs / leading dimension");
    return 0;
}

int smslu_solve(smslu_handle_t h, double* x, int64_t nx, const double* b, int64_t nb,
                int64_t nrhs, int64_t ldx, int64_t ldb) {
    DeviceGuard device_guard_;
    if (!h || !x || !b) return SMSLU_E_ARG;
    int rc;
    if ((rc = check_vec(h, nx, nrhs, ldx, "x"))) return rc;
    if ((rc = check_vec(h, nb, nrhs, ldb, "b"))) return rc;
    if (h->pending_refactor && (rc = smslu_sync(h)) && rc != SMSLU_E_REPIVOT) return rc;
    if (!h->factored) return fail(h, SMSLU_E_ARG, "no valid factorization (call smslu_refactor)");
    if ((rc = ensure_uploaded(h))) return rc;
    const int n = h->n;
    const bool xdev = is_device_ptr(x), bdev = is_device_ptr(b);
    float h2d = 0, dev = 0, d2h = 0, ms;
    int64_t nsweeps = 0;
    for (int64_t c = 0; c < nrhs;) {
        const int rb = (nrhs - c >= WIDE_MIN && ensure_wide(h)) ? RB_WIDE : chunk_rb(nrhs - c);     // 32, 8, 4 or 1 slots per sweep
        const int nv = (int)std::min<int64_t>(rb, nrhs - c);
        double* const stage = rb == RB_WIDE ? h->d_xb32 : h->d_xb;      // host columns pass through this device block
        const double* bc = b + c * ldb;
        double* xc = x + c * ldx;
        int64_t lb = ldb, lx = ldx;
        CU(cudaEventRecord(h->ev0, h->stream));
        if (!bdev) {                                     // host columns -> contiguous device block, ld = n
            CU(cudaMemcpy2DAsync(stage, sizeof(double) * n, bc, sizeof(double) * ldb, sizeof(double) * n, nv,
                                 cudaMemcpyHostToDevice, h->stream));
            bc = stage; lb = n;
        }
        CU(cudaEventRecord(h->ev1, h->stream));
        if ((rc = enqueue_solve(h, xdev ? xc : stage, xdev ? lx : n, bc, lb, rb, nv))) return rc;
        CU(cudaEventRecord(h->ev2, h->stream));
        if (!xdev) CU(cudaMemcpy2DAsync(xc, sizeof(double) * ldx, stage, sizeof(double) * n, sizeof(double) * n, nv,
                                        cudaMemcpyDeviceToHost, h->stream));
        CU(cudaEventRecord(h->ev3, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1)); h2d += ms;
        CU(cudaEventElapsedTime(&ms, h->ev1, h->ev2)); dev += ms;
        CU(cudaEventElapsedTime(&ms, h->ev2, h->ev3)); d2h += ms;
        c += nv; ++nsweeps;
    }
    {   // the persistent chain kernels report a timed-out wait through the pivot flag word
        int f0 = 0;
        CU(cudaMemcpy(&f0, h->cx.flag, sizeof(int), cudaMemcpyDeviceToHost));
        if (f0 == -4) return fail(h, SMSLU_E_INTERNAL, "a chain solve kernel timed out waiting for another CTA");
    }
    h->st.ms_solve_h2d = h2d; h->st.ms_solve = dev; h->st.ms_solve_d2h = d2h;
    h->st.launches_solve = nrhs == 1 ? (int64_t)h->fwd1.size() + (int64_t)h->bwd1.size() + (int64_t)h->fwd_top1.size() + (int64_t)h->bwd_top1.size() + 2
                                      : nsweeps * ((int64_t)h->fwd.size() + (int64_t)h->bwd.size() + (int64_t)h->fwd_top.size() + (int64_t)h->bwd_top.size() + 2);
    h->st.n_solve++;
    return prof_collect(h);
}

static int tri_solve(smslu_handle_t h, double* x, int64_t nx, int64_t nrhs, int64_t ld, bool lower) {
    DeviceGuard device_guard_;
    if (!h || !x) return SMSLU_E_ARG;
    int rc;
    if ((rc = check_vec(h, nx, nrhs, ld, "x"))) return rc;
    if (h->pending_refactor && (rc = smslu_sync(h)) && rc != SMSLU_E_REPIVOT) return rc;
    if (!h->factored) return fail(h, SMSLU_E_ARG, "no valid factorization (call smslu_refactor)");
    if ((rc = ensure_uploaded(h))) return rc;
    const int n = h->n;
    const bool dev = is_device_ptr(x);
    for (int64_t c = 0; c < nrhs;) {
        const int rb = (nrhs - c >= WIDE_MIN && ensure_wide(h)) ? RB_WIDE : chunk_rb(nrhs - c);
        const int nv = (int)std::min<int64_t>(rb, nrhs - c);
        double* const stage = rb == RB_WIDE ? h->d_xb32 : h->d_xb;
        double* xc = x + c * ld;
        const double* src = xc; int64_t lsrc = ld;
        if (!dev) {
            CU(cudaMemcpy2DAsync(stage, sizeof(double) * n, xc, sizeof(double) * ld, sizeof(double) * n, nv,
                                 cudaMemcpyHostToDevice, h->stream));
            src = stage; lsrc = n;
        }
        if (rb == RB_WIDE) {                 // one GPU, 32 slots: the same sweep on the tensor-pipe kernels
            launch_permute_scale(h->stream, n, h->d_post, nullptr, src, lsrc, lower ? h->d_w32 : h->d_z32, rb, nv);
            if ((rc = run_schedule(h, lower ? h->fwd32 : h->bwd32, lower ? h->d_w32 : nullptr, h->d_z32, rb))) return rc;
            launch_unpermute(h->stream, n, h->d_post, h->d_z32, dev ? xc : stage, dev ? ld : n, rb, nv);
            if (!dev) CU(cudaMemcpy2DAsync(xc, sizeof(double) * ld, stage, sizeof(double) * n, sizeof(double) * n, nv,
                                           cudaMemcpyDeviceToHost, h->stream));
            CU(cudaStreamSynchronize(h->stream));
            c += nv;
            continue;
        }
        // interleave the block (identity permutation, no scaling), sweep, de-interleave
        if (lower) {
            launch_permute_scale(h->stream, n, h->d_post, nullptr, src, lsrc, h->d_w, rb, nv);
            if ((rc = run_schedule(h, rb == 1 ? h->fwd1 : h->fwd, h->d_w, h->d_z, rb))) return rc;
            if (h->nranks > 1 && (rc = enqueue_top_forward(h, rb))) return rc;
        } else {
            launch_permute_scale(h->stream, n, h->d_post, nullptr, src, lsrc, h->d_z, rb, nv);
            if (h->nranks > 1 && (rc = run_schedule(h, rb == 1 ? h->bwd_top1 : h->bwd_top, nullptr, h->d_z, rb))) return rc;
            if ((rc = run_schedule(h, rb == 1 ? h->bwd1 : h->bwd, nullptr, h->d_z, rb))) return rc;
        }
        if (h->nranks > 1 && (rc = enqueue_gather_solution(h, rb))) return rc;
        launch_unpermute(h->stream, n, h->d_post, h->d_z, dev ? xc : h->d_xb, dev ? ld : n, rb, nv);
        if (!dev) CU(cudaMemcpy2DAsync(xc, sizeof(double) * ld, h->d_xb, sizeof(double) * n, sizeof(double) * n, nv,
                                       cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        c += nv;
    }
    return 0;
}

int smslu_lsolve(smslu_handle_t h, double* x, int64_t nx, int64_t nrhs, int64_t ld) { return tri_solve(h, x, nx, nrhs, ld, true); }
int smslu_rsolve(smslu_handle_t h, double* x, int64_t nx, int64_t nrhs, int64_t ld) { return tri_solve(h, x, nx, nrhs, ld, false); }

int smslu_get_nnz(smslu_handle_t h, int64_t* nnz_l, int64_t* nnz_u) {
    if (!h || !h->analyzed) return SMSLU_E_ARG;
    if (nnz_l) *nnz_l = h->S.nnzL_exact;
    if (nnz_u) *nnz_u = h->S.nnzL_exact;
    return 0;
}

int smslu_get_factors(smslu_handle_t h, int64_t* lp, int64_t* li, double* lx, int64_t* up, int64_t* ui,
                      double* ux, int64_t* p, int64_t* q, double* Rs, int32_t index_base) {
    DeviceGuard device_guard_;
    if (!h || !h->analyzed) return SMSLU_E_ARG;
    if (index_base != 0 && index_base != 1) return SMSLU_E_ARG;
    const Symbolic& S = h->S;
    const bool relabel = !S.post.empty();        // ORD_GIVEN: answer in the caller's labelling (see Symbolic::post)
    if (p) for (int i = 0; i < S.n; ++i) p[i] = (relabel ? S.p_given[i] : S.p[i]) + index_base;
    if (q) for (int i = 0; i < S.n; ++i) q[i] = (relabel ? S.q_given[i] : S.q[i]) + index_base;
    const bool want_vals = lx || ux || Rs;
    if (want_vals && h->pending_refactor) { const int rcs = smslu_sync(h); if (rcs && rcs != SMSLU_E_REPIVOT) return rcs; }
    if (want_vals && !h->factored) return fail(h, SMSLU_E_ARG, "no valid factorization (call smslu_refactor)");
    if (Rs) {
        CU(cudaSetDevice(h->device));
        CU(cudaMemcpy(Rs, h->d_Rs, sizeof(double) * S.n, cudaMemcpyDeviceToHost));
    }
    if (!(lp || li || lx || up || ui || ux)) return 0;
    if (!h->have_exact) {
        exact_structure(S, h->Ap.data(), h->Ai.data(), h->ex_ptr, h->ex_idx);
        h->have_exact = true;
    }
    std::vector<double> lu;
    if (lx || ux) {
        CU(cudaSetDevice(h->device));
        lu.resize((size_t)S.lu_size);
        CU(cudaMemcpy(lu.data(), h->cx.lu, sizeof(double) * S.lu_size, cudaMemcpyDeviceToHost));
    }
    std::vector<char> col_mine;
    if (h->nranks > 1) {
        // this rank's share: its own supernodes, plus the (replicated) top on rank 0; zero elsewhere, so
        // that the value arrays summed over the ranks are F.L and F.U
        col_mine.assign(S.n, 0);
        for (int s2 = 0; s2 < S.nsn; ++s2) {
            const bool mine = S.owner[s2] == h->rank || (S.owner[s2] == -1 && h->rank == 0);
            const int64_t k = S.sn_start[s2 + 1] - S.sn_start[s2], r = S.rows_ptr[s2 + 1] - S.rows_ptr[s2];
            if (mine) { for (int j = S.sn_start[s2]; j < S.sn_start[s2 + 1]; ++j) col_mine[j] = 1; continue; }
            if (!lu.empty()) {
                std::fill(lu.begin() + S.Loff[s2], lu.begin() + S.Loff[s2] + (k + r) * k, 0.0);
                std::fill(lu.begin() + S.Uoff[s2], lu.begin() + S.Uoff[s2] + r * k, 0.0);
            }
        }
    }
    if (relabel) {
        // the relabelling permutes rows inside the columns, so it needs colptr + rowval next to the values
        const int64_t nz = S.nnzL_exact;
        std::vector<int64_t> tlp, tli, tup, tui;
        if (!lp) { tlp.resize(S.n + 1); lp = tlp.data(); }
        if (!li) { tli.resize(nz); li = tli.data(); }
        if (!up) { tup.resize(S.n + 1); up = tup.data(); }
        if (!ui) { tui.resize(nz); ui = tui.data(); }
        export_factors(S, h->ex_ptr, h->ex_idx, lu.empty() ? nullptr : lu.data(), index_base, lp, li,
                       lu.empty() ? nullptr : lx, up, ui, lu.empty() ? nullptr : ux,
                       col_mine.empty() ? nullptr : col_mine.data());
        relabel_csc(S.n, S.post.data(), index_base, lp, li, lu.empty() ? nullptr : lx);
        relabel_csc(S.n, S.post.data(), index_base, up, ui, lu.empty() ? nullptr : ux);
        return 0;
    }
    export_factors(S, h->ex_ptr, h->ex_idx, lu.empty() ? nullptr : lu.data(), index_base, lp, li,
                   lu.empty() ? nullptr : lx, up, ui, lu.empty() ? nullptr : ux,
                   col_mine.empty() ? nullptr : col_mine.data());
    return 0;
}

int smslu_get_stats(smslu_handle_t h, smslu_stats_t* st) {
    if (!h || !st) return SMSLU_E_ARG;
    *st = h->st;
    return 0;
}

const char* smslu_last_error(smslu_handle_t h) { return h ? h->err.c_str() : "null handle"; }

int smslu_get_symbolic(smslu_handle_t h, int64_t* sn_start, int64_t* rows_ptr, int64_t* rows,
                       int64_t* sn_parent, int64_t* sn_level, int64_t* etree_parent, int64_t* colcount) {
    if (!h || !h->analyzed) return SMSLU_E_ARG;
    const Symbolic& S = h->S;
    if (sn_start) for (int s = 0; s <= S.nsn; ++s) sn_start[s] = S.sn_start[s];
    if (rows_ptr) for (int s = 0; s <= S.nsn; ++s) rows_ptr[s] = S.rows_ptr[s];
    if (rows) for (size_t t = 0; t < S.rows.size(); ++t) rows[t] = S.rows[t];
    if (sn_parent) for (int s = 0; s < S.nsn; ++s) sn_parent[s] = S.sn_parent[s];
    if (sn_level) for (int s = 0; s < S.nsn; ++s) sn_level[s] = S.sn_level[s];
    if (etree_parent) for (int j = 0; j < S.n; ++j) etree_parent[j] = S.parent[j];
    if (colcount) for (int j = 0; j < S.n; ++j) colcount[j] = S.colcount[j];
    return 0;
}

int smslu_destroy(smslu_handle_t h) {
    DeviceGuard device_guard_;
    if (!h) return 0;
    if (h->comm) { cudaSetDevice(h->device); nccl_api().CommDestroy(h->comm); h->comm = nullptr; }
    release_device_state(h);
    delete h;
    return 0;
}

}  // extern "C"

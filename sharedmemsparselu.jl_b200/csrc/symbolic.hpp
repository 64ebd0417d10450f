// Host-side symbolic analysis: done once per sparsity pattern, result uploaded once.
//
// Replaces the symbolic half of UMFPACK that the reference enters through `lu(A)`
// (reference src/SharedMemSparseLU.jl:74) and the reference's own solve "layout" step
// (get_chunking_parameters, src:101-149): instead of dense column chunks over the row
// envelope it produces a supernodal multifrontal layout -- fill-reducing ordering,
// elimination tree, supernodes, per-supernode row lists, child->parent index maps,
// the A->factor scatter map, storage offsets and the level schedule.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace smslu {

enum Ordering { ORD_AUTO = 0, ORD_NATURAL = 1, ORD_GIVEN = 2, ORD_ND_GRAPH = 3, ORD_ND_GRID = 4 };

struct SymOptions {
    int ordering = ORD_AUTO;
    int grid[3] = {0, 0, 0};      // nx, ny, nz hint for ORD_ND_GRID (idx = i + nx*(j + ny*k))
    int nd_leaf = 48;             // stop dissecting below this many vertices
    int max_width = 128;          // pivot-block width of a front; wider supernodes are chained
    int relax_width = 32;         // relaxed amalgamation never builds a pivot block wider than this
    int small_k_max = 32;         // ... and the fused kernel only takes pivot blocks up to this wide
    int relax = 1;                // relaxed supernode amalgamation on/off
    int relax_always = 4;         // merged width <= this: always merge
    int relax_k1 = 16;            // width <= k1: allow zero fraction relax_f1
    int relax_k2 = 48;            // width <= k2: allow zero fraction relax_f2, else relax_f3
    double relax_f1 = 0.5, relax_f2 = 0.15, relax_f3 = 0.05;
    int dense_factor = 10;        // vertices with degree > dense_factor*sqrt(n) are ordered last
    int small_front_max = 96;     // fronts up to this size run in the fused shared-memory kernel
    int nranks = 1;               // GPUs the tree is partitioned over (one process per GPU)
};

struct Symbolic {
    int n = 0;
    int64_t annz = 0;
    std::vector<int> p, q;            // B = (Rs .* A)[p, q], 0-based, postordered
    // ORD_GIVEN only (empty otherwise): the caller's (p0, q0) and post[k] = position in the caller's order of
    // internal index k (p[k] = p0[post[k]]).  The postorder is an equivalent reordering (a topological order of
    // the elimination tree), so L_caller(post[i], post[j]) = L(i, j): the C ABI reports p0, q0 and the factors
    // in the caller's labelling, and lsolve/rsolve permute their vector through post.
    std::vector<int> p_given, q_given, post;
    std::vector<int> parent;          // column elimination tree of pattern(B + B')
    std::vector<int> colcount;        // exact |struct(L(:,j))| incl. diagonal

    int nsn = 0;
    std::vector<int> sn_start;        // nsn+1 column boundaries
    std::vector<int> col2sn;          // n
    std::vector<int64_t> rows_ptr;    // nsn+1
    std::vector<int> rows;            // off-diagonal-block row indices, sorted, permuted numbering
    std::vector<int> rel;             // same shape as rows: local index in the PARENT's front
    std::vector<int> sn_parent;       // -1 for roots
    std::vector<int> child_ptr, child_idx;   // children in ascending order
    // direct[c] = 1: c is the only child of its parent and a 'big' front, so its Schur update is
    // written straight into the parent's panels / contribution block (no assembly pass).
    // cb_assigned[s] = 1: every entry of C_s is assigned by that direct child (no zeroing pass).
    std::vector<char> direct, cb_assigned;
    std::vector<int> sn_level;        // 0 = leaves
    int nlevels = 0;
    std::vector<int> level_ptr, level_sn;
    int max_children = 0;

    // factor storage (doubles): L panel f x k column-major (ld = f) holding the k x k diagonal
    // block (L strictly below the diagonal, U on and above) on top of L21 (r x k);
    // U panel r x k column-major (ld = r) holding U12 transposed.
    std::vector<int64_t> Loff, Uoff;
    int64_t lu_size = 0;
    int64_t lu_big_size = 0;          // [0, lu_big_size) holds the top and big fronts, the small ones follow
    // partition over nranks GPUs (nranks == 1: everything is owned by rank 0, no top set)
    int nranks = 1;
    std::vector<int> owner;           // rank that factors the supernode; -1 = top of the tree (distributed by columns)
    std::vector<char> iface;          // top front that receives contributions from below the cut
    // Top of the tree, distributed: every top supernode has a panel owner (it factors the pivot block and L21 and
    // replicates them); every top COLUMN j belongs to the panel owner of its supernode, and that rank holds / updates
    // column j of every trailing matrix it appears in (rows of U12', columns of contribution blocks).  A child's
    // contribution-block column and the parent entry it is added to have the same global column, hence the same owner:
    // extend-add inside the top needs no communication.
    std::vector<int> top_owner;       // per supernode: panel owner of a top supernode, else -1
    std::vector<int> col_owner;       // per permuted column: owner of a top column, else -1
    std::vector<char> xroot;          // subtree root (owner >= 0, parent in the top): its contribution block is
                                      // delivered column by column to the column owners (slots [0, cb_xchg_size))
    int64_t cb_xchg_size = 0;
    std::vector<char> small;          // handled by the shared-memory kernels (never a top front)
    int64_t lu_top_size = 0;          // [0, lu_top_size): panels of the top fronts (all-reduced)
    std::vector<int64_t> lu_big_begin, lu_big_end;   // per rank: its big fronts' panels
    int64_t cb_iface_size = 0;        // [0, cb_iface_size): contribution blocks of the interface fronts
    // contribution blocks r x r (ld = r), lifetime level(s)..level(parent(s))
    std::vector<int64_t> CBoff;
    int64_t cb_size = 0;
    // for every nonzero of the caller's CSC, in the caller's order: offset into factor storage
    std::vector<int64_t> a_dst;
    // ... the front it lands in, and its position inside that front: row | col << 16 of the f x f
    // frontal matrix [pivot block, U12; L21, contribution block] (0 when f >= 65536)
    std::vector<int> a_sn, a_loc;

    int64_t nnzL_exact = 0;           // incl. unit diagonal; nnz(U) is the same number
    int64_t nnzL_stored = 0;          // k(k+1)/2 + k*r summed (what the panels hold per factor)
    double flops_exact = 0, flops_stored = 0;
    int max_front = 0, max_k = 0;
    int64_t sum_r = 0;
};

// Ap/Ai: CSC pattern, 0-based.  p_in/q_in: only for ORD_GIVEN (0-based, B = A[p,q]).
// Returns 0 or a negative SMSLU_E_* code with a message in err.
int analyze(int n, const int64_t* Ap, const int64_t* Ai, const int* p_in, const int* q_in,
            const SymOptions& opt, Symbolic& S, std::string& err);

// Exact (unrelaxed) structure of L by columns, strictly below the diagonal, permuted numbering.
// By symmetry of the pattern this is also the structure of the rows of U right of the diagonal.
void exact_structure(const Symbolic& S, const int64_t* Ap, const int64_t* Ai,
                     std::vector<int64_t>& ptr, std::vector<int>& idx);

// Gather F.L / F.U (CSC, sorted rows, exact pattern, L with explicit unit diagonal) out of a host
// copy of the factor storage.  colptr arrays have n+1 entries; any output may be null.
// col_mine (optional, per permuted column): L's unit diagonal is emitted as 0 where it is 0.
void export_factors(const Symbolic& S, const std::vector<int64_t>& ptr, const std::vector<int>& idx,
                    const double* lu, int64_t base, int64_t* Lp, int64_t* Li, double* Lx,
                    int64_t* Up, int64_t* Ui, double* Ux, const char* col_mine = nullptr);

// Relabel a square CSC matrix in place: entry (i, j) moves to (post[i], post[j]); rows sorted inside every column.
// colptr keeps `base`; any of idx / val may be null (the other is still permuted consistently).
void relabel_csc(int n, const int* post, int64_t base, int64_t* colptr, int64_t* idx, double* val);

}  // namespace smslu

// Host-side cell topology of the sampling tree: neighbour pointers and shared node ids exactly as the reference keeps
// them on its Cell objects (sparseSpatialSampling/s_cube.py: Cell :30-85, _assign_neighbors :904-1186, _assign_indices
// :1188-1536, check_nb_node :1739-1755, parent_or_child :1758-1775, _check_nb :447-464, the neighbour reset in
// _remove_invalid_cells :721-731, _resort_nodes_and_indices_of_grid :734-772 + renumber_node_indices_parallel
// :1695-1736).
//
// Why this exists: the reference's vertex numbering and its max_delta_level closure are HISTORY dependent. A child gets
// its 8 / 26 neighbour pointers when it is created (the same-position child of the parent's neighbour if that one has
// children at that moment, else the coarser neighbour itself), pointers are only refreshed for the siblings of cells
// that get selected, cells removed by a geometry are unhooked from the cells they point to, and a node is shared only
// with a neighbour that -- through these pointers, at that moment -- exists, is a leaf and has the same level. None of
// that can be recovered from the final geometry, so the drop-in keeps the same integer bookkeeping. It is pure integer
// work on a few arrays (no device involvement, ~100 table look-ups per refined cell) and lives here in C++ because a
// Python loop over 10^5..10^6 cells would dominate the grid generation.
//
// The two hand-written case ladders of the reference are data here:
//   * neighbour rule (geometric): child at offset o in {0,1}^d looking in direction delta in {-1,0,1}^d reaches position
//     t = o + delta; per axis t in {0,1} stays inside the parent, t = -1 / 2 leaves it through the parent's neighbour
//     in that direction and enters the child with coordinate 1 / 0. All axes inside -> the sibling at t, otherwise
//     parent_or_child(parent.nb[direction], child t').  tests/golden/make_golden.py extracts the reference's own
//     table by instrumenting _assign_neighbors and asserts that it equals this rule (all 4*8 and 8*26 entries).
//   * node rule: SURVEY.md appendix C (priority lists "neighbour slot . node" per child and node, else a new node;
//     copies from already numbered siblings), checked against the faces / vertices of the reference's golden runs.
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "common.cuh"
#include "../../include/s3b200.h"

namespace s3 {

namespace {

// child / node order CH (s_cube.py:29): swu nwu neu seu | swl nwl nel sel; u = upper (z+), l = lower (z-)
const int kChildDir[8][3] = {{-1, -1, 1}, {-1, 1, 1}, {1, 1, 1}, {1, -1, 1}, {-1, -1, -1}, {-1, 1, -1}, {1, 1, -1}, {1, -1, -1}};
// neighbour slots NB (s_cube.py:22-26): same plane w nw n ne e se s sw; lower plane the same eight + cl; upper plane + cu
const int kPlane[8][2] = {{-1, 0}, {-1, 1}, {0, 1}, {1, 1}, {1, 0}, {1, -1}, {0, -1}, {-1, -1}};

enum : int { W = 0, NW = 1, N = 2, NE = 3, E = 4, SE = 5, S = 6, SW = 7, WL = 8, NWL = 9, NL = 10, NEL = 11, EL = 12,
             SEL = 13, SL = 14, SWL = 15, CL = 16, WU = 17, NWU = 18, NU = 19, NEU = 20, EU = 21, SEU = 22, SU = 23,
             SWU = 24, CU = 25 };

struct Source { int slot, node; };                 // "neighbour slot . node"
struct NodeRule { int node; int n_src; Source src[3]; };
struct CopyRule { int node, sibling, from; };

// appendix C of SURVEY.md, rules in the order the reference evaluates them (= order in which new nodes are appended)
const NodeRule kRules2D[4][3] = {
    {{1, 1, {{W, 2}}}, {2, 0, {}}, {3, 1, {{S, 2}}}},
    {{2, 1, {{N, 3}}}, {-1, 0, {}}, {-1, 0, {}}},
    {{3, 1, {{E, 0}}}, {-1, 0, {}}, {-1, 0, {}}},
    {{-1, 0, {}}, {-1, 0, {}}, {-1, 0, {}}},
};
const CopyRule kCopies2D[4][3] = {
    {{-1, 0, 0}, {-1, 0, 0}, {-1, 0, 0}},
    {{0, 0, 1}, {3, 0, 2}, {-1, 0, 0}},
    {{0, 0, 2}, {1, 1, 2}, {-1, 0, 0}},
    {{0, 0, 3}, {1, 0, 2}, {2, 2, 3}},
};
const NodeRule kRules3D[8][7] = {
    {{1, 3, {{W, 2}, {WU, 6}, {CU, 5}}}, {2, 1, {{CU, 6}}}, {3, 3, {{S, 2}, {SU, 6}, {CU, 7}}},
     {4, 3, {{W, 7}, {SW, 6}, {S, 5}}}, {5, 1, {{W, 6}}}, {6, 0, {}}, {7, 1, {{S, 6}}}},
    {{2, 3, {{N, 3}, {NU, 7}, {CU, 6}}}, {5, 3, {{W, 6}, {NW, 7}, {N, 4}}}, {6, 1, {{N, 7}}}, {-1, 0, {}}, {-1, 0, {}},
     {-1, 0, {}}, {-1, 0, {}}},
    {{3, 3, {{E, 0}, {EU, 4}, {CU, 7}}}, {6, 3, {{E, 5}, {NE, 4}, {N, 7}}}, {7, 1, {{E, 4}}}, {-1, 0, {}}, {-1, 0, {}},
     {-1, 0, {}}, {-1, 0, {}}},
    {{7, 3, {{E, 4}, {SE, 5}, {S, 6}}}, {-1, 0, {}}, {-1, 0, {}}, {-1, 0, {}}, {-1, 0, {}}, {-1, 0, {}}, {-1, 0, {}}},
    {{5, 3, {{W, 6}, {WL, 2}, {CL, 1}}}, {6, 1, {{CL, 2}}}, {7, 3, {{S, 6}, {SL, 2}, {CL, 3}}}, {-1, 0, {}}, {-1, 0, {}},
     {-1, 0, {}}, {-1, 0, {}}},
    {{6, 3, {{N, 7}, {NL, 3}, {CL, 2}}}, {-1, 0, {}}, {-1, 0, {}}, {-1, 0, {}}, {-1, 0, {}}, {-1, 0, {}}, {-1, 0, {}}},
    {{7, 3, {{E, 4}, {EL, 0}, {CL, 3}}}, {-1, 0, {}}, {-1, 0, {}}, {-1, 0, {}}, {-1, 0, {}}, {-1, 0, {}}, {-1, 0, {}}},
    {{-1, 0, {}}, {-1, 0, {}}, {-1, 0, {}}, {-1, 0, {}}, {-1, 0, {}}, {-1, 0, {}}, {-1, 0, {}}},
};
const CopyRule kCopies3D[8][7] = {
    {{-1, 0, 0}, {-1, 0, 0}, {-1, 0, 0}, {-1, 0, 0}, {-1, 0, 0}, {-1, 0, 0}, {-1, 0, 0}},
    {{0, 0, 1}, {3, 0, 2}, {4, 0, 5}, {7, 0, 6}, {-1, 0, 0}, {-1, 0, 0}, {-1, 0, 0}},
    {{0, 0, 2}, {1, 1, 2}, {4, 0, 6}, {5, 1, 6}, {-1, 0, 0}, {-1, 0, 0}, {-1, 0, 0}},
    {{0, 0, 3}, {1, 0, 2}, {2, 2, 3}, {4, 0, 7}, {5, 0, 6}, {6, 2, 7}, {-1, 0, 0}},
    {{0, 0, 4}, {1, 0, 5}, {2, 0, 6}, {3, 0, 7}, {-1, 0, 0}, {-1, 0, 0}, {-1, 0, 0}},
    {{0, 1, 4}, {1, 1, 5}, {2, 1, 6}, {3, 1, 7}, {4, 4, 5}, {7, 4, 6}, {-1, 0, 0}},
    {{0, 2, 4}, {1, 2, 5}, {2, 2, 6}, {3, 2, 7}, {4, 5, 7}, {5, 5, 6}, {-1, 0, 0}},
    {{0, 3, 4}, {1, 3, 5}, {2, 3, 6}, {3, 3, 7}, {4, 4, 7}, {5, 4, 6}, {6, 6, 7}},
};

constexpr int32_t kNone = -1;
constexpr int32_t kLeaf = -1;       // children is None
constexpr int32_t kEmpty = -2;      // children == [] (removed by a geometry)

struct NbEntry { int8_t sibling; int8_t slot; int8_t child; };   // sibling >= 0: that sibling; else parent_or_child

}  // namespace

struct Topology {
    int dim, nch, nnb;
    double width;
    std::vector<int32_t> parent, children, level;   // children: first child index | kLeaf | kEmpty
    std::vector<int32_t> nb;                         // [n_cells, nnb]
    std::vector<int32_t> node;                       // [n_cells, nch]
    std::vector<double> center;                      // [n_cells, dim]
    std::vector<double> nodes;                       // [n_nodes, dim]
    NbEntry table[8][26];

    int slot_of(const int* dir) const {
        int p = -1;
        for (int i = 0; i < 8; ++i)
            if (kPlane[i][0] == dir[0] && kPlane[i][1] == dir[1]) p = i;
        if (dim == 2) return p;
        if (dir[2] == 0) return p;
        if (dir[2] < 0) return p >= 0 ? 8 + p : 16;
        return p >= 0 ? 17 + p : 25;
    }
    void slot_dir(int slot, int* dir) const {
        dir[2] = 0;
        int p = slot;
        if (slot >= 17) { dir[2] = 1; p = slot - 17; }
        else if (slot >= 8) { dir[2] = -1; p = slot - 8; }
        if (p == 8) { dir[0] = 0; dir[1] = 0; }
        else { dir[0] = kPlane[p][0]; dir[1] = kPlane[p][1]; }
    }
    int child_of(const int* off) const {             // off in {0,1}^d -> CH index
        for (int c = 0; c < nch; ++c) {
            bool ok = true;
            for (int a = 0; a < dim; ++a) ok = ok && ((kChildDir[c][a] > 0) == (off[a] == 1));
            if (ok) return c;
        }
        return -1;
    }
    void build_table() {
        for (int c = 0; c < nch; ++c)
            for (int s = 0; s < nnb; ++s) {
                int delta[3], pdir[3] = {0, 0, 0}, off[3] = {0, 0, 0};
                slot_dir(s, delta);
                bool inside = true;
                for (int a = 0; a < dim; ++a) {
                    const int t = (kChildDir[c][a] > 0 ? 1 : 0) + delta[a];
                    if (t < 0) { pdir[a] = -1; off[a] = 1; inside = false; }
                    else if (t > 1) { pdir[a] = 1; off[a] = 0; inside = false; }
                    else off[a] = t;
                }
                NbEntry e;
                e.child = (int8_t)child_of(off);
                if (inside) { e.sibling = e.child; e.slot = -1; }
                else { e.sibling = -1; e.slot = (int8_t)slot_of(pdir); }
                table[c][s] = e;
            }
    }
    int64_t n_cells() const { return (int64_t)parent.size(); }
    int64_t n_nodes() const { return (int64_t)(nodes.size() / dim); }
    bool is_leaf(int32_t c) const { return children[c] == kLeaf; }

    // _assign_neighbors(cell, children=cell.children): (re)evaluate the neighbour pointers of the children of p
    void assign_neighbors(int32_t p) {
        const int32_t first = children[p];
        if (first < 0) return;
        const int32_t* pnb = &nb[(size_t)p * nnb];
        for (int c = 0; c < nch; ++c) {
            int32_t* cnb = &nb[(size_t)(first + c) * nnb];
            for (int s = 0; s < nnb; ++s) {
                const NbEntry e = table[c][s];
                if (e.sibling >= 0) { cnb[s] = first + e.sibling; continue; }
                const int32_t q = pnb[e.slot];
                if (q == kNone) cnb[s] = kNone;
                else if (children[q] >= 0) cnb[s] = children[q] + e.child;     // bool(n and n.children)
                else cnb[s] = q;
            }
        }
    }
    // check_nb_node: the neighbour exists, is a leaf cell and has the same level
    bool shares(int32_t cell, int slot) const {
        const int32_t q = nb[(size_t)cell * nnb + slot];
        return q != kNone && children[q] == kLeaf && level[q] == level[cell];
    }
    int32_t new_node(int32_t cell, int j) {
        // _compute_cell_centers(_factor=0.5, _cell=cell): centre + dir * 0.5 * width / 2^level
        const double h = 0.5 * width / (double)((int64_t)1 << level[cell]);
        for (int a = 0; a < dim; ++a) nodes.push_back(center[(size_t)cell * dim + a] + (double)kChildDir[j][a] * h);
        return (int32_t)(n_nodes() - 1);
    }
    void assign_indices(int32_t p) {
        const int32_t first = children[p];
        for (int c = 0; c < nch; ++c) {
            const int32_t cell = first + c;
            int32_t* nd = &node[(size_t)cell * nch];
            nd[c] = node[(size_t)p * nch + c];
            const NodeRule* rules = dim == 2 ? kRules2D[c] : kRules3D[c];
            const int n_rules = dim == 2 ? 3 : 7;
            for (int r = 0; r < n_rules && rules[r].node >= 0; ++r) {
                const NodeRule& rule = rules[r];
                int32_t id = kNone;
                for (int t = 0; t < rule.n_src; ++t)
                    if (shares(cell, rule.src[t].slot)) {
                        id = node[(size_t)nb[(size_t)cell * nnb + rule.src[t].slot] * nch + rule.src[t].node];
                        break;
                    }
                nd[rule.node] = id != kNone ? id : new_node(cell, rule.node);
            }
            const CopyRule* copies = dim == 2 ? kCopies2D[c] : kCopies3D[c];
            const int n_copies = dim == 2 ? 3 : 7;
            for (int r = 0; r < n_copies && copies[r].node >= 0; ++r)
                nd[copies[r].node] = node[(size_t)(first + copies[r].sibling) * nch + copies[r].from];
        }
    }
    // one parent of _refine_cells / _refine_uniform: children get the next 2^d indices
    void refine(int32_t p) {
        const int32_t first = (int32_t)n_cells();
        const double q = 0.25 * width / (double)((int64_t)1 << level[p]);
        for (int c = 0; c < nch; ++c) {
            parent.push_back(p);
            children.push_back(kLeaf);
            level.push_back(level[p] + 1);
            for (int s = 0; s < nnb; ++s) nb.push_back(kNone);
            for (int j = 0; j < nch; ++j) node.push_back(0);
            for (int a = 0; a < dim; ++a) center.push_back(center[(size_t)p * dim + a] + (double)kChildDir[c][a] * q);
        }
        children[p] = first;
        assign_neighbors(p);
        assign_indices(p);
    }
    // _remove_invalid_cells (:721-731): children = [], and every cell the removed one points to forgets it
    void mark_invalid(int32_t cell) {
        children[cell] = kEmpty;
        for (int s = 0; s < nnb; ++s) {
            const int32_t q = nb[(size_t)cell * nnb + s];
            if (q == kNone) continue;
            int32_t* qnb = &nb[(size_t)q * nnb];
            for (int t = 0; t < nnb; ++t)
                if (qnb[t] == cell) qnb[t] = kNone;
        }
    }
};

}  // namespace s3

using s3::Topology;

// The handle. Updates (refine / refresh / mark_invalid) are applied in call order. In asynchronous mode they are queued
// and applied by a native worker thread, so the replay runs next to the device work of the refinement loop without
// touching the Python interpreter; every read (check_nb, cell, final, counts) first waits for the queue to drain.
struct s3_topo {
    Topology t;
    enum Kind { kRefine, kRefreshCells, kRefreshParents, kInvalid };
    struct Job { Kind kind; std::vector<int64_t> cells; };
    bool async = false;
    std::thread worker;
    std::mutex mu;
    std::condition_variable cv_work, cv_idle;
    std::deque<Job> queue;
    bool busy = false, stop = false;
    std::string error;                     // first failure of a queued update

    int apply(const Job& job, std::string* err) {
        char buf[160];
        for (int64_t c : job.cells) {
            if (c < 0 || c >= t.n_cells()) {
                snprintf(buf, sizeof(buf), "topology update: cell %lld out of range", (long long)c);
                *err = buf;
                return S3_ERR_INVALID;
            }
            switch (job.kind) {
                case kRefine:
                    if (t.children[c] >= 0 || t.n_cells() + t.nch >= ((int64_t)1 << 31)) {
                        snprintf(buf, sizeof(buf), "s3_topo_refine: cell %lld already has children (or too many cells)",
                                 (long long)c);
                        *err = buf;
                        return S3_ERR_INVALID;
                    }
                    t.refine((int32_t)c);
                    break;
                case kRefreshCells:
                    if (t.parent[c] != s3::kNone) t.assign_neighbors(t.parent[c]);
                    break;
                case kRefreshParents:
                    t.assign_neighbors((int32_t)c);
                    break;
                case kInvalid:
                    t.mark_invalid((int32_t)c);
                    break;
            }
        }
        return S3_OK;
    }
    void run() {
        std::unique_lock<std::mutex> lock(mu);
        while (true) {
            cv_work.wait(lock, [&] { return stop || !queue.empty(); });
            if (queue.empty()) return;             // stop requested and nothing left
            Job job = std::move(queue.front());
            queue.pop_front();
            busy = true;
            lock.unlock();
            std::string err;
            if (error.empty()) apply(job, &err);   // after a failure the remaining updates are dropped
            lock.lock();
            if (!err.empty() && error.empty()) error = err;
            busy = false;
            if (queue.empty()) cv_idle.notify_all();
        }
    }
    // wait until every queued update has been applied; reports a failure of the queue once
    int drain() {
        if (!async) return S3_OK;
        std::unique_lock<std::mutex> lock(mu);
        cv_idle.wait(lock, [&] { return queue.empty() && !busy; });
        if (!error.empty()) {
            s3::set_error("%s", error.c_str());
            return S3_ERR_INVALID;
        }
        return S3_OK;
    }
    int submit(Kind kind, const int64_t* cells, int64_t n) {
        Job job{kind, std::vector<int64_t>(cells, cells + n)};
        if (!async) {
            std::string err;
            const int rc = apply(job, &err);
            if (rc != S3_OK) s3::set_error("%s", err.c_str());
            return rc;
        }
        {
            std::lock_guard<std::mutex> lock(mu);
            queue.push_back(std::move(job));
        }
        cv_work.notify_one();
        return S3_OK;
    }
    ~s3_topo() {
        if (async) {
            {
                std::lock_guard<std::mutex> lock(mu);
                stop = true;
            }
            cv_work.notify_one();
            if (worker.joinable()) worker.join();
        }
    }
};

extern "C" int s3_topo_create(int dim, const double* root_center, double width, int async, s3_topo_t** out) {
    S3_REQUIRE(out && root_center, "s3_topo_create: NULL argument");
    S3_REQUIRE(dim == 2 || dim == 3, "s3_topo_create: dim must be 2 or 3");
    s3_topo* h = new s3_topo();
    Topology& t = h->t;
    t.dim = dim;
    t.nch = 1 << dim;
    t.nnb = dim == 2 ? 8 : 26;
    t.width = width;
    t.build_table();
    // _create_first_cell (:338-397): root with 2^d nodes at centre +- width/2, no neighbours
    t.parent.push_back(s3::kNone);
    t.children.push_back(s3::kLeaf);
    t.level.push_back(0);
    for (int s = 0; s < t.nnb; ++s) t.nb.push_back(s3::kNone);
    for (int a = 0; a < dim; ++a) t.center.push_back(root_center[a]);
    for (int j = 0; j < t.nch; ++j) {
        t.node.push_back(j);
        for (int a = 0; a < dim; ++a) t.nodes.push_back(root_center[a] + (double)s3::kChildDir[j][a] * 0.5 * width);
    }
    if (async) {
        h->async = true;
        h->worker = std::thread([h] { h->run(); });
    }
    *out = h;
    return S3_OK;
}

extern "C" int s3_topo_free(s3_topo_t* h) {
    delete h;
    return S3_OK;
}

extern "C" int s3_topo_sync(s3_topo_t* h) {
    S3_REQUIRE(h, "s3_topo_sync: NULL argument");
    return h->drain();
}

extern "C" int64_t s3_topo_n_cells(s3_topo_t* h) {
    if (!h) return 0;
    h->drain();
    return h->t.n_cells();
}
extern "C" int64_t s3_topo_n_nodes(s3_topo_t* h) {
    if (!h) return 0;
    h->drain();
    return h->t.n_nodes();
}

extern "C" int s3_topo_refine(s3_topo_t* h, const int64_t* parents, int64_t n) {
    S3_REQUIRE(h && (parents || n == 0), "s3_topo_refine: NULL argument");
    return h->submit(s3_topo::kRefine, parents, n);
}

// cell.parent.children = _assign_neighbors(cell.parent, children=cell.parent.children) for every cell of the list
// (:611, :489-490, :826); of_parents != 0: the entries are the parents themselves (_refine_uniform, :547-549)
extern "C" int s3_topo_refresh(s3_topo_t* h, const int64_t* cells, int64_t n, int of_parents) {
    S3_REQUIRE(h && (cells || n == 0), "s3_topo_refresh: NULL argument");
    return h->submit(of_parents ? s3_topo::kRefreshParents : s3_topo::kRefreshCells, cells, n);
}

extern "C" int s3_topo_mark_invalid(s3_topo_t* h, const int64_t* cells, int64_t n) {
    S3_REQUIRE(h && (cells || n == 0), "s3_topo_mark_invalid: NULL argument");
    return h->submit(s3_topo::kInvalid, cells, n);
}

// _check_nb (:447-464): neighbours (slot order) that exist, are leaves and have a lower level; returns the count
extern "C" int64_t s3_topo_check_nb(s3_topo_t* h, int64_t cell, int64_t* out) {
    if (!h || !out || h->drain() != S3_OK || cell < 0 || cell >= h->t.n_cells()) return -1;
    const Topology& t = h->t;
    int64_t n = 0;
    for (int s = 0; s < t.nnb; ++s) {
        const int32_t q = t.nb[(size_t)cell * t.nnb + s];
        if (q != s3::kNone && t.children[q] == s3::kLeaf && t.level[q] < t.level[cell]) out[n++] = q;
    }
    return n;
}

// raw views for tests: neighbour pointers [nnb] and node ids [2^d] of one cell
extern "C" int s3_topo_cell(s3_topo_t* h, int64_t cell, int32_t* nb_out, int32_t* node_out, int32_t* state_out) {
    S3_REQUIRE(h, "s3_topo_cell: NULL argument");
    S3_TRY(h->drain());
    S3_REQUIRE(cell >= 0 && cell < h->t.n_cells(), "s3_topo_cell: cell out of range");
    const Topology& t = h->t;
    if (nb_out) memcpy(nb_out, &t.nb[(size_t)cell * t.nnb], sizeof(int32_t) * t.nnb);
    if (node_out) memcpy(node_out, &t.node[(size_t)cell * t.nch], sizeof(int32_t) * t.nch);
    if (state_out) { state_out[0] = t.parent[cell]; state_out[1] = t.children[cell]; state_out[2] = t.level[cell]; }
    return S3_OK;
}

// _resort_nodes_and_indices_of_grid + renumber_node_indices_parallel: faces of the leaf cells in cell-list order with
// the unused node ids squeezed out. Two calls: with faces == NULL it only returns the sizes.
extern "C" int s3_topo_final(s3_topo_t* h, int64_t* n_leaf_out, int64_t* n_vertices_out, int32_t* faces,
                             double* vertices, double* centers_by_index) {
    S3_REQUIRE(h && n_leaf_out && n_vertices_out, "s3_topo_final: NULL argument");
    S3_TRY(h->drain());
    const Topology& t = h->t;
    const int64_t nc = t.n_cells(), nn = t.n_nodes();
    std::vector<uint8_t> used((size_t)nn, 0);
    int64_t n_leaf = 0;
    int32_t lo = INT32_MAX, hi = -1;
    for (int64_t c = 0; c < nc; ++c) {
        if (t.children[c] != s3::kLeaf) continue;
        ++n_leaf;
        for (int j = 0; j < t.nch; ++j) {
            const int32_t id = t.node[(size_t)c * t.nch + j];
            used[id] = 1;
            lo = std::min(lo, id);
            hi = std::max(hi, id);
        }
    }
    // unused = ids of the initial cell and of [min, max] that no leaf references (:752-760); everything else stays
    std::vector<int32_t> mapping((size_t)nn, -1);
    int64_t counter = 0;
    for (int64_t i = 0; i < nn; ++i) {
        const bool candidate = i < t.nch || (n_leaf > 0 && i >= lo && i <= hi);
        if (candidate && !used[i]) continue;
        mapping[i] = (int32_t)counter++;
    }
    *n_leaf_out = n_leaf;
    *n_vertices_out = counter;
    if (!faces) return S3_OK;
    S3_REQUIRE(vertices, "s3_topo_final: vertices is NULL");
    for (int64_t i = 0; i < nn; ++i)
        if (mapping[i] >= 0) memcpy(vertices + (size_t)mapping[i] * t.dim, &t.nodes[(size_t)i * t.dim], sizeof(double) * t.dim);
    int64_t row = 0;
    for (int64_t c = 0; c < nc; ++c) {
        if (t.children[c] != s3::kLeaf) continue;
        for (int j = 0; j < t.nch; ++j) faces[row * t.nch + j] = mapping[t.node[(size_t)c * t.nch + j]];
        ++row;
    }
    if (centers_by_index) memcpy(centers_by_index, t.center.data(), sizeof(double) * t.center.size());
    return S3_OK;
}

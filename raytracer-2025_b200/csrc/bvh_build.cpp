// bvh_build.cpp — binned-SAH BVH2 builder (host, OpenMP tasks).
//
// The reference's BVH is a median split on the longest axis (bvh.rs:16-46); only its
// *semantics* (closest hit, tie order) are kept — the tie order is carried by per-primitive
// ranks, so the acceleration structure is free to be a surface-area-heuristic tree laid out
// for the GPU: 64-byte nodes that hold both children's boxes.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "compile.h"

namespace rt {

namespace {

// The builder permutes RECORDS (box + primitive index, 28 bytes), not an index list: every pass over a range then streams through
// memory instead of gathering 24-byte boxes at random - on 10^7 primitives the passes of the upper levels were bound by the
// latency of those gathers.
struct Rec {
    float lo[3], hi[3];
    uint32_t id;
};
struct Ctx {
    Rec* recs;
    Rec* scratch;  // as long as recs: the chunked partition of a big range copies through it
    Node* nodes;
    std::atomic<uint32_t>* node_count;
    std::atomic<uint32_t>* max_depth;
    uint32_t base;
};

struct Box3 {
    float lo[3], hi[3];
    void reset() {
        for (int a = 0; a < 3; a++) lo[a] = INFINITY, hi[a] = -INFINITY;
    }
    void grow(const float* l, const float* h) {
        for (int a = 0; a < 3; a++) lo[a] = std::min(lo[a], l[a]), hi[a] = std::max(hi[a], h[a]);
    }
    float half_area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (!(dx >= 0.f) || !(dy >= 0.f) || !(dz >= 0.f)) return 0.f;
        return dx * dy + dy * dz + dz * dx;
    }
};

constexpr int NBINS = 16;
// SAH constants; RT2025_SAH_CPRIM / RT2025_SAH_LEAF override them for tuning experiments
float C_TRAV = 1.0f, C_PRIM = 4.0f;  // a binary64 primitive test costs several binary32 slab tests
uint32_t LEAF_TARGET = 2;  // SAH may stop at <= this many primitives
constexpr uint32_t TASK_MIN = 8192;

inline float centroid(const Rec& b, int a) { return 0.5f * (b.lo[a] + b.hi[a]); }

uint32_t make_leaf(const Ctx& c, uint32_t begin, uint32_t end) {
    return LEAF_FLAG | ((c.base + begin) << 3) | (end - begin - 1);
}

// Big ranges (the top of a 10^7..10^8-primitive tree) are processed by the whole team: the range is cut into chunks that run as
// tasks (the caller is already inside the build's task region, so a nested parallel loop would run on one thread), the partial
// results are merged by the caller.  With one task per node the first four levels ran on 1, 2, 4 and 8 cores.
constexpr uint32_t PAR_MIN = 1u << 21;
constexpr int PAR_CHUNKS = 64;
struct Bins {
    uint32_t cnt[3][NBINS];
    Box3 bb[3][NBINS];
    void reset() {
        for (int a = 0; a < 3; a++)
            for (int k = 0; k < NBINS; k++) cnt[a][k] = 0, bb[a][k].reset();
    }
};
inline int bin_of(float ce, float lo, float scale) { return std::min(NBINS - 1, std::max(0, (int)((ce - lo) * scale))); }

// returns the child reference for [begin,end) and its bounds
uint32_t build_range(const Ctx& c, uint32_t begin, uint32_t end, uint32_t depth, Box3& bounds) {
    const uint32_t n = end - begin;
    const bool par = n >= PAR_MIN;
    auto chunk_at = [&](int k) { return begin + (uint32_t)((uint64_t)n * (uint64_t)k / PAR_CHUNKS); };
    bounds.reset();
    Box3 cb;
    cb.reset();
    if (par) {
        std::vector<Box3> pb(PAR_CHUNKS), pc(PAR_CHUNKS);
#pragma omp taskloop grainsize(1) default(shared)
        for (int k = 0; k < PAR_CHUNKS; k++) {
            Box3 b0, c0;
            b0.reset(), c0.reset();
            for (uint32_t i = chunk_at(k); i < chunk_at(k + 1); i++) {
                const Rec& b = c.recs[i];
                b0.grow(b.lo, b.hi);
                float ce[3] = {centroid(b, 0), centroid(b, 1), centroid(b, 2)};
                c0.grow(ce, ce);
            }
            pb[k] = b0, pc[k] = c0;
        }
        for (int k = 0; k < PAR_CHUNKS; k++) bounds.grow(pb[k].lo, pb[k].hi), cb.grow(pc[k].lo, pc[k].hi);
    } else {
        for (uint32_t i = begin; i < end; i++) {
            const Rec& b = c.recs[i];
            bounds.grow(b.lo, b.hi);
            float ce[3] = {centroid(b, 0), centroid(b, 1), centroid(b, 2)};
            cb.grow(ce, ce);
        }
    }
    uint32_t d = c.max_depth->load(std::memory_order_relaxed);
    while (depth > d && !c.max_depth->compare_exchange_weak(d, depth)) {
    }
    if (n == 1) return make_leaf(c, begin, end);

    uint32_t mid = 0;
    bool have_split = false;
    if (depth < 32) {
        float best_cost = INFINITY;
        int best_axis = -1, best_bin = -1;
        // the three axes are binned in ONE pass over the range (the boxes are read through the index list, i.e. at random)
        float ext3[3], scale3[3];
        for (int a = 0; a < 3; a++) ext3[a] = cb.hi[a] - cb.lo[a], scale3[a] = ext3[a] > 0.f ? NBINS / ext3[a] : 0.f;
        // Most nodes of a tree hold a handful of primitives, and for those the 3 x 16 bins (1.3 KB to clear and to sweep) cost far more
        // than the primitives: a small range is sorted by bin index instead and only the boundaries between OCCUPIED bins are
        // evaluated.  Same candidates, same costs (unions and counts do not depend on the order), same first minimum: same tree.
        constexpr uint32_t SMALL_N = 12;
        if (n <= SMALL_N) {
            for (int a = 0; a < 3; a++) {
                if (!(ext3[a] > 0.f)) continue;
                int kb[SMALL_N];
                uint8_t ord[SMALL_N];
                for (uint32_t i = 0; i < n; i++) {
                    kb[i] = bin_of(centroid(c.recs[begin + i], a), cb.lo[a], scale3[a]);
                    uint32_t j = i;  // insertion sort by bin index
                    while (j > 0 && kb[ord[j - 1]] > kb[i]) ord[j] = ord[j - 1], j--;
                    ord[j] = (uint8_t)i;
                }
                float right_area[SMALL_N + 1];
                Box3 acc;
                acc.reset();
                for (uint32_t j = n; j-- > 1;) {
                    const Rec& b = c.recs[begin + ord[j]];
                    acc.grow(b.lo, b.hi);
                    right_area[j] = acc.half_area();
                }
                acc.reset();
                for (uint32_t j = 0; j + 1 < n; j++) {
                    const Rec& b = c.recs[begin + ord[j]];
                    acc.grow(b.lo, b.hi);
                    if (kb[ord[j]] == kb[ord[j + 1]]) continue;  // not a bin boundary
                    const float cost = acc.half_area() * (float)(j + 1) + right_area[j + 1] * (float)(n - j - 1);
                    if (cost < best_cost) best_cost = cost, best_axis = a, best_bin = kb[ord[j]];
                }
            }
        }
        Bins bins;
        if (n > SMALL_N) bins.reset();
        auto bin_range = [&](uint32_t i0, uint32_t i1, Bins& out) {
            for (uint32_t i = i0; i < i1; i++) {
                const Rec& b = c.recs[i];
                for (int a = 0; a < 3; a++) {
                    if (!(ext3[a] > 0.f)) continue;
                    const int k = bin_of(centroid(b, a), cb.lo[a], scale3[a]);
                    out.cnt[a][k]++;
                    out.bb[a][k].grow(b.lo, b.hi);
                }
            }
        };
        if (par) {
            std::vector<Bins> part(PAR_CHUNKS);
#pragma omp taskloop grainsize(1) default(shared)
            for (int k = 0; k < PAR_CHUNKS; k++) {
                part[k].reset();
                bin_range(chunk_at(k), chunk_at(k + 1), part[k]);
            }
            for (int k = 0; k < PAR_CHUNKS; k++)
                for (int a = 0; a < 3; a++)
                    for (int q = 0; q < NBINS; q++) bins.cnt[a][q] += part[k].cnt[a][q], bins.bb[a][q].grow(part[k].bb[a][q].lo, part[k].bb[a][q].hi);
        } else if (n > SMALL_N) {
            bin_range(begin, end, bins);
        }
        for (int a = 0; a < 3 && n > SMALL_N; a++) {
            if (!(ext3[a] > 0.f)) continue;
            const uint32_t* cnt = bins.cnt[a];
            const Box3* bb = bins.bb[a];
            float right_area[NBINS];
            uint32_t right_cnt[NBINS];
            Box3 acc;
            acc.reset();
            uint32_t ac = 0;
            for (int k = NBINS - 1; k > 0; k--) {
                acc.grow(bb[k].lo, bb[k].hi);
                ac += cnt[k];
                right_area[k] = acc.half_area();
                right_cnt[k] = ac;
            }
            acc.reset();
            ac = 0;
            for (int k = 0; k < NBINS - 1; k++) {
                acc.grow(bb[k].lo, bb[k].hi);
                ac += cnt[k];
                if (ac == 0 || right_cnt[k + 1] == 0) continue;
                float cost = acc.half_area() * ac + right_area[k + 1] * right_cnt[k + 1];
                if (cost < best_cost) best_cost = cost, best_axis = a, best_bin = k;
            }
        }
        if (best_axis >= 0) {
            float area = bounds.half_area();
            float split_cost = C_TRAV + (area > 0.f ? best_cost / area : (float)n) * C_PRIM;
            float leaf_cost = (float)n * C_PRIM;
            if (n <= LEAF_TARGET && leaf_cost <= split_cost) return make_leaf(c, begin, end);
            const float scale = scale3[best_axis], lo = cb.lo[best_axis];
            auto goes_left = [&](const Rec& b) { return bin_of(centroid(b, best_axis), lo, scale) <= best_bin; };
            if (par) {
                // chunked partition through the scratch copy: count per chunk (while copying), then every chunk writes its two parts
                // at their offsets
                const Rec* src = c.scratch;
                std::vector<uint32_t> n_left(PAR_CHUNKS + 1, 0);
#pragma omp taskloop grainsize(1) default(shared)
                for (int k = 0; k < PAR_CHUNKS; k++) {
                    uint32_t nl = 0;
                    for (uint32_t i = chunk_at(k); i < chunk_at(k + 1); i++) {
                        c.scratch[i] = c.recs[i];
                        nl += goes_left(c.recs[i]);
                    }
                    n_left[k + 1] = nl;
                }
                for (int k = 0; k < PAR_CHUNKS; k++) n_left[k + 1] += n_left[k];
                mid = begin + n_left[PAR_CHUNKS];
#pragma omp taskloop grainsize(1) default(shared)
                for (int k = 0; k < PAR_CHUNKS; k++) {
                    Rec* l = c.recs + begin + n_left[k];
                    Rec* r = c.recs + mid + (chunk_at(k) - begin - n_left[k]);
                    for (uint32_t i = chunk_at(k); i < chunk_at(k + 1); i++) {
                        if (goes_left(src[i]))
                            *l++ = src[i];
                        else
                            *r++ = src[i];
                    }
                }
            } else {
                Rec* p = std::partition(c.recs + begin, c.recs + end, goes_left);
                mid = (uint32_t)(p - c.recs);
            }
            have_split = mid > begin && mid < end;
        }
    }
    if (!have_split) {
        if (n <= MAX_LEAF_PRIMS && depth < 32) return make_leaf(c, begin, end);
        // identical centroids or depth guard: object median on the widest centroid axis
        int a = 0;
        for (int k = 1; k < 3; k++)
            if (cb.hi[k] - cb.lo[k] > cb.hi[a] - cb.lo[a]) a = k;
        mid = begin + n / 2;
        std::nth_element(c.recs + begin, c.recs + mid, c.recs + end, [&](const Rec& x, const Rec& y) { return centroid(x, a) < centroid(y, a); });
    }

    uint32_t me = c.node_count->fetch_add(1);
    Box3 bl, br;
    uint32_t cl, cr;
    if (n >= TASK_MIN) {
#pragma omp task shared(bl, cl) firstprivate(begin, mid, depth)
        cl = build_range(c, begin, mid, depth + 1, bl);
        cr = build_range(c, mid, end, depth + 1, br);
#pragma omp taskwait
    } else {
        cl = build_range(c, begin, mid, depth + 1, bl);
        cr = build_range(c, mid, end, depth + 1, br);
    }
    Node& nd = c.nodes[me];
    for (int a = 0; a < 3; a++) nd.lo0[a] = bl.lo[a], nd.hi0[a] = bl.hi[a], nd.lo1[a] = br.lo[a], nd.hi1[a] = br.hi[a];
    nd.child0 = cl;
    nd.child1 = cr;
    nd.pad0 = nd.pad1 = 0;
    return me;
}

}  // namespace

uint32_t build_bvh(const RawVec<BuildBox>& boxes, uint32_t first_prim_base, RawVec<Node>& nodes,
                   RawVec<uint32_t>& order, uint32_t& depth_out) {
    if (const char* e = getenv("RT2025_SAH_CPRIM")) C_PRIM = (float)atof(e);
    if (const char* e = getenv("RT2025_SAH_LEAF")) LEAF_TARGET = (uint32_t)atoi(e);
    const uint32_t n = (uint32_t)boxes.size();
    order.resize(n);
    if (n == 0) return INVALID_REF;
    RawVec<Rec> recs(n), scratch(n >= PAR_MIN ? n : 0);
#pragma omp parallel for schedule(static) if (n > 65536)
    for (uint32_t i = 0; i < n; i++) {
        for (int a = 0; a < 3; a++) recs[i].lo[a] = boxes[i].lo[a], recs[i].hi[a] = boxes[i].hi[a];
        recs[i].id = i;
    }
    const uint32_t node_base = (uint32_t)nodes.size();
    nodes.resize(node_base + (n > 1 ? n - 1 : 0));
    std::atomic<uint32_t> count{node_base}, depth{0};
    Ctx c{recs.data(), scratch.data(), nodes.data(), &count, &depth, first_prim_base};
    Box3 b;
    uint32_t root = 0;
#pragma omp parallel if (n > 32768)  // no team for book-sized scenes (the build itself is ~1 us per primitive)
#pragma omp single
    root = build_range(c, 0, n, 1, b);
#pragma omp parallel for schedule(static) if (n > 65536)
    for (uint32_t i = 0; i < n; i++) order[i] = recs[i].id;
    nodes.resize(count.load());
    depth_out = std::max(depth_out, depth.load());
    return root;
}

// ---- four-wide collapse -------------------------------------------------------------------------
// Each wide node starts from the two children of a binary node and repeatedly opens the interior child with
// the largest surface area until it has four entries (or only leaves are left): the standard SAH-guided
// collapse.  Boxes are copied, never recomputed, so they stay the conservative bounds of the binary tree.
uint32_t collapse_bvh4(const RawVec<Node>& nodes, uint32_t root, RawVec<Node4>& out, uint32_t& depth_out) {
    depth_out = 0;
    if (root == INVALID_REF || (root & LEAF_FLAG)) return root;
    struct Entry {
        uint32_t ref;
        float b[6];
    };
    auto area = [](const Entry& e) {
        const float dx = e.b[3] - e.b[0], dy = e.b[4] - e.b[1], dz = e.b[5] - e.b[2];
        return dx * dy + dy * dz + dz * dx;
    };
    auto children_of = [&](uint32_t n2, Entry* two) {
        const Node& nd = nodes[n2];
        two[0].ref = nd.child0, two[1].ref = nd.child1;
        for (int k = 0; k < 3; k++) {
            two[0].b[k] = nd.lo0[k], two[0].b[3 + k] = nd.hi0[k];
            two[1].b[k] = nd.lo1[k], two[1].b[3 + k] = nd.hi1[k];
        }
    };
    // level by level: every node of a level is expanded independently (all host cores), the interior children of
    // the level get consecutive indices from a prefix sum, so the array comes out breadth first
    struct Pending {
        uint32_t n2, n4;
    };
    auto expand = [&](uint32_t n2, Entry* e) {  // -> number of entries
        int n = 0;
        Entry two[2];
        children_of(n2, two);
        for (int k = 0; k < 2; k++)
            if (two[k].ref != INVALID_REF) e[n++] = two[k];
        while (n < 4) {
            int best = -1;
            float best_area = -1.f;
            for (int k = 0; k < n; k++)
                if (!(e[k].ref & LEAF_FLAG) && area(e[k]) > best_area) best = k, best_area = area(e[k]);
            if (best < 0) break;
            children_of(e[best].ref, two);
            int m = 0;
            for (int k = 0; k < 2; k++)
                if (two[k].ref != INVALID_REF) {
                    if (m == 0) e[best] = two[k]; else e[n++] = two[k];
                    m++;
                }
            if (m == 0) {  // a binary node without children: drop the entry
                e[best] = e[n - 1];
                n--;
            }
        }
        return n;
    };
    std::vector<Pending> level{{root, 0u}}, next;
    out.emplace_back();
    std::vector<uint32_t> first_child;  // per node of the level: index of its first interior child (after the scan)
    while (!level.empty()) {
        depth_out++;
        const size_t m = level.size();
        first_child.assign(m + 1, 0);
#pragma omp parallel for schedule(static) if (m > 4096)
        for (size_t i = 0; i < m; i++) {
            Entry e[4];
            const int n = expand(level[i].n2, e);
            uint32_t inner = 0;
            for (int k = 0; k < n; k++) inner += !(e[k].ref & LEAF_FLAG);
            first_child[i + 1] = inner;
        }
        for (size_t i = 0; i < m; i++) first_child[i + 1] += first_child[i];
        const uint32_t base = (uint32_t)out.size();
        out.resize((size_t)base + first_child[m]);
        next.resize(first_child[m]);
#pragma omp parallel for schedule(static) if (m > 4096)
        for (size_t i = 0; i < m; i++) {
            Entry e[4];
            const int n = expand(level[i].n2, e);
            uint32_t slot = first_child[i];
            Node4 w;
            for (int k = 0; k < 4; k++) {
                if (k < n) {
                    for (int q = 0; q < 6; q++) w.box[k][q] = e[k].b[q];
                    if (e[k].ref & LEAF_FLAG) {
                        w.child[k] = e[k].ref;
                    } else {
                        w.child[k] = base + slot;
                        next[slot] = Pending{e[k].ref, base + slot};
                        slot++;
                    }
                } else {
                    for (int q = 0; q < 3; q++) w.box[k][q] = INFINITY, w.box[k][3 + q] = -INFINITY;
                    w.child[k] = INVALID_REF;
                }
                w.pad[k] = 0;
            }
            out[level[i].n4] = w;
        }
        level.swap(next);
        next.clear();
    }
    return 0;  // the root is the first wide node
}

}  // namespace rt

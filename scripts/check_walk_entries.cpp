// CPU check of the scene compiler's entry leaves for optically thick media (csrc/compile.cpp, Medium::entry): the leaves the
// recursive box-overlap walk collects must be ALL the leaves of the world tree whose box meets the (padded) box of the
// medium's boundary - found here by a flat scan over every node, which also checks that a parent's box contains its
// children's (the property the walk relies on) - and a thin medium must not be flagged.  book2_final over several seeds.
// Host only, no CUDA:  make check_walk_entries && build/check_walk_entries
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <set>
#include <string>

#include "compile.h"

struct rth_scene;
extern "C" rth_scene* rth_scene_named(const char* name, uint64_t seed, const double* params, int n_params);
extern "C" const rt_scene_desc* rth_scene_desc(const rth_scene* s);

using namespace rt;

int main() {
    int bad = 0, thick_seen = 0;
    for (uint64_t seed = 1; seed <= 12; seed++) {
        const double params[3] = {64, 4, 40};
        rth_scene* s = rth_scene_named("book2_final", seed, params, 3);
        if (!s) return 2;
        CompiledScene cs;
        std::string err;
        if (compile_scene(*rth_scene_desc(s), 0, cs, err) != 0) {
            printf("compile failed: %s\n", err.c_str());
            return 2;
        }
        for (const Medium& m : cs.media) {
            const double radius = std::fabs(cs.geom[m.single_sphere].d[6]);
            const double optical_radius = radius / std::fabs(m.neg_inv_density);
            const bool thick = (m.flags & MEDIUM_THICK) != 0;
            if (thick != (optical_radius >= 1.0)) printf("seed %llu: thick flag %d at optical radius %.3f\n", (unsigned long long)seed, (int)thick, optical_radius), bad++;
            if (!thick) continue;
            thick_seen++;
            if (m.n_entry == MEDIUM_NO_ENTRIES) continue;
            const double* g = cs.geom[m.single_sphere].d;
            // every leaf of the world tree whose box meets the ball must be an entry (flat scan, exact ball distance, no padding:
            // the compiler's padded test may add leaves, never drop one)
            std::set<uint32_t> entries(m.entry, m.entry + m.n_entry);
            for (uint32_t e : entries)
                if (!(e & LEAF_FLAG)) printf("seed %llu: entry %08x is not a leaf\n", (unsigned long long)seed, e), bad++;
            auto meets = [&](const float* lo, const float* hi) {
                double d2 = 0.0;
                for (int k = 0; k < 3; k++) {
                    const double c = g[k], dk = c < lo[k] ? lo[k] - c : (c > hi[k] ? c - hi[k] : 0.0);
                    d2 += dk * dk;
                }
                return d2 <= radius * radius;
            };
            // nodes reachable from the world root only (the media groups have trees of their own in the same array)
            std::vector<uint32_t> todo{cs.world_root};
            while (!todo.empty()) {
                const uint32_t ref = todo.back();
                todo.pop_back();
                if (ref == INVALID_REF || (ref & LEAF_FLAG)) continue;
                const Node& nd = cs.nodes[ref];
                const float* lo[2] = {nd.lo0, nd.lo1};
                const float* hi[2] = {nd.hi0, nd.hi1};
                const uint32_t child[2] = {nd.child0, nd.child1};
                for (int c = 0; c < 2; c++) {
                    if (child[c] == INVALID_REF) continue;
                    if (child[c] & LEAF_FLAG) {
                        if (meets(lo[c], hi[c]) && !entries.count(child[c]))
                            printf("seed %llu: leaf %08x meets the ball but is no entry\n", (unsigned long long)seed, child[c]), bad++;
                    } else {
                        const Node& ch = cs.nodes[child[c]];  // containment: the walk prunes on the parent's copy of the box
                        for (int k = 0; k < 3; k++) {
                            const float clo = std::min(ch.child0 != INVALID_REF ? ch.lo0[k] : INFINITY, ch.child1 != INVALID_REF ? ch.lo1[k] : INFINITY);
                            const float chi = std::max(ch.child0 != INVALID_REF ? ch.hi0[k] : -INFINITY, ch.child1 != INVALID_REF ? ch.hi1[k] : -INFINITY);
                            if (clo < lo[c][k] || chi > hi[c][k]) printf("seed %llu: node %u not inside its parent's box\n", (unsigned long long)seed, child[c]), bad++;
                        }
                        todo.push_back(child[c]);
                    }
                }
            }
        }
    }
    if (thick_seen == 0) printf("no thick medium seen\n"), bad++;
    printf("check_walk_entries: %d thick media checked, %d problems\n", thick_seen, bad);
    return bad ? 1 : 0;
}

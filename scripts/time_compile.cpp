// Host-only timing of the scene compile step (no CUDA needed):
//   g++ -O3 -fopenmp -std=c++17 -Iinclude -Iraytracer-2025_b200/csrc -Iraytracer-2025_b200/host -I/usr/local/cuda/include \
//       scripts/time_compile.cpp raytracer-2025_b200/csrc/compile.cpp raytracer-2025_b200/csrc/bvh_build.cpp \
//       raytracer-2025_b200/host/host_capi.cpp -o /tmp/time_compile && RT2025_TIMING=1 /tmp/time_compile tri_soup 2000000
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <string>
#include <vector>

#include "compile.h"

struct rth_scene;
extern "C" rth_scene* rth_scene_named(const char* name, uint64_t seed, const double* params, int n_params);
extern "C" const rt_scene_desc* rth_scene_desc(const rth_scene* s);

int main(int argc, char** argv) {
    const char* name = argc > 1 ? argv[1] : "tri_soup";
    double n = argc > 2 ? atof(argv[2]) : 1e6;
    auto t0 = std::chrono::steady_clock::now();
    rth_scene* s = rth_scene_named(name, 5, &n, 1);
    if (!s) return 1;
    auto t1 = std::chrono::steady_clock::now();
    rt::CompiledScene cs;
    std::string err;
    int rc = rt::compile_scene(*rth_scene_desc(s), argc > 3 ? (uint32_t)atoi(argv[3]) : 0, cs, err);
    auto t2 = std::chrono::steady_clock::now();
    printf("%s N=%.0f: host scene %.2f s, compile %.2f s (rc %d %s), %zu nodes, depth %u\n", name, n,
           std::chrono::duration<double>(t1 - t0).count(), std::chrono::duration<double>(t2 - t1).count(), rc, err.c_str(),
           cs.nodes.size(), cs.bvh_depth);
    if (rc == 0 && getenv("RT2025_TREE_HASH")) {  // FNV-1a over the node array and the primitive order: equal trees, equal hashes
        auto fnv = [](const void* p, size_t n, uint64_t h) {
            const unsigned char* b = (const unsigned char*)p;
            for (size_t i = 0; i < n; i++) h = (h ^ b[i]) * 1099511628211ull;
            return h;
        };
        uint64_t h = fnv(cs.nodes.data(), cs.nodes.size() * sizeof(rt::Node), 1469598103934665603ull);
        h = fnv(cs.meta.data(), cs.meta.size() * sizeof(rt::PrimMeta), h);
        printf("tree hash %016llx\n", (unsigned long long)h);
    }
    if (rc == 0 && getenv("RT2025_VERIFY_TREE")) {
        // every world primitive in exactly one leaf, every child box inside its parent's copy of it
        std::vector<unsigned char> seen(cs.geom.size(), 0);
        std::vector<uint32_t> todo{cs.world_root};
        size_t bad = 0, leaves = 0;
        auto visit_leaf = [&](uint32_t ref) {
            const uint32_t first = (ref & ~rt::LEAF_FLAG) >> 3, count = (ref & 7u) + 1;
            for (uint32_t i = first; i < first + count; i++) bad += i >= seen.size() || seen[i]++;
            leaves++;
        };
        while (!todo.empty()) {
            const uint32_t ref = todo.back();
            todo.pop_back();
            if (ref == rt::INVALID_REF) continue;
            if (ref & rt::LEAF_FLAG) {
                visit_leaf(ref);
                continue;
            }
            const rt::Node& nd = cs.nodes[ref];
            const float* lo[2] = {nd.lo0, nd.lo1};
            const float* hi[2] = {nd.hi0, nd.hi1};
            const uint32_t child[2] = {nd.child0, nd.child1};
            for (int c = 0; c < 2; c++) {
                if (child[c] != rt::INVALID_REF && !(child[c] & rt::LEAF_FLAG)) {
                    const rt::Node& ch = cs.nodes[child[c]];
                    for (int k = 0; k < 3; k++)
                        bad += std::min(ch.lo0[k], ch.lo1[k]) < lo[c][k] || std::max(ch.hi0[k], ch.hi1[k]) > hi[c][k];
                }
                todo.push_back(child[c]);
            }
        }
        size_t missing = 0;
        for (unsigned char v : seen) missing += v != 1;
        printf("tree check: %zu leaves, %zu primitives not in exactly one leaf, %zu other problems\n", leaves, missing, bad);
        if (missing || bad) return 3;
    }
    return rc;
}

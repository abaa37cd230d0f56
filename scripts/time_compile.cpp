// Host-only timing of the scene compile step (no CUDA needed):
//   g++ -O3 -fopenmp -std=c++17 -Iinclude -Iraytracer-2025_b200/csrc -Iraytracer-2025_b200/host -I/usr/local/cuda/include \
//       scripts/time_compile.cpp raytracer-2025_b200/csrc/compile.cpp raytracer-2025_b200/csrc/bvh_build.cpp \
//       raytracer-2025_b200/host/host_capi.cpp -o /tmp/time_compile && RT2025_TIMING=1 /tmp/time_compile tri_soup 2000000
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>

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
    return rc;
}

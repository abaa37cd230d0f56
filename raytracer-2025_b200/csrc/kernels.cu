// kernels.cu — the wavefront path tracer and the closest-hit batch kernel (sm_100a).
//
// Pipeline per iteration (all queues and counters live in HBM; no host round trip is needed to
// size a launch — kernels are persistent grid-stride loops that read the counts themselves):
//
//   generate : tops the current ray stream up to capacity with camera rays (Camera::get_ray, camera.rs:247-273)
//   media    : constant-medium sampling (volume.rs:37-73).  When every boundary is a sphere it runs BEFORE extend and
//              leaves the nearest scatter point in the hit stream as the incumbent; otherwise after it, against the
//              surface hit
//   extend   : closest surface hit, persistent traversal (world.hit, camera.rs:286)
//   bin      : the append of every queue position to the shade queue of the winner's material class
//   shade    : one kernel per class - emitted + scatter + mixture-pdf light sampling (camera.rs:290-321);
//              survivors are appended to the other copy of the ray/state streams
//   walk     : the class of scatter points inside optically thick media: the path stays in registers from one scatter point to
//              the next (Isotropic::scatter, then world.hit of the next segment in place) until it leaves the medium or ends
//
// Paths flow through dense, position-indexed streams (kernels.h); terminated paths simply are not
// re-appended.  Radiance is accumulated with binary64 atomics into a per-pixel framebuffer.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "kernels.h"
#include "shade.cuh"

namespace rt {

// ------------------------------------------------------------------------------------------
// closest-hit batch kernel (rt_closest_hit)
// ------------------------------------------------------------------------------------------
struct RayArrayIO {  // rt_closest_hit batches: rays in, rt_hit out
    const SceneView& sv;
    const rt_ray* __restrict__ rays;
    rt_hit* __restrict__ out;
    double tmin, tmax;
    __device__ __forceinline__ double t_min() const { return tmin; }
    __device__ __forceinline__ bool load(uint32_t i, RayD& r, double& t1, uint32_t& prim0, uint32_t& rank0) const {
        const double* rp = reinterpret_cast<const double*>(rays + i);
        r.o = D3{rp[0], rp[1], rp[2]};
        r.d = D3{rp[3], rp[4], rp[5]};
        r.time = rp[6];
        t1 = tmax;
        prim0 = 0xFFFFFFFFu, rank0 = 0xFFFFFFFFu;
        return true;
    }
    __device__ __forceinline__ void prefetch(uint32_t i) const { asm volatile("prefetch.global.L2 [%0];" ::"l"(rays + i)); }
    __device__ __forceinline__ void store(uint32_t i, bool hit, double t, uint32_t prim, uint32_t) const {
        rt_hit h;
        if (hit) {
            RayD r;
            double b;
            uint32_t p0, r0;
            load(i, r, b, p0, r0);
            HitInfo hi;
            surface_hit_info(sv, prim, t, r, true, hi);
            const PrimMeta m = sv.meta[prim];
            h.t = t;
            h.prim_id = m.object;
            h.inst_id = m.xform == RT_NONE ? RT_NONE : sv.xforms[m.xform].inst_object;
            h.u = (float)hi.u, h.v = (float)hi.v;
        } else {
            h.t = INFINITY;
            h.prim_id = RT_NONE, h.inst_id = RT_NONE;
            h.u = 0.f, h.v = 0.f;
        }
        out[i] = h;
    }
};

template <bool COUNT, bool PARK, int WIDE, bool XF = true, int KINDS = 0>
__global__ void __launch_bounds__(EXTEND_BLOCK, EXTEND_MIN_BLOCKS) k_closest_hit(SceneView sv, const rt_ray* __restrict__ rays, uint32_t n, double tmin,
                                                                 double tmax, rt_hit* __restrict__ out, unsigned long long* counters) {
    extern __shared__ float4 s_mem[];  // [cached nodes | traversal stacks | per-warp ray FIFOs]
    __shared__ uint32_t s_cursor;
    uint32_t* stack_base = reinterpret_cast<uint32_t*>(s_mem + cached_tree_float4(sv));
    uint32_t* stack = stack_base + threadIdx.x;
    uint32_t* fifo = stack_base + sv.stack_entries * EXTEND_BLOCK + (threadIdx.x >> 5) * sv.fifo_slots * FIFO_SLOT_WORDS;
    if (threadIdx.x == 0) s_cursor = 0;
    stage_nodes(sv, s_mem);
    TraceCounters cnt{0, 0};
    RayArrayIO io{sv, rays, out, tmin, tmax};
    double* ray_s = reinterpret_cast<double*>(stack_base + sv.stack_entries * EXTEND_BLOCK + (EXTEND_BLOCK / 32) * sv.fifo_slots * FIFO_SLOT_WORDS) + threadIdx.x;
    trace_persistent<COUNT, true, PARK, WIDE, XF, KINDS>(sv, io, n, &s_cursor, s_mem, stack, EXTEND_BLOCK, fifo, sv.fifo_slots, &cnt, ray_s);
    if (COUNT) {
        atomicAdd(&counters[0], (unsigned long long)cnt.nodes);
        atomicAdd(&counters[1], (unsigned long long)cnt.prims);
    }
}

// Launch with the top of the BVH pinned in L2.  A scene whose nodes + primitives exceed the L2 makes almost every
// warp-wide node visit wait for at least one lane's DRAM miss; nodes are stored breadth first, so a persisting
// access-policy window over a prefix of the array is "the top of the tree" and keeps it resident while the
// primitive records stream through the rest of the cache.  sv.l2_window_bytes == 0 (cache-resident scenes) is
// a plain launch.  The window is a per-launch attribute: the caller's stream is left untouched.
template <class K, class... Args>
static void launch_with_l2_window(K kernel, const SceneView& sv, int grid, int block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid), cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem, cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    if (sv.l2_window_bytes) {
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = sv.nodes4 ? (void*)const_cast<Node4*>(sv.nodes4) : (void*)const_cast<Node*>(sv.nodes);
        attr[0].val.accessPolicyWindow.num_bytes = sv.l2_window_bytes;
        attr[0].val.accessPolicyWindow.hitRatio = 1.0f;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cfg.attrs = attr, cfg.numAttrs = 1;
    }
    cudaLaunchKernelEx(&cfg, kernel, args...);
}

void launch_closest_hit(const SceneView& sv, const rt_ray* d_rays, uint32_t n, double tmin, double tmax, bool count, rt_hit* d_out,
                        unsigned long long* d_counters, int grid, size_t stack_bytes, cudaStream_t stream) {  // stack_bytes = whole dynamic smem
    // scenes without any Transform that hold one kind of primitive (a mesh, the soups) have variants without the detransform and
    // without the other primitive test
    auto soup = [&]() {
        if (sv.kinds == 1) return count ? k_closest_hit<true, true, 1, false, 1> : k_closest_hit<false, true, 1, false, 1>;
        if (sv.kinds == 2) return count ? k_closest_hit<true, true, 1, false, 2> : k_closest_hit<false, true, 1, false, 2>;
        return count ? k_closest_hit<true, true, 1, false, 0> : k_closest_hit<false, true, 1, false, 0>;
    };
    auto k = (sv.nodes4 && !sv.n_cached_nodes && !sv.n_xforms) ? soup()
             : sv.nodes4 ? (sv.n_cached_nodes && !sv.fifo_slots ? (count ? k_closest_hit<true, true, 2> : k_closest_hit<false, true, 2>)
                                            : (count ? k_closest_hit<true, true, 1> : k_closest_hit<false, true, 1>))
             : (!sv.park_leaves && !sv.fifo_slots && sv.n_cached_nodes == sv.n_nodes) ? (count ? k_closest_hit<true, false, 3> : k_closest_hit<false, false, 3>)
             : count   ? (sv.park_leaves ? k_closest_hit<true, true, 0> : k_closest_hit<true, false, 0>)
                       : (sv.park_leaves ? k_closest_hit<false, true, 0> : k_closest_hit<false, false, 0>);
    launch_with_l2_window(k, sv, grid, EXTEND_BLOCK, stack_bytes, stream, sv, d_rays, n, tmin, tmax, d_out, d_counters);
}

// ------------------------------------------------------------------------------------------
// wavefront state
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t pack_ids(uint32_t pixel, uint32_t sample, uint32_t segment) {
    return (uint64_t)pixel | ((uint64_t)sample << IDS_PIXEL_BITS) | ((uint64_t)segment << (IDS_PIXEL_BITS + IDS_SAMPLE_BITS));
}
__device__ __forceinline__ void unpack_ids(uint64_t ids, uint32_t& pixel, uint32_t& sample, uint32_t& segment) {
    pixel = (uint32_t)(ids & ((1ull << IDS_PIXEL_BITS) - 1ull));
    sample = (uint32_t)((ids >> IDS_PIXEL_BITS) & ((1ull << IDS_SAMPLE_BITS) - 1ull));
    segment = (uint32_t)(ids >> (IDS_PIXEL_BITS + IDS_SAMPLE_BITS));
}
__device__ __forceinline__ void load_ray(const RayRec* __restrict__ rec, RayD& r, uint64_t& ids) {
    const double2* p = reinterpret_cast<const double2*>(rec);
    double2 a0 = p[0], a1 = p[1], a2 = p[2], a3 = p[3];
    r.o = D3{a0.x, a0.y, a1.x};
    r.d = D3{a1.y, a2.x, a2.y};
    r.time = a3.x;
    ids = (uint64_t)__double_as_longlong(a3.y);
}
__device__ __forceinline__ void store_ray(RayRec* __restrict__ rec, const RayD& r, uint64_t ids) {
    double2* p = reinterpret_cast<double2*>(rec);
    p[0] = make_double2(r.o.x, r.o.y);
    p[1] = make_double2(r.o.z, r.d.x);
    p[2] = make_double2(r.d.y, r.d.z);
    p[3] = make_double2(r.time, __longlong_as_double((long long)ids));
}
__device__ __forceinline__ void store_beta(BetaRec* __restrict__ rec, D3 beta) {
    double2* p = reinterpret_cast<double2*>(rec);
    p[0] = make_double2(beta.x, beta.y);
    p[1] = make_double2(beta.z, 0.0);
}

// warp-aggregated reservation of one entry in queue `q` (lanes with q < 0 reserve nothing);
// returns the reserved position
__device__ __forceinline__ uint32_t queue_reserve(uint32_t* counts, int q) {
    unsigned active = __activemask();
    unsigned peers = __match_any_sync(active, q);
    uint32_t pos = 0;
    if (q >= 0) {
        int leader = __ffs(peers) - 1;
        int lane = threadIdx.x & 31;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(&counts[q], (uint32_t)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        pos = base + __popc(peers & ((1u << lane) - 1));
    }
    return pos;
}

// pixel list of this partition: 8x8 tiles dealt round-robin (rt_render_opts.part_index/part_count)
__global__ void k_pixel_list(uint32_t W, uint32_t H, uint32_t part_index, uint32_t part_count, uint32_t* list, uint32_t* count) {
    uint32_t tiles_x = (W + 7) / 8, tiles_y = (H + 7) / 8;
    uint32_t n_tiles = tiles_x * tiles_y;
    // one warp handles one tile = 64 pixels, two per lane; order inside the list is tile-major
    uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x & 31;
    uint32_t n_warps = gridDim.x * blockDim.x / 32;
    for (uint32_t tile = warp; tile < n_tiles; tile += n_warps) {
        if (tile % part_count != part_index) continue;
        uint32_t tx = tile % tiles_x, ty = tile / tiles_x;
        // positions in the list are claimed with one atomic per tile (order only affects scheduling)
        uint32_t valid = 0;
        uint32_t px[2], ok[2];
        for (int k = 0; k < 2; k++) {
            uint32_t p = lane + 32 * k;
            uint32_t x = tx * 8 + (p & 7), y = ty * 8 + (p >> 3);
            ok[k] = x < W && y < H;
            px[k] = y * W + x;
            valid += ok[k];
        }
        uint32_t total = valid;
        for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xFFFFFFFFu, total, o);
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(count, total);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        // exclusive prefix of `valid` over lanes
        uint32_t incl = valid;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= (uint32_t)o) incl += v;
        }
        uint32_t pos = base + incl - valid;
        for (int k = 0; k < 2; k++)
            if (ok[k]) list[pos++] = px[k];
    }
}

// after shade the current stream is consumed: its length and the class queues are reset (one thread)
__device__ __forceinline__ void reset_consumed_queues(const WavefrontState& W) {
    Counters* c = W.counters;
    c->n_extend[W.parity] = 0;
    c->walk_cursor = 0;
    for (int k = 0; k < SC_COUNT; k++) c->n_shade[k] = 0;
}
__global__ void k_step(WavefrontState W) { reset_consumed_queues(W); }

// ------------------------------------------------------------------------------------------
// extend
// ------------------------------------------------------------------------------------------
// the free-flight draw of one medium (rt2025_rng.h: one Philox call serves two media; `cache` keeps the pair between the
// iterations of a loop over the media)
struct MediumDraws {
    Rand2 pair;
    uint32_t slot = 0xFFFFFFFFu;
    __device__ __forceinline__ double get(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t segment, uint32_t medium_index) {
        const uint32_t s = RT_MEDIUM_SLOT(medium_index);
        if (s != slot) {
            pair = philox_pair(seed, pixel, sample, segment, s);
            slot = s;
        }
        return (medium_index & 1u) ? pair.b : pair.a;
    }
};

constexpr uint32_t MEDIUM_INCUMBENT = 0x80000000u;  // `prim` of a traversal whose incumbent is a medium scatter point: flag | medium index

// ConstantMedium::hit (volume.rs:37-73) for a medium whose boundary is a single Sphere, on the caller's interval
// [1e-8, inf): true and the scatter distance when the path scatters inside.  XF: the medium or its sphere may sit
// under a Transform.
template <bool COUNT, bool XF>
__device__ __forceinline__ bool sample_sphere_medium(const SceneView& sv, const Medium& med, const RayD& r, double xi, double& tm, TraceCounters* cnt) {
    RayD lr = r;
    if (XF && med.xform != RT_NONE) lr = ray_to_local(sv, med.xform, r);
    RayD br = lr;
    if (XF) {
        const uint32_t bx = sv.meta[med.single_sphere].xform;
        if (bx != med.xform) br = bx == RT_NONE ? r : ray_to_local(sv, bx, r);
    }
    double t1, t2;
    if (COUNT) cnt->prims += 2;
    if (!sphere_entry_exit(sv.geom[med.single_sphere].d, br, t1, t2)) return false;
    if (t1 < 1e-8) t1 = 1e-8;  // clamp to the caller's interval [1e-8, inf)
    if (t1 >= t2) return false;
    if (t1 < 0.0) t1 = 0.0;
    const double ray_length = length(lr.d);
    const double distance_inside_boundary = (t2 - t1) * ray_length;
    // binary32 screen: when even a pessimistic free-flight estimate clears the segment inside the boundary by a wide
    // margin the path does not scatter here and the binary64 ln is not needed (the exact comparison below is unchanged)
    const float hf = (float)med.neg_inv_density * logf((float)xi);
    const float hf_err = 4e-7f * fabsf((float)med.neg_inv_density);
    if (hf > (float)distance_inside_boundary * 1.001f + hf_err) return false;
    const double hit_distance = med.neg_inv_density * log(xi);
    if (hit_distance > distance_inside_boundary) return false;
    tm = t1 + hit_distance / ray_length;
    return true;
}

// MEDIA: 0 = the ray load does not look at media (none, or they are sampled by k_media after extend);
// 3 = a k_media<PRE> pass ahead of extend left the nearest scatter point of every ray in the hit stream;
// 1 / 2 = every ConstantMedium boundary is a single Sphere and the media are sampled HERE, while the ray is being
// prepared (2: some of them under a Transform).  The nearest scatter point becomes the incumbent of the traversal -
// its distance bounds the search, its tie rank decides an exact tie - so a path scattering inside a dense medium
// (book2's trapped paths bounce there up to max_depth times) costs a traversal of a few units instead of the whole
// scene, and no separate pass re-reads the ray stream.  The winner is the same as with Hittables::hit's order
// (hits.rs:39-46 takes the minimum t over all children, media included).
template <bool COUNT, int MEDIA>
struct PathIO {  // k_extend: rays come from the current ray stream, hits go to the hit stream
    const SceneView& sv;
    const RayRec* __restrict__ rays;
    HitRec* __restrict__ hits;
    uint8_t* __restrict__ cls;  // nullptr: a media pass follows and writes the class bytes
    uint64_t seed;
    bool bin_by_class, drop_misses;
    TraceCounters* cnt;
    __device__ __forceinline__ double t_min() const { return 1e-8; }  // camera.rs:286
    __device__ __forceinline__ bool load(uint32_t j, RayD& r, double& t1, uint32_t& prim0, uint32_t& rank0) const {
        uint64_t ids;
        load_ray(rays + j, r, ids);
        t1 = INFINITY;
        prim0 = 0xFFFFFFFFu, rank0 = 0xFFFFFFFFu;
        if (MEDIA == 3) {
            const double2 hw = *reinterpret_cast<const double2*>(hits + j);
            if ((uint32_t)__double2loint(hw.y) == HIT_MEDIUM) {  // a surface must beat this point (or tie it with a lower rank)
                const uint32_t m = (uint32_t)__double2hiint(hw.y);
                t1 = hw.x;
                prim0 = MEDIUM_INCUMBENT | m;
                rank0 = sv.media[m].rank;
            }
        } else if (MEDIA) {
            uint32_t pixel, sample, segment;
            unpack_ids(ids, pixel, sample, segment);
            MediumDraws draws;
            for (uint32_t m = 0; m < sv.n_media; m++) {
                const Medium& med = sv.media[m];
                const double xi = draws.get(seed, pixel, sample, segment, med.medium_index);
                double tm;
                if (!sample_sphere_medium<COUNT, MEDIA == 2>(sv, med, r, xi, tm, cnt)) continue;
                // the media compete with each other like any children of a container: nearest, then lower rank
                if (prim0 == 0xFFFFFFFFu || tm < t1 || (tm == t1 && med.rank < rank0)) {
                    t1 = tm;
                    prim0 = MEDIUM_INCUMBENT | m;
                    rank0 = med.rank;
                }
            }
        }
        return true;
    }
    __device__ __forceinline__ void prefetch(uint32_t j) const {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(rays + j));
        if (MEDIA == 3) asm volatile("prefetch.global.L2 [%0];" ::"l"(hits + j));
    }
    __device__ __forceinline__ void store(uint32_t j, bool hit, double t, uint32_t prim, uint32_t prim_meta) const {
        uint32_t kind = HIT_MISS, c = SC_MISS;
        if (hit) {
            if (MEDIA && (prim & MEDIUM_INCUMBENT)) {  // the scatter point kept its place
                kind = HIT_MEDIUM, prim &= ~MEDIUM_INCUMBENT;
                // the phase function of a ConstantMedium is an Isotropic (validated by the scene compiler)
                c = (sv.media[prim].flags & MEDIUM_THICK) ? SC_WALK : SC_ISOTROPIC;
            } else {
                kind = HIT_SURFACE;
                c = (prim_meta >> META_CLASS_SHIFT) & 15u;  // PrimMeta::kind_mat of the winner, kept by the traversal
            }
        }
        *reinterpret_cast<double2*>(hits + j) = make_double2(hit ? t : INFINITY, __hiloint2double((int)prim, (int)kind));
        if (cls) cls[j] = (c == SC_MISS && drop_misses) ? CLASS_DROPPED : (uint8_t)(bin_by_class || c == SC_MISS ? c : SC_DIFFUSE);
    }
};

// closest hit of every path in the extend queue: world.hit (camera.rs:286) over the surfaces and, with MEDIA, the media
template <bool COUNT, bool PARK, int WIDE, int MEDIA>
__global__ void __launch_bounds__(EXTEND_BLOCK, EXTEND_MIN_BLOCKS) k_extend(SceneView sv, RenderParams P, WavefrontState W) {
    extern __shared__ float4 s_mem[];  // [cached nodes | traversal stacks | per-warp ray FIFOs]
    __shared__ uint32_t s_cursor;
    uint32_t* stack_base = reinterpret_cast<uint32_t*>(s_mem + cached_tree_float4(sv));
    uint32_t* stack = stack_base + threadIdx.x;
    uint32_t* fifo = stack_base + sv.stack_entries * EXTEND_BLOCK + (threadIdx.x >> 5) * sv.fifo_slots * FIFO_SLOT_WORDS;
    const uint32_t n = W.counters->n_extend[W.parity];
    if (n == 0) return;
    if (threadIdx.x == 0) s_cursor = 0;
    stage_nodes(sv, s_mem);
    TraceCounters cnt{0, 0};
    // a media pass after extend (classic order, or boundaries that are not spheres) owns the class bytes
    const bool media_pass_follows = sv.n_media != 0 && MEDIA == 0;  // (P.media_first == 0)
    PathIO<COUNT, MEDIA> io{sv, W.ray_q[W.parity], W.hit_q, media_pass_follows ? nullptr : W.cls_q, P.seed, P.bin_by_class != 0, P.drop_misses != 0, &cnt};
    double* ray_s = reinterpret_cast<double*>(stack_base + sv.stack_entries * EXTEND_BLOCK + (EXTEND_BLOCK / 32) * sv.fifo_slots * FIFO_SLOT_WORDS) + threadIdx.x;
    trace_persistent<COUNT, true, PARK, WIDE, true, 0>(sv, io, n, &s_cursor, s_mem, stack, EXTEND_BLOCK, fifo, sv.fifo_slots, &cnt, ray_s);
    if (COUNT) {
        atomicAdd(&W.counters->node_visits, (unsigned long long)cnt.nodes);
        atomicAdd(&W.counters->prim_tests, (unsigned long long)cnt.prims);
    }
}

// ConstantMedium::hit of every medium against the hit known so far (t, prim, kind; kind == HIT_MISS: none): the nearest scatter
// point replaces it when it wins (nearer, or an exact tie with the lower rank).  Returns whether it did.
// `first`: the medium tried first (the winner is the minimum of (t, rank) and does not depend on the order; the one the path
// is most likely to scatter in goes first so that its scatter point screens the others).  FILTER: 0 = every medium,
// 1 = only MEDIUM_THICK ones, 2 = only the others (k_walk asks in two steps and shares the draws between them).
template <bool COUNT, bool GENERIC, bool XF, int FILTER = 0>
__device__ __forceinline__ bool sample_media(const SceneView& sv, uint64_t seed, const RayD& r, uint32_t pixel, uint32_t sample, uint32_t segment, double& t,
                                             uint32_t& prim, uint32_t& kind, const float4* s_mem, uint32_t* stack, int stride, TraceCounters* cntp,
                                             uint32_t first, MediumDraws& draws) {
    uint32_t rank = kind == HIT_SURFACE ? sv.meta[prim].rank : (kind == HIT_MEDIUM ? sv.media[prim].rank : 0xFFFFFFFFu);
    bool changed = false;
        for (uint32_t k = 0; k < sv.n_media; k++) {
            uint32_t m = first + k;
            if (m >= sv.n_media) m -= sv.n_media;
            if (FILTER && ((sv.media[m].flags & MEDIUM_THICK) != 0u) != (FILTER == 1)) continue;
            const Medium& med = sv.media[m];
            const double xi = draws.get(seed, pixel, sample, segment, med.medium_index);
            double tm;
            RayD lr = r;
            if (XF && med.xform != RT_NONE) lr = ray_to_local(sv, med.xform, r);
            // binary32 screen: the scatter point lies at t1 + dist/len with t1 >= 0, so a free flight that clearly
            // overshoots the known hit cannot win whatever the boundary does: skip the boundary test
            if (RT_MEDIA_EARLY_SCREEN && kind != HIT_MISS) {
                const float hf = (float)med.neg_inv_density * logf((float)xi);
                const float hf_err = 4e-7f * fabsf((float)med.neg_inv_density);
                const float len_f = sqrtf((float)lr.d.x * (float)lr.d.x + (float)lr.d.y * (float)lr.d.y + (float)lr.d.z * (float)lr.d.z);
                if (hf > (float)t * len_f * 1.001f + hf_err) continue;
            }
            if (med.single_sphere != RT_NONE) {
                if (!sample_sphere_medium<COUNT, XF>(sv, med, r, xi, tm, cntp)) continue;
            } else if (GENERIC) {
                double t1, t2;
                uint32_t bp;
                if (!closest_hit<COUNT, false>(sv, med.root, r, -INFINITY, INFINITY, s_mem, stack, stride, t1, bp, cntp)) continue;
                if (!closest_hit<COUNT, false>(sv, med.root, r, t1 + 0.0001, INFINITY, s_mem, stack, stride, t2, bp, cntp)) continue;
                if (t1 < 1e-8) t1 = 1e-8;  // clamp to the caller's interval [1e-8, inf)
                if (t1 >= t2) continue;
                if (t1 < 0.0) t1 = 0.0;
                const double ray_length = length(lr.d);
                const double distance_inside_boundary = (t2 - t1) * ray_length;
                const double hit_distance = med.neg_inv_density * log(xi);
                if (hit_distance > distance_inside_boundary) continue;
                tm = t1 + hit_distance / ray_length;
            } else {
                continue;  // unreachable: the host launches the GENERIC instantiation for such scenes
            }
            // the medium competes with the other children of its container like any hit
            if (kind == HIT_MISS || tm < t || (tm == t && med.rank < rank)) {
                t = tm;
                prim = m;
                kind = HIT_MEDIUM;
                rank = med.rank;
                changed = true;
            }
        }
    return changed;
}

template <bool COUNT, bool GENERIC, bool XF>
__device__ __forceinline__ bool sample_media(const SceneView& sv, uint64_t seed, const RayD& r, uint32_t pixel, uint32_t sample, uint32_t segment, double& t,
                                             uint32_t& prim, uint32_t& kind, const float4* s_mem, uint32_t* stack, int stride, TraceCounters* cntp) {
    MediumDraws draws;
    return sample_media<COUNT, GENERIC, XF, 0>(sv, seed, r, pixel, sample, segment, t, prim, kind, s_mem, stack, stride, cntp, 0u, draws);
}

// Camera::get_ray, camera.rs:247-273: tops the current ray stream up to capacity with camera rays
// MEDIA: the scene's media all have sphere boundaries and are sampled ahead of extend (RenderParams::media_first == 1): the
// free-flight draws of segment 0 are taken here, while the ray is in registers, and the nearest scatter point goes to the hit
// stream as extend's incumbent - the sampling pass then only reads the survivors of the last shade stage (a camera ray is a
// quarter to a third of all segments, and this kernel is bound by its stores).  XF as in k_media.
#ifndef RT_GEN_MEDIA_MIN_BLOCKS
#define RT_GEN_MEDIA_MIN_BLOCKS 4
#endif
template <bool MEDIA, bool XF>
__global__ void __launch_bounds__(256, MEDIA ? RT_GEN_MEDIA_MIN_BLOCKS : 4) k_generate(SceneView sv, RenderParams P, WavefrontState W) {
    const uint64_t remaining = W.total_paths - W.counters->next_path;
    const uint32_t extend_base = W.counters->n_extend[W.parity];
    const uint32_t room = W.capacity - extend_base;
    const uint32_t n_new = (uint32_t)(remaining < (uint64_t)room ? remaining : (uint64_t)room);
    const uint64_t first = W.counters->next_path;
    const rt_camera& cam = P.cam;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n_new; j += gridDim.x * blockDim.x) {
        uint64_t g = first + j;
        uint32_t sidx = P.sample_begin + (uint32_t)(g / W.n_pixels);
        uint32_t pixel = W.pixel_list[g % W.n_pixels];
        uint32_t i = pixel % cam.image_width, jj = pixel / cam.image_width;
        uint32_t s_i = sidx / cam.sqrt_spp, s_j = sidx % cam.sqrt_spp;
        Rand2 jit = philox_pair(P.seed, pixel, sidx, 0, RT_SLOT_CAM_JITTER);
        double px = (((double)s_i + jit.a) * cam.recip_sqrt_spp) - 0.5;
        double py = (((double)s_j + jit.b) * cam.recip_sqrt_spp) - 0.5;
        D3 pixel_sample = ld3(cam.pixel00_loc) + (((double)i + px) * ld3(cam.pixel_delta_u)) + (((double)jj + py) * ld3(cam.pixel_delta_v));
        RayD r;
        if (cam.defocus_angle_in_degrees <= 0.0) {
            r.o = ld3(cam.center);
        } else {  // defocus_disk_sample, vec3.rs:63-69
            Rand2 dk = philox_pair(P.seed, pixel, sidx, 0, RT_SLOT_CAM_DISK);
            double theta = (2.0 * RT_PI) * dk.a;
            double rr = sqrt(dk.b);
            double s, c;
            sincos(theta, &s, &c);
            double p0 = rr * c, p1 = rr * s;
            r.o = ld3(cam.center) + (p0 * ld3(cam.defocus_disk_u)) + (p1 * ld3(cam.defocus_disk_v));
        }
        r.d = pixel_sample - r.o;
        r.time = philox_pair(P.seed, pixel, sidx, 0, RT_SLOT_CAM_TIME).a;
        store_ray(W.ray_q[W.parity] + extend_base + j, r, pack_ids(pixel, sidx, 0u));
        store_beta(W.beta_q[W.parity] + extend_base + j, D3{1.0, 1.0, 1.0});
        if (MEDIA) {
            double t = INFINITY;
            uint32_t prim = 0xFFFFFFFFu, kind = HIT_MISS;
            TraceCounters cnt{0, 0};
            sample_media<false, false, XF>(sv, P.seed, r, pixel, sidx, 0u, t, prim, kind, nullptr, nullptr, 0, &cnt);
            *reinterpret_cast<double2*>(W.hit_q + extend_base + j) = make_double2(t, __hiloint2double((int)prim, (int)kind));
        }
    }
    // Bookkeeping of the stage (it used to be a launch of its own): every block read the counters above before it got here,
    // so the LAST block to arrive may advance them - the next kernel of the stream sees the topped-up queue length.
    __syncthreads();
    if (threadIdx.x == 0) {
        Counters* c = W.counters;
        __threadfence();
        if (atomicAdd(&c->gen_done, 1u) == gridDim.x - 1u) {
            c->gen_done = 0;
            c->next_path += n_new;
            c->n_extend[W.parity] = extend_base + n_new;
            c->n_unsampled = MEDIA ? extend_base : extend_base + n_new;
            c->segments += extend_base + n_new;
            c->iterations += (extend_base + n_new > 0);
            __threadfence();
        }
    }
}

// Media pass AFTER extend (the order of Hittables::hit): ConstantMedium::hit for every medium (volume.rs:37-73)
// against the surface hit k_extend found, one thread per extend-queue entry.  Used when some boundary is not a
// single Sphere (GENERIC: the boundary needs a BVH traversal per lane, kept out of the common instantiation because
// it doubles the register footprint) and for the classic-order A/B switch; sphere-bounded media are otherwise sampled
// by k_extend itself (PathIO<MEDIA>).  Writes the winner back to the hit stream and the class byte of every entry.
//   MODE 1: every boundary is a single Sphere; MODE 2: general boundaries.
//   XF = some medium (or its sphere boundary) sits under a Transform (without the detransform code the single-sphere
//   pass needs 81 registers instead of 126 and runs three CTAs per SM).
//   PRE = the pass runs BEFORE extend (sphere boundaries only): it knows no surface hit and leaves the nearest scatter
//   point (or a miss) in the hit stream; extend takes it as the incumbent (PathIO<3>) and writes the class bytes.
template <bool COUNT, int MODE, bool XF, bool PRE>
__global__ void __launch_bounds__(MEDIA_BLOCK, MODE == 2 ? RT_MEDIA_GENERIC_MIN_BLOCKS : (XF ? RT_MEDIA_MIN_BLOCKS : RT_MEDIA_MIN_BLOCKS_NOXF))
    k_media(SceneView sv, RenderParams P, WavefrontState W) {
    constexpr bool GENERIC = MODE == 2;
    extern __shared__ float4 s_mem[];  // traversal stacks for boundaries that are not a single sphere
    uint32_t* stack = reinterpret_cast<uint32_t*>(s_mem) + threadIdx.x;
    TraceCounters cnt{0, 0};
    const uint32_t n = PRE ? W.counters->n_unsampled : W.counters->n_extend[W.parity];
    const uint32_t n_round = (n + 31u) & ~31u;  // whole warps iterate together
    const RayRec* __restrict__ rays = W.ray_q[W.parity];
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n_round; j += gridDim.x * blockDim.x) {
        {  // pull the next iteration's records towards L1 while this one computes
            const uint32_t jn = j + gridDim.x * blockDim.x;
            if (jn < n) {
                asm volatile("prefetch.global.L1 [%0];" ::"l"(rays + jn));
                if (!PRE) asm volatile("prefetch.global.L1 [%0];" ::"l"(W.hit_q + jn));
            }
        }
        if (j >= n) continue;
        double t = INFINITY;
        uint32_t prim = 0xFFFFFFFFu, kind = HIT_MISS;
        if (!PRE) {
            const double2 hw = *reinterpret_cast<const double2*>(W.hit_q + j);
            t = hw.x;
            prim = (uint32_t)__double2hiint(hw.y), kind = (uint32_t)__double2loint(hw.y);
        }
        RayD r;
        uint64_t ids64;
        load_ray(rays + j, r, ids64);
        uint32_t pixel, sample, segment;
        unpack_ids(ids64, pixel, sample, segment);
        const uint32_t surface_meta = kind == HIT_SURFACE ? __ldg(&sv.meta[prim].kind_mat) : 0u;
        const bool changed = sample_media<COUNT, GENERIC, XF>(sv, P.seed, r, pixel, sample, segment, t, prim, kind, s_mem, stack, MEDIA_BLOCK, &cnt);
        if (PRE || changed) *reinterpret_cast<double2*>(W.hit_q + j) = make_double2(t, __hiloint2double((int)prim, (int)kind));
        if (PRE) continue;  // extend writes the class bytes
        uint32_t c = kind == HIT_MISS     ? (uint32_t)SC_MISS
                     : kind == HIT_MEDIUM ? ((sv.media[prim].flags & MEDIUM_THICK) ? (uint32_t)SC_WALK : (uint32_t)SC_ISOTROPIC)
                                          : ((surface_meta >> META_CLASS_SHIFT) & 15u);
        if (!P.bin_by_class && c != SC_MISS) c = SC_DIFFUSE;
        W.cls_q[j] = (c == SC_MISS && P.drop_misses) ? CLASS_DROPPED : (uint8_t)c;
    }
    if (COUNT) {
        atomicAdd(&W.counters->node_visits, (unsigned long long)cnt.nodes);
        atomicAdd(&W.counters->prim_tests, (unsigned long long)cnt.prims);
    }
}

// Binning: the append of every queue position to the shade queue of its class byte.  A memory-bound pass (1 byte in,
// 4 bytes out per entry) with one global atomic per class per 256 entries: same-address atomics are what bounds it.
__global__ void __launch_bounds__(MEDIA_BLOCK, 4) k_bin(WavefrontState W) {
    __shared__ uint32_t s_cnt[2][SC_COUNT], s_base[SC_COUNT];
    if (threadIdx.x < 2 * SC_COUNT) (&s_cnt[0][0])[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t n = W.counters->n_extend[W.parity];
    const uint32_t n_round = (n + MEDIA_BLOCK - 1u) / MEDIA_BLOCK * MEDIA_BLOCK;  // whole CTAs iterate together
    uint32_t it = 0;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n_round; j += gridDim.x * blockDim.x, it++) {
        int q = j < n ? (int)W.cls_q[j] : -1;
        if (q == (int)CLASS_DROPPED) q = -1;
        uint32_t* cntb = s_cnt[it & 1];
        const uint32_t local = queue_reserve(cntb, q);
        __syncthreads();
        if (threadIdx.x < SC_COUNT) {
            const uint32_t c = cntb[threadIdx.x];
            if (c) s_base[threadIdx.x] = atomicAdd(&W.counters->n_shade[threadIdx.x], c);
            s_cnt[(it & 1) ^ 1][threadIdx.x] = 0;
        }
        __syncthreads();
        if (q >= 0) W.q_shade[q][s_base[q] + local] = j;
    }
}

// ------------------------------------------------------------------------------------------
// shade
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void contribute(const RenderParams& P, const WavefrontState& W, uint32_t pixel, D3 c) {
    if (isnan(c.x) || isnan(c.y) || isnan(c.z)) {  // camera.rs:323 would panic
        atomicAdd(&W.counters->errors, 1ull);
        return;
    }
    double* a = W.accum + (size_t)pixel * 3;
    if (c.x != 0.0) atomicAdd(a + 0, c.x);
    if (c.y != 0.0) atomicAdd(a + 1, c.y);
    if (c.z != 0.0) atomicAdd(a + 2, c.z);
}

// emitted + scatter + mixture-pdf light sampling for ONE hit (camera.rs:288-321).  Returns whether the path goes on, with the
// next ray and throughput in nr / nbeta.  One instantiation per shade class: everything another class would need is compiled
// out; CLS == SC_OTHER is the fully general version (Mix, Portal, Transparent, lights that wrap a material, misses, media).
template <uint32_t CLS>
__device__ __forceinline__ bool shade_one(const SceneView& sv, const RenderParams& P, const WavefrontState& W, const RayD& r, D3 beta, uint32_t pixel,
                                          uint32_t sidx, uint32_t segment, double t, uint32_t prim, uint32_t kind, RayD& nr, D3& nbeta) {
    constexpr bool GENERIC = CLS == SC_OTHER;
    constexpr bool DO_MISS = CLS == SC_MISS;
    constexpr bool DO_MEDIUM = GENERIC || CLS == SC_ISOTROPIC;
    constexpr bool DO_EMIT = GENERIC || CLS == SC_EMISSIVE || CLS == SC_DISNEY;  // Disney below DiffuseLight wrappers (the OBJ loader's `Ke`)
    constexpr bool DO_PDF = GENERIC || CLS == SC_DIFFUSE || CLS == SC_TEXTURED || CLS == SC_ISOTROPIC || CLS == SC_DISNEY;
    constexpr bool DO_REMAP = GENERIC || CLS == SC_DISNEY;
    constexpr bool DO_METAL = GENERIC || CLS == SC_METAL;
    constexpr bool DO_DIELECTRIC = GENERIC || CLS == SC_DIELECTRIC;
            bool alive = false, error = false;
            nr = r;
            nbeta = beta;

            if (DO_MISS || (GENERIC && kind == HIT_MISS)) {
                D3 bg;
                if (background_value(sv, P.cam.background_tex, r.d, bg))
                    contribute(P, W, pixel, beta * bg);
                else
                    error = true;
            } else {
                HitInfo h;
                bool hit_ok = true;
                if (DO_MEDIUM && kind == HIT_MEDIUM) {  // volume.rs:66-72, in the medium's local space
                    const Medium& med = sv.media[prim];
                    const RayD lr = med.xform == RT_NONE ? r : ray_to_local(sv, med.xform, r);
                    h.p = lr.o + t * lr.d;
                    D3 nrm = D3{1.0, 0.0, 0.0};
                    h.front_face = dot(lr.d, nrm) < 0.0;
                    h.normal = h.front_face ? nrm : -nrm;
                    h.u = 0.0, h.v = 0.0;
                    h.material = med.material;
                    if (med.xform != RT_NONE) hit_ok = hit_to_world(sv, med.xform, h);
                } else {
                    uint32_t mat = sv.meta[prim].kind_mat & META_MAT_MASK;
                    hit_ok = surface_hit_info(sv, prim, t, r, sv.materials[mat].needs_uv != 0, h);
                }
                if (!hit_ok) error = true;
                uint32_t mat = h.material;
                // RemappedMaterial is the per-face wrapper the OBJ loader puts outermost (obj.rs:165-176)
                while (DO_REMAP && sv.materials[mat].kind == RT_MAT_REMAPPED) {
                    if (!remap_record(sv, sv.remaps[sv.materials[mat].inner2], h)) error = true;
                    mat = sv.materials[mat].inner;
                }
                // emitted (camera.rs:290) — only light-carrying materials can return non-black
                uint32_t mk = sv.materials[mat].kind;
                if (DO_EMIT && (mk == RT_MAT_DIFFUSE_LIGHT || mk == RT_MAT_MIX)) {
                    D3 e = material_emitted(sv, mat, h);
                    contribute(P, W, pixel, beta * e);
                }
                // resolve DiffuseLight wrappers and Mix picks down to the scattering material
                uint32_t level = 0;
                while (DO_EMIT && mat != RT_NONE) {
                    const Material& M = sv.materials[mat];
                    if (M.kind == RT_MAT_DIFFUSE_LIGHT)
                        mat = M.inner;
                    else if (M.kind == RT_MAT_MIX) {  // material.rs:249-257
                        double ratio = M.tex == RT_NONE ? M.param
                                                        : (sv.textures[M.tex].a == RT_NONE ? 1.0 : (double)image_get_pixel(sv, sv.textures[M.tex].a, h.u, h.v).w);
                        double xi = philox_pair(P.seed, pixel, sidx, segment, RT_SLOT_MIX + 256u * level).a;
                        mat = xi > ratio ? M.inner : M.inner2;
                        level++;
                    } else
                        break;
                }
                if (mat != RT_NONE) {
                    const Material& M = sv.materials[mat];
                    nr.o = h.p;
                    // the material kinds this instantiation can meet
                    const uint32_t mkind = M.kind;
                    const bool is_metal = DO_METAL && (CLS == SC_METAL || mkind == RT_MAT_METAL);
                    const bool is_dielectric = DO_DIELECTRIC && (CLS == SC_DIELECTRIC || mkind == RT_MAT_DIELECTRIC);
                    const bool is_pdf = DO_PDF && (!GENERIC || mkind == RT_MAT_EMPTY || mkind == RT_MAT_LAMBERTIAN || mkind == RT_MAT_ISOTROPIC ||
                                                   mkind == RT_MAT_DISNEY);
                    switch (is_metal ? RT_MAT_METAL : is_dielectric ? RT_MAT_DIELECTRIC : is_pdf ? RT_MAT_LAMBERTIAN : (GENERIC ? mkind : 0xFFFFu)) {
                        case RT_MAT_METAL: {  // material.rs:82-95
                            D3 ud, ur;
                            if (unit_vector(r.d, ud) && unit_vector(reflect(ud, h.normal), ur)) {
                                Rand2 xi = philox_pair(P.seed, pixel, sidx, segment, RT_SLOT_MATERIAL);
                                nr.d = ur + (M.param * random_unit_vector(xi.a, xi.b));
                                nbeta = beta * ld3(M.color);
                                alive = true;
                            }
                            break;
                        }
                        case RT_MAT_DIELECTRIC: {  // material.rs:117-144
                            double ri = h.front_face ? 1.0 / M.param : M.param;
                            D3 ud;
                            if (!unit_vector(r.d, ud)) {
                                error = true;
                                break;
                            }
                            double cos_theta = rmin(dot(-ud, h.normal), 1.0);
                            double sin_theta = sqrt(1.0 - cos_theta * cos_theta);
                            bool cannot_refract = ri * sin_theta > 1.0;
                            double r0 = (1.0 - ri) / (1.0 + ri);
                            double r0s = r0 * r0;
                            double x = 1.0 - cos_theta;
                            double x2 = x * x;
                            double reflectance = r0s + (1.0 - r0s) * (x * (x2 * x2));
                            D3 dir;
                            if (cannot_refract || reflectance > philox_pair(P.seed, pixel, sidx, segment, RT_SLOT_MATERIAL).a) {
                                dir = reflect(ud, h.normal);
                            } else if (!refract(ud, h.normal, ri, dir)) {
                                error = true;
                                break;
                            }
                            nr.d = dir;
                            nbeta = beta * texture_value(sv, M.tex, h.u, h.v, h.p);
                            alive = true;
                            break;
                        }
                        case RT_MAT_TRANSPARENT:  // material.rs:211-218
                            alive = GENERIC;
                            break;
                        case RT_MAT_PORTAL: {  // material/portal.rs:21-30
                            if (!GENERIC) break;
                            nr.o = h.p + ld3(M.v);
                            // quaternion sandwich product, quaternion.rs:72-104
                            double qw = M.v[3], qx = M.v[4], qy = M.v[5], qz = M.v[6];
                            double aw = qw * 0.0 - qx * r.d.x - qy * r.d.y - qz * r.d.z;
                            double ax = qw * r.d.x + qx * 0.0 + qy * r.d.z - qz * r.d.y;
                            double ay = qw * r.d.y - qx * r.d.z + qy * 0.0 + qz * r.d.x;
                            double az = qw * r.d.z + qx * r.d.y - qy * r.d.x + qz * 0.0;
                            double cx = -qx, cy = -qy, cz = -qz;
                            nr.d.x = aw * cx + ax * qw + ay * cz - az * cy;
                            nr.d.y = aw * cy - ax * cz + ay * qw + az * cx;
                            nr.d.z = aw * cz + ax * cy - ay * cx + az * qw;
                            nbeta = beta * ld3(M.color);
                            alive = true;
                            break;
                        }
                        case RT_MAT_LAMBERTIAN: {  // Empty / Lambertian / Isotropic / Disney
                            // ScatterRecord::PDF branch, camera.rs:297-312
                            const bool iso = CLS == SC_ISOTROPIC || (GENERIC && M.kind == RT_MAT_ISOTROPIC);
                            const bool dis = CLS == SC_DISNEY || (GENERIC && M.kind == RT_MAT_DISNEY);
                            D3 albedo = (M.kind == RT_MAT_EMPTY || dis) ? D3{0.75, 0.75, 0.75} : texture_value(sv, M.tex, h.u, h.v, h.p);
                            ONB uvw;
                            if (!iso && !make_onb(h.normal, uvw)) {
                                error = true;
                                break;
                            }
                            disney::Params DP;
                            D3 v_out_l = D3{0.0, 1.0, 0.0};
                            if (dis) {  // Disney::scatter + DisneyPDF::new, disney.rs:71-91, 522-538
                                DP.base_color = M.tex == RT_NONE ? ld3(M.color) : texture_value(sv, M.tex, h.u, h.v, h.p);
                                DP.roughness = M.v[RT_DISNEY_ROUGHNESS], DP.anisotropic = M.v[RT_DISNEY_ANISOTROPIC], DP.sheen = M.v[RT_DISNEY_SHEEN];
                                DP.sheen_tint = M.v[RT_DISNEY_SHEEN_TINT], DP.clearcoat = M.v[RT_DISNEY_CLEARCOAT];
                                DP.clearcoat_gloss = M.v[RT_DISNEY_CLEARCOAT_GLOSS], DP.specular_tint = M.v[RT_DISNEY_SPECULAR_TINT];
                                DP.metallic = M.v[RT_DISNEY_METALLIC], DP.ior = M.v[RT_DISNEY_IOR], DP.flatness = M.v[RT_DISNEY_FLATNESS];
                                DP.spec_trans = M.v[RT_DISNEY_SPEC_TRANS], DP.diff_trans = M.v[RT_DISNEY_DIFF_TRANS], DP.thin = M.v[RT_DISNEY_THIN] != 0.0;
                                D3 vo;
                                if (!unit_vector(-r.d, vo)) {
                                    error = true;
                                    break;
                                }
                                v_out_l = D3{dot(vo, uvw.u), dot(vo, uvw.v), dot(vo, uvw.w)};  // world_to_onb, onb.rs:40-45
                            }
                            Rand2 pick = philox_pair(P.seed, pixel, sidx, segment, RT_SLOT_MIXTURE);
                            Rand2 dx = philox_pair(P.seed, pixel, sidx, segment, RT_SLOT_DIRECTION);
                            D3 gen = D3{0.0, 0.0, 0.0};
                            const bool have_lights = sv.n_lights > 0;
                            // No `break` inside the divergent sampling branch below: with an early-exit edge the warp would
                            // only reconverge at the end of the switch and run the whole pdf evaluation once per side
                            // (measured: 16 of 32 lanes active from here on).  stop: 1 = the reference panics, 2 = None (black).
                            int stop = 0;
                            if (have_lights && !(pick.a < 0.5)) {
                                if (!lights_random(sv, h.p, pick.b, dx.a, dx.b, gen)) stop = 1;
                            } else if (dis) {
                                Rand2 dpick = philox_pair(P.seed, pixel, sidx, segment, RT_SLOT_DISNEY);
                                D3 v_in_l;
                                bool derr = false;
                                if (!disney::generate_local(DP, v_out_l, h.front_face, dpick, dx, v_in_l, derr))
                                    stop = derr ? 1 : 2;  // None: black (camera.rs:313-315)
                                else if (!unit_vector(onb_to_world(uvw, v_in_l), gen))
                                    stop = 1;
                            } else {
                                gen = iso ? random_unit_vector(dx.a, dx.b) : onb_to_world(uvw, random_cosine_direction(dx.a, dx.b));
                            }
                            // PDF::value, pdf.rs:22-29, 51-57, disney.rs:656-666
                            D3 axp = D3{0.0, 0.0, 0.0};
                            double value0 = 0.0;
                            if (!stop) {
                                if (iso) {
                                    value0 = 1.0 / (4.0 * RT_PI);
                                    axp = albedo / (4.0 * RT_PI);
                                } else {
                                    D3 ud;
                                    if (!unit_vector(gen, ud)) {
                                        stop = 1;
                                    } else if (dis) {
                                        D3 v_in_l = D3{dot(ud, uvw.u), dot(ud, uvw.v), dot(ud, uvw.w)};
                                        bool derr = false;
                                        disney::evaluate_disney(DP, v_out_l, v_in_l, h.front_face, axp, value0, derr);
                                        if (derr) stop = 1;
                                    } else {
                                        double cosine_theta = dot(ud, uvw.v);
                                        value0 = rmax(0.0, cosine_theta / RT_PI);
                                        axp = albedo * rmax(cosine_theta, 0.0) / RT_PI;
                                    }
                                }
                            }
                            double pdf_val = value0;
                            if (!stop && have_lights) {
                                double value1 = lights_pdf_value(sv, P.lights_flat, h.p, gen);
                                if (isnan(value1) || (value0 == 0.0 && value1 == 0.0))  // hits.rs:64, pdf.rs:105-109
                                    stop = 1;
                                else
                                    pdf_val = value0 * 0.5 + value1 * 0.5;
                            }
                            if (!stop && pdf_val == 0.0) stop = 1;  // camera.rs:309
                            if (!stop) {
                                nr.d = gen;
                                nbeta = beta * (axp / pdf_val);
                                alive = true;
                            }
                            if (stop == 1) error = true;
                            break;
                        }
                        default: break;
                    }
                }
            }
            if (error) {
                atomicAdd(&W.counters->errors, 1ull);
                alive = false;
            }
            // a zero throughput can never contribute again; depth == 0 returns black (camera.rs:282)
            if (alive && (segment + 1 >= P.cam.max_depth || (nbeta.x == 0.0 && nbeta.y == 0.0 && nbeta.z == 0.0))) alive = false;
            return alive;
}

// One instantiation per shade class (scene_types.h ShadeClass): the queue it reads only holds hits
// of that class, so everything another class would need is compiled out and the kernel keeps few
// registers.  CLS == SC_OTHER is the fully general version (Mix, Portal, Transparent, lights that
// wrap a material) and is also what runs when binning is switched off.
template <uint32_t CLS>
__global__ void __launch_bounds__(SHADE_BLOCK, ((RT_SHADE_3BLOCK_MASK >> CLS) & 1u) ? 3 : RT_SHADE_MIN_BLOCKS) k_shade(SceneView sv, RenderParams P, WavefrontState W, uint32_t queue) {
    __shared__ uint32_t s_warp_count[SHADE_BLOCK / 32];
    __shared__ uint32_t s_base;
    const uint32_t n = W.counters->n_shade[queue];
    const RayRec* __restrict__ rays = W.ray_q[W.parity];
    const BetaRec* __restrict__ betas = W.beta_q[W.parity];
    // the loop bound is uniform over the block: the survivor append is aggregated per block
    for (uint32_t j0 = blockIdx.x * blockDim.x; j0 < n; j0 += gridDim.x * blockDim.x) {
        const uint32_t j = j0 + threadIdx.x;
        int q = -1;
        RayD nr;
        D3 nbeta;
        uint32_t pixel = 0, sidx = 0, segment = 0;
#if RT_SHADE_PREFETCH
        {  // the records are gathered through the class queue: start pulling the next round's towards L2/L1 now
            const uint32_t jn = j + gridDim.x * blockDim.x;
            if (jn < n) {
                const uint32_t pn = W.q_shade[queue][jn];
                asm volatile("prefetch.global.L2 [%0];" ::"l"(rays + pn));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(betas + pn));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(W.hit_q + pn));
            }
        }
#endif
        if (j < n) {
            const uint32_t pos = W.q_shade[queue][j];
            RayD r;
            uint64_t ids64;
            load_ray(rays + pos, r, ids64);
            unpack_ids(ids64, pixel, sidx, segment);
            const double2* sp2 = reinterpret_cast<const double2*>(betas + pos);
            const double2 b0 = sp2[0], b1 = sp2[1];
            D3 beta = D3{b0.x, b0.y, b1.x};
            const double2 hw = *reinterpret_cast<const double2*>(W.hit_q + pos);
            const double t = hw.x;
            const uint32_t prim = (uint32_t)__double2hiint(hw.y), kind = (uint32_t)__double2loint(hw.y);
            const bool alive = shade_one<CLS>(sv, P, W, r, beta, pixel, sidx, segment, t, prim, kind, nr, nbeta);
            q = alive ? 0 : -1;
        }
        // survivors are appended to the other copy of the streams: one atomic per block
        const unsigned alive_mask = __ballot_sync(0xFFFFFFFFu, q == 0);
        const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
        if (lane == 0) s_warp_count[warp] = __popc(alive_mask);
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t total = 0;
#pragma unroll
            for (int w = 0; w < SHADE_BLOCK / 32; w++) total += s_warp_count[w];
            s_base = total ? atomicAdd(&W.counters->n_extend[W.parity ^ 1u], total) : 0u;
        }
        __syncthreads();
        uint32_t npos = s_base + __popc(alive_mask & ((1u << lane) - 1u));
        for (uint32_t w = 0; w < warp; w++) npos += s_warp_count[w];
        __syncthreads();  // s_warp_count / s_base are reused by the next iteration
        if (q == 0) {
            store_ray(W.ray_q[W.parity ^ 1u] + npos, nr, pack_ids(pixel, sidx, segment + 1));
            store_beta(W.beta_q[W.parity ^ 1u] + npos, nbeta);
        }
    }
}

// ------------------------------------------------------------------------------------------
// random walk inside an optically thick medium
// ------------------------------------------------------------------------------------------
// A path that scatters inside a dense ConstantMedium (book2's blue subsurface sphere: mean free path 5 in a sphere of radius 70)
// scatters there again and again until max_depth ends it: on book2_final a third of ALL segments are such scatter points, and
// each went through the sampling pass, extend, the binning and the isotropic shade kernel - four trips through HBM and a
// traversal at extend's 15 active lanes for a segment that is a few units long.  k_walk takes the queue of scatter points inside
// MEDIUM_THICK media and keeps every path in registers from one scatter point to the next: Isotropic::scatter + the mixture pdf
// (the SC_ISOTROPIC code of shade_one), then world.hit of the NEXT segment in place - the free-flight draw of every medium
// (volume.rs:37-73, the medium the path is in first, so that its scatter point screens the others) and a traversal bounded by
// the nearest scatter point, with the tie rule of the wavefront (a surface wins with a smaller t, or the same t and a lower
// rank).  While a thick medium wins again the loop goes on; otherwise the new ray is appended to the ray stream like any
// survivor of a shade kernel and the wavefront evaluates that segment itself (one look-ahead per walk is computed twice).
// Every draw is addressed by (pixel, sample, segment, slot), so the radiance is the very sum the wavefront alone produces.
// A lane whose walk ends takes the next queue entry at once (static striding, no atomics): the loop body is the same for every
// lane, so the warp stays converged however long the individual walks are.
#ifndef RT_WALK_BLOCK
#define RT_WALK_BLOCK 128
#endif
constexpr int WALK_BLOCK = RT_WALK_BLOCK;
#ifndef RT_WALK_MIN_BLOCKS
#define RT_WALK_MIN_BLOCKS 4
#endif
// ENTRIES: every thick medium of the scene has its entry leaves and there is only one of them, so the traversal from the world
// root (another medium's scatter point, more leaves than slots) is never needed and is compiled out.
template <bool XF, bool ENTRIES>
__global__ void __launch_bounds__(WALK_BLOCK, RT_WALK_MIN_BLOCKS) k_walk(SceneView sv, RenderParams P, WavefrontState W) {
    extern __shared__ float4 s_mem[];  // traversal stacks (global-memory nodes only)
    uint32_t* stack = reinterpret_cast<uint32_t*>(s_mem) + threadIdx.x;
    const uint32_t n = W.counters->n_shade[SC_WALK];
    const RayRec* __restrict__ rays = W.ray_q[W.parity];
    const BetaRec* __restrict__ betas = W.beta_q[W.parity];
    const uint32_t* __restrict__ queue = W.q_shade[SC_WALK];
    const uint32_t lane = threadIdx.x & 31u;
    const unsigned FULL = 0xFFFFFFFFu, lt_mask = (1u << lane) - 1u;
    TraceCounters cnt{0, 0};
    unsigned long long walked = 0;
    bool have = false;
    RayD r;
    D3 beta;
    uint32_t pixel = 0, sidx = 0, segment = 0, medium = 0;
    double t = 0.0;
    // A walk is a serial chain (about 8 us per segment for a warp that has the SM to itself): once the queue is too short to fill the
    // GPU - the drain of a frame, where every iteration would wait some 300 us for the handful of paths that have just entered the
    // medium - a walk is cut after `drain_steps` segments and its path goes back to the ray stream; the wavefront evaluates the
    // next segment, finds the same scatter point and queues the path for the walk of the next iteration.
    const uint32_t max_steps = n <= P.walk_drain_queue ? P.walk_drain_steps : 0xFFFFFFFFu;
    uint32_t steps = 0;
    // queue entries are claimed 32 at a time per warp (one atomic, records prefetched) and handed to the lanes as their walks end
    uint32_t pool_next = 0, pool_end = 0;  // warp-uniform
    bool exhausted = false;                // warp-uniform: the cursor ran past the queue
    while (true) {
        unsigned need = __ballot_sync(FULL, !have);
        while (need && !(exhausted && pool_next == pool_end)) {
            if (pool_next == pool_end) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&W.counters->walk_cursor, 32u);
                base = __shfl_sync(FULL, base, 0);
                if (base >= n) {
                    exhausted = true;
                    break;
                }
                pool_next = base, pool_end = min(base + 32u, n);
                if (base + lane < pool_end) {
                    const uint32_t pos = queue[base + lane];
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(rays + pos));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(betas + pos));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(W.hit_q + pos));
                }
            }
            const uint32_t take = min((uint32_t)__popc(need), pool_end - pool_next);
            const uint32_t my = __popc(need & lt_mask);
            if (!have && my < take) {
                const uint32_t pos = queue[pool_next + my];
                uint64_t ids64;
                load_ray(rays + pos, r, ids64);
                unpack_ids(ids64, pixel, sidx, segment);
                const double2* sp2 = reinterpret_cast<const double2*>(betas + pos);
                const double2 b0 = sp2[0], b1 = sp2[1];
                beta = D3{b0.x, b0.y, b1.x};
                const double2 hw = *reinterpret_cast<const double2*>(W.hit_q + pos);
                t = hw.x;
                medium = (uint32_t)__double2hiint(hw.y);
                have = true;
                steps = 0;
            }
            pool_next += take;
            need = __ballot_sync(FULL, !have);
        }
        if (need == FULL) break;  // nothing left to claim and every walk of the warp has ended
        if (have) {
            RayD nr;
            D3 nbeta;
            if (!shade_one<SC_ISOTROPIC>(sv, P, W, r, beta, pixel, sidx, segment, t, medium, HIT_MEDIUM, nr, nbeta)) {
                have = false;
            } else {
                // world.hit of the next segment: the thick media first (the one the path is in leads); when none of them scatters
                // the walk is over whatever the thin ones do, otherwise they are screened by the scatter point found
                double tn = INFINITY;
                uint32_t pn = 0xFFFFFFFFu, kn = HIT_MISS;
                MediumDraws draws;
                sample_media<false, false, XF, 1>(sv, P.seed, nr, pixel, sidx, segment + 1u, tn, pn, kn, s_mem, stack, WALK_BLOCK, &cnt, medium, draws);
                bool stay = kn == HIT_MEDIUM && ++steps < max_steps;
                if (stay) {
                    sample_media<false, false, XF, 2>(sv, P.seed, nr, pixel, sidx, segment + 1u, tn, pn, kn, s_mem, stack, WALK_BLOCK, &cnt, 0u, draws);
                    stay = (sv.media[pn].flags & MEDIUM_THICK) != 0u;
                }
                if (stay) {
                    const Medium& mn = sv.media[pn];
                    const uint32_t med_rank = mn.rank;
                    double ts;
                    uint32_t ps;
                    if (ENTRIES || (pn == medium && mn.n_entry != MEDIUM_NO_ENTRIES)) {
                        // both ends of the segment lie inside the same boundary: only the world leaves that overlap it can be met
                        for (uint32_t e = 0; e < mn.n_entry && stay; e++) stay = !leaf_beats_scatter(sv, mn.entry[e], nr, 1e-8, tn, med_rank);
                    } else if (closest_hit<false, true>(sv, sv.world_root, nr, 1e-8, tn, s_mem, stack, WALK_BLOCK, ts, ps, &cnt)) {
                        stay = !(ts < tn || sv.meta[ps].rank < med_rank);
                    }
                }
                if (stay) {
                    r = nr, beta = nbeta, segment++, t = tn, medium = pn;
                    walked++;
                } else {
                    const uint32_t npos = queue_reserve(&W.counters->n_extend[W.parity ^ 1u], 0);
                    store_ray(W.ray_q[W.parity ^ 1u] + npos, nr, pack_ids(pixel, sidx, segment + 1u));
                    store_beta(W.beta_q[W.parity ^ 1u] + npos, nbeta);
                    have = false;
                }
            }
        }
    }
    if (walked) {
        atomicAdd(&W.counters->segments, walked);
        atomicAdd(&W.counters->walk_segments, walked);
    }
}
static void launch_walk(const SceneView& sv, const RenderParams& P, const WavefrontState& W, int grid, cudaStream_t s) {
    SceneView wv = sv;
    wv.n_cached_nodes = 0;  // the non-persistent closest_hit() reads the binary tree from global memory
    const size_t smem = (size_t)std::max(sv.tail_stack_entries, 4u) * WALK_BLOCK * sizeof(uint32_t);
    auto k = sv.media_xform ? (sv.walk_entries_only ? k_walk<true, true> : k_walk<true, false>) : (sv.walk_entries_only ? k_walk<false, true> : k_walk<false, false>);
    k<<<grid, WALK_BLOCK, smem, s>>>(wv, P, W);
}

// ------------------------------------------------------------------------------------------
// tail: the last few thousand paths of a frame, traced to completion by one launch
// ------------------------------------------------------------------------------------------
// Once every camera path has been generated the wavefront only drains: depth-40 stragglers keep it alive for dozens of
// iterations of ~14 launches each that move a few hundred paths - a fixed cost per frame that does not shrink when the frame
// is split over GPUs (round 1: 5.6 % of the 8-GPU step).  k_tail is launched after every iteration and returns at once unless
// generation is over and at most `threshold` paths are left; then one thread takes one path and runs generate-less
// extend -> media -> shade in a loop until the path ends.  The draws are addressed by (pixel, sample, segment, slot), so the
// radiance is the very sum the wavefront would have produced.  The last CTA to finish marks the queue empty.
constexpr int TAIL_BLOCK = MEDIA_BLOCK;
__global__ void __launch_bounds__(TAIL_BLOCK, 1) k_tail(SceneView sv, RenderParams P, WavefrontState W, uint32_t threshold) {
    extern __shared__ float4 s_mem[];  // traversal stacks (global-memory nodes only)
    uint32_t* stack = reinterpret_cast<uint32_t*>(s_mem) + threadIdx.x;
    Counters* c = W.counters;
    // the bookkeeping after shade (nobody in this kernel reads what it resets: the consumed stream's length and the class queues)
    if (blockIdx.x == 0 && threadIdx.x == 0) reset_consumed_queues(W);
    const uint32_t np = W.parity ^ 1u;  // shade has just appended the survivors to the other copy of the streams
    const uint32_t n = c->n_extend[np];
    if (n == 0 || n > threshold || c->next_path < W.total_paths) return;
    const RayRec* __restrict__ rays = W.ray_q[np];
    const BetaRec* __restrict__ betas = W.beta_q[np];
    TraceCounters cnt{0, 0};
    unsigned long long my_segments = 0;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        RayD r;
        uint64_t ids64;
        load_ray(rays + j, r, ids64);
        uint32_t pixel, sidx, segment;
        unpack_ids(ids64, pixel, sidx, segment);
        const double2* sp2 = reinterpret_cast<const double2*>(betas + j);
        const double2 b0 = sp2[0], b1 = sp2[1];
        D3 beta = D3{b0.x, b0.y, b1.x};
        while (true) {
            my_segments++;
            double t = INFINITY;
            uint32_t prim = 0xFFFFFFFFu, kind = HIT_MISS;
            if (closest_hit<false, true>(sv, sv.world_root, r, 1e-8, INFINITY, s_mem, stack, TAIL_BLOCK, t, prim, &cnt)) kind = HIT_SURFACE;
            if (sv.n_media) sample_media<false, true, true>(sv, P.seed, r, pixel, sidx, segment, t, prim, kind, s_mem, stack, TAIL_BLOCK, &cnt);
            RayD nr;
            D3 nbeta;
            if (!shade_one<SC_OTHER>(sv, P, W, r, beta, pixel, sidx, segment, t, prim, kind, nr, nbeta)) break;
            r = nr, beta = nbeta, segment++;
        }
    }
    if (my_segments) atomicAdd(&c->segments, my_segments);
    // every CTA read `n` before it got here, and the last one to arrive does so after all the others have finished
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&c->pad, 1u) == gridDim.x - 1u) {
            c->n_extend[np] = 0;
            c->pad = 0;
            c->iterations += 1;
        }
    }
}
void launch_tail(const SceneView& sv, const RenderParams& P, const WavefrontState& W, uint32_t threshold, int grid, cudaStream_t s) {
    SceneView tv = sv;
    tv.n_cached_nodes = 0;  // the non-persistent closest_hit() reads the binary tree from global memory
    const size_t smem = (size_t)std::max(sv.tail_stack_entries, 4u) * TAIL_BLOCK * sizeof(uint32_t);
    k_tail<<<grid, TAIL_BLOCK, smem, s>>>(tv, P, W, threshold);
}

__global__ void k_finalize(const double* __restrict__ accum, uint64_t n, double scale, void* out, int out_f64) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        double v = accum[i] * scale;
        if (out_f64)
            reinterpret_cast<double*>(out)[i] = v;
        else
            reinterpret_cast<float*>(out)[i] = (float)v;
    }
}


// The multi-GPU reduce: GPU 0 sums the partial framebuffers where they lie.  parts.p[k] for k > 0 are PEER pointers (memory of
// the other GPUs mapped through cudaDeviceEnablePeerAccess), so the loads below cross NVLink / NVSwitch; every value is read once
// and only GPU 0 writes.  The partitions are disjoint pixel tiles, so most addends are zero and the sum is also the gather.
template <class T>
__global__ void __launch_bounds__(256) k_sum_parts(PartList parts, T* __restrict__ out, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        T acc = reinterpret_cast<const T*>(parts.p[0])[i];
        for (uint32_t k = 1; k < parts.n; k++) acc += reinterpret_cast<const T*>(parts.p[k])[i];
        out[i] = acc;
    }
}
void launch_sum_parts(const PartList& parts, void* out, uint64_t n_values, bool f64, int grid, cudaStream_t s) {
    if (f64)
        k_sum_parts<double><<<grid, 256, 0, s>>>(parts, reinterpret_cast<double*>(out), n_values);
    else
        k_sum_parts<float><<<grid, 256, 0, s>>>(parts, reinterpret_cast<float*>(out), n_values);
}

// Color::to_rgb, utils/color.rs:14-36: optional ACES fit, then linear -> sRGB 8 bit.  palette's
// encoder is not vendored with the reference; this is the standard piecewise curve rounded to
// nearest (palette's f32 fast path may differ by one code at rounding boundaries).
__global__ void k_tonemap(const void* __restrict__ accum, int f64, uint64_t n_pixels, uint32_t toon_map, uint8_t* __restrict__ rgb, int* error_flag) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pixels; i += (uint64_t)gridDim.x * blockDim.x) {
        D3 c;
        if (f64) {
            const double* a = reinterpret_cast<const double*>(accum) + 3 * i;
            c = D3{a[0], a[1], a[2]};
        } else {
            const float* a = reinterpret_cast<const float*>(accum) + 3 * i;
            c = D3{(double)a[0], (double)a[1], (double)a[2]};
        }
        if (isnan(c.x) || isnan(c.y) || isnan(c.z)) *error_flag = 1;  // color.rs:28 assert
        if (toon_map == 1) {  // aces_tonemap, color.rs:14-25
            const double A = 2.51, C = 2.43;
            const D3 B = D3{0.03, 0.03, 0.03}, Dd = D3{0.59, 0.59, 0.59}, E = D3{0.14, 0.14, 0.14};
            D3 num = c * (A * c + B), den = c * (C * c + Dd) + E;
            D3 m = D3{num.x / den.x, num.y / den.y, num.z / den.z};
            c = D3{fmin(fmax(m.x, 0.0), 1.0), fmin(fmax(m.y, 0.0), 1.0), fmin(fmax(m.z, 0.0), 1.0)};
        }
        double v[3] = {c.x, c.y, c.z};
        for (int k = 0; k < 3; k++) {
            double x = v[k];
            double enc = x <= 0.0031308 ? 12.92 * x : 1.055 * pow(x, 1.0 / 2.4) - 0.055;
            if (!(enc > 0.0)) enc = 0.0;
            if (enc > 1.0) enc = 1.0;
            rgb[3 * i + k] = (uint8_t)llrint(floor(enc * 255.0 + 0.5));
        }
    }
}
void launch_tonemap(const void* accum, bool f64, uint64_t n_pixels, uint32_t toon_map, uint8_t* rgb, int* error_flag, cudaStream_t s) {
    int grid = (int)((n_pixels + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    if (grid < 1) grid = 1;
    k_tonemap<<<grid, 256, 0, s>>>(accum, f64 ? 1 : 0, n_pixels, toon_map, rgb, error_flag);
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
void launch_init(const WavefrontState& W, const RenderParams& P, int grid, cudaStream_t s) {
    k_pixel_list<<<grid, 256, 0, s>>>(P.cam.image_width, P.cam.image_height, P.part_index, P.part_count, W.pixel_list, &W.counters->n_pixels);
}
void launch_generate(const SceneView& sv, const RenderParams& P, const WavefrontState& W, int grid, cudaStream_t s) {
    if (P.media_first == 1 && sv.n_media && P.sample_in_generate) {
        grid = grid / 4 * RT_GEN_MEDIA_MIN_BLOCKS;  // api.cu sizes the grid for four resident CTAs per SM
        sv.media_xform ? k_generate<true, true><<<grid, 256, 0, s>>>(sv, P, W) : k_generate<true, false><<<grid, 256, 0, s>>>(sv, P, W);
    } else
        k_generate<false, false><<<grid, 256, 0, s>>>(sv, P, W);
}
template <int MEDIA>
static void launch_extend_media(const SceneView& sv, const RenderParams& P, const WavefrontState& W, bool count, int grid, size_t stack_bytes, cudaStream_t s) {
    auto k = sv.nodes4 ? (sv.n_cached_nodes && !sv.fifo_slots ? (count ? k_extend<true, true, 2, MEDIA> : k_extend<false, true, 2, MEDIA>)
                                            : (count ? k_extend<true, true, 1, MEDIA> : k_extend<false, true, 1, MEDIA>))
             : (!sv.park_leaves && !sv.fifo_slots && sv.n_cached_nodes == sv.n_nodes) ? (count ? k_extend<true, false, 3, MEDIA> : k_extend<false, false, 3, MEDIA>)
             : count   ? (sv.park_leaves ? k_extend<true, true, 0, MEDIA> : k_extend<true, false, 0, MEDIA>)
                       : (sv.park_leaves ? k_extend<false, true, 0, MEDIA> : k_extend<false, false, 0, MEDIA>);
    launch_with_l2_window(k, sv, grid, EXTEND_BLOCK, stack_bytes, s, sv, P, W);
}
void launch_extend(const SceneView& sv, const RenderParams& P, const WavefrontState& W, bool count, int grid, size_t stack_bytes, cudaStream_t s) {
    if (P.media_first == 2 && sv.n_media)
        sv.media_xform ? launch_extend_media<2>(sv, P, W, count, grid, stack_bytes, s) : launch_extend_media<1>(sv, P, W, count, grid, stack_bytes, s);
    else if (P.media_first == 1 && sv.n_media)
        launch_extend_media<3>(sv, P, W, count, grid, stack_bytes, s);
    else
        launch_extend_media<0>(sv, P, W, count, grid, stack_bytes, s);
}
int launch_media_bin(const SceneView& sv, const RenderParams& P, const WavefrontState& W, bool count, bool generic, int grid, cudaStream_t s, int phase) {
    if (phase == 1) {  // media-first order: the sampling pass ahead of extend (sphere boundaries)
        const bool xf = sv.media_xform != 0;
        auto k = count ? (xf ? k_media<true, 1, true, true> : k_media<true, 1, false, true>) : (xf ? k_media<false, 1, true, true> : k_media<false, 1, false, true>);
        k<<<grid, MEDIA_BLOCK, 0, s>>>(sv, P, W);
        return 1;
    }
    int launches = 1;
    if (phase == 0 && sv.n_media) {  // the media pass after extend: global-memory nodes only (its shared memory holds just the stacks)
        SceneView mv = sv;
        mv.n_cached_nodes = 0;
        const size_t media_smem = generic ? (size_t)sv.media_stack_entries * MEDIA_BLOCK * sizeof(uint32_t) : 0;  // <= 64 KB, opted in by kernel_setup
        const bool xf = sv.media_xform != 0;
        if (generic) {
            auto k = count ? (xf ? k_media<true, 2, true, false> : k_media<true, 2, false, false>) : (xf ? k_media<false, 2, true, false> : k_media<false, 2, false, false>);
            k<<<grid, MEDIA_BLOCK, media_smem, s>>>(mv, P, W);
        } else {
            auto k = count ? (xf ? k_media<true, 1, true, false> : k_media<true, 1, false, false>) : (xf ? k_media<false, 1, true, false> : k_media<false, 1, false, false>);
            k<<<grid, MEDIA_BLOCK, 0, s>>>(mv, P, W);
        }
        launches++;
    }
    k_bin<<<grid, MEDIA_BLOCK, 0, s>>>(W);
    return launches;
}
template <uint32_t CLS>
static void launch_shade_cls(const SceneView& sv, const RenderParams& P, const WavefrontState& W, uint32_t queue, int grid, cudaStream_t s) {
    // `grid` is sized for RT_SHADE_MIN_BLOCKS resident CTAs per SM; classes compiled for three get the matching grid
    if ((RT_SHADE_3BLOCK_MASK >> CLS) & 1u) grid = grid / RT_SHADE_MIN_BLOCKS * 3;
    k_shade<CLS><<<grid, SHADE_BLOCK, 0, s>>>(sv, P, W, queue);
}
// class_mask: bit c set when the scene can produce hits of shade class c (the miss queue always runs).
// The class kernels are independent (own queue each, atomics on the shared outputs): with `fan` they are dealt
// over side streams between a fork and a join event so that the tail of one overlaps the start of the next.
int launch_shade(const SceneView& sv, const RenderParams& P, const WavefrontState& W, uint32_t class_mask, int grid, cudaStream_t s,
                 const ShadeFan* fan, bool tail_follows, int walk_grid) {
    int launches = tail_follows ? 1 : 2;  // the miss kernel (and k_step)
    if (!P.bin_by_class) {  // everything but misses sits in the SC_DIFFUSE queue: general kernel
        launch_shade_cls<SC_MISS>(sv, P, W, SC_MISS, grid, s);
        launch_shade_cls<SC_OTHER>(sv, P, W, SC_DIFFUSE, grid, s);
        launches++;
    } else {
        const int n_side = fan ? fan->n_side : 0;
        if (n_side) {
            cudaEventRecord(fan->fork, s);
            for (int i = 0; i < n_side; i++) cudaStreamWaitEvent(fan->side[i], fan->fork, 0);
        }
        int next = 0;
        auto lane = [&]() {  // main stream first, then the side streams, round robin
            const int k = next++ % (n_side + 1);
            return k == 0 ? s : fan->side[k - 1];
        };
        // the random walk runs longest (one launch takes its paths through dozens of segments): start it first
        if (class_mask & (1u << SC_WALK)) launch_walk(sv, P, W, walk_grid, lane()), launches++;
        // the classes that usually hold most hits first, so they start on different streams
        if (class_mask & (1u << SC_DIFFUSE)) launch_shade_cls<SC_DIFFUSE>(sv, P, W, SC_DIFFUSE, grid, lane()), launches++;
        if (class_mask & (1u << SC_ISOTROPIC)) launch_shade_cls<SC_ISOTROPIC>(sv, P, W, SC_ISOTROPIC, grid, lane()), launches++;
        if (class_mask & (1u << SC_DISNEY)) launch_shade_cls<SC_DISNEY>(sv, P, W, SC_DISNEY, grid, lane()), launches++;
        if (class_mask & (1u << SC_TEXTURED)) launch_shade_cls<SC_TEXTURED>(sv, P, W, SC_TEXTURED, grid, lane()), launches++;
        launch_shade_cls<SC_MISS>(sv, P, W, SC_MISS, grid, lane());
        if (class_mask & (1u << SC_DIELECTRIC)) launch_shade_cls<SC_DIELECTRIC>(sv, P, W, SC_DIELECTRIC, grid, lane()), launches++;
        if (class_mask & (1u << SC_EMISSIVE)) launch_shade_cls<SC_EMISSIVE>(sv, P, W, SC_EMISSIVE, grid, lane()), launches++;
        if (class_mask & (1u << SC_METAL)) launch_shade_cls<SC_METAL>(sv, P, W, SC_METAL, grid, lane()), launches++;
        if (class_mask & (1u << SC_OTHER)) launch_shade_cls<SC_OTHER>(sv, P, W, SC_OTHER, grid, lane()), launches++;
        for (int i = 0; i < n_side; i++) {
            cudaEventRecord(fan->join[i], fan->side[i]);
            cudaStreamWaitEvent(s, fan->join[i], 0);
        }
    }
    if (!tail_follows) k_step<<<1, 1, 0, s>>>(W);  // otherwise k_tail, launched next, resets the consumed queues
    return launches;
}
void launch_finalize(const double* accum, uint64_t n, double scale, void* out, bool out_f64, int grid, cudaStream_t s) {
    k_finalize<<<grid, 256, 0, s>>>(accum, n, scale, out, out_f64 ? 1 : 0);
}

int kernel_setup(size_t smem_bytes, int* extend_blocks_per_sm, int* shade_blocks_per_sm, int* walk_blocks_per_sm) {
    // dynamic shared memory above 48 KB is opt-in
    cudaError_t e = cudaSuccess;
    const void* big_smem[] = {
#define RT_EXTEND_VARIANTS(M)                                                                                                     \
    (const void*)k_extend<false, false, 0, M>, (const void*)k_extend<false, true, 0, M>, (const void*)k_extend<false, true, 1, M>, \
        (const void*)k_extend<false, true, 2, M>, (const void*)k_extend<true, false, 0, M>, (const void*)k_extend<true, true, 0, M>, \
        (const void*)k_extend<true, true, 1, M>, (const void*)k_extend<true, true, 2, M>, (const void*)k_extend<false, false, 3, M>, \
        (const void*)k_extend<true, false, 3, M>
        RT_EXTEND_VARIANTS(0), RT_EXTEND_VARIANTS(1), RT_EXTEND_VARIANTS(2), RT_EXTEND_VARIANTS(3),
#undef RT_EXTEND_VARIANTS
        (const void*)k_closest_hit<false, false, 0>, (const void*)k_closest_hit<false, true, 0>, (const void*)k_closest_hit<false, true, 1>,
        (const void*)k_closest_hit<false, true, 2>,  (const void*)k_closest_hit<true, false, 0>, (const void*)k_closest_hit<true, true, 0>,
        (const void*)k_closest_hit<true, true, 1>,   (const void*)k_closest_hit<true, true, 2>,  (const void*)k_closest_hit<false, false, 3>,
        (const void*)k_closest_hit<true, false, 3>,  (const void*)k_closest_hit<false, true, 1, false, 0>,  (const void*)k_closest_hit<true, true, 1, false, 0>,
        (const void*)k_closest_hit<false, true, 1, false, 1>,  (const void*)k_closest_hit<true, true, 1, false, 1>,
        (const void*)k_closest_hit<false, true, 1, false, 2>,  (const void*)k_closest_hit<true, true, 1, false, 2>};
    for (const void* f : big_smem)
        if ((e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(EXTEND_SMEM_MAX - 1024))) != cudaSuccess) return (int)e;
    // the general-boundary media pass: TRAVERSAL_STACK entries per thread is 64 KB
    if ((e = cudaFuncSetAttribute((const void*)k_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, TRAVERSAL_STACK * TAIL_BLOCK * (int)sizeof(uint32_t))) != cudaSuccess) return (int)e;
    for (const void* f : {(const void*)k_walk<false, false>, (const void*)k_walk<true, false>, (const void*)k_walk<false, true>, (const void*)k_walk<true, true>})
        if ((e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, TRAVERSAL_STACK * WALK_BLOCK * (int)sizeof(uint32_t))) != cudaSuccess) return (int)e;
    const void* media_smem[] = {(const void*)k_media<false, 2, false, false>, (const void*)k_media<false, 2, true, false>,
                                (const void*)k_media<true, 2, false, false>, (const void*)k_media<true, 2, true, false>};
    for (const void* f : media_smem)
        if ((e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, TRAVERSAL_STACK * MEDIA_BLOCK * (int)sizeof(uint32_t))) != cudaSuccess) return (int)e;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(extend_blocks_per_sm, k_extend<false, false, 0, 0>, EXTEND_BLOCK, smem_bytes);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(shade_blocks_per_sm, k_shade<SC_OTHER>, SHADE_BLOCK, 0);  // a class never compiled for three
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(walk_blocks_per_sm, k_walk<true, false>, WALK_BLOCK, 24 * WALK_BLOCK * sizeof(uint32_t));
    return 0;
}

}  // namespace rt

// lbvh.cu — device BVH builder for the world group (the reference has no counterpart,
// bvh.rs:16-46 only fixes the SEMANTICS, which the tie ranks carry).
//
// Linear BVH after Karras 2012: 63-bit Morton codes of the box centres, one radix sort (CUB), the binary radix
// tree built with one thread per interior node, then a bottom-up pass that unions the conservative binary32 boxes
// and writes the 64-byte nodes of scene_types.h (both children's boxes in the parent).  One primitive per leaf.
// Selected with RT_BUILD_DEVICE_LBVH: the tree is built in a fraction of a second where the host SAH builder
// takes 17 s (100 M triangles), at the price of a tree that traverses slower (profiles/README.md); ids and t do not
// depend on the tree.  Any failure (CUDA error, a tree deeper than the traversal stack) makes the caller fall back
// to the host builder.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/device/device_radix_sort.cuh>

#include "compile.h"

namespace rt {
namespace {

struct Bounds6 {
    int lo[3], hi[3];  // order-preserving integer images of binary32 values
};
__device__ __forceinline__ int f2ord(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

__global__ void k_centroid_bounds(const BuildBox* __restrict__ boxes, uint32_t n, Bounds6* out) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        for (int a = 0; a < 3; a++) {
            const float c = 0.5f * boxes[i].lo[a] + 0.5f * boxes[i].hi[a];
            lo[a] = fminf(lo[a], c), hi[a] = fmaxf(hi[a], c);
        }
    for (int a = 0; a < 3; a++) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xFFFFFFFFu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xFFFFFFFFu, hi[a], o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&out->lo[a], f2ord(lo[a]));
            atomicMax(&out->hi[a], f2ord(hi[a]));
        }
    }
}

__device__ __forceinline__ uint64_t spread21(uint64_t x) {  // bit i -> bit 3i
    x &= 0x1FFFFFull;
    x = (x | x << 32) & 0x1F00000000FFFFull;
    x = (x | x << 16) & 0x1F0000FF0000FFull;
    x = (x | x << 8) & 0x100F00F00F00F00Full;
    x = (x | x << 4) & 0x10C30C30C30C30C3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void k_morton(const BuildBox* __restrict__ boxes, uint32_t n, const Bounds6* __restrict__ b, uint64_t* codes, uint32_t* idx) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint64_t code = 0;
        for (int a = 0; a < 3; a++) {
            const float lo = ord2f(b->lo[a]), hi = ord2f(b->hi[a]);
            const float c = 0.5f * boxes[i].lo[a] + 0.5f * boxes[i].hi[a];
            const float ext = hi - lo;
            float u = ext > 0.f ? (c - lo) / ext : 0.f;
            u = fminf(fmaxf(u, 0.f), 1.f);
            const uint32_t q = min(2097151u, (uint32_t)(u * 2097152.f));
            code |= spread21(q) << (2 - a);
        }
        codes[i] = code;
        idx[i] = i;
    }
}

// common prefix length of the keys of sorted positions i and j (ties broken by the position itself)
__device__ __forceinline__ int delta(const uint64_t* __restrict__ codes, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t a = codes[i], b = codes[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll((long long)(a ^ b));
}

// Karras 2012, one thread per interior node.  children: bit 31 set = leaf (sorted position), else interior index.
__global__ void k_hierarchy(const uint64_t* __restrict__ codes, int n, uint32_t* child0, uint32_t* child1, int* node_parent, int* leaf_parent) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n - 1; i += gridDim.x * blockDim.x) {
        const int d = delta(codes, n, i, i + 1) - delta(codes, n, i, i - 1) >= 0 ? 1 : -1;
        const int dmin = delta(codes, n, i, i - d);
        int lmax = 2;
        while (delta(codes, n, i, i + lmax * d) > dmin) lmax <<= 1;
        int l = 0;
        for (int t = lmax >> 1; t >= 1; t >>= 1)
            if (delta(codes, n, i, i + (l + t) * d) > dmin) l += t;
        const int j = i + l * d;
        const int dnode = delta(codes, n, i, j);
        int s = 0, t = l;
        do {
            t = (t + 1) >> 1;
            if (delta(codes, n, i, i + (s + t) * d) > dnode) s += t;
        } while (t > 1);
        const int gamma = i + s * d + min(d, 0);
        const int first = min(i, j), last = max(i, j);
        if (first == gamma) {
            child0[i] = 0x80000000u | (uint32_t)gamma;
            leaf_parent[gamma] = i;
        } else {
            child0[i] = (uint32_t)gamma;
            node_parent[gamma] = i;
        }
        if (last == gamma + 1) {
            child1[i] = 0x80000000u | (uint32_t)(gamma + 1);
            leaf_parent[gamma + 1] = i;
        } else {
            child1[i] = (uint32_t)(gamma + 1);
            node_parent[gamma + 1] = i;
        }
        if (i == 0) node_parent[0] = -1;
    }
}

struct Box6 {
    float lo[3], hi[3];
};

// bottom-up: the second thread to reach a node finds both children finished, writes the node and goes on
__global__ void k_refit(const BuildBox* __restrict__ boxes, const uint32_t* __restrict__ idx, int n, const uint32_t* __restrict__ child0,
                        const uint32_t* __restrict__ child1, const int* __restrict__ node_parent, const int* __restrict__ leaf_parent,
                        uint32_t* visits, Box6* node_box, uint8_t* height, Node* nodes, uint32_t node_base, uint32_t prim_base) {
    for (int leaf = blockIdx.x * blockDim.x + threadIdx.x; leaf < n; leaf += gridDim.x * blockDim.x) {
        int p = leaf_parent[leaf];
        while (p >= 0) {
            if (atomicAdd(&visits[p], 1u) == 0u) break;  // the sibling subtree is not finished yet
            __threadfence();
            Node nd;
            Box6 me;
            uint32_t h = 0;
            const uint32_t c[2] = {child0[p], child1[p]};
            for (int k = 0; k < 2; k++) {
                Box6 b;
                uint32_t ref;
                if (c[k] & 0x80000000u) {
                    const uint32_t pos = c[k] & 0x7FFFFFFFu;
                    const BuildBox& bb = boxes[idx[pos]];
                    for (int a = 0; a < 3; a++) b.lo[a] = bb.lo[a], b.hi[a] = bb.hi[a];
                    ref = LEAF_FLAG | ((prim_base + pos) << 3);
                } else {
                    // written by another SM a moment ago: read past the (incoherent) L1
                    const float* src = reinterpret_cast<const float*>(node_box + c[k]);
                    for (int a = 0; a < 3; a++) b.lo[a] = __ldcg(src + a), b.hi[a] = __ldcg(src + 3 + a);
                    ref = node_base + c[k];
                    h = max(h, (uint32_t)__ldcg(height + c[k]));
                }
                float* lo = k == 0 ? nd.lo0 : nd.lo1;
                float* hi = k == 0 ? nd.hi0 : nd.hi1;
                for (int a = 0; a < 3; a++) {
                    lo[a] = b.lo[a], hi[a] = b.hi[a];
                    me.lo[a] = k == 0 ? b.lo[a] : fminf(me.lo[a], b.lo[a]);
                    me.hi[a] = k == 0 ? b.hi[a] : fmaxf(me.hi[a], b.hi[a]);
                }
                (k == 0 ? nd.child0 : nd.child1) = ref;
            }
            nd.pad0 = nd.pad1 = 0;
            nodes[p] = nd;
            node_box[p] = me;
            height[p] = (uint8_t)min(255u, h + 1u);
            __threadfence();
            p = node_parent[p];
        }
    }
}

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    bool alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1) == cudaSuccess; }
    template <class T>
    T* as() const {
        return reinterpret_cast<T*>(p);
    }
};

}  // namespace

bool build_bvh_device(const RawVec<BuildBox>& boxes, uint32_t first_prim_base, RawVec<Node>& nodes, RawVec<uint32_t>& order,
                      uint32_t& depth_out, uint32_t& root_out) {
    const uint32_t n = (uint32_t)boxes.size();
    if (n < 2 || n >= (1u << 28)) return false;  // trivial inputs go to the host builder
    const uint32_t node_base = (uint32_t)nodes.size();
    DevBuf d_boxes, d_bounds, d_codes, d_codes2, d_idx, d_idx2, d_tmp, d_c0, d_c1, d_np, d_lp, d_visits, d_nbox, d_height, d_nodes;
    if (!d_boxes.alloc((size_t)n * sizeof(BuildBox)) || !d_bounds.alloc(sizeof(Bounds6)) || !d_codes.alloc((size_t)n * 8) ||
        !d_codes2.alloc((size_t)n * 8) || !d_idx.alloc((size_t)n * 4) || !d_idx2.alloc((size_t)n * 4) || !d_c0.alloc((size_t)n * 4) ||
        !d_c1.alloc((size_t)n * 4) || !d_np.alloc((size_t)n * 4) || !d_lp.alloc((size_t)n * 4) || !d_visits.alloc((size_t)n * 4) ||
        !d_nbox.alloc((size_t)n * sizeof(Box6)) || !d_height.alloc(n) || !d_nodes.alloc((size_t)n * sizeof(Node))) {
        cudaGetLastError();
        return false;
    }
    const int block = 256;
    int sms = 148, device = 0;
    cudaGetDevice(&device);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int grid = sms * 8;
    Bounds6 init;
    for (int a = 0; a < 3; a++) init.lo[a] = 0x7FFFFFFF, init.hi[a] = (int)0x80000000;
    bool ok = cudaMemcpy(d_boxes.p, boxes.data(), (size_t)n * sizeof(BuildBox), cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(d_bounds.p, &init, sizeof(init), cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) return false;
    k_centroid_bounds<<<grid, block>>>(d_boxes.as<BuildBox>(), n, d_bounds.as<Bounds6>());
    k_morton<<<grid, block>>>(d_boxes.as<BuildBox>(), n, d_bounds.as<Bounds6>(), d_codes.as<uint64_t>(), d_idx.as<uint32_t>());
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_codes.as<uint64_t>(), d_codes2.as<uint64_t>(), d_idx.as<uint32_t>(), d_idx2.as<uint32_t>(),
                                    (int)n, 0, 63);
    if (!d_tmp.alloc(tmp_bytes)) return false;
    cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, d_codes.as<uint64_t>(), d_codes2.as<uint64_t>(), d_idx.as<uint32_t>(), d_idx2.as<uint32_t>(),
                                    (int)n, 0, 63);
    cudaMemset(d_visits.p, 0, (size_t)n * 4);
    cudaMemset(d_height.p, 0, n);
    k_hierarchy<<<grid, block>>>(d_codes2.as<uint64_t>(), (int)n, d_c0.as<uint32_t>(), d_c1.as<uint32_t>(), d_np.as<int>(), d_lp.as<int>());
    k_refit<<<grid, block>>>(d_boxes.as<BuildBox>(), d_idx2.as<uint32_t>(), (int)n, d_c0.as<uint32_t>(), d_c1.as<uint32_t>(), d_np.as<int>(),
                             d_lp.as<int>(), d_visits.as<uint32_t>(), d_nbox.as<Box6>(), d_height.as<uint8_t>(), d_nodes.as<Node>(), node_base,
                             first_prim_base);
    if (cudaDeviceSynchronize() != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    uint8_t root_height = 0;
    if (cudaMemcpy(&root_height, d_height.p, 1, cudaMemcpyDeviceToHost) != cudaSuccess) return false;
    if (root_height == 0 || root_height + 2 > TRAVERSAL_STACK) return false;  // deeper than the traversal stack: host builder
    nodes.resize((size_t)node_base + n - 1);
    order.resize(n);
    if (cudaMemcpy(nodes.data() + node_base, d_nodes.p, (size_t)(n - 1) * sizeof(Node), cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(order.data(), d_idx2.p, (size_t)n * 4, cudaMemcpyDeviceToHost) != cudaSuccess) {
        nodes.resize(node_base);
        return false;
    }
    depth_out = std::max<uint32_t>(depth_out, root_height);
    root_out = node_base;  // interior node 0 of the radix tree is the root
    return true;
}

}  // namespace rt

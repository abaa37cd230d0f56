// traverse.cuh — closest-hit traversal of the flat BVH2 with binary64 primitive tests.
//
// Box tests run in binary32 but are CONSERVATIVE: boxes are rounded outwards at build time and
// every slab distance is widened by a bound on its rounding error, so a box that the exact test
// would enter is never culled.  The primitive tests are the reference's own arithmetic in
// binary64 (sphere.rs:77-108, quad.rs:71-102, triangle.rs:56-98), which makes the winning t
// bit-identical to the CPU restatement for unbaked primitives.  Exact ties are resolved by the
// per-primitive rank (hits.rs:42: first child wins; bvh.rs:78-84: right child wins).
#pragma once
#include "device_math.cuh"
#include "scene_types.h"

namespace rt {

struct RayD {
    D3 o, d;
    double time;
};

// binary32 image of the ray for the slab test: t = fma(plane, idf, noidf)
struct RayF {
    float idx, idy, idz;     // 1/d, clamped to +-1e30
    float nox, noy, noz;     // -o * idf (rounded once from binary64)
    float ex, ey, ez;        // per-axis absolute error bound of the slab distances
};

__device__ __forceinline__ void make_rayf(const RayD& r, RayF& f) {
    auto inv = [](double d) {
        double i = 1.0 / d;
        if (!(fabs(i) <= 1e30)) i = copysign(1e30, d);
        return (float)i;
    };
    f.idx = inv(r.d.x), f.idy = inv(r.d.y), f.idz = inv(r.d.z);
    f.nox = (float)(-r.o.x * (double)f.idx);
    f.noy = (float)(-r.o.y * (double)f.idy);
    f.noz = (float)(-r.o.z * (double)f.idz);
    // |t_computed - t_exact| <= (|t| + |o*idf|) * 2^-24 ; the |t| part is applied as a relative slack
    const float k = 1.0f / 4194304.0f;  // 2^-22
    f.ex = fabsf(f.nox) * k, f.ey = fabsf(f.noy) * k, f.ez = fabsf(f.noz) * k;
}

// returns true when the box may intersect the ray within [tmin_f, tmax_f]; tnear for ordering
__device__ __forceinline__ bool slab(const RayF& f, const float* lo, const float* hi, float tmin_f, float tmax_f, float& tnear) {
    float x0 = __fmaf_rn(lo[0], f.idx, f.nox), x1 = __fmaf_rn(hi[0], f.idx, f.nox);
    float y0 = __fmaf_rn(lo[1], f.idy, f.noy), y1 = __fmaf_rn(hi[1], f.idy, f.noy);
    float z0 = __fmaf_rn(lo[2], f.idz, f.noz), z1 = __fmaf_rn(hi[2], f.idz, f.noz);
    float tn = fmaxf(fmaxf(fminf(x0, x1) - f.ex, fminf(y0, y1) - f.ey), fminf(z0, z1) - f.ez);
    float tf = fminf(fminf(fmaxf(x0, x1) + f.ex, fmaxf(y0, y1) + f.ey), fmaxf(z0, z1) + f.ez);
    const float rel = 1.0f / 2097152.0f;  // 2^-21
    tn = __fmaf_rn(-fabsf(tn), rel, tn);
    tf = __fmaf_rn(fabsf(tf), rel, tf);
    tnear = tn;
    // written with negations so that a NaN anywhere keeps the box (conservative)
    return !(tn > tf) && !(tn > tmax_f) && !(tf < tmin_f);
}

__device__ __forceinline__ bool contains(double mn, double mx, double x) { return x >= mn && x <= mx; }  // interval.rs:65-67

// Sphere::hit, sphere.rs:77-108 (geometry only)
__device__ __forceinline__ bool sphere_hit(const double* __restrict__ g, const RayD& r, double tmin, double tmax, double& t_out) {
    const double2 a0 = __ldg(reinterpret_cast<const double2*>(g));
    const double2 a1 = __ldg(reinterpret_cast<const double2*>(g) + 1);
    const double2 a2 = __ldg(reinterpret_cast<const double2*>(g) + 2);
    const double2 a3 = __ldg(reinterpret_cast<const double2*>(g) + 3);
    D3 center = D3{a0.x, a0.y, a1.x}, cvec = D3{a1.y, a2.x, a2.y};
    double radius = a3.x;
    D3 current_center = center + r.time * cvec;
    D3 oc = current_center - r.o;
    double a = length_squared(r.d);
    double h = dot(r.d, oc);
    double c = length_squared(oc) - radius * radius;
    double discriminant = h * h - a * c;
    if (discriminant < 0.0) return false;
    double sqrtd = sqrt(discriminant);
    double root = (h - sqrtd) / a;
    if (!contains(tmin, tmax, root)) {
        root = (h + sqrtd) / a;
        if (!contains(tmin, tmax, root)) return false;
    }
    t_out = root;
    return true;
}

struct Planar {
    D3 q, u, v, n, w;
    double D;
};
__device__ __forceinline__ void load_planar(const double* __restrict__ g, Planar& p) {
    const double2* s = reinterpret_cast<const double2*>(g);
    double2 a0 = __ldg(s), a1 = __ldg(s + 1), a2 = __ldg(s + 2), a3 = __ldg(s + 3);
    double2 a4 = __ldg(s + 4), a5 = __ldg(s + 5), a6 = __ldg(s + 6), a7 = __ldg(s + 7);
    p.q = D3{a0.x, a0.y, a1.x};
    p.u = D3{a1.y, a2.x, a2.y};
    p.v = D3{a3.x, a3.y, a4.x};
    p.n = D3{a4.y, a5.x, a5.y};
    p.D = a6.x;
    p.w = D3{a6.y, a7.x, a7.y};
}
// Quad::hit / Triangle::hit, quad.rs:71-102, triangle.rs:56-98 (geometry only)
__device__ __forceinline__ bool planar_hit_loaded(const Planar& p, bool triangle, const RayD& r, double tmin, double tmax,
                                                  double& t_out, double& alpha, double& beta) {
    double denom = dot(p.n, r.d);
    if (fabs(denom) < 1e-8) return false;
    double t = (p.D - dot(p.n, r.o)) / denom;
    if (!contains(tmin, tmax, t)) return false;
    D3 intersection = r.o + t * r.d;
    D3 hp = intersection - p.q;
    alpha = dot(p.w, cross(hp, p.v));
    beta = dot(p.w, cross(p.u, hp));
    if (!contains(0.0, 1.0, alpha) || !contains(0.0, 1.0, beta)) return false;
    if (triangle && !contains(0.0, 1.0, alpha + beta)) return false;
    t_out = t;
    return true;
}
__device__ __forceinline__ bool planar_hit(const double* __restrict__ g, bool triangle, const RayD& r, double tmin, double tmax, double& t_out) {
    // plane first: most candidates are rejected before the rest of the record is needed
    const double2* s = reinterpret_cast<const double2*>(g);
    double2 a4 = __ldg(s + 4), a5 = __ldg(s + 5), a6 = __ldg(s + 6);
    D3 n = D3{a4.y, a5.x, a5.y};
    double denom = dot(n, r.d);
    if (fabs(denom) < 1e-8) return false;
    double t = (a6.x - dot(n, r.o)) / denom;
    if (!contains(tmin, tmax, t)) return false;
    double2 a0 = __ldg(s), a1 = __ldg(s + 1), a2 = __ldg(s + 2), a3 = __ldg(s + 3), a7 = __ldg(s + 7);
    D3 q = D3{a0.x, a0.y, a1.x}, u = D3{a1.y, a2.x, a2.y}, v = D3{a3.x, a3.y, a4.x}, w = D3{a6.y, a7.x, a7.y};
    D3 intersection = r.o + t * r.d;
    D3 hp = intersection - q;
    double alpha = dot(w, cross(hp, v));
    double beta = dot(w, cross(u, hp));
    if (!contains(0.0, 1.0, alpha) || !contains(0.0, 1.0, beta)) return false;
    if (triangle && !contains(0.0, 1.0, alpha + beta)) return false;
    t_out = t;
    return true;
}

struct TraceCounters {
    uint32_t nodes, prims;
};

// Closest hit below `root` within [tmin, tmax] (both inclusive).  stack[] is this thread's column
// of the shared-memory traversal stack (entry i at stack[i * stride]).
template <bool COUNT, bool USE_RANK>
__device__ __forceinline__ bool closest_hit(const SceneView& sv, uint32_t root, const RayD& r, double tmin, double tmax,
                                            uint32_t* __restrict__ stack, int stride, double& best_t, uint32_t& best_prim,
                                            TraceCounters* cnt) {
    if (root == INVALID_REF) return false;
    RayF f;
    make_rayf(r, f);
    const float tmin_f = __double2float_rd(tmin);
    float tmax_f = __double2float_ru(tmax);
    double tbest = tmax;
    uint32_t prim = 0xFFFFFFFFu, prim_rank = 0xFFFFFFFFu;
    int sp = 0;
    uint32_t cur = root;
    while (true) {
        if (cur & LEAF_FLAG) {
            uint32_t first = (cur & ~LEAF_FLAG) >> 3, count = (cur & 7u) + 1;
            for (uint32_t i = 0; i < count; i++) {
                uint32_t pi = first + i;
                const uint2 km = __ldg(reinterpret_cast<const uint2*>(&sv.meta[pi]));  // kind_mat, rank
                const uint32_t kind = km.x >> 30;
                const double* g = sv.geom[pi].d;
                double t;
                bool hit = kind == PRIM_SPHERE ? sphere_hit(g, r, tmin, tbest, t) : planar_hit(g, kind == PRIM_TRIANGLE, r, tmin, tbest, t);
                if (COUNT) cnt->prims++;
                if (hit) {
                    // t <= tbest here; an exact tie keeps the lower rank
                    bool better = prim == 0xFFFFFFFFu || t < tbest || (USE_RANK && km.y < prim_rank);
                    if (better) {
                        tbest = t;
                        prim = pi;
                        prim_rank = km.y;
                        tmax_f = __double2float_ru(t);
                    }
                }
            }
            if (sp == 0) break;
            cur = stack[(--sp) * stride];
        } else {
            const float4* np = reinterpret_cast<const float4*>(sv.nodes + cur);
            float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3);
            if (COUNT) cnt->nodes++;
            float lo0[3] = {n0.x, n0.y, n0.z}, hi0[3] = {n0.w, n1.x, n1.y};
            float lo1[3] = {n1.z, n1.w, n2.x}, hi1[3] = {n2.y, n2.z, n2.w};
            uint32_t c0 = __float_as_uint(n3.x), c1 = __float_as_uint(n3.y);
            float t0, t1;
            bool h0 = slab(f, lo0, hi0, tmin_f, tmax_f, t0);
            bool h1 = slab(f, lo1, hi1, tmin_f, tmax_f, t1);
            if (h0 && h1) {
                bool swap = t1 < t0;
                uint32_t nearc = swap ? c1 : c0, farc = swap ? c0 : c1;
                stack[(sp++) * stride] = farc;
                cur = nearc;
            } else if (h0) {
                cur = c0;
            } else if (h1) {
                cur = c1;
            } else {
                if (sp == 0) break;
                cur = stack[(--sp) * stride];
            }
        }
    }
    if (prim == 0xFFFFFFFFu) return false;
    best_t = tbest;
    best_prim = prim;
    return true;
}

}  // namespace rt

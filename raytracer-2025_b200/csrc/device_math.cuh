// device_math.cuh — binary64 vector algebra, Philox and the samplers, written in the reference's
// operation order (utils/vec3.rs, utils/onb.rs) so that, compiled with -fmad=false, the device
// produces the same bits as the Rust code for everything except libm calls.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rt2025_rng.h"

namespace rt {

#define RT_PI 3.14159265358979323846264338327950288

struct D3 {
    double x, y, z;
};
__device__ __forceinline__ D3 mk3(double x, double y, double z) { return D3{x, y, z}; }
__device__ __forceinline__ D3 ld3(const double* p) { return D3{p[0], p[1], p[2]}; }
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return D3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return D3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ D3 operator*(D3 a, D3 b) { return D3{a.x * b.x, a.y * b.y, a.z * b.z}; }
__device__ __forceinline__ D3 operator-(D3 a) { return D3{-a.x, -a.y, -a.z}; }
__device__ __forceinline__ D3 operator*(double s, D3 a) { return D3{s * a.x, s * a.y, s * a.z}; }  // vec3.rs:140-146
__device__ __forceinline__ D3 operator*(D3 a, double s) { return D3{a.x * s, a.y * s, a.z * s}; }  // vec3.rs:156-162
__device__ __forceinline__ D3 operator/(D3 a, double s) { return (1.0 / s) * a; }                  // vec3.rs:222-228
__device__ __forceinline__ double dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }    // vec3.rs:107-109
__device__ __forceinline__ D3 cross(D3 a, D3 b) {                                                  // vec3.rs:111-117
    return D3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ double length_squared(D3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
__device__ __forceinline__ double length(D3 a) { return sqrt(length_squared(a)); }
__device__ __forceinline__ bool finite3(D3 a) { return isfinite(a.x) && isfinite(a.y) && isfinite(a.z); }
// UnitVec3::from_vec3 (vec3.rs:303-310)
__device__ __forceinline__ bool unit_vector(D3 v, D3& out) {
    out = v / length(v);
    return finite3(out);
}
// Rust f64::min/max semantics (NaN operand ignored) == IEEE minNum/maxNum == CUDA fmin/fmax
__device__ __forceinline__ double rmin(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ double rmax(double a, double b) { return fmax(a, b); }

// 3x3 row-major times vector
__device__ __forceinline__ D3 mul33(const double* A, D3 v) {
    return D3{A[0] * v.x + A[1] * v.y + A[2] * v.z, A[3] * v.x + A[4] * v.y + A[5] * v.z, A[6] * v.x + A[7] * v.y + A[8] * v.z};
}

// ---- Philox4x32-10 (include/rt2025_rng.h) ----------------------------------------------------
struct Rand2 {
    double a, b;
};
__device__ __forceinline__ Rand2 philox_pair(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t segment, uint32_t slot) {
    uint32_t c0 = pixel, c1 = sample, c2 = segment, c3 = slot;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(RT_PHILOX_M0, c0), lo0 = RT_PHILOX_M0 * c0;
        uint32_t hi1 = __umulhi(RT_PHILOX_M1, c2), lo1 = RT_PHILOX_M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0, c1 = lo1, c2 = n2, c3 = lo0;
        k0 += RT_PHILOX_W0;
        k1 += RT_PHILOX_W1;
    }
    Rand2 r;
    r.a = (double)(((uint64_t)(c0 >> 5) << 26) | (uint64_t)(c1 >> 6)) * 0x1.0p-53;
    r.b = (double)(((uint64_t)(c2 >> 5) << 26) | (uint64_t)(c3 >> 6)) * 0x1.0p-53;
    return r;
}

// ---- samplers ----------------------------------------------------------------------------
__device__ __forceinline__ D3 random_unit_vector(double r1, double r2) {  // vec3.rs:313-322
    double s, c;
    sincos(2.0 * RT_PI * r1, &s, &c);
    double x = c * 2.0 * sqrt(r2 * (1.0 - r2));
    double y = s * 2.0 * sqrt(r2 * (1.0 - r2));
    double z = 1.0 - 2.0 * r2;
    return D3{x, y, z};
}
__device__ __forceinline__ D3 random_cosine_direction(double r1, double r2) {  // vec3.rs:333-343
    double phi = 2.0 * RT_PI * r1;
    double s, c;
    sincos(phi, &s, &c);
    double x = s * sqrt(r2);
    double y = sqrt(1.0 - r2);
    double z = c * sqrt(r2);
    return D3{x, y, z};
}
__device__ __forceinline__ D3 reflect(D3 v, D3 n) { return v - 2.0 * dot(v, n) * n; }  // vec3.rs:71-73
__device__ __forceinline__ bool refract(D3 uv, D3 n, double relative_eta, D3& out) {  // vec3.rs:345-354
    double cos_theta = rmin(dot(-uv, n), 1.0);
    D3 out_perp = relative_eta * (uv + cos_theta * n);
    double out_parallel_length = sqrt(1.0 - length_squared(out_perp));
    if (isnan(out_parallel_length)) return false;
    D3 out_parallel = -out_parallel_length * n;
    out = out_perp + out_parallel;
    return true;
}

// utils/onb.rs:8-22
struct ONB {
    D3 u, v, w;
};
__device__ __forceinline__ bool make_onb(D3 n, ONB& o) {
    D3 a = fabs(n.x) > 0.9 ? D3{0.0, 1.0, 0.0} : D3{1.0, 0.0, 0.0};
    bool ok = unit_vector(cross(n, a), o.u);
    o.v = n;
    o.w = cross(o.u, n);
    return ok;
}
__device__ __forceinline__ D3 onb_to_world(const ONB& o, D3 v) { return v.x * o.u + v.y * o.v + v.z * o.w; }  // onb.rs:34-38

}  // namespace rt

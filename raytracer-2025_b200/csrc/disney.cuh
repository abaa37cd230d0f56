// disney.cuh — the Disney BSDF of material/disney.rs and the helpers of utils/fresnel.rs on the device,
// in the reference's operation order, quirks included (see the notes on cos_theta2 / cos_phi below).
#pragma once
#include "device_math.cuh"
#include "rt2025.h"

namespace rt {
namespace disney {

struct Params {  // DisneyParameters, disney.rs:18-35
    D3 base_color;
    double roughness, anisotropic, sheen, sheen_tint, clearcoat, clearcoat_gloss, specular_tint, metallic, ior, flatness, spec_trans, diff_trans;
    bool thin;
};

__device__ __forceinline__ double lerp(double a, double b, double t) { return a * (1.0 - t) + b * t; }  // utils.rs:14-19
__device__ __forceinline__ D3 lerp(D3 a, D3 b, double t) { return a * (1.0 - t) + b * t; }
__device__ __forceinline__ double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }
__device__ __forceinline__ double pow5(double x) {
    double x2 = x * x;
    return x * (x2 * x2);
}
// UnitVec3 helpers, vec3.rs:372-420.  Reproduced as written: cos_theta2() returns y (not y*y), and
// cos_phi()/sin_phi() test `sin_theta.abs() < 1e8`, which always holds, so both are 1.
__device__ __forceinline__ double cos_theta(D3 w) { return w.y; }
__device__ __forceinline__ double sin_theta2(D3 w) { return clampd(1.0 - w.y, 0.0, 1.0); }
__device__ __forceinline__ double sin_theta(D3 w) { return sqrt(sin_theta2(w)); }
__device__ __forceinline__ double tan_theta(D3 w) { return sin_theta(w) / cos_theta(w); }
__device__ __forceinline__ double cos_phi2(D3 w) {
    double st = sin_theta(w);
    double c = fabs(st) < 1e8 ? 1.0 : w.x / st;
    return c * c;
}
__device__ __forceinline__ double sin_phi2(D3 w) {
    double st = sin_theta(w);
    double c = fabs(st) < 1e8 ? 1.0 : w.z / st;
    return c * c;
}
__device__ __forceinline__ D3 reflect2(D3 v, D3 n) { return -v + 2.0 * dot(v, n) * n; }  // vec3.rs:76-78
__device__ __forceinline__ bool refract2(D3 v, D3 n, double eta, D3& out) {             // vec3.rs:357-366
    double ct = rmin(dot(v, n), 1.0);
    D3 out_perp = eta * (-v + ct * n);
    double len = sqrt(1.0 - length_squared(out_perp));
    if (isnan(len)) return false;
    out = out_perp + (-len * n);
    return true;
}
// utils/fresnel.rs
__device__ __forceinline__ D3 schlick(D3 r0, double radians) {
    double e = pow5(1.0 - radians);
    return r0 + (D3{1.0, 1.0, 1.0} - r0) * e;
}
__device__ __forceinline__ double schlick_weight(double u) { return pow5(clampd(1.0 - u, 0.0, 1.0)); }
__device__ __forceinline__ double schlick_f64(double r0, double radians) { return lerp(1.0, schlick_weight(radians), r0); }
__device__ __forceinline__ double schlick_r0_from_relative_ior(double eta) { return ((eta - 1.0) * (eta - 1.0)) / ((eta + 1.0) * (eta + 1.0)); }
__device__ inline double dielectric(double cos_theta_in, double n_in, double n_out) {  // fresnel.rs:22-46
    cos_theta_in = clampd(cos_theta_in, -1.0, 1.0);
    if (cos_theta_in < 0.0) {
        double t = n_in;
        n_in = n_out, n_out = t;
        cos_theta_in = -cos_theta_in;
    }
    double sin_in = sqrt(rmax(1.0 - cos_theta_in * cos_theta_in, 0.0));
    double sin_out = n_in / n_out * sin_in;
    if (sin_out >= 1.0) return 1.0;
    double cos_out = sqrt(rmax(1.0 - sin_out * sin_out, 0.0));
    double r_par = (n_out * cos_theta_in - n_in * cos_out) / (n_out * cos_theta_in + n_in * cos_out);
    double r_perp = (n_in * cos_theta_in - n_out * cos_out) / (n_in * cos_theta_in + n_out * cos_out);
    return (r_par * r_par + r_perp * r_perp) / 2.0;
}
__device__ __forceinline__ D3 calculate_tint(D3 base) {  // disney.rs:425-433
    double lum = dot(D3{0.3, 0.6, 1.0}, base);
    return lum > 0.0 ? base * (1.0 / lum) : D3{1.0, 1.0, 1.0};
}
__device__ __forceinline__ double gtr1(double dot_hl, double a) {  // :435-443
    if (a >= 1.0) return 1.0 / RT_PI;
    double a2 = a * a;
    return (a2 - 1.0) / (RT_PI * log(a2) * (1.0 + (a2 - 1.0) * dot_hl * dot_hl));
}
__device__ __forceinline__ double separable_smith_ggxg1(D3 w, double a) {  // :445-450
    double a2 = a * a, nv = w.y;
    return 2.0 / (1.0 + sqrt(a2 + (1.0 - a2) * nv * nv));
}
__device__ __forceinline__ double ggx_anisotropic_d(D3 h, double ax, double ay) {  // :452-460
    double hx2 = h.x * h.x, hy2 = h.z * h.z, ct2 = h.y * h.y;
    double ax2 = ax * ax, ay2 = ay * ay;
    double q = hx2 / ax2 + hy2 / ay2 + ct2;
    return 1.0 / (RT_PI * ax * ay * (q * q));
}
__device__ inline double aniso_smith_g1(D3 w, D3 h, double ax, double ay, bool& error) {  // :462-480
    if (dot(w, h) <= 0.0) return 0.0;
    double att = fabs(tan_theta(w));
    if (isnan(att)) error = true;  // assert!
    if (isinf(att)) return 0.0;
    double a = sqrt(cos_phi2(w) * ax * ax + sin_phi2(w) * ay * ay);
    double at = a * att;
    double lambda = 0.5 * (-1.0 + sqrt(1.0 + at * at));
    return 1.0 / (1.0 + lambda);
}
__device__ __forceinline__ void aniso_params(double roughness, double anisotropic, double& ax, double& ay) {  // :482-488
    double aspect = sqrt(1.0 - 0.9 * anisotropic);
    double r2 = roughness * roughness;
    ax = rmax(0.001, r2 / aspect);
    ay = rmax(0.001, r2 * aspect);
}
__device__ inline void vndf_pdf(D3 v_in, D3 h, D3 v_out, double ax, double ay, double& fwd, double& rev, bool& error) {  // :490-510
    double d = ggx_anisotropic_d(h, ax, ay);
    double g1v = aniso_smith_g1(v_out, h, ax, ay, error);
    fwd = g1v * fabs(dot(h, v_out)) * d / fabs(cos_theta(v_out));
    double g1l = aniso_smith_g1(v_in, h, ax, ay, error);
    rev = g1l * fabs(dot(h, v_in)) * d / fabs(cos_theta(v_in));
}
__device__ __forceinline__ double thin_transmission_roughness(double ior, double roughness) { return clampd((0.65 * ior - 0.35) * roughness, 0.0, 1.0); }

__device__ inline D3 disney_fresnel(const Params& P, D3 v_out, D3 h, D3 v_in, double relative_ior) {  // :177-200
    double dot_hv = dot(h, v_out);
    D3 tint = calculate_tint(P.base_color);
    D3 r0 = schlick_r0_from_relative_ior(relative_ior) * lerp(D3{1.0, 1.0, 1.0}, tint, P.specular_tint);
    r0 = lerp(r0, P.base_color, P.metallic);
    double df = dielectric(dot_hv, 1.0, P.ior);
    D3 mf = schlick(r0, dot(v_in, h));
    return lerp(D3{df, df, df}, mf, P.metallic);
}
__device__ inline void evaluate_brdf(const Params& P, D3 v_out, D3 h, D3 v_in, double relative_ior, D3& value, double& fwd, bool& error) {  // :100-130
    double nl = cos_theta(v_in), nv = cos_theta(v_out);
    value = D3{0.0, 0.0, 0.0}, fwd = 0.0;
    if (nl <= 0.0 || nv <= 0.0) return;
    double ax, ay;
    aniso_params(P.roughness, P.anisotropic, ax, ay);
    double d = ggx_anisotropic_d(h, ax, ay);
    double gl = aniso_smith_g1(v_in, h, ax, ay, error), gv = aniso_smith_g1(v_out, h, ax, ay, error);
    D3 f = disney_fresnel(P, v_out, h, v_in, relative_ior);
    double rev;
    vndf_pdf(v_in, h, v_out, ax, ay, fwd, rev, error);
    fwd = fwd / (4.0 * fabs(dot(v_in, h)));
    value = (d * gl * gv * f) / (4.0 * nl * nv);
}
__device__ inline D3 evaluate_sheen(const Params& P, D3 h, D3 v_in) {  // :132-146
    if (P.sheen <= 0.0) return D3{0.0, 0.0, 0.0};
    double dot_hl = dot(h, v_in);
    D3 tint = calculate_tint(P.base_color);
    return (P.sheen * lerp(D3{1.0, 1.0, 1.0}, tint, P.sheen_tint)) * schlick_weight(dot_hl);
}
__device__ inline void evaluate_clearcoat(const Params& P, D3 v_out, D3 h, D3 v_in, double& value, double& fwd) {  // :148-175
    value = fwd = 0.0;
    if (P.clearcoat <= 0.0) return;
    double dot_nh = h.y, dot_hl = dot(h, v_in);
    double d = gtr1(dot_nh, lerp(0.1, 0.001, P.clearcoat_gloss));
    double f = schlick_f64(0.04, dot_hl);
    double gl = separable_smith_ggxg1(v_in, 0.25), gv = separable_smith_ggxg1(v_out, 0.25);
    value = 0.25 * P.clearcoat * d * f * gl * gv;
    fwd = d / (4.0 * fabs(dot(v_in, h)));
}
__device__ inline D3 evaluate_spec_transmission(const Params& P, D3 v_out, D3 h, D3 v_in, double ax, double ay, double relative_ior, bool& error) {  // :202-236
    double n2 = relative_ior * relative_ior;
    double anl = fabs(cos_theta(v_in)), anv = fabs(cos_theta(v_out));
    double dot_hl = dot(h, v_in), dot_hv = dot(h, v_out);
    double d = ggx_anisotropic_d(h, ax, ay);
    double gl = aniso_smith_g1(v_in, h, ax, ay, error), gv = aniso_smith_g1(v_out, h, ax, ay, error);
    double f = dielectric(dot_hv, 1.0, 1.0 / relative_ior);
    D3 color = P.base_color;
    if (P.thin) {
        color = D3{sqrt(color.x), sqrt(color.y), sqrt(color.z)};
        if (isnan(color.x) || isnan(color.y) || isnan(color.z)) error = true;  // Vec3::sqrt panics
    }
    double c = (fabs(dot_hl) * fabs(dot_hv)) / (anl * anv);
    double q = dot_hl + relative_ior * dot_hv;
    double t = n2 / (q * q);
    return (c * t * (1.0 - f) * gl * gv * d) * color;
}
__device__ inline double evaluate_retro_diffuse(const Params& P, D3 v_out, D3 v_in) {  // :272-290
    double anl = fabs(cos_theta(v_in)), anv = fabs(cos_theta(v_out));
    double roughness = P.roughness * P.roughness;
    double rr = 0.5 + 2.0 * anl * anl * roughness;
    double fl = schlick_weight(anl), fv = schlick_weight(anv);
    return rr * (fl + fv + fl * fv * (rr - 1.0));
}
__device__ inline double evaluate_diffuse(const Params& P, D3 v_out, D3 h, D3 v_in, bool thin) {  // :238-270
    double anl = fabs(cos_theta(v_in)), anv = fabs(cos_theta(v_out));
    double fl = schlick_weight(anl), fv = schlick_weight(anv);
    double hk = 0.0;
    if (thin && P.flatness > 0.0) {
        double roughness = P.roughness * P.roughness;
        double dot_hl = dot(h, v_in);
        double fss90 = dot_hl * dot_hl * roughness;
        double fss = lerp(1.0, fss90, fl) * lerp(1.0, fss90, fv);
        hk = 1.25 * (fss * (1.0 / (anl + anv) - 0.5) + 0.5);
    }
    double retro = evaluate_retro_diffuse(P, v_out, v_in);
    double subsurface = lerp(1.0, hk, thin ? P.flatness : 0.0);
    return 1.0 / RT_PI * (retro + subsurface * (1.0 - 0.5 * fl) * (1.0 - 0.5 * fv));
}
__device__ __forceinline__ void lobe_pdfs(const Params& P, double& p_spec, double& p_diff, double& p_clear, double& p_trans) {  // :403-422
    double metallic_brdf = P.metallic;
    double specular_bsdf = (1.0 - P.metallic) * P.spec_trans;
    double dielectric_brdf = (1.0 - P.spec_trans) * (1.0 - P.metallic);
    double sw = metallic_brdf + dielectric_brdf, tw = specular_bsdf, dw = dielectric_brdf, cw = 1.0 * clampd(P.clearcoat, 0.0, 1.0);
    double norm = 1.0 / (sw + tw + dw + cw);
    p_spec = sw * norm, p_trans = tw * norm, p_diff = dw * norm, p_clear = cw * norm;
}
// Disney::evaluate_disney, disney.rs:292-401 (forward pdf only: the reverse pdf is never used by camera.rs)
__device__ __noinline__ void evaluate_disney(const Params& P, D3 v_out, D3 v_in, bool front_face, D3& reflectance, double& forward_pdf, bool& error) {
    double relative_ior = front_face ? P.ior : 1.0 / P.ior;
    double nv = cos_theta(v_out), nl = cos_theta(v_in);
    bool is_transmission = nv * nl < 0.0;
    D3 h;
    if (!unit_vector(is_transmission ? v_in - v_out : v_in + v_out, h)) error = true;  // .expect
    reflectance = D3{0.0, 0.0, 0.0};
    forward_pdf = 0.0;
    double p_brdf, p_diffuse, p_clearcoat, p_spec_trans;
    lobe_pdfs(P, p_brdf, p_diffuse, p_clearcoat, p_spec_trans);
    double diffuse_weight = (1.0 - P.metallic) * (1.0 - P.spec_trans);
    double trans_weight = (1.0 - P.metallic) * P.spec_trans;
    bool upper = nl > 0.0 && nv > 0.0;
    if (upper && P.clearcoat > 0.0) {
        double cc, f;
        evaluate_clearcoat(P, v_out, h, v_in, cc, f);
        reflectance = reflectance + D3{cc, cc, cc};
        forward_pdf += p_clearcoat * f;
    }
    if (diffuse_weight > 0.0) {
        double fwd = fabs(cos_theta(v_in));
        double diffuse = evaluate_diffuse(P, v_out, h, v_in, P.thin);
        D3 sheen = evaluate_sheen(P, h, v_in);
        reflectance = reflectance + diffuse_weight * (diffuse * P.base_color + sheen);
        forward_pdf += p_diffuse * fwd;
    }
    if (trans_weight > 0.0) {
        double rscaled = P.thin ? thin_transmission_roughness(P.ior, P.roughness) : P.roughness;
        double tax, tay;
        aniso_params(rscaled, P.anisotropic, tax, tay);
        D3 t_v_out = is_transmission ? -v_out : v_out;
        D3 transmission = evaluate_spec_transmission(P, t_v_out, h, v_in, tax, tay, relative_ior, error);
        reflectance = reflectance + trans_weight * transmission;
        double fwd, rev;
        vndf_pdf(v_in, h, t_v_out, tax, tay, fwd, rev, error);
        double dot_lh = dot(h, v_in), dot_vh = dot(h, t_v_out);
        double q = dot_lh + relative_ior * dot_vh;
        double jacobian = (relative_ior * relative_ior * dot_lh) / (q * q);
        forward_pdf += p_spec_trans * fwd * fabs(jacobian);
    }
    if (upper) {
        D3 spec;
        double f;
        evaluate_brdf(P, v_out, h, v_in, relative_ior, spec, f, error);
        reflectance = reflectance + spec;
        forward_pdf += p_brdf * f;
    }
    reflectance = reflectance * fabs(nl);
    if (forward_pdf == 0.0) forward_pdf = INFINITY;
}
// sample_ggx_vndf_anisotropic, disney.rs:690-716
__device__ inline bool sample_vndf(D3 v_out, double ax, double ay, double u1, double u2, D3& out) {
    D3 v;
    if (!unit_vector(D3{v_out.x * ax, v_out.y, v_out.z * ay}, v)) return false;
    D3 t1 = v.y < 0.9999999 ? cross(v, D3{0.0, 1.0, 0.0}) : D3{1.0, 0.0, 0.0};
    D3 t2 = cross(t1, v);
    double a = 1.0 / (1.0 + v.y);
    double r = sqrt(u1);
    double phi = u2 < a ? (u2 / a) * RT_PI : RT_PI + (u2 - a) / (1.0 - a) * RT_PI;
    double s, c;
    sincos(phi, &s, &c);
    double p1 = r * c;
    double p2 = r * s * (u2 < a ? 1.0 : v.y);
    D3 n = p1 * t1 + p2 * t2 + sqrt(rmax(1.0 - p1 * p1 - p2 * p2, 0.0)) * v;
    return unit_vector(D3{ax * n.x, n.y, ay * n.z}, out);
}
// DisneyPDF::generate, disney.rs:668-688 in the local frame; returns false for None.  pick = the
// RT_SLOT_DISNEY pair, u = the RT_SLOT_DIRECTION pair.
__device__ __noinline__ bool generate_local(const Params& P, D3 v_out, bool front_face, Rand2 pick, Rand2 u, D3& v_in, bool& error) {
    double p_spec, p_diff, p_clear, p_trans;
    lobe_pdfs(P, p_spec, p_diff, p_clear, p_trans);
    const double p = pick.a;
    if (p <= p_spec) {  // sample_disney_brdf :540-556
        double ax, ay;
        aniso_params(P.roughness, P.anisotropic, ax, ay);
        D3 h;
        if (!sample_vndf(v_out, ax, ay, u.a, u.b, h) || !unit_vector(reflect2(v_out, h), v_in)) {
            error = true;
            return false;
        }
        return !(cos_theta(v_in) <= 0.0);
    }
    if (p <= p_spec + p_clear) {  // sample_disney_clearcoat :558-587
        double a = 0.25, a2 = a * a;
        double ct = sqrt(rmax((1.0 - pow(a2, 1.0 - u.a)) / (1.0 - a2), 0.0));
        double st = sqrt(rmax(1.0 - ct * ct, 0.0));
        double phi = 2.0 * RT_PI * u.b;
        double s, c;
        sincos(phi, &s, &c);
        D3 h = D3{st * c, ct, st * s};
        if (dot(h, v_out) < 0.0) h = -h;
        v_in = reflect2(v_out, h);
        return !(dot(v_in, v_out) < 0.0);
    }
    if (p <= p_spec + p_diff + p_clear) {  // sample_disney_diffuse :589-605
        double y = cos_theta(v_out);
        double sign = isnan(y) ? y : (signbit(y) ? -1.0 : 1.0);  // f64::signum
        v_in = sign * random_cosine_direction(u.a, u.b);
        if (pick.b <= P.diff_trans) v_in = -v_in;
        return !(cos_theta(v_in) == 0.0);
    }
    if (p_trans >= 0.0) {  // disney_spec_transmission :607-664
        double ior = front_face ? P.ior : 1.0 / P.ior;
        if (cos_theta(v_out) == 0.0) return false;
        double rscaled = P.thin ? thin_transmission_roughness(ior, P.roughness) : P.roughness;
        double tax, tay;
        aniso_params(rscaled, P.anisotropic, tax, tay);
        D3 h;
        if (!sample_vndf(v_out, tax, tay, u.a, u.b, h)) {
            error = true;
            return false;
        }
        double dot_vh = dot(v_out, h);
        if (h.y < 0.0) dot_vh = -dot_vh;
        double ni = v_out.y > 0.0 ? 1.0 : ior, nt = v_out.y > 0.0 ? ior : 1.0;
        double relative_ior = ni / nt;
        double f = dielectric(dot_vh, 1.0, P.ior);
        bool ok = true;
        if (pick.b <= f) {
            ok = unit_vector(reflect2(v_out, h), v_in);
        } else if (P.thin) {
            D3 wi = reflect2(v_out, h);
            ok = unit_vector(D3{wi.x, -wi.y, wi.z}, v_in);
        } else if (!refract2(v_out, h, relative_ior, v_in)) {
            ok = unit_vector(reflect2(v_out, h), v_in);
        }
        if (!ok) {
            error = true;
            return false;
        }
        return !(cos_theta(v_in) == 0.0);
    }
    error = true;  // panic!("The conditions should be exhausted!")
    return false;
}

}  // namespace disney
}  // namespace rt

"""Disney::evaluate_disney and DisneyPDF::generate of the oracle against an independent restatement in plain Python, written from the reference's source
(material/disney.rs:102-520, utils/fresnel.rs, the UnitVec3 trigonometry of utils/vec3.rs:376-426 with its quirks: cos_theta2()
returns y, cos_phi() / sin_phi() are 1 for every unit vector).  The reference holds no vectors for the BSDF; this pins the
oracle's lobes - clearcoat, diffuse + retro-reflection + sheen, thin-surface subsurface, specular transmission with its Jacobian,
the GGX specular lobe with the Disney Fresnel blend - and the lobe probabilities over parameter sets that switch every branch."""
import math

import numpy as np
import pytest

PI = math.pi


def lerp(a, b, t):  # utils.rs:14-19
    return a * (1.0 - t) + b * t


def lerp3(a, b, t):
    return tuple(lerp(x, y, t) for x, y in zip(a, b))


def dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def unit(v):
    n = math.sqrt(dot(v, v))
    return (v[0] / n, v[1] / n, v[2] / n)


def clamp(x, lo, hi):
    return max(lo, min(hi, x))


# ---- utils/vec3.rs:376-426 -------------------------------------------------------------------------
def cos_theta(w): return w[1]
def cos_theta2(w): return w[1]  # (sic)
def sin_theta2(w): return clamp(1.0 - cos_theta2(w), 0.0, 1.0)
def sin_theta(w): return math.sqrt(sin_theta2(w))
def tan_theta(w): return sin_theta(w) / cos_theta(w) if cos_theta(w) != 0.0 else math.copysign(math.inf, sin_theta(w)) if sin_theta(w) != 0.0 else math.nan
def cos_phi(w): return 1.0 if abs(sin_theta(w)) < 1e8 else w[0] / sin_theta(w)
def sin_phi(w): return 1.0 if abs(sin_theta(w)) < 1e8 else w[2] / sin_theta(w)


# ---- utils/fresnel.rs ------------------------------------------------------------------------------
def schlick_weight(u):
    return clamp(1.0 - u, 0.0, 1.0) ** 5


def schlick3(r0, radians):
    e = (1.0 - radians) ** 5
    return tuple(r + (1.0 - r) * e for r in r0)


def schlick_f64(r0, radians):
    return lerp(1.0, schlick_weight(radians), r0)


def schlick_r0_from_relative_ior(eta):
    return (eta - 1.0) ** 2 / (eta + 1.0) ** 2


def dielectric(cos_in, n_in, n_out):
    cos_in = clamp(cos_in, -1.0, 1.0)
    if cos_in < 0.0:
        n_in, n_out = n_out, n_in
        cos_in = -cos_in
    sin_in = math.sqrt(max(1.0 - cos_in * cos_in, 0.0))
    sin_out = n_in / n_out * sin_in
    if sin_out >= 1.0:
        return 1.0
    cos_out = math.sqrt(max(1.0 - sin_out * sin_out, 0.0))
    r_par = (n_out * cos_in - n_in * cos_out) / (n_out * cos_in + n_in * cos_out)
    r_perp = (n_in * cos_in - n_out * cos_out) / (n_in * cos_in + n_out * cos_out)
    return (r_par * r_par + r_perp * r_perp) / 2.0


# ---- material/disney.rs:425-514 --------------------------------------------------------------------
def calculate_tint(c):
    lum = dot((0.3, 0.6, 1.0), c)
    return tuple(x * (1.0 / lum) for x in c) if lum > 0.0 else (1.0, 1.0, 1.0)


def gtr1(dot_hl, a):
    if a >= 1.0:
        return 1.0 / PI
    a2 = a * a
    return (a2 - 1.0) / (PI * math.log(a2) * (1.0 + (a2 - 1.0) * dot_hl * dot_hl))


def separable_smith_ggxg1(w, a):
    a2 = a * a
    return 2.0 / (1.0 + math.sqrt(a2 + (1.0 - a2) * w[1] * w[1]))


def ggx_anisotropic_d(h, ax, ay):
    return 1.0 / (PI * ax * ay * (h[0] ** 2 / (ax * ax) + h[2] ** 2 / (ay * ay) + h[1] ** 2) ** 2)


def aniso_g1(w, h, ax, ay):
    if dot(w, h) <= 0.0:
        return 0.0
    t = abs(tan_theta(w))
    assert not math.isnan(t)
    if math.isinf(t):
        return 0.0
    a = math.sqrt(cos_phi(w) ** 2 * ax * ax + sin_phi(w) ** 2 * ay * ay)
    return 1.0 / (1.0 + 0.5 * (-1.0 + math.sqrt(1.0 + (a * t) ** 2)))


def aniso_params(roughness, anisotropic):
    aspect = math.sqrt(1.0 - 0.9 * anisotropic)
    r2 = roughness * roughness
    return max(0.001, r2 / aspect), max(0.001, r2 * aspect)


def vndf_pdf(v_in, h, v_out, ax, ay):
    d = ggx_anisotropic_d(h, ax, ay)
    fwd = aniso_g1(v_out, h, ax, ay) * abs(dot(h, v_out)) * d / abs(cos_theta(v_out))
    rev = aniso_g1(v_in, h, ax, ay) * abs(dot(h, v_in)) * d / abs(cos_theta(v_in))
    return fwd, rev


def thin_transmission_roughness(ior, roughness):
    return clamp((0.65 * ior - 0.35) * roughness, 0.0, 1.0)


# ---- material/disney.rs:102-420 --------------------------------------------------------------------
class P:
    def __init__(s, base_color=(0.8, 0.8, 0.8), roughness=0.5, anisotropic=0.0, sheen=0.0, sheen_tint=0.0, clearcoat=0.0, clearcoat_gloss=0.0,
                 specular_tint=0.0, metallic=0.0, ior=1.45, flatness=0.0, spec_trans=0.0, diff_trans=0.0, thin=False):
        s.__dict__.update(locals())

    def flat(s):
        return list(s.base_color) + [s.roughness, s.anisotropic, s.sheen, s.sheen_tint, s.clearcoat, s.clearcoat_gloss, s.specular_tint, s.metallic,
                                     s.ior, s.flatness, s.spec_trans, s.diff_trans]


def disney_fresnel(p, v_out, h, v_in, relative_ior):
    tint = calculate_tint(p.base_color)
    k = schlick_r0_from_relative_ior(relative_ior)
    r0 = tuple(k * x for x in lerp3((1.0, 1.0, 1.0), tint, p.specular_tint))
    r0 = lerp3(r0, p.base_color, p.metallic)
    fd = dielectric(dot(h, v_out), 1.0, p.ior)
    fm = schlick3(r0, dot(v_in, h))
    return lerp3((fd, fd, fd), fm, p.metallic)


def evaluate_brdf(p, v_out, h, v_in, relative_ior):
    nl, nv = cos_theta(v_in), cos_theta(v_out)
    if nl <= 0.0 or nv <= 0.0:
        return (0.0, 0.0, 0.0), 0.0
    ax, ay = aniso_params(p.roughness, p.anisotropic)
    d = ggx_anisotropic_d(h, ax, ay)
    gl, gv = aniso_g1(v_in, h, ax, ay), aniso_g1(v_out, h, ax, ay)
    f = disney_fresnel(p, v_out, h, v_in, relative_ior)
    fwd, _ = vndf_pdf(v_in, h, v_out, ax, ay)
    fwd = fwd / (4.0 * abs(dot(v_in, h)))
    return tuple(d * gl * gv * x / (4.0 * nl * nv) for x in f), fwd


def evaluate_sheen(p, h, v_in):
    if p.sheen <= 0.0:
        return (0.0, 0.0, 0.0)
    w = schlick_weight(dot(h, v_in))
    return tuple(p.sheen * x * w for x in lerp3((1.0, 1.0, 1.0), calculate_tint(p.base_color), p.sheen_tint))


def evaluate_clearcoat(p, v_out, h, v_in):
    if p.clearcoat <= 0.0:
        return 0.0, 0.0
    d = gtr1(h[1], lerp(0.1, 0.001, p.clearcoat_gloss))
    f = schlick_f64(0.04, dot(h, v_in))
    value = 0.25 * p.clearcoat * d * f * separable_smith_ggxg1(v_in, 0.25) * separable_smith_ggxg1(v_out, 0.25)
    return value, d / (4.0 * abs(dot(v_in, h)))


def retro_diffuse(p, v_out, v_in):
    nl, nv = abs(cos_theta(v_in)), abs(cos_theta(v_out))
    rr = 0.5 + 2.0 * nl * nl * (p.roughness * p.roughness)
    fl, fv = schlick_weight(nl), schlick_weight(nv)
    return rr * (fl + fv + fl * fv * (rr - 1.0))


def evaluate_diffuse(p, v_out, h, v_in, thin):
    nl, nv = abs(cos_theta(v_in)), abs(cos_theta(v_out))
    fl, fv = schlick_weight(nl), schlick_weight(nv)
    hk = 0.0
    if thin and p.flatness > 0.0:
        hl = dot(h, v_in)
        fss90 = hl * hl * (p.roughness * p.roughness)
        fss = lerp(1.0, fss90, fl) * lerp(1.0, fss90, fv)
        hk = 1.25 * (fss * (1.0 / (nl + nv) - 0.5) + 0.5)
    subsurface = lerp(1.0, hk, p.flatness if thin else 0.0)
    return 1.0 / PI * (retro_diffuse(p, v_out, v_in) + subsurface * (1.0 - 0.5 * fl) * (1.0 - 0.5 * fv))


def evaluate_spec_transmission(p, v_out, h, v_in, ax, ay, relative_ior):
    n2 = relative_ior * relative_ior
    hl, hv = dot(h, v_in), dot(h, v_out)
    d = ggx_anisotropic_d(h, ax, ay)
    gl, gv = aniso_g1(v_in, h, ax, ay), aniso_g1(v_out, h, ax, ay)
    f = dielectric(hv, 1.0, 1.0 / relative_ior)
    color = tuple(math.sqrt(x) for x in p.base_color) if p.thin else p.base_color
    c = (abs(hl) * abs(hv)) / (abs(cos_theta(v_in)) * abs(cos_theta(v_out)))
    t = n2 / (hl + relative_ior * hv) ** 2
    return tuple(c * t * (1.0 - f) * gl * gv * d * x for x in color)


def lobe_pdfs(p):
    metallic_brdf = p.metallic
    specular_bsdf = (1.0 - p.metallic) * p.spec_trans
    dielectric_brdf = (1.0 - p.spec_trans) * (1.0 - p.metallic)
    sw, tw, dw, cw = metallic_brdf + dielectric_brdf, specular_bsdf, dielectric_brdf, clamp(p.clearcoat, 0.0, 1.0)
    norm = 1.0 / (sw + tw + dw + cw)
    return sw * norm, dw * norm, cw * norm, tw * norm  # specular, diffuse, clearcoat, transmission


def evaluate_disney(p, v_out, v_in, front_face):
    relative_ior = p.ior if front_face else 1.0 / p.ior
    nv, nl = cos_theta(v_out), cos_theta(v_in)
    is_transmission = nv * nl < 0.0
    h = unit(tuple(a - b for a, b in zip(v_in, v_out))) if is_transmission else unit(tuple(a + b for a, b in zip(v_in, v_out)))
    refl, fwd = [0.0, 0.0, 0.0], 0.0
    p_brdf, p_diffuse, p_clearcoat, p_trans = lobe_pdfs(p)
    diffuse_weight = (1.0 - p.metallic) * (1.0 - p.spec_trans)
    trans_weight = (1.0 - p.metallic) * p.spec_trans
    upper = nl > 0.0 and nv > 0.0
    if upper and p.clearcoat > 0.0:
        cc, pw = evaluate_clearcoat(p, v_out, h, v_in)
        refl = [x + cc for x in refl]
        fwd += p_clearcoat * pw
    if diffuse_weight > 0.0:
        diffuse = evaluate_diffuse(p, v_out, h, v_in, p.thin)
        sheen = evaluate_sheen(p, h, v_in)
        refl = [x + diffuse_weight * (diffuse * b + s) for x, b, s in zip(refl, p.base_color, sheen)]
        fwd += p_diffuse * abs(cos_theta(v_in))
    if trans_weight > 0.0:
        rscaled = thin_transmission_roughness(p.ior, p.roughness) if p.thin else p.roughness
        tax, tay = aniso_params(rscaled, p.anisotropic)
        t_v_out = tuple(-x for x in v_out) if is_transmission else v_out
        tr = evaluate_spec_transmission(p, t_v_out, h, v_in, tax, tay, relative_ior)
        refl = [x + trans_weight * t for x, t in zip(refl, tr)]
        pw, _ = vndf_pdf(v_in, h, t_v_out, tax, tay)
        lh, vh = dot(h, v_in), dot(h, t_v_out)
        jac = (relative_ior * relative_ior * lh) / (lh + relative_ior * vh) ** 2
        fwd += p_trans * pw * abs(jac)
    if upper:
        spec, pw = evaluate_brdf(p, v_out, h, v_in, relative_ior)
        refl = [x + s for x, s in zip(refl, spec)]
        fwd += p_brdf * pw
    refl = [x * abs(nl) for x in refl]
    if fwd == 0.0:
        fwd = math.inf
    return refl, fwd


CASES = {
    "default": P(),
    "metal_aniso": P(base_color=(0.9, 0.6, 0.2), roughness=0.3, anisotropic=0.7, metallic=1.0, specular_tint=0.5),
    "coated_sheen": P(base_color=(0.2, 0.5, 0.7), roughness=0.6, sheen=0.8, sheen_tint=0.4, clearcoat=0.9, clearcoat_gloss=0.7, specular_tint=0.3, metallic=0.25),
    "glass": P(base_color=(0.95, 0.97, 1.0), roughness=0.15, ior=1.5, spec_trans=1.0),
    "half_glass": P(base_color=(0.7, 0.3, 0.3), roughness=0.4, anisotropic=0.3, ior=1.33, spec_trans=0.5, clearcoat=0.3, metallic=0.1),
    "thin_flat": P(base_color=(0.6, 0.8, 0.3), roughness=0.7, flatness=0.6, spec_trans=0.4, ior=1.4, thin=True, sheen=0.2),
    "black": P(base_color=(0.0, 0.0, 0.0), roughness=0.9, sheen=0.5, sheen_tint=1.0),
    "mirror": P(base_color=(1.0, 1.0, 1.0), roughness=0.0, metallic=1.0),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_disney_evaluate_matches_a_plain_restatement(orc, name):
    p = CASES[name]
    rng = np.random.default_rng(sum(map(ord, name)))
    n = checked = 0
    while n < 300:
        v_out, v_in = unit(tuple(rng.normal(size=3))), unit(tuple(rng.normal(size=3)))
        if abs(v_out[1]) < 1e-3 or abs(v_in[1]) < 1e-3:
            continue
        n += 1
        for front in (True, False):
            got = orc.disney_evaluate(p.flat(), p.thin, v_out, v_in, front)
            assert got is not None
            want_refl, want_pdf = evaluate_disney(p, v_out, v_in, front)
            if math.isinf(want_pdf):
                assert math.isinf(got[1])
            else:
                assert got[1] == pytest.approx(want_pdf, rel=1e-11, abs=1e-300)
            assert np.allclose(got[0], want_refl, rtol=1e-11, atol=1e-300), (name, v_out, v_in, front, got[0], want_refl)
            checked += 1
    assert checked == 600


# ---- DisneyPDF::generate and its samplers, material/disney.rs:542-720; vec3.rs:76-78, 333-343, 357-366 ----------------------
def reflect2(v, n):
    k = 2.0 * dot(v, n)
    return tuple(-a + k * b for a, b in zip(v, n))


def refract2(v, n, relative_eta):
    cos_t = min(dot(v, n), 1.0)
    perp = tuple(relative_eta * (-a + cos_t * b) for a, b in zip(v, n))
    arg = 1.0 - dot(perp, perp)
    if arg < 0.0:
        return None  # sqrt -> NaN -> None
    par = math.sqrt(arg)
    return tuple(a + (-par) * b for a, b in zip(perp, n))


def cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def sample_ggx_vndf_anisotropic(v_out, ax, ay, u1, u2):
    v = unit((v_out[0] * ax, v_out[1], v_out[2] * ay))
    t1 = cross(v, (0.0, 1.0, 0.0)) if v[1] < 0.9999999 else (1.0, 0.0, 0.0)
    t2 = cross(t1, v)
    a = 1.0 / (1.0 + v[1])
    r = math.sqrt(u1)
    phi = (u2 / a) * PI if u2 < a else PI + (u2 - a) / (1.0 - a) * PI
    p1 = r * math.cos(phi)
    p2 = r * math.sin(phi) * (1.0 if u2 < a else v[1])
    s = math.sqrt(max(1.0 - p1 * p1 - p2 * p2, 0.0))
    n = tuple(p1 * x + p2 * y + s * z for x, y, z in zip(t1, t2, v))
    return unit((ax * n[0], n[1], ay * n[2]))


def disney_generate(p, v_out, front_face, pick, u):
    """-> local direction or None.  The draws of the reference's thread RNG are the addressed pairs of this implementation: pick.a chooses
    the lobe, pick.b is the coin of the second decision inside a lobe, u = (r0, r1) of the lobe's sampler (include/rt2025_rng.h)."""
    p_spec, p_diff, p_clear, p_trans = lobe_pdfs(p)
    q = pick[0]
    if q <= p_spec:  # sample_disney_brdf
        ax, ay = aniso_params(p.roughness, p.anisotropic)
        h = sample_ggx_vndf_anisotropic(v_out, ax, ay, u[0], u[1])
        v_in = unit(reflect2(v_out, h))
        return None if cos_theta(v_in) <= 0.0 else v_in
    if q <= p_spec + p_clear:  # sample_disney_clearcoat
        a2 = 0.25 * 0.25
        ct = math.sqrt(max((1.0 - a2 ** (1.0 - u[0])) / (1.0 - a2), 0.0))
        st = math.sqrt(max(1.0 - ct * ct, 0.0))
        phi = 2.0 * PI * u[1]
        h = (st * math.cos(phi), ct, st * math.sin(phi))
        if dot(h, v_out) < 0.0:
            h = tuple(-x for x in h)
        v_in = reflect2(v_out, h)
        return None if dot(v_in, v_out) < 0.0 else unit(v_in)
    if q <= p_spec + p_diff + p_clear:  # sample_disney_diffuse
        sign = math.copysign(1.0, cos_theta(v_out))
        phi = 2.0 * PI * u[0]
        c = (math.sin(phi) * math.sqrt(u[1]), math.sqrt(1.0 - u[1]), math.cos(phi) * math.sqrt(u[1]))  # random_cosine_direction
        v_in = tuple(sign * x for x in c)
        if pick[1] <= p.diff_trans:
            v_in = tuple(-x for x in v_in)
        return None if cos_theta(v_in) == 0.0 else unit(v_in)
    assert p_trans >= 0.0  # disney_spec_transmission
    ior = p.ior if front_face else 1.0 / p.ior
    if cos_theta(v_out) == 0.0:
        return None
    rscaled = thin_transmission_roughness(ior, p.roughness) if p.thin else p.roughness
    tax, tay = aniso_params(rscaled, p.anisotropic)
    h = sample_ggx_vndf_anisotropic(v_out, tax, tay, u[0], u[1])
    dot_vh = dot(v_out, h)
    if h[1] < 0.0:
        dot_vh = -dot_vh
    ni, nt = (1.0, ior) if v_out[1] > 0.0 else (ior, 1.0)
    f = dielectric(dot_vh, 1.0, p.ior)
    if pick[1] <= f:
        v_in = unit(reflect2(v_out, h))
    elif p.thin:
        wi = reflect2(v_out, h)
        v_in = unit((wi[0], -wi[1], wi[2]))
    else:
        v_in = refract2(v_out, h, ni / nt)
        if v_in is None:
            v_in = unit(reflect2(v_out, h))
    return None if cos_theta(v_in) == 0.0 else unit(v_in)


@pytest.mark.parametrize("name", sorted(CASES))
def test_disney_generate_matches_a_plain_restatement(orc, name):
    p = CASES[name]
    rng = np.random.default_rng(1000 + sum(map(ord, name)))
    some = none = 0
    for _ in range(500):
        v_out = unit(tuple(rng.normal(size=3)))
        if abs(v_out[1]) < 1e-3:
            continue
        pick, u = tuple(rng.uniform(size=2)), tuple(rng.uniform(size=2))
        for front in (True, False):
            got = orc.disney_generate(p.flat(), p.thin, v_out, front, pick, u)
            want = disney_generate(p, v_out, front, pick, u)
            assert not isinstance(got, str), "the oracle reports a panic"
            if want is None:
                assert got is None, (name, v_out, front, pick, u, got)
                none += 1
            else:
                assert got is not None, (name, v_out, front, pick, u, want)
                assert np.allclose(got, want, rtol=0.0, atol=1e-11), (name, v_out, front, pick, u, got, want)
                some += 1
    assert some > 300

"""Hand-derived known-answer vectors for the functions of the hot path the reference's own unit tests do NOT cover
("hits, scatter, pdfs, media: parity unpinned by the reference").

Every expected value below was derived on paper from the reference source cited next to it, with inputs chosen so that
each intermediate is exactly representable in binary64 (small integers and dyadic fractions); where a libm call is
unavoidable (ln, sqrt of a non-square, pi) the expected value is the reference's formula restated inline with Python
floats - an implementation independent of oracle/oracle.cpp.  The same hit vectors are run against the CUDA core in
tests/test_gpu_parity.py::test_hand_derived_hit_vectors.
"""
import math

import pytest

SCATTER_NONE, SCATTER_COSINE, SCATTER_SPHERE, SCATTER_RAY = 0, 1, 2, 3

# (name, builder, [(origin, direction, time, t_min, t_max, expected or None)]): expected = (t, u, v[, p, normal, front_face])
INF = float("inf")


def sphere_scene(rt):
    b = rt.Builder(1)
    hs = b.finish(b.list([b.sphere([0, 0, 0], 2.0, b.empty())]))
    hs._b = b
    return hs


def moving_sphere_scene(rt):
    b = rt.Builder(1)
    hs = b.finish(b.list([b.sphere_moving([0, 0, 0], [4, 0, 0], 2.0, b.empty())]))
    hs._b = b
    return hs


def quad_scene(rt):
    b = rt.Builder(1)
    hs = b.finish(b.list([b.quad([-1, -1, 0], [2, 0, 0], [0, 2, 0], b.empty())]))
    hs._b = b
    return hs


def triangle_scene(rt):
    b = rt.Builder(1)
    hs = b.finish(b.list([b.triangle([0, 0, 0], [4, 0, 0], [0, 4, 0], b.empty())]))
    hs._b = b
    return hs


def transform_scene(rt):
    # Transform { offset (1,2,3), quaternion (w,x,y,z) = (0,0,1,0) = half a turn about y, scale (2,1,4) } over the quad above
    b = rt.Builder(1)
    q = b.quad([-1, -1, 0], [2, 0, 0], [0, 2, 0], b.empty())
    s = b.transform(b.sphere([0, 0, 0], 1.0, b.empty()), offset=[10, 0, 0], scale=[2, 2, 2])
    hs = b.finish(b.list([b.transform(q, offset=[1, 2, 3], quat=[0, 0, 1, 0], scale=[2, 1, 4]), s]))
    hs._b = b
    return hs


HIT_VECTORS = {
    # Sphere::hit, shapes/sphere.rs:77-108: oc = c - o, a = d.d, h = d.oc, c = oc.oc - r^2, roots (h -+ sqrt(h^2 - a c)) / a
    "sphere": (sphere_scene, [
        # oc=(0,0,-5) a=1 h=5 c=21 disc=4: near root (5-2)/1 = 3; p=(0,0,2) n=(0,0,1); uv of +z is (0.25, 0.5) (sphere.rs:152-169)
        ((0, 0, 5), (0, 0, -1), 0, 1e-8, INF, dict(t=3.0, u=0.25, v=0.5, p=(0, 0, 2), normal=(0, 0, 1), front_face=True)),
        # near root outside [3.5, 10] -> far root (5+2)/1 = 7: exit point, outward normal -z, back face -> normal flipped to +z
        ((0, 0, 5), (0, 0, -1), 0, 3.5, 10.0, dict(t=7.0, u=0.75, v=0.5, p=(0, 0, -2), normal=(0, 0, 1), front_face=False)),
        # both roots outside [3.5, 6.5] -> None
        ((0, 0, 5), (0, 0, -1), 0, 3.5, 6.5, None),
        # the direction is not normalised: a=4 h=10 c=21 disc=16, root (10-4)/4 = 1.5
        ((0, 0, 5), (0, 0, -2), 0, 1e-8, INF, dict(t=1.5, u=0.25, v=0.5, p=(0, 0, 2), normal=(0, 0, 1), front_face=True)),
        # oc=(-3,0,-5): h=5 c=30 disc=-5 < 0 -> None
        ((3, 0, 5), (0, 0, -1), 0, 1e-8, INF, None),
        # grazing: oc=(-2,0,-5) h=5 c=25 disc=0 (not < 0) -> root 5, p=(2,0,0), outward +x, d.n = 0 is not < 0 -> back face
        ((2, 0, 5), (0, 0, -1), 0, 1e-8, INF, dict(t=5.0, u=0.5, v=0.5, p=(2, 0, 0), normal=(-1, 0, 0), front_face=False)),
        # Interval::contains is inclusive (interval.rs:65-67): t_max exactly 3 still hits
        ((0, 0, 5), (0, 0, -1), 0, 1e-8, 3.0, dict(t=3.0, u=0.25, v=0.5)),
    ]),
    # centre(time) = c0 + time * (c1 - c0), sphere.rs:78: at time 0.5 the centre is (2,0,0)
    "moving_sphere": (moving_sphere_scene, [
        ((2, 0, 5), (0, 0, -1), 0.5, 1e-8, INF, dict(t=3.0, u=0.25, v=0.5, p=(2, 0, 2), normal=(0, 0, 1), front_face=True)),
        ((2, 0, 5), (0, 0, -1), 0.0, 1e-8, INF, dict(t=5.0, u=0.5, v=0.5, p=(2, 0, 0))),   # time 0: the grazing case above
        ((7, 0, 5), (0, 0, -1), 1.0, 1e-8, INF, None),                                       # time 1: centre (4,0,0), x = 7 misses
    ]),
    # Quad::new / hit, quad.rs:31-47,71-102: n = u x v = (0,0,4), normal (0,0,1), D = 0, w = n/(n.n) = (0,0,1/4);
    # t = (D - n.o)/(n.d); alpha = w.(hp x v), beta = w.(u x hp)
    "quad": (quad_scene, [
        # hp = (1.5, 0.5, 0): alpha = (1.5*2)/4 = 0.75, beta = (2*0.5)/4 = 0.25
        ((0.5, -0.5, 3), (0, 0, -1), 0, 1e-8, INF, dict(t=3.0, u=0.75, v=0.25, p=(0.5, -0.5, 0), normal=(0, 0, 1), front_face=True)),
        # from behind: d.n = 1 -> back face, normal flipped (hit.rs:24-43)
        ((0.5, -0.5, -3), (0, 0, 1), 0, 1e-8, INF, dict(t=3.0, u=0.75, v=0.25, normal=(0, 0, -1), front_face=False)),
        # the far corner: alpha = beta = 1 is interior (inclusive unit interval, quad.rs:60-68)
        ((1, 1, 3), (0, 0, -1), 0, 1e-8, INF, dict(t=3.0, u=1.0, v=1.0)),
        ((1.5, 0, 3), (0, 0, -1), 0, 1e-8, INF, None),                 # alpha = 1.25
        ((0, 0, 3), (1, 0, 0), 0, 1e-8, INF, None),                    # parallel: denom = 0
        # |denom| = 2^-30 < 1e-8 -> None although the ray meets the quad at t = 0.5 (quad.rs:74-76)
        ((-0.5, 0, 2.0 ** -31), (1, 0, -2.0 ** -30), 0, 1e-8, INF, None),
        ((0.5, -0.5, 3), (0, 0, -1), 0, 1e-8, 2.5, None),              # t = 3 outside the interval
    ]),
    # Triangle::new / hit, triangle.rs:29-46,56-98: n = (0,0,16), w = (0,0,1/16); additionally 0 <= alpha + beta <= 1
    "triangle": (triangle_scene, [
        ((1, 1, 2), (0, 0, -1), 0, 1e-8, INF, dict(t=2.0, u=0.25, v=0.25, p=(1, 1, 0), normal=(0, 0, 1), front_face=True)),
        ((2, 2, 2), (0, 0, -1), 0, 1e-8, INF, dict(t=2.0, u=0.5, v=0.5)),      # on the hypotenuse: alpha + beta = 1 is inside
        ((3, 3, 2), (0, 0, -1), 0, 1e-8, INF, None),                           # inside the parallelogram, outside the triangle
        ((-1, 1, 2), (0, 0, -1), 0, 1e-8, INF, None),                          # alpha = -0.25
    ]),
    # Transform::hit, shapes.rs:88-111: local(x) = conj(q) (x - offset) / scale; local ray from local(o) to local(o + d): t unchanged;
    # p re-transformed, normal = unit(q (n / scale)).  Half a turn about y maps (x,y,z) to (-x,y,-z) exactly.
    "transform": (transform_scene, [
        # local o = ((2,2.5,8)-(1,2,3)) -> (1,.5,5) -> (-1,.5,-5) -> /(2,1,4) = (-.5,.5,-1.25); local d = (0,0,.25); t = 1.25/.25 = 5;
        # hp = (.5,1.5,0): alpha = (.5*2)/4 = .25, beta = (2*1.5)/4 = .75; local back face -> (0,0,-1) -> /scale, rotated, unit = (0,0,1)
        ((2, 2.5, 8), (0, 0, -1), 0, 1e-8, INF, dict(t=5.0, u=0.25, v=0.75, p=(2, 2.5, 3), normal=(0, 0, 1), front_face=False)),
        # unit sphere scaled by 2 at (10,0,0): local o = (0,0,2.5), local d = (0,0,-.5): a=.25 h=1.25 c=5.25 disc=.25 root (1.25-.5)/.25 = 3
        ((10, 0, 5), (0, 0, -1), 0, 1e-8, INF, dict(t=3.0, u=0.25, v=0.5, p=(10, 0, 2), normal=(0, 0, 1), front_face=True)),
        ((6, 2.5, 8), (0, 0, -1), 0, 1e-8, INF, None),
    ]),
}


def check_hit(got, want):
    if want is None:
        assert got is None
        return
    assert got is not None
    assert got["t"] == want["t"] and got["u"] == want["u"] and got["v"] == want["v"]
    if "p" in want:
        assert tuple(got["p"]) == tuple(float(x) for x in want["p"])
    if "normal" in want:
        assert tuple(got["normal"]) == tuple(float(x) for x in want["normal"])  # -0.0 == 0.0
    if "front_face" in want:
        assert got["front_face"] == want["front_face"]


@pytest.mark.parametrize("name", sorted(HIT_VECTORS))
def test_hit_vectors(rt, orc, name):
    make, vectors = HIT_VECTORS[name]
    osc = orc.OracleScene(make(rt))
    for o, d, time, t_min, t_max, want in vectors:
        check_hit(osc.world_hit(o, d, time, t_min, t_max), want)


def test_constant_medium_hit_with_fixed_draws(rt, orc):
    """ConstantMedium::hit, volume.rs:37-73, boundary = Sphere(0, r=2), density 0.5 (neg_inv_density = -2)."""
    b = rt.Builder(1)
    hs = b.finish(b.list([b.medium(b.sphere([0, 0, 0], 2.0, b.empty()), 0.5, b.solid(1, 1, 1))]))
    osc = orc.OracleScene(hs)
    ln_half = math.log(0.5)
    # entry 3 (near root), exit 7 (the far root is the first one inside [3.0001, inf)); length 1; inside 4 units;
    # xi = 1: hit_distance = -2 ln 1 = -0 -> t = 3 + (-0)/1 = 3 exactly; Vec3(1,0,0) normal, d.n = 0 is not < 0 -> flipped
    h = osc.world_hit((0, 0, 5), (0, 0, -1), xi=(1.0, 0.0))
    assert h["t"] == 3.0 and tuple(h["p"]) == (0.0, 0.0, 2.0) and tuple(h["normal"]) == (-1.0, 0.0, 0.0) and not h["front_face"]
    assert (h["u"], h["v"]) == (0.0, 0.0)
    # xi = 1/2: hit_distance = -2 ln(1/2) = 1.386..., t = 3 + hit_distance / 1
    assert osc.world_hit((0, 0, 5), (0, 0, -1), xi=(0.5, 0.0))["t"] == 3.0 + (-2.0 * ln_half) / 1.0
    # xi = 2^-10: hit_distance = 13.86 > 4 -> None
    assert osc.world_hit((0, 0, 5), (0, 0, -1), xi=(2.0 ** -10, 0.0)) is None
    # unnormalised direction: roots 1.5 and 3.5, |d| = 2, inside (3.5 - 1.5) * 2 = 4; t = 1.5 + hit_distance / 2
    assert osc.world_hit((0, 0, 5), (0, 0, -2), xi=(0.5, 0.0))["t"] == 1.5 + (-2.0 * ln_half) / 2.0
    # origin inside: rec1.t = -2 is clamped to the interval minimum 1e-8 (volume.rs:46), exit 2
    assert osc.world_hit((0, 0, 0), (0, 0, 1), xi=(1.0, 0.0))["t"] == 1e-8
    assert osc.world_hit((0, 0, 0), (0, 0, 1), xi=(0.5, 0.0))["t"] == 1e-8 + (-2.0 * ln_half) / 1.0
    # exit clamped by interval.max (volume.rs:47): inside (3.5 - 3) = 0.5 < 1.386 -> None; and entry >= exit -> None
    assert osc.world_hit((0, 0, 5), (0, 0, -1), t_max=3.5, xi=(0.5, 0.0)) is None
    assert osc.world_hit((0, 0, 5), (0, 0, -1), t_max=2.5, xi=(1.0, 0.0)) is None
    # the boundary is missed / only grazed (exit needs a second hit beyond entry + 1e-4)
    assert osc.world_hit((3, 0, 5), (0, 0, -1), xi=(1.0, 0.0)) is None
    assert osc.world_hit((2, 0, 5), (0, 0, -1), xi=(1.0, 0.0)) is None


def test_dielectric_scatter(rt, orc):
    """Dielectric::scatter, material.rs:117-144 + reflect / refract, vec3.rs:71-73,345-354; refraction index 1.5."""
    b = rt.Builder(1)
    glass = b.dielectric(b.solid(0.5, 0.25, 1.0), 1.5)
    hs = b.finish(b.list([b.sphere([0, 0, 0], 1.0, glass)]))
    osc = orc.OracleScene(hs)
    glass = int(hs.objects()[0]["material"])  # index in the flattened description
    ri = 1.0 / 1.5
    r0 = (1.0 - ri) / (1.0 + ri)
    # normal incidence on the front face: cos = 1, sin = 0, reflectance = r0^2 + (1 - r0^2) * 0^5 = r0^2 (about 0.04)
    kind, att, dr, err = osc.scatter(glass, (0, 1, 0), (0, -1, 0), (0, 0, 0), (0, 1, 0), True, xi=(0.5, 0.0))
    assert kind == SCATTER_RAY and not err and att == [0.5, 0.25, 1.0]
    assert dr == [0.0, -1.0, 0.0]      # xi = 0.5 > r0^2: refract; out_perp = ri * (ud + 1 * n) = 0, out_parallel = -sqrt(1 - 0) n
    kind, att, dr, err = osc.scatter(glass, (0, 1, 0), (0, -1, 0), (0, 0, 0), (0, 1, 0), True, xi=(r0 * r0 / 2.0, 0.0))
    assert dr == [0.0, 1.0, 0.0]       # xi < r0^2: reflect = ud - 2 (ud.n) n
    # the draw is compared with `>`: xi exactly r0^2 refracts
    assert osc.scatter(glass, (0, 1, 0), (0, -1, 0), (0, 0, 0), (0, 1, 0), True, xi=(r0 * r0, 0.0))[2] == [0.0, -1.0, 0.0]
    # back face (ri = 1.5) at grazing incidence: cos = 0, sin = 1, 1.5 > 1 -> total internal reflection whatever the draw
    assert osc.scatter(glass, (-1, 0, 0), (1, 0, 0), (0, 0, 0), (0, 1, 0), False, xi=(0.999, 0.0))[2] == [1.0, 0.0, 0.0]
    # 3-4-5 incidence on the front face, formula restated: ud = (1/5) d
    inv = 1.0 / 5.0
    ud = (inv * 3.0, inv * -4.0, 0.0)
    cos_t = min(-ud[1], 1.0)
    x = 1.0 - cos_t
    refl = r0 * r0 + (1.0 - r0 * r0) * (x * ((x * x) * (x * x)))
    perp = (ri * (ud[0] + cos_t * 0.0), ri * (ud[1] + cos_t * 1.0), ri * (ud[2] + cos_t * 0.0))
    par = -math.sqrt(1.0 - (perp[0] * perp[0] + perp[1] * perp[1] + perp[2] * perp[2]))
    want = [perp[0] + par * 0.0, perp[1] + par * 1.0, perp[2] + par * 0.0]
    assert 0.04 < refl < 0.06
    assert osc.scatter(glass, (-3, 4, 0), (3, -4, 0), (0, 0, 0), (0, 1, 0), True, xi=(0.5, 0.0))[2] == want
    assert osc.scatter(glass, (-3, 4, 0), (3, -4, 0), (0, 0, 0), (0, 1, 0), True, xi=(refl / 2, 0.0))[2] == \
        [ud[0] - 2.0 * (ud[1] * 1.0) * 0.0, ud[1] - 2.0 * ud[1] * 1.0, 0.0]


def test_metal_lambertian_isotropic_scatter(rt, orc):
    b = rt.Builder(1)
    metal = b.metal([0.5, 0.5, 1.0], 0.0)
    lam = b.lambertian(b.solid(0.25, 0.5, 0.75))
    light = b.diffuse_light(b.solid(4, 4, 4))
    hs = b.finish(b.list([b.sphere([0, 0, 0], 1.0, metal), b.sphere([3, 0, 0], 1.0, lam), b.sphere([6, 0, 0], 1.0, light),
                          b.medium(b.sphere([9, 0, 0], 1.0, b.empty()), 1.0, b.solid(0.125, 0.25, 0.5))]))
    osc = orc.OracleScene(hs)
    # material indices of the FLATTENED description (the builder's handles are not): three spheres, then the medium's Isotropic
    objs = hs.objects()
    metal, lam, light = [int(o["material"]) for o in objs if o["kind"] == 1][:3]
    iso = [int(o["material"]) for o in objs if o["kind"] == 7][0]
    # Metal::scatter, material.rs:82-95: unit(reflect(unit d, n)) + fuzz * unit_sphere; fuzz 0; never absorbs
    kind, att, dr, err = osc.scatter(metal, (0, 2, 0), (0, -2, 0), (0, 0, 0), (0, 1, 0), True)
    assert kind == SCATTER_RAY and att == [0.5, 0.5, 1.0] and dr == [0.0, 1.0, 0.0]
    assert osc.scatter(metal, (0, -2, 0), (0, 2, 0), (0, 0, 0), (0, 1, 0), True)[0] == SCATTER_RAY  # leaves into the surface: kept
    # Lambertian -> ScatterRecord::PDF(CosinePDF) with the texture's albedo; Isotropic -> SpherePDF; DiffuseLight -> None
    assert osc.scatter(lam, (0, 2, 0), (0, -2, 0), (0, 0, 0), (0, 1, 0), True)[:2] == (SCATTER_COSINE, [0.25, 0.5, 0.75])
    assert osc.scatter(iso, (0, 2, 0), (0, -2, 0), (0, 0, 0), (0, 1, 0), True)[:2] == (SCATTER_SPHERE, [0.125, 0.25, 0.5])
    assert osc.scatter(light, (0, 2, 0), (0, -2, 0), (0, 0, 0), (0, 1, 0), True)[0] == SCATTER_NONE


def test_cosine_pdf(orc):
    """CosinePDF, pdf.rs:36-64 over OrthonormalBasis::new, onb.rs:8-24: normal (0,1,0) -> u = unit(n x (1,0,0)) = (0,0,-1), w = u x n = (1,0,0)."""
    inv_pi = 1.0 / math.pi
    brdf, pdf, gen = orc.cosine_pdf((0.5, 0.25, 1.0), (0, 1, 0), (0, 2, 0), xi=(0.25, 0.25))
    assert pdf == max(0.0, 1.0 / math.pi)                       # cos = 1
    assert brdf == [inv_pi * (0.5 * 1.0), inv_pi * (0.25 * 1.0), inv_pi * (1.0 * 1.0)]   # Vec3 / f64 is (1/rhs) * v (vec3.rs:222-228)
    # random_cosine_direction (vec3.rs:333-343) with r1 = r2 = 1/4: (sin(pi/2) sqrt(1/4), sqrt(3/4), cos(pi/2) sqrt(1/4)) in the basis (u, n, w)
    phi = 2.0 * math.pi * 0.25
    x, y, z = math.sin(phi) * math.sqrt(0.25), math.sqrt(1.0 - 0.25), math.cos(phi) * math.sqrt(0.25)
    assert gen == [x * 0.0 + y * 0.0 + z * 1.0, x * 0.0 + y * 1.0 + z * 0.0, x * -1.0 + y * 0.0 + z * 0.0]
    brdf, pdf, _ = orc.cosine_pdf((0.5, 0.25, 1.0), (0, 1, 0), (0, -1, 0))
    assert pdf == 0.0 and brdf == [0.0, 0.0, 0.0]               # below the surface: max(0, cos/pi), albedo * max(cos, 0) / pi
    brdf, pdf, _ = orc.cosine_pdf((1.0, 1.0, 1.0), (0, 1, 0), (3, 4, 0))
    cos = (1.0 / 5.0) * 4.0
    assert pdf == cos / math.pi and brdf[0] == inv_pi * (1.0 * cos)
    # |n.x| > 0.9 switches the helper axis to (0,1,0) (onb.rs:9-13): normal (1,0,0) -> u = unit((1,0,0) x (0,1,0)) = (0,0,1), w = u x n = (0,1,0)
    _, _, gen = orc.cosine_pdf((1, 1, 1), (1, 0, 0), (1, 0, 0), xi=(0.25, 0.25))
    assert gen == [x * 0.0 + y * 1.0 + z * 0.0, x * 0.0 + y * 0.0 + z * 1.0, x * 1.0 + y * 0.0 + z * 0.0]


def test_light_pdfs(rt, orc):
    """Quad / Sphere / Transform / Hittables pdf_value and random: quad.rs:108-125, sphere.rs:63-73,114-144, shapes.rs:117-132, hits.rs:52-75."""
    b = rt.Builder(1)
    e = b.empty()
    quad = b.quad([-1, 4, -1], [2, 0, 0], [0, 0, 2], e)     # n = u x v = (0,-4,0): normal (0,-1,0), area 4
    hs = b.finish(b.list([b.sphere([0, -50, 0], 1.0, e)]), b.list([quad]))
    osc = orc.OracleScene(hs)
    # t = (D - n.o)/(n.d) = (-4 - 0)/(-2) = 2; distance^2 = t^2 |d|^2 = 16; cosine = |d.n / |d|| = 1; pdf = 16 / (1 * 4) = 4
    assert osc.lights_pdf_value((0, 0, 0), (0, 2, 0)) == 4.0
    assert osc.lights_pdf_value((0, 0, 0), (1, 0, 0)) == 0.0
    # 3-4-5 direction towards (0.75, 4, 0) = 4 * (3/16, 1, 0): t = 4 with d = (3/16, 1, 0); restated
    d = (0.1875, 1.0, 0.0)
    dl2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2]
    assert osc.lights_pdf_value((0, 0, 0), d) == (4.0 * 4.0 * dl2) / (abs(-1.0 / math.sqrt(dl2)) * 4.0)
    # random: anchor + r1 u + r2 v - origin, normalised: r1 = r2 = 1/2 -> (0,4,0) -> (0,1,0)
    assert osc.lights_random((0, 0, 0), 0, 0.5, 0.5) == [0.0, 1.0, 0.0]
    # two lights: Hittables::pdf_value is the mean (hits.rs:52-67); Sphere::pdf_value = 1 / (2 pi (1 - sqrt(1 - r^2/dist^2)))
    b = rt.Builder(1)
    e = b.empty()
    lights = b.list([b.quad([-1, 4, -1], [2, 0, 0], [0, 0, 2], e), b.sphere([0, 0, 4], 2.0, e),
                     b.transform(b.quad([-1, 0, -1], [2, 0, 0], [0, 0, 2], e), offset=[0, -8, 0], scale=[2, 2, 2])])
    osc = orc.OracleScene(b.finish(b.list([b.sphere([0, -50, 0], 1.0, e)]), lights))
    sphere_pdf = 1.0 / (2.0 * math.pi * (1.0 - math.sqrt(1.0 - 2.0 * 2.0 / 16.0)))
    assert osc.lights_pdf_value((0, 0, 0), (0, 2, 0)) == (4.0 + 0.0 + 0.0) / 3.0
    assert osc.lights_pdf_value((0, 0, 0), (0, 0, 1)) == (0.0 + sphere_pdf + 0.0) / 3.0
    # inside the sphere: cos_theta_max is NaN -> uniform 1/(4 pi) (sphere.rs:123-126)
    assert osc.lights_pdf_value((0, 0, 4), (0, 0, 1)) == (0.0 + 1.0 / (4.0 * math.pi) + 0.0) / 3.0
    # Transform::pdf_value delegates in local space WITHOUT the Jacobian (shapes.rs:117-123): local origin (0,4,0), local d (0,-0.5,0):
    # local quad at y = 0, normal (0,-1,0), area 4: t = (0 - (-4)) / 0.5 = 8, distance^2 = 64 * 0.25 = 16, cosine 1 -> 4 (the world-space value is 1)
    assert osc.lights_pdf_value((0, 0, 0), (0, -1, 0)) == (0.0 + 0.0 + 4.0) / 3.0
    assert osc.lights_random((0, 0, 0), 2, 0.5, 0.5) == [0.0, -1.0, 0.0]

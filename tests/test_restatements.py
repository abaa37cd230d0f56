"""CheckerTexture and ImageTexture of the oracle against an independent restatement in Python written from the reference's source
(texture.rs:38-174, utils/image.rs:63-82): the checker's floor / parity rule with negative coordinates, the nearest lookup with its
flipped v and the clamp at v == 1, the sRGB decode of non-raw images (palette's Srgba::into_linear is the standard piecewise curve),
the bilinear lookup of raw images in binary32, and the cyan of a missing image."""
import math

import numpy as np

f32 = np.float32


def flat_material(hs, obj):
    """(material index, texture index) of object `obj` in the FLATTENED description - the host mirror numbers materials and textures in
    the order the flattening meets them, not in the order the builder created them."""
    import ctypes as C
    mi = int(hs.objects()[obj]["material"])
    tex = C.c_uint32.from_address(hs.desc.contents.materials + 176 * mi + 4).value
    return mi, tex


def srgb_to_linear(c):
    c = float(c)
    return c / 12.92 if c <= 0.04045 else ((c + 0.055) / 1.055) ** 2.4


def abs_fract(x):
    return x - math.floor(x)


def nearest(img, u, v, raw):  # texture.rs:108-116, image.rs:63-82
    h, w, _ = img.shape
    uu, vv = abs_fract(u), 1.0 - abs_fract(v)
    i, j = min(int(uu * w), w - 1), min(int(vv * h), h - 1)
    px = img[j, i]
    return [float(px[k]) if raw else srgb_to_linear(px[k]) for k in range(3)]


def bilinear(img, u, v):  # texture.rs:118-148 (raw images: no decode)
    h, w, _ = img.shape
    uu, vv = abs_fract(u), 1.0 - abs_fract(v)
    x, y = uu * w - 0.5, vv * h - 0.5
    x0, y0 = int(max(math.floor(x), 0.0)), int(max(math.floor(y), 0.0))
    x1, y1 = min(x0 + 1, w - 1), min(y0 + 1, h - 1)
    dx, dy = f32(x - x0), f32(y - y0)
    cx0, cy0 = min(x0, w - 1), min(y0, h - 1)  # pixel_data clamps
    out = []
    for k in range(3):
        p00, p10, p01, p11 = img[cy0, cx0, k], img[cy0, x1, k], img[y1, cx0, k], img[y1, x1, k]
        v0 = p00 * (f32(1.0) - dx) + p10 * dx
        v1 = p01 * (f32(1.0) - dx) + p11 * dx
        out.append(float(v0 * (f32(1.0) - dy) + v1 * dy))
    return out


def test_checker_and_image_textures_match_a_plain_restatement(rt, orc):
    rng = np.random.default_rng(23)
    img = rng.uniform(0.0, 1.0, (5, 7, 4)).astype(np.float32)
    b = rt.Builder(3)
    even, odd = b.solid(0.9, 0.1, 0.2), b.solid(0.05, 0.6, 0.3)
    scale = 0.32
    chk = b.checker(scale, even, odd)
    t_srgb = b.image(img)                       # ImageTexture::new on a PNG-like image: decoded
    t_lin = b.image(img, linear_format=True)    # .hdr / .exr: used as stored
    t_raw = b.image(img, raw=True)              # new_raw_image: bilinear, no decode
    t_missing = b.image_missing()
    hs = b.finish(b.list([b.sphere([0, 0, 0], 1.0, b.lambertian(chk)), b.sphere([3, 0, 0], 1.0, b.lambertian(t_srgb)),
                          b.sphere([6, 0, 0], 1.0, b.lambertian(t_lin)), b.sphere([9, 0, 0], 1.0, b.lambertian(t_raw)),
                          b.sphere([12, 0, 0], 1.0, b.lambertian(t_missing))]), width=8, spp=1)
    osc = orc.OracleScene(hs)
    chk, t_srgb, t_lin, t_raw, t_missing = (flat_material(hs, k)[1] for k in range(5))
    inv = 1.0 / scale
    for _ in range(500):
        p = list(rng.uniform(-9.0, 9.0, 3))
        s = math.floor(inv * p[0]) + math.floor(inv * p[1]) + math.floor(inv * p[2])
        want = (0.9, 0.1, 0.2) if s % 2 == 0 else (0.05, 0.6, 0.3)
        assert tuple(osc.texture_value(chk, 0.3, 0.4, p)) == want
    uvs = [tuple(rng.uniform(-2.0, 3.0, 2)) for _ in range(400)] + [(0.0, 0.0), (1.0, 1.0), (0.0, 1.0), (-1.0, 2.0), (0.999999, 1e-9), (0.5, -0.0)]
    for u, v in uvs:
        assert np.allclose(osc.texture_value(t_srgb, u, v, [0, 0, 0]), nearest(img, u, v, raw=False), rtol=2e-6, atol=1e-7)
        assert np.allclose(osc.texture_value(t_lin, u, v, [0, 0, 0]), nearest(img, u, v, raw=True), rtol=0, atol=0)
        assert np.allclose(osc.texture_value(t_raw, u, v, [0, 0, 0]), bilinear(img, u, v), rtol=3e-7, atol=1e-7)
        assert tuple(osc.texture_value(t_missing, u, v, [0, 0, 0])) == (0.0, 1.0, 1.0)


def test_tone_map_matches_a_plain_restatement(orc):
    """Color::to_rgb (utils/color.rs:14-36): the ACES fit with its constants and clamp, then the sRGB encode.  The `palette` crate is not
    vendored with the reference, so the encode is the standard piecewise curve rounded to nearest - what this pins is the ACES arithmetic
    and that both tone-map settings go through the same encode (the 8-bit codes are stated to +-1 against the real crate)."""
    rng = np.random.default_rng(31)
    c = np.concatenate([rng.uniform(0.0, 1.5, (300, 3)), rng.uniform(0.0, 30.0, (100, 3)), np.array([[0.0, 0.0, 0.0], [1.0, 1.0, 1.0], [0.0031308, 0.5, 100.0]])])

    def encode(x):
        e = 12.92 * x if x <= 0.0031308 else 1.055 * x ** (1.0 / 2.4) - 0.055
        return int(math.floor(min(max(e, 0.0), 1.0) * 255.0 + 0.5))

    def aces(x):
        return min(max((x * (2.51 * x + 0.03)) / (x * (2.43 * x + 0.59) + 0.14), 0.0), 1.0)

    plain = orc.tonemap(c, toon_map=0)
    mapped = orc.tonemap(c, toon_map=1)
    for row, p_row, m_row in zip(c, plain, mapped):
        assert [encode(x) for x in row] == list(p_row)
        assert [encode(aces(x)) for x in row] == list(m_row)


def test_portal_transparent_and_mix_scatter(rt, orc):
    """Portal::scatter (material/portal.rs:14-30: direction turned by q v q*, quaternion.rs:72-104), Transparent (material.rs:209-218) and
    Mix (material.rs:249-257: `Random::f64() > ratio` takes mat1) with the draw forced."""
    def qmul(a, b):
        return (a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0])

    b = rt.Builder(4)
    q = b.quat_axis_angle([0.3, 1.0, -0.2], 77.0)
    portal = b.portal([0.9, 0.8, 0.7], [1.0, -2.0, 0.5], q)
    clear = b.transparent()
    red = b.metal([0.8, 0.1, 0.1], 0.0)
    mix = b.mix(red, clear, 0.25)
    hs = b.finish(b.list([b.sphere([0, 0, 0], 1.0, portal), b.sphere([3, 0, 0], 1.0, clear), b.sphere([6, 0, 0], 1.0, mix)]), width=8, spp=1)
    osc = orc.OracleScene(hs)
    portal, clear, mix = (flat_material(hs, k)[0] for k in range(3))
    rng = np.random.default_rng(9)
    RAY = 3  # ScatterKind::SCATTER_RAY of the oracle
    for _ in range(50):
        dvec = list(rng.normal(size=3))
        n = list(rng.normal(size=3))
        n = [x / math.sqrt(sum(y * y for y in n)) for x in n]
        kind, att, out, err = osc.scatter(portal, [0, 0, 5], dvec, [0.1, 0.2, 0.3], n, True)
        r = qmul(qmul(tuple(q), (0.0, *dvec)), (q[0], -q[1], -q[2], -q[3]))
        assert kind == RAY and not err and att == [0.9, 0.8, 0.7] and out == list(r[1:])
        kind, att, out, err = osc.scatter(clear, [0, 0, 5], dvec, [0.1, 0.2, 0.3], n, False)
        assert kind == RAY and att == [1.0, 1.0, 1.0] and out == dvec
        for xi0 in (0.0, 0.25, 0.2500001, 0.9):
            kind, att, out, err = osc.scatter(mix, [0, 0, 5], dvec, [0.1, 0.2, 0.3], n, True, xi=(xi0, 0.5))
            assert kind == RAY and att == ([0.8, 0.1, 0.1] if xi0 > 0.25 else [1.0, 1.0, 1.0])


def test_remapped_material_remaps_uv_and_normal(rt, orc):
    """RemappedMaterial::remap_record (shapes/obj.rs:32-62) without a normal map: the inner material sees the face's texture coordinate
    tex_ori + u tex_u + v tex_v and the interpolated, normalised vertex normal.  A mirror inside reveals the normal, an image-textured
    Lambertian the texture coordinate."""
    rng = np.random.default_rng(41)
    img = rng.uniform(0.0, 1.0, (6, 5, 4)).astype(np.float32)
    b = rt.Builder(6)
    mirror = b.metal([1.0, 1.0, 1.0], 0.0)
    painted = b.lambertian(b.image(img, linear_format=True))
    pos = [[0.0, 0.0, 0.0], [2.0, 0.0, 0.0], [0.0, 0.0, -2.0]]
    uv = [[0.1, 0.2], [0.9, 0.3], [0.2, 0.8]]
    nrm = [[0.1, 1.0, 0.0], [-0.2, 0.9, 0.3], [0.0, 0.8, -0.4]]
    m_mirror, m_painted = b.remapped(mirror, pos, uv, nrm), b.remapped(painted, pos, uv, nrm)
    hs = b.finish(b.list([b.obj_face(mirror, pos, uv, nrm), b.sphere([5, 0, 0], 1.0, m_mirror), b.sphere([8, 0, 0], 1.0, m_painted)]), width=8, spp=1)
    osc = orc.OracleScene(hs)
    m_mirror, m_painted = flat_material(hs, 1)[0], flat_material(hs, 2)[0]
    COSINE, RAY = 1, 3
    h, w, _ = img.shape
    for _ in range(100):
        u = rng.uniform(0.0, 1.0)
        v = rng.uniform(0.0, 1.0 - u)
        dvec = rng.normal(size=3)
        n = (1.0 - u - v) * np.array(nrm[0]) + u * np.array(nrm[1]) + v * np.array(nrm[2])
        n = n / np.sqrt(n @ n)
        ud = dvec / np.sqrt(dvec @ dvec)
        refl = ud - 2.0 * (ud @ n) * n
        refl = refl / np.sqrt(refl @ refl)
        kind, att, out, err = osc.scatter(m_mirror, [0, 3, 0], list(dvec), [0.5, 0.0, -0.5], [0.0, 1.0, 0.0], True, u=u, v=v)
        assert kind == RAY and not err and np.allclose(out, refl, rtol=0, atol=1e-12)
        tu = uv[0][0] + u * (uv[1][0] - uv[0][0]) + v * (uv[2][0] - uv[0][0])
        tv = uv[0][1] + u * (uv[1][1] - uv[0][1]) + v * (uv[2][1] - uv[0][1])
        want = img[min(int((1.0 - (tv - math.floor(tv))) * h), h - 1), min(int((tu - math.floor(tu)) * w), w - 1), :3]
        kind, att, out, err = osc.scatter(m_painted, [0, 3, 0], list(dvec), [0.5, 0.0, -0.5], [0.0, 1.0, 0.0], True, u=u, v=v)
        assert kind == COSINE and not err and att == [float(x) for x in want]


def test_normal_mapped_remap_and_triangle_light(rt, orc):
    """The normal-map branch of remap_record with uv_local_to_world (shapes/obj.rs:44-53, 196-210): the mapped normal is
    unit(u_vec c.x + v_vec c.y + n c.z) with c = 2 texel - 1, revealed by a mirror inside.  Triangle::pdf_value / random
    (shapes/triangle.rs:104-128: area |u x v| / 2, the fold of the unit square onto the triangle)."""
    rng = np.random.default_rng(57)
    nmap = rng.uniform(0.3, 1.0, (4, 4, 4)).astype(np.float32)
    b = rt.Builder(8)
    t_n = b.image(nmap, linear_format=True)
    mirror = b.metal([1.0, 1.0, 1.0], 0.0)
    pos = [[0.0, 0.0, 0.0], [2.0, 0.5, 0.0], [0.3, 0.0, -2.0]]
    uv = [[0.1, 0.2], [0.9, 0.3], [0.2, 0.8]]
    nrm = [[0.1, 1.0, 0.0], [-0.2, 0.9, 0.3], [0.0, 0.8, -0.4]]
    mm = b.remapped(mirror, pos, uv, nrm, normal_tex=t_n)
    tri = b.triangle([1.0, 4.0, -1.0], [2.0, 0.0, 0.5], [0.0, 0.5, 2.0], b.diffuse_light(b.solid(5, 5, 5)))
    lights = b.list([b.triangle([1.0, 4.0, -1.0], [2.0, 0.0, 0.5], [0.0, 0.5, 2.0], b.empty())])
    hs = b.finish(b.list([b.sphere([5, 0, 0], 1.0, mm), tri]), lights, width=8, spp=1)
    osc = orc.OracleScene(hs)
    mm = flat_material(hs, 0)[0]
    P = [np.array(p) for p in pos]
    tex_u, tex_v = np.array(uv[1] + [0.0]) - np.array(uv[0] + [0.0]), np.array(uv[2] + [0.0]) - np.array(uv[0] + [0.0])
    world_u, world_v = P[1] - P[0], P[2] - P[0]
    ua = tex_v[1] / (-tex_u[1] * tex_v[0] + tex_u[0] * tex_v[1])
    ub = tex_u[1] / (tex_u[1] * tex_v[0] - tex_u[0] * tex_v[1])
    va = tex_v[0] / (tex_u[1] * tex_v[0] - tex_u[0] * tex_v[1])
    vb = tex_u[0] / (-tex_u[1] * tex_v[0] + tex_u[0] * tex_v[1])
    unit_ = lambda x: x / np.sqrt(x @ x)
    u_vec, v_vec = unit_(world_u * ua + world_v * ub), unit_(world_u * va + world_v * vb)
    for _ in range(100):
        u = rng.uniform(0.0, 1.0)
        v = rng.uniform(0.0, 1.0 - u)
        dvec = rng.normal(size=3)
        n = unit_((1.0 - u - v) * np.array(nrm[0]) + u * np.array(nrm[1]) + v * np.array(nrm[2]))
        tc = np.array(uv[0] + [0.0]) + u * tex_u + v * tex_v
        c = np.array(nearest(nmap, tc[0], tc[1], raw=True)) * 2.0 - 1.0
        n2 = unit_(u_vec * c[0] + v_vec * c[1] + n * c[2])
        ud = unit_(dvec)
        refl = unit_(ud - 2.0 * (ud @ n2) * n2)
        kind, att, out, err = osc.scatter(mm, [0, 3, 0], list(dvec), [0.5, 0.0, -0.5], [0.0, 1.0, 0.0], True, u=u, v=v)
        assert kind == 3 and not err and np.allclose(out, refl, rtol=0, atol=1e-11)
    # the triangle light
    A, U, V = np.array([1.0, 4.0, -1.0]), np.array([2.0, 0.0, 0.5]), np.array([0.0, 0.5, 2.0])
    nvec = np.cross(U, V)
    area, normal = np.sqrt(nvec @ nvec) / 2.0, unit_(nvec)
    origin = np.array([1.5, 0.0, 0.0])
    for _ in range(100):
        r1, r2 = rng.uniform(size=2)
        ul, vl = (1.0 - r2, 1.0 - r1) if r1 + r2 > 1.0 else (r1, r2)
        want_dir = unit_(A + ul * U + vl * V - origin)
        got = osc.lights_random(list(origin), 0, r1, r2)
        assert np.allclose(got, want_dir, rtol=0, atol=1e-12)
        # a direction through the triangle: pdf = t^2 |d|^2 / (|d . n| / |d| * area), with t from the plane equation
        d = want_dir * rng.uniform(0.5, 3.0)
        t = (normal @ A - normal @ origin) / (normal @ d)
        want_pdf = t * t * (d @ d) / (abs(d @ normal / np.sqrt(d @ d)) * area)
        assert osc.lights_pdf_value(list(origin), list(d)) == __import__("pytest").approx(want_pdf, rel=1e-12)
    assert osc.lights_pdf_value(list(origin), [0.0, -1.0, 0.0]) == 0.0  # away from the light


def test_camera_get_ray_matches_a_plain_restatement(rt, orc):
    """Camera::get_ray (camera.rs:247-273) with sample_square_stratified and defocus_disk_sample (vec3.rs:63-69): the oracle's rays
    against the formulas restated on the camera block, the draws taken through the addressed-Philox hook (slots 0, 1, 2 of
    include/rt2025_rng.h: jitter, defocus disk, time; pixel = j * width + i, segment 0)."""
    import ctypes as C
    for name, params in (("book1_final", [120, 25, 5]), ("book2_final", [80, 16, 5])):  # book 1 has a defocus disk, book 2 does not
        hs = rt.named_scene(name, seed=3, params=params)
        cam = hs.camera
        L = orc.lib()
        v3 = lambda a: np.array(list(a))
        rng = np.random.default_rng(77)
        px = [(int(rng.integers(0, cam.image_width)), int(rng.integers(0, cam.image_height))) for _ in range(60)]
        spp = cam.sqrt_spp * cam.sqrt_spp
        for sidx in (0, 1, spp // 2, spp - 1):
            got = orc.camera_rays(cam, 9, px, sidx)
            for (i, j), ray in zip(px, got):
                pixel = j * cam.image_width + i
                pair = (C.c_double * 2)()
                draw = lambda slot: (L.orc_kat_draw(9, pixel, sidx, 0, slot, pair), (pair[0], pair[1]))[1]
                s_i, s_j = sidx // cam.sqrt_spp, sidx % cam.sqrt_spp
                ja, jb = draw(0)
                ox = ((s_i + ja) * cam.recip_sqrt_spp) - 0.5
                oy = ((s_j + jb) * cam.recip_sqrt_spp) - 0.5
                sample = v3(cam.pixel00_loc) + ((i + ox) * v3(cam.pixel_delta_u)) + ((j + oy) * v3(cam.pixel_delta_v))
                if cam.defocus_angle_in_degrees <= 0.0:
                    origin = v3(cam.center)
                else:
                    da, db = draw(1)
                    theta, r = (2.0 * math.pi) * da, math.sqrt(db)
                    origin = v3(cam.center) + ((r * math.cos(theta)) * v3(cam.defocus_disk_u)) + ((r * math.sin(theta)) * v3(cam.defocus_disk_v))
                assert np.array_equal(ray["origin"], origin)
                assert np.array_equal(ray["direction"], sample - origin)
                assert ray["time"] == draw(2)[0]
        assert (cam.defocus_angle_in_degrees > 0.0) == (name == "book1_final")

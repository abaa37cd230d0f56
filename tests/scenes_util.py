"""Small scene builders shared by the CPU and GPU tests (all through the host mirror)."""
import numpy as np


def random_graph_scene(rt, seed, n_prims=60, with_transforms=True, with_media=False, with_lights=True, width=24, spp=4, depth=6):
    """A random object graph mixing every container the reference has."""
    rng = np.random.default_rng(seed)
    b = rt.Builder(seed)
    mats = [
        b.lambertian(b.solid(*rng.uniform(0.2, 0.9, 3))),
        b.lambertian(b.checker(0.7, b.solid(0.9, 0.9, 0.9), b.solid(0.1, 0.3, 0.1))),
        b.metal(rng.uniform(0.5, 1, 3), 0.2),
        b.dielectric(b.solid(1, 1, 1), 1.5),
        b.lambertian(b.noise(1.3)),
        b.empty(),
    ]
    light_mat = b.diffuse_light(b.solid(6, 6, 6))

    def prim():
        k = rng.integers(0, 3)
        m = mats[rng.integers(0, len(mats))]
        c = rng.uniform(-4, 4, 3)
        if k == 0:
            return b.sphere(c, rng.uniform(0.2, 0.9), m)
        u, v = rng.uniform(-1.5, 1.5, 3), rng.uniform(-1.5, 1.5, 3)
        if k == 1:
            return b.quad(c, u, v, m)
        return b.triangle(c, u, v, m)

    prims = [prim() for _ in range(n_prims)]
    third = n_prims // 3
    top = []
    top.append(b.bvh(prims[:third]))
    inner = b.list(prims[third:2 * third])
    if with_transforms:
        q = b.quat_axis_angle(rng.normal(size=3), rng.uniform(0, 360))
        inner = b.transform(inner, offset=rng.uniform(-1, 1, 3), quat=q, scale=[1.7, 1.7, 1.7])
    top.append(inner)
    rest = prims[2 * third:]
    if with_transforms:
        q2 = b.quat_axis_angle([0, 1, 0], 33.0)
        sub = b.transform(b.bvh(rest[: len(rest) // 2]), offset=[0.5, -0.5, 1.0], quat=q2)
        top.append(sub)
        top.extend(rest[len(rest) // 2:])
    else:
        top.extend(rest)
    light_quad = b.quad([-1.5, 5.5, -1.5], [3, 0, 0], [0, 0, 3], light_mat)
    top.append(light_quad)
    if with_media:
        top.append(b.medium(b.sphere([0, 0, 0], 3.0, b.empty()), 0.4, b.solid(0.8, 0.8, 0.9)))
        top.append(b.medium(b.box([-6, -6, -6], [6, 6, 6], b.empty()), 0.02, b.solid(1, 1, 1)))
    world = b.list(top)
    lights = rt.RT_NONE
    if with_lights:
        ls = [b.quad([-1.5, 5.5, -1.5], [3, 0, 0], [0, 0, 3], b.empty()), b.sphere([2.5, 2.0, 0.0], 0.6, b.empty())]
        lights = b.list(ls)
    bg = b.gradient([0.05, 0.05, 0.05], [0.1, 0.15, 0.3])
    hs = b.finish(world, lights, width=width, spp=spp, max_depth=depth, vfov=55.0, look_from=(0, 1, 11), look_at=(0, 0, 0),
                  background=bg)
    hs._builder = b
    return hs


def random_rays(rng, n, extent=6.0, time=True):
    o = rng.uniform(-extent, extent, (n, 3))
    d = rng.normal(size=(n, 3))
    t = rng.uniform(0, 1, n) if time else None
    return o, d, t


def scene_rays(rt, orc, hs, n, seed):
    """Half camera rays, half incoherent rays through the scene volume."""
    rng = np.random.default_rng(seed)
    cam = hs.camera
    px = np.stack([rng.integers(0, cam.image_width, n // 2), rng.integers(0, cam.image_height, n // 2)], axis=1)
    prim = orc.camera_rays(cam, 5, px, 0)
    return prim


def compare_hits(rt, got, want, rel=0.0):
    """IDs bit-exact; t bit-exact by default (the north star allows 1e-5 relative): every primitive,
    transformed or not, is intersected with the reference's own binary64 arithmetic."""
    assert np.array_equal(got["prim_id"], want["prim_id"]), \
        f"{int((got['prim_id'] != want['prim_id']).sum())} primitive id mismatches"
    assert np.array_equal(got["inst_id"], want["inst_id"])
    hit = want["prim_id"] != rt.RT_NONE
    if hit.any():
        err = np.abs(got["t"][hit] - want["t"][hit]) / np.abs(want["t"][hit])
        assert err.max() <= rel, f"t relative error {err.max():.3e}"
        # u, v go through acos/atan2 for spheres: libm last-bit differences only
        assert np.abs(got["u"][hit] - want["u"][hit]).max() < 1e-6
        assert np.abs(got["v"][hit] - want["v"][hit]).max() < 1e-6
    assert np.all(np.isinf(got["t"][~hit]))


def disney_scene(rt, with_lights=True, width=64, spp=16, depth=8, seed=2):
    """Five Disney spheres covering every lobe + one with a base-colour image texture (obj.rs:271-293)."""
    rng = np.random.default_rng(seed)
    b = rt.Builder(seed)
    white = b.lambertian(b.solid(0.7, 0.7, 0.7))
    light = b.diffuse_light(b.solid(8, 8, 8))
    tex = b.image(rng.uniform(0.1, 1, (8, 16, 4)).astype(np.float32))
    mats = [b.disney((0.8, 0.3, 0.2), roughness=0.4),
            b.disney((0.9, 0.8, 0.3), metallic=1.0, roughness=0.3, anisotropic=0.5),
            b.disney((0.9, 0.9, 0.9), spec_trans=1.0, roughness=0.1, ior=1.5),
            b.disney((0.2, 0.4, 0.8), clearcoat=1.0, clearcoat_gloss=0.9, sheen=0.5, sheen_tint=0.5, roughness=0.6, specular_tint=0.3),
            b.disney((0.6, 0.8, 0.6), thin=1.0, flatness=0.5, diff_trans=0.5, spec_trans=0.3, roughness=0.5),
            b.disney((1, 1, 1), tex=tex, roughness=0.7, metallic=0.2, spec_trans=0.2, clearcoat=0.3)]
    objs = [b.quad([-9, -1, -8], [18, 0, 0], [0, 0, 16], white), b.quad([-2, 5, -2], [4, 0, 0], [0, 0, 4], light)]
    for k, m in enumerate(mats):
        objs.append(b.sphere([k * 2.2 - 5.5, 0, 0], 1.0, m))
    lights = b.list([b.quad([-2, 5, -2], [4, 0, 0], [0, 0, 4], b.empty())]) if with_lights else rt.RT_NONE
    hs = b.finish(b.list([b.bvh(objs)]), lights, width=width, aspect=2.0, spp=spp, max_depth=depth, vfov=38, look_from=(0, 3, 13),
                  look_at=(0, 0, 0), background=b.solid(0.1, 0.1, 0.15))
    hs._builder = b
    return hs


def obj_mesh_scene(rt, width=56, spp=16, depth=8, seed=3):
    """A small mesh built face by face the way shapes/obj.rs load_object does: every triangle carries its
    own RemappedMaterial (uv frame, smooth vertex normals, raw normal map), models in their own BVH."""
    rng = np.random.default_rng(seed)
    b = rt.Builder(seed)
    albedo = b.image(rng.uniform(0.2, 1, (16, 16, 4)).astype(np.float32))
    nm = rng.uniform(0.35, 0.65, (8, 8, 4)).astype(np.float32)
    nm[..., 2] = rng.uniform(0.8, 1.0, (8, 8))
    normal_map = b.image(nm, raw=True)
    alpha = rng.uniform(0, 1, (8, 8, 4)).astype(np.float32)
    alpha[..., 3] = (rng.uniform(0, 1, (8, 8)) > 0.3)
    alpha_tex = b.image(alpha)
    inner = [b.lambertian(albedo), b.disney((1, 1, 1), tex=albedo, roughness=0.4, metallic=0.3),
             b.diffuse_light(b.solid(0.4, 0.3, 0.2), inner=b.disney((0.7, 0.7, 0.9), roughness=0.5))]
    # Mix::from_image(Transparent, mat, alpha texture) as for `dissolve_texture` (obj.rs:318-325) and the
    # constant-ratio Mix of `dissolve < 1` (obj.rs:326-330)
    inner.append(b.mix_image(b.transparent(), inner[0], alpha_tex))
    inner.append(b.mix(b.transparent(), inner[1], 0.6))
    n = 6
    xs = np.linspace(-3, 3, n + 1)

    def height(x, z):
        return 0.4 * np.sin(1.3 * x) * np.cos(1.1 * z)

    def vertex(i, j):
        x, z = xs[i], xs[j]
        p = np.array([x, height(x, z), z])
        dx, dz = 0.4 * 1.3 * np.cos(1.3 * x) * np.cos(1.1 * z), -0.4 * 1.1 * np.sin(1.3 * x) * np.sin(1.1 * z)
        nrm = np.array([-dx, 1.0, -dz])
        return p, np.array([i / n, j / n]), nrm / np.linalg.norm(nrm)

    models = []
    for k, mat in enumerate(inner):
        faces = []
        for i in range(n):
            for j in range(n):
                if (i + j) % len(inner) != k:
                    continue
                quad = [vertex(i, j), vertex(i + 1, j), vertex(i + 1, j + 1), vertex(i, j + 1)]
                for tri in ((0, 1, 2), (0, 2, 3)):
                    pos = [quad[t][0] for t in tri]
                    uv = [quad[t][1] for t in tri]
                    nr = [quad[t][2] for t in tri]
                    f = b.obj_face(mat, pos, uv, nr, normal_map if k % 2 == 0 else rt.RT_NONE)
                    if f != rt.RT_NONE:
                        faces.append(f)
        models.append(b.bvh(faces))
    wavefont = b.list(models)
    light = b.quad([-2, 5, -2], [4, 0, 0], [0, 0, 4], b.diffuse_light(b.solid(9, 9, 9)))
    floor = b.quad([-8, -2, -8], [16, 0, 0], [0, 0, 16], b.lambertian(b.solid(0.6, 0.6, 0.6)))
    lights = b.list([b.quad([-2, 5, -2], [4, 0, 0], [0, 0, 4], b.empty())])
    hs = b.finish(b.list([wavefont, light, floor]), lights, width=width, spp=spp, max_depth=depth, vfov=40, look_from=(0, 5, 9), look_at=(0, 0, 0),
                  background=b.solid(0.15, 0.15, 0.2))
    hs._builder = b
    return hs

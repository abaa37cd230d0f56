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

"""Debug helper: print details for failing GPU parity cases."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc  # noqa: E402
from scenes_util import random_graph_scene, random_rays  # noqa: E402

rt = orc.rt


def show_hits(tag, rays, got, want):
    bad = np.nonzero((got["prim_id"] != want["prim_id"]) | ~((got["t"] == want["t"]) | (np.isinf(got["t"]) & np.isinf(want["t"]))))[0]
    print(f"[{tag}] mismatches {len(bad)} / {len(rays)}")
    for i in bad[:12]:
        print("   ray", i, rays[i]["origin"], rays[i]["direction"], rays[i]["time"], "got", got[i], "want", want[i])


def ellipsoids():
    b = rt.Builder(9)
    m = b.empty()
    e1 = b.transform(b.sphere([0, 0, 0], 1.0, m), offset=[-2, 0, 0], quat=b.quat_axis_angle([0, 0, 1], 30.0), scale=[2.0, 0.5, 1.0])
    inner = b.transform(b.bvh([b.sphere([0, 0, 0], 0.7, m), b.quad([-1, -1, 1], [2, 0, 0], [0, 2, 0], m),
                               b.triangle([0, 1, -1], [1, 0, 0], [0, 1, 1], m)]), offset=[0.5, 0, 0], scale=[1, 2, 1])
    e2 = b.transform(b.list([inner, b.sphere_moving([0, -2, 0], [1, -2, 0], 0.5, m)]), offset=[2.5, 0.5, 0],
                     quat=b.quat_axis_angle([1, 1, 0], 50.0), scale=[0.8, 0.8, 1.6])
    e3 = b.transform(b.sphere([0, 0, 0], 1.0, m), offset=[0, 3, 0], scale=[-1.0, 1.0, 1.5])
    hs = b.finish(b.list([e1, e2, e3]))
    sc, osc = rt.Scene(hs), orc.OracleScene(hs)
    rng = np.random.default_rng(4)
    o, d, t = random_rays(rng, 40000, extent=5.0)
    rays = rt.make_rays(o, d, t)
    got, _ = sc.closest_hit(rays)
    want = osc.closest_hit(rays, mode=0)
    show_hits("ellipsoids", rays, got, want)
    print("   objects:", hs.objects()[["kind", "bbox"]])


def edge():
    b = rt.Builder(1)
    m = b.empty()
    hs = b.finish(b.list([b.bvh([b.sphere([0, 0, 0], 1.0, m), b.quad([-1, -1, -3], [2, 0, 0], [0, 2, 0], m),
                                 b.triangle([-1, -1, 3], [2, 0, 0], [0, 2, 0], m), b.sphere([4, 0, 0], 0.5, m)])]))
    sc, osc = rt.Scene(hs), orc.OracleScene(hs)
    o = [[0, 0, 5], [0, 0, 5], [-5, 0, 0], [0, 0, 0], [0, 0, 5], [float("nan"), 0, 5], [0.25, 0.25, -10], [0, 0, 5]]
    d = [[0, 0, -1], [0, 0, 1], [1, 0, 0], [0, 1, 0], [0, 0, 0], [0, 0, -1], [0, 0, 1e-30], [0, 0, -1e300]]
    rays = rt.make_rays(o, d)
    got, _ = sc.closest_hit(rays)
    want = osc.closest_hit(rays, mode=0)
    show_hits("edge", rays, got, want)
    print(got, want)
    b = rt.Builder(1)
    hs = b.finish(b.list([]), width=8, spp=1, background=b.solid(0.25, 0.5, 1.0))
    sc = rt.Scene(hs)
    img, st = sc.render(seed=1)
    print("empty world image", img[0, 0], st.segments, st.paths)


def golden_random_graph():
    fx = np.load(os.path.join(ROOT, "tests", "golden", "random_graph.npz"))
    hs = random_graph_scene(rt, 11, n_prims=72, with_media=True, width=32, spp=4, depth=8)
    sc = rt.Scene(hs)
    img, st = sc.render(seed=int(fx["render_seed"]))
    ref = fx["image"]
    diff = np.abs(img - ref)
    bad = (diff > 1e-6 * (1 + np.abs(ref))).any(axis=2)
    print("[random_graph] errors", st.errors, "vs", int(fx["errors"]), "bad pixels", int(bad.sum()), "of", bad.size, "max diff", diff.max(),
          "means", img.mean(), ref.mean())
    ys, xs = np.nonzero(bad)
    for y, x in list(zip(ys, xs))[:8]:
        print("   px", x, y, img[y, x], ref[y, x])


if __name__ == "__main__":
    ellipsoids()
    edge()
    golden_random_graph()

"""Exploratory GPU run: closest-hit parity and same-seed image parity against the oracle."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import orc  # noqa: E402

rt = orc.rt


def rays_for(hs, n, seed):
    rng = np.random.default_rng(seed)
    cam = hs.camera
    W, H = cam.image_width, cam.image_height
    px = np.stack([rng.integers(0, W, n // 2), rng.integers(0, H, n // 2)], axis=1)
    prim = orc.camera_rays(cam, 5, px, 0)
    # incoherent: origins at primary hit-ish points (random points in the scene box), random directions
    o = rng.uniform(-100, 600, (n - n // 2, 3))
    d = rng.normal(size=(n - n // 2, 3))
    sec = rt.make_rays(o, d, rng.uniform(0, 1, n - n // 2))
    return np.concatenate([prim, sec])


def main():
    out = {}
    for name, params, spp_note in [("book2_final", [96, 16, 40], ""), ("cornell_glass", [96, 16, 50], ""),
                                   ("book1_final", [160, 9, 50], "")]:
        hs = rt.named_scene(name, seed=7, params=params)
        t0 = time.time()
        sc = rt.Scene(hs)
        info = sc.info()
        print(f"[{name}] scene_create {time.time() - t0:.3f}s prims={info.n_prims} nodes={info.n_nodes} depth={info.bvh_depth} "
              f"media={info.n_media} lights={info.n_lights}")
        osc = orc.OracleScene(hs)
        assert np.array_equal(sc.ranks(), osc.ranks()), "tie ranks differ"
        rays = rays_for(hs, 20000, 3)
        g, st = sc.closest_hit(rays, flags=rt.RT_OPT_COUNT)
        o0 = osc.closest_hit(rays, mode=0)
        o1 = osc.closest_hit(rays, mode=1)
        for nm, o in (("ref", o0), ("brute", o1)):
            ids = int((g["prim_id"] != o["prim_id"]).sum())
            inst = int((g["inst_id"] != o["inst_id"]).sum())
            hit = g["prim_id"] != rt.RT_NONE
            both = hit & (o["prim_id"] == g["prim_id"])
            rel = np.abs(g["t"][both] - o["t"][both]) / np.abs(o["t"][both])
            exact = int((g["t"][both] == o["t"][both]).sum())
            print(f"  closest-hit vs {nm}: id mismatches {ids}/{len(rays)}, inst mismatches {inst}, hits {int(hit.sum())}, "
                  f"max rel t err {rel.max() if len(rel) else 0:.3e}, bit-equal t {exact}/{int(both.sum())}, "
                  f"uv max err {np.abs(g['u'][both]-o['u'][both]).max():.2e}")
        print(f"  nodes/ray {st.node_visits / len(rays):.1f} prims/ray {st.prim_tests / len(rays):.1f} ms {st.ms_total:.3f}")
        t0 = time.time()
        img, rst = sc.render(seed=11)
        tg = time.time() - t0
        t0 = time.time()
        ref, ost = osc.render(seed=11)
        to = time.time() - t0
        diff = np.abs(img - ref)
        tol = 1e-6 * (1 + np.abs(ref))
        bad = (diff > tol).any(axis=2)
        print(f"  render: gpu {tg:.3f}s ({rst.ms_total:.1f} ms dev, {rst.iterations} iters, {rst.kernel_launches} launches) oracle {to:.2f}s; "
              f"paths {rst.paths} vs {ost.paths}; segments {rst.segments} vs {ost.segments}; errors {rst.errors} vs {ost.errors}")
        print(f"  image: max abs diff {diff.max():.3e}, mean abs diff {diff.mean():.3e}, pixels off (>1e-6 rel) {int(bad.sum())}/{bad.size}, "
              f"mean gpu {img.mean():.6f} mean ref {ref.mean():.6f}")
        out[name] = (img, ref)
        np.save(os.path.join(ROOT, "gpurun_out", f"probe_{name}_gpu.npy"), img)
        np.save(os.path.join(ROOT, "gpurun_out", f"probe_{name}_ref.npy"), ref)
        sc.close()
    # throughput sample at full size
    hs = rt.named_scene("book2_final", seed=7, params=[800, 36, 40])
    sc = rt.Scene(hs)
    for k in range(2):
        img, rst = sc.render(seed=11, accum_type=rt.RT_ACCUM_F32, flags=rt.RT_OPT_STAGE_TIMES if k else 0)
        print(f"book2 800x800x36spp: {rst.ms_total:.1f} ms, {rst.paths / rst.ms_total * 1e3 / 1e6:.1f} Mpaths/s, "
              f"{rst.segments / rst.ms_total * 1e3 / 1e6:.1f} Msegments/s, seg/path {rst.segments / rst.paths:.2f}, iters {rst.iterations}, "
              f"stages gen/ext/shade/other {rst.ms_raygen:.1f}/{rst.ms_extend:.1f}/{rst.ms_shade:.1f}/{rst.ms_other:.1f}")


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    main()

"""Generate the committed golden fixtures with the CPU oracle.

The reference (a Rust crate) cannot be built or imported in this environment, so these vectors
come from oracle/oracle.cpp — the restatement that tests/test_oracle_kat.py pins against the
reference's own known-answer tests.  Each fixture stores INPUTS and OUTPUTS, so the GPU tests
that read them do not need the oracle at run time:

    <scene>.npz: rays (rt_ray records), hits (rt_hit records, reference container semantics),
                 image (H,W,3 float64 mean radiance at `seed`), paths, errors

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc  # noqa: E402
from scenes_util import final_reduced_scene, random_graph_scene  # noqa: E402

rt = orc.rt
RENDER_SEED = 2025
N_RAYS = 2048

SCENES = {
    "book2_final": lambda: rt.named_scene("book2_final", seed=7, params=[32, 4, 12]),
    "cornell_glass": lambda: rt.named_scene("cornell_glass", seed=7, params=[32, 4, 12]),
    "book1_final": lambda: rt.named_scene("book1_final", seed=7, params=[48, 4, 12]),
    "random_graph": lambda: random_graph_scene(rt, 11, n_prims=72, with_media=True, width=32, spp=4, depth=8),
    "final_reduced": lambda: final_reduced_scene(rt, width=96, spp=4, depth=12),
}


def fixture_rays(hs, seed):
    """Camera rays + rays leaving first-hit points (what a path tracer actually traces)."""
    rng = np.random.default_rng(seed)
    cam = hs.camera
    n = N_RAYS // 2
    px = np.stack([rng.integers(0, cam.image_width, n), rng.integers(0, cam.image_height, n)], axis=1)
    prim = orc.camera_rays(cam, 5, px, 0)
    osc = orc.OracleScene(hs)
    h = osc.closest_hit(prim)
    hit = h["prim_id"] != rt.RT_NONE
    p = prim["origin"] + h["t"][:, None].clip(0, 1e6) * prim["direction"]
    p[~hit] = prim["origin"][~hit]
    d = rng.normal(size=(n, 3))
    sec = rt.make_rays(p, d, prim["time"])
    return np.concatenate([prim, sec])


def main():
    only = sys.argv[1:]
    for name, make in SCENES.items():
        if only and name not in only:
            continue
        hs = make()
        osc = orc.OracleScene(hs)
        rays = fixture_rays(hs, 99)
        hits = osc.closest_hit(rays, mode=0)
        brute = osc.closest_hit(rays, mode=1)
        assert np.array_equal(hits["prim_id"], brute["prim_id"]), name
        img, st = osc.render(seed=RENDER_SEED)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), rays=rays, hits=hits, image=img,
                            paths=np.uint64(st.paths), errors=np.uint64(st.errors), render_seed=np.uint64(RENDER_SEED))
        print(name, "hits", int((hits["prim_id"] != rt.RT_NONE).sum()), "/", len(rays), "image mean", img.mean(), "errors", st.errors)


if __name__ == "__main__":
    main()

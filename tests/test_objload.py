"""The Python OBJ/MTL loader must build what shapes/obj.rs builds (structure), and the resulting
config-4-like scene must render identically on the GPU and in the oracle."""
import importlib.util
import os

import numpy as np
import pytest

from scenes_util import compare_hits, random_rays, synthetic_obj_scene, write_synthetic_assets

KIND = dict(sphere=1, quad=2, tri=3, list=4, bvh=5, transform=6, medium=7)


def _objload(rt):
    spec = importlib.util.spec_from_file_location("objload", os.path.join(os.path.dirname(rt.__file__), "objload.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_loader_structure(rt, tmp_path):
    root = write_synthetic_assets(str(tmp_path), n=8)
    objload = _objload(rt)
    models, libs, pos, tex, nrm = objload.parse_obj(os.path.join(root, "Synth", "terrain.obj"))
    assert libs == ["terrain.mtl"] and [m["material"] for m in models] == ["ground", "glow", "leaves"]  # one model per usemtl run
    assert sum(len(m["faces"]) for m in models) == 2 * 8 * 8 and pos.shape == (81, 3) and tex.shape == (81, 2)
    mats = objload.parse_mtl(os.path.join(root, "Synth", "terrain.mtl"))
    assert mats[0]["normal_texture"].endswith("normal.png") and mats[1]["unknown_param"]["map_Ke"] == "emit.png"
    assert mats[2]["dissolve"] == 0.8 and mats[2]["dissolve_texture"] == "alpha.png"
    quads, _, _, _, _ = objload.parse_obj(os.path.join(root, "Synth", "ball.obj"))
    assert len(quads[0]["faces"]) == 2 * 8 * 16  # quads are fan-triangulated like tobj's `triangulate`
    b = rt.Builder(1)
    wf = objload.Wavefont(b, root)
    assert wf.new("missing.obj", "Synth", False) is None  # Wavefont::new -> None
    top = wf.new("terrain.obj", "Synth", False)
    hs = b.finish(b.list([top]))
    o = hs.objects()
    assert (o["kind"] == KIND["bvh"]).sum() == 3 and (o["kind"] == KIND["tri"]).sum() == 128
    d = hs.desc.contents
    assert d.n_remaps == 128 and d.n_images == 4  # albedo, normal (raw), emit, alpha
    # the reference zips models with the materials' normal maps (obj.rs:129): 2 materials -> the third model is dropped
    with open(os.path.join(root, "Synth", "terrain.mtl")) as f:
        text = f.read().split("newmtl leaves")[0]
    with open(os.path.join(root, "Synth", "terrain.mtl"), "w") as f:
        f.write(text)
    b2 = rt.Builder(1)
    hs2 = b2.finish(b2.list([objload.Wavefont(b2, root).new("terrain.obj", "Synth", False)]))
    assert (hs2.objects()["kind"] == KIND["bvh"]).sum() == 2


def test_synthetic_obj_scene_oracle(rt, orc, tmp_path):
    hs = synthetic_obj_scene(rt, write_synthetic_assets(str(tmp_path), n=12), width=40, spp=9)
    img, st = orc.OracleScene(hs).render(seed=2)
    assert st.errors == 0 and np.isfinite(img).all() and 0.1 < img.mean() < 3.0


@pytest.mark.gpu
def test_synthetic_obj_scene_gpu_parity(gpu, rt, orc, tmp_path):
    hs = synthetic_obj_scene(rt, write_synthetic_assets(str(tmp_path), n=24), width=64, spp=16)
    sc, osc = rt.Scene(hs), orc.OracleScene(hs)
    assert sc.info().n_media == 1
    rng = np.random.default_rng(2)
    o, d, t = random_rays(rng, 30000, extent=4.5)
    rays = rt.make_rays(o, d, t)
    got, _ = sc.closest_hit(rays)
    compare_hits(rt, got, osc.closest_hit(rays, mode=0))
    img, st = sc.render(seed=8)
    ref, ost = osc.render(seed=8)
    assert st.paths == ost.paths and int(st.errors) <= int(ost.errors)
    diff = np.abs(img - ref)
    bad = (diff > 1e-6 * (1 + np.abs(ref))).any(axis=2)
    assert bad.mean() <= 5e-3, f"{int(bad.sum())} pixels differ"

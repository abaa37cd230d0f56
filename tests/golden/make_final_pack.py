"""Writes tests/golden/final_assets_pack.npz from the reference's own asset directory.

Run in the build container (where /root/reference exists):  python tests/golden/make_final_pack.py
The GPU box has no /root/reference, so config 4 (BASELINE.json: "assets/Final triangle-mesh scene") is rebuilt
there from this pack: the tables `parse_obj` / `parse_mtl` produce for the 13 OBJ files the reference ships
(src/main.rs:208-223 names 15; 初音未来.obj and 卒.obj are listed in .MISSING_LARGE_BLOBS), and the decoded RGBA8
texels of the images their MTL files name.  Geometry and materials are stored verbatim (binary64); images are
box-reduced to <= 512 texels on the longer side to keep the fixture small - texture resolution does not change
which code runs (the reference samples them nearest-texel, texture.rs:111-119).
"""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ASSETS = "/root/reference/assets"
# order of obj_scene(), src/main.rs:208-223 (+ the fog mesh, :224)
OBJ_FILES = ["初音未来.obj", "玻璃球.obj", "外框.obj", "声匣.obj", "镜子门.obj", "镜子.obj", "环.obj", "传送门框.obj", "水下.obj", "水面.obj",
             "文字.obj", "mc.obj", "伞.obj", "卒.obj", "雾.obj"]

if __name__ == "__main__":
    spec = importlib.util.spec_from_file_location("objload", os.path.join(ROOT, "raytracer-2025_b200", "objload.py"))
    objload = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(objload)
    out = os.path.join(ROOT, "tests", "golden", "final_assets_pack.npz")
    meta = objload.write_pack(out, ASSETS, [("Final", f) for f in OBJ_FILES], max_image_side=512)
    print("objs:", len(meta["objs"]), "mtls:", len(meta["mtls"]), "images:", len(meta["images"]), "bytes:", os.path.getsize(out))
    for k in OBJ_FILES:
        if "Final/" + k not in meta["objs"]:
            print("  missing:", k)
    sys.exit(0)

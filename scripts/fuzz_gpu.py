"""GPU-side robustness: single-word corruptions of every table of a valid description (numbers as well as
indices), then create + closest-hit + a tiny render.  Every call must return a status; a CUDA fault
(illegal address, launch failure) would show up as RT_ERR_CUDA on this or the next call and is reported
with the corruption that caused it.   python scripts/fuzz_gpu.py [seed] [iterations]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc
from scenes_util import disney_scene, obj_mesh_scene, random_graph_scene, random_rays
rt = orc.rt
L = rt.product_lib()
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
rng = np.random.default_rng(seed)
scenes = [random_graph_scene(rt, 3, n_prims=40, with_media=True, width=8, spp=1, depth=4), disney_scene(rt, True, width=8, spp=1),
          obj_mesh_scene(rt, width=8, spp=1), rt.named_scene("book2_final", seed=7, params=[8, 1, 6])]
sizes = {"objects": 72, "children": 4, "spheres": 64, "planars": 144, "materials": 176, "textures": 80, "media": 16, "transforms": 80,
         "images": 24, "remaps": 200, "perlins": 9216}
o, d, t = random_rays(np.random.default_rng(1), 256)
rays = rt.make_rays(o, d, t)
hits = np.zeros(256, dtype=rt.rt_hit_dtype)
counts = {}
for it in range(iters):
    hs = scenes[it % len(scenes)]
    good = hs.desc.contents
    name = list(sizes)[rng.integers(len(sizes))]
    n = getattr(good, "n_" + name)
    if n == 0:
        continue
    raw = np.ctypeslib.as_array((C.c_uint32 * (n * sizes[name] // 4)).from_address(getattr(good, name))).copy()
    k = int(rng.integers(raw.size))
    val = int(rng.choice([0, 1, 2, 3, 7, 0xFFFFFFFF, 0x7FF80000, 0x7FF00000, 0xFFF00000, 0x7FEFFFFF, 0x00000001, 0x80000000, 1000, 1 << 20,
                          int(rng.integers(1 << 32))]))
    raw[k] = val
    bad = rt.rt_scene_desc.from_buffer_copy(good)
    setattr(bad, name, raw.ctypes.data)
    opts = rt.rt_build_opts(C.sizeof(rt.rt_build_opts), 0, -1, 0)
    h = C.c_void_p()
    rc = L.rt_scene_create(C.byref(bad), C.byref(opts), C.byref(h))
    tag = f"create={rc}"
    if rc == 0:
        rc2 = L.rt_closest_hit(h, rays.ctypes.data, 256, 1e-8, float("inf"), 0, hits.ctypes.data, None)
        cam = hs.camera
        ro = rt.rt_render_opts()
        ro.struct_size = C.sizeof(rt.rt_render_opts)
        ro.seed = 1
        img = np.zeros((cam.image_height, cam.image_width, 3), dtype=np.float64)
        ro.accum_type = rt.RT_ACCUM_F64
        rc3 = L.rt_render(h, C.byref(cam), C.byref(ro), img.ctypes.data, None)
        L.rt_scene_destroy(h)
        tag = f"create=0 hit={rc2} render={rc3}"
        if rc2 == -4 or rc3 == -4:  # RT_ERR_CUDA
            print(f"CUDA FAULT at iteration {it}: {name}[{k}] (byte {k * 4 % sizes[name]} of record {k * 4 // sizes[name]}) = {val:#x}: {L.rt_last_error().decode()}")
            sys.exit(1)
    counts[tag] = counts.get(tag, 0) + 1
print(counts)

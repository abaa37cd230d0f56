"""The C-ABI library: loads without a GPU, exports every symbol include/rt2025.h declares,
validates descriptions, and refuses to compute without a device (there is no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "rt2025.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(rt):
    L = rt.product_lib()
    names = declared_symbols()
    assert set(names) == set(rt.ABI_SYMBOLS)
    for n in names:
        assert hasattr(L, n), f"librt2025.so does not export {n}"
    assert L.rt_abi_version() == 1


def test_struct_layouts_match_the_header(rt, tmp_path):
    # compile a C program against include/rt2025.h and compare sizeof() with the ctypes mirrors
    src = tmp_path / "sz.c"
    src.write_text("""
#include <stdio.h>
#include "rt2025.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(rt_camera), sizeof(rt_render_opts), sizeof(rt_build_opts),
         sizeof(rt_stats), sizeof(rt_scene_info), sizeof(rt_scene_desc), sizeof(rt_ray), sizeof(rt_hit), sizeof(rt_object));
  return 0;
}
""")
    exe = tmp_path / "sz"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(rt.rt_camera), C.sizeof(rt.rt_render_opts), C.sizeof(rt.rt_build_opts), C.sizeof(rt.rt_stats),
            C.sizeof(rt.rt_scene_info), C.sizeof(rt.rt_scene_desc), rt.rt_ray_dtype.itemsize, rt.rt_hit_dtype.itemsize,
            rt.rt_object_dtype.itemsize]
    assert got == want


def _create(rt, desc, flags=0):
    L = rt.product_lib()
    opts = rt.rt_build_opts(C.sizeof(rt.rt_build_opts), flags, -1, 0)
    h = C.c_void_p()
    rc = L.rt_scene_create(C.byref(desc), C.byref(opts), C.byref(h))
    return rc, h, L.rt_last_error().decode()


def test_description_is_validated_before_any_device_work(rt):
    hs = rt.named_scene("cornell_glass", seed=1, params=[32, 4, 4])
    good = hs.desc.contents
    bad = rt.rt_scene_desc.from_buffer_copy(good)
    bad.version = 99
    rc, h, msg = _create(rt, bad)
    assert rc == -6 and "version" in msg
    bad = rt.rt_scene_desc.from_buffer_copy(good)
    bad.world_root = good.n_objects + 5
    rc, h, msg = _create(rt, bad)
    assert rc == -1 and "world_root" in msg
    bad = rt.rt_scene_desc.from_buffer_copy(good)
    bad.spheres = None
    rc, h, msg = _create(rt, bad)
    assert rc == -1 and "spheres" in msg
    # corrupt a child index
    objs = hs.objects().copy()
    kids = hs.children().copy()
    kids[0] = 10 ** 6
    bad = rt.rt_scene_desc.from_buffer_copy(good)
    bad.children = kids.ctypes.data
    rc, h, msg = _create(rt, bad)
    assert rc == -1
    del objs


def test_corrupted_descriptions_never_crash(rt):
    """Single-word corruptions of every index-bearing table: rt_scene_create must answer with a status (invalid,
    unsupported, or - when the damage is only numeric - the no-device / success path), never fault."""
    from scenes_util import disney_scene, obj_mesh_scene, random_graph_scene
    rng = np.random.default_rng(7)
    scenes = [random_graph_scene(rt, 3, n_prims=40, with_media=True, width=8, spp=1, depth=2), disney_scene(rt, True, width=8, spp=1),
              obj_mesh_scene(rt, width=8, spp=1)]
    sizes = {"objects": 72, "children": 4, "materials": 176, "textures": 80, "media": 16, "transforms": 80, "images": 24, "remaps": 200}
    seen = set()
    for it in range(3000):
        good = scenes[it % len(scenes)].desc.contents
        name = list(sizes)[rng.integers(len(sizes))]
        n = getattr(good, "n_" + name)
        if n == 0:
            continue
        raw = np.ctypeslib.as_array((C.c_uint32 * (n * sizes[name] // 4)).from_address(getattr(good, name))).copy()
        raw[rng.integers(raw.size)] = rng.choice([0, 1, 2, 3, 7, 0xFFFFFFFF, 0xFFFFFFFE, 1000, 1 << 20, int(rng.integers(1 << 32))])
        bad = rt.rt_scene_desc.from_buffer_copy(good)
        setattr(bad, name, raw.ctypes.data)
        rc, h, msg = _create(rt, bad)
        seen.add(rc)
        if rc == 0:
            rt.product_lib().rt_scene_destroy(h)
    assert seen <= {0, -1, -2, -3} and -1 in seen
    # a ConstantMedium whose phase function is not an Isotropic (volume.rs:23-35 cannot build one)
    b = rt.Builder(1)
    med = b.medium(b.sphere([0, 0, 0], 1.0, b.empty()), 0.5, b.solid(1, 1, 1))
    hs = b.finish(b.list([med]))
    mats = np.ctypeslib.as_array((C.c_uint32 * (hs.desc.contents.n_materials * 44)).from_address(hs.desc.contents.materials)).copy()
    objs = hs.objects()
    mats[int(objs["material"][objs["kind"] == 7][0]) * 44] = 1  # RT_OBJ_MEDIUM -> its material becomes RT_MAT_LAMBERTIAN
    bad = rt.rt_scene_desc.from_buffer_copy(hs.desc.contents)
    bad.materials = mats.ctypes.data
    rc, h, msg = _create(rt, bad)
    assert rc == -1 and "Isotropic" in msg


def test_unsupported_constructs_are_reported(rt):
    # more than 4 nested Transforms exceed the device's chain table -> RT_ERR_UNSUPPORTED
    b = rt.Builder(1)
    t = b.sphere([0, 0, 0], 1.0, b.empty())
    for k in range(5):
        t = b.transform(t, offset=[1, 0, 0])
    hs = b.finish(b.list([t]))
    rc, h, msg = _create(rt, hs.desc.contents)
    assert rc == -2 and "nested" in msg
    # a zero scale cannot be inverted
    b = rt.Builder(1)
    t = b.transform(b.sphere([0, 0, 0], 1.0, b.empty()), scale=[1, 0, 1])
    hs = b.finish(b.list([t]))
    rc, h, msg = _create(rt, hs.desc.contents)
    assert rc == -2 and "singular" in msg
    # lights containing a BVH: pdf_value/random are unimplemented!() in the reference (hit.rs:51-59)
    b = rt.Builder(1)
    s = b.sphere([0, 0, 0], 1.0, b.empty())
    hs = b.finish(b.list([s]), b.list([b.bvh([b.sphere([0, 3, 0], 1.0, b.empty())])]))
    rc, h, msg = _create(rt, hs.desc.contents)
    assert rc == -2 and "unimplemented" in msg


def test_no_device_means_no_result(rt):
    L = rt.product_lib()
    if L.rt_device_count() > 0:
        pytest.skip("a GPU is present; the no-device path is exercised on the CPU box")
    hs = rt.named_scene("cornell_glass", seed=1, params=[32, 4, 4])
    rc, h, msg = _create(rt, hs.desc.contents)
    assert rc == -3 and "no CPU path" in msg and not h.value
    with pytest.raises(rt.RtError):
        rt.Scene(hs)
    with pytest.raises(rt.RtError):
        rt.tonemap(np.zeros((2, 2, 3)))


def test_product_does_not_touch_the_oracle():
    # the product tree must not reference oracle/ in any way
    pkg = os.path.join(ROOT, "raytracer-2025_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in text and "oracle/" not in text and "import orc" not in text, f

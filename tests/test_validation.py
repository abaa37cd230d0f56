"""rt_scene_create validates the whole description on the host BEFORE it touches a device, so malformed input is
checked here without a GPU: a bad description must come back as RT_ERR_INVALID / RT_ERR_UNSUPPORTED (never a crash), a
good one reaches the device check (RT_ERR_NO_DEVICE on this box, RT_OK on a B200).  Covers the ADVICE findings of
round 1: images no texture references, objects with two parents, unbounded nesting."""
import ctypes as C

import numpy as np
import pytest

RT_OK, RT_ERR_INVALID, RT_ERR_UNSUPPORTED, RT_ERR_NO_DEVICE = 0, -1, -2, -3
rt_image_dtype = np.dtype([("width", "<u4"), ("height", "<u4"), ("flags", "<u4"), ("reserved", "<u4"), ("texel_offset", "<u8")])


def create(rt, desc):
    L = rt.product_lib()
    h = C.c_void_p()
    opts = rt.rt_build_opts(C.sizeof(rt.rt_build_opts), 0, -1, 0)
    rc = L.rt_scene_create(C.byref(desc), C.byref(opts), C.byref(h))
    if rc == RT_OK:
        L.rt_scene_destroy(h)
    return rc, L.rt_last_error().decode()


def copy_desc(rt, hs):
    d = rt.rt_scene_desc()
    C.memmove(C.byref(d), hs.desc, C.sizeof(d))
    return d


def test_valid_description_reaches_the_device_check(rt):
    hs = rt.named_scene("cornell_glass", seed=7, params=[16, 1, 4])
    rc, _ = create(rt, copy_desc(rt, hs))
    assert rc in (RT_OK, RT_ERR_NO_DEVICE)


@pytest.mark.parametrize("width,height,offset,n_texels", [
    (4, 4, 1 << 40, 64),          # offset far beyond the pool
    (0x10000, 0x10000, 0, 64),    # width * height * 4 wraps 32 bits
    (0xFFFFFFFF, 0xFFFFFFFF, 0, 64),
    (0, 4, 0, 64), (4, 0, 0, 64),  # empty image: min(x, width - 1) would wrap on the device
    (4, 4, 2, 64),                # not a whole texel
    (4, 4, 4, 64),                # 4 + 64 > 64
])
def test_unreferenced_bad_image_is_rejected(rt, width, height, offset, n_texels):
    """copy_tables() converts EVERY image, referenced or not: each one is bounded before anything is copied."""
    hs = rt.named_scene("cornell_glass", seed=7, params=[16, 1, 4])
    d = copy_desc(rt, hs)
    assert d.n_images == 0
    images = np.zeros(1, dtype=rt_image_dtype)
    images[0] = (width, height, 0, 0, offset)
    texels = np.zeros(n_texels, dtype=np.float32)
    d.n_images, d.images, d.n_texels, d.texels = 1, images.ctypes.data, n_texels, texels.ctypes.data
    rc, msg = create(rt, d)
    assert rc == RT_ERR_INVALID and "image" in msg


def test_good_unreferenced_image_is_accepted(rt):
    hs = rt.named_scene("cornell_glass", seed=7, params=[16, 1, 4])
    d = copy_desc(rt, hs)
    images = np.zeros(1, dtype=rt_image_dtype)
    images[0] = (4, 4, 0, 0, 0)
    texels = np.zeros(64, dtype=np.float32)
    d.n_images, d.images, d.n_texels, d.texels = 1, images.ctypes.data, 64, texels.ctypes.data
    assert create(rt, d)[0] in (RT_OK, RT_ERR_NO_DEVICE)


def test_shared_child_is_rejected(rt):
    """Box<dyn Hittable> ownership cannot express a DAG; 40 lists that each name the previous one twice would expand to 2^40 leaves."""
    b = rt.Builder(1)
    s = b.sphere([0, 0, 0], 1.0, b.empty())
    hs = b.finish(b.list([b.list([s])]))
    d = copy_desc(rt, hs)
    objs = hs.objects().copy()
    kids = hs.children().copy()
    outer = objs[d.world_root]
    assert outer["child_count"] == 1
    # the world list names its child twice
    kids2 = np.concatenate([kids, kids[outer["first_child"]:outer["first_child"] + 1].repeat(2)]).astype(np.uint32)
    objs[d.world_root]["first_child"], objs[d.world_root]["child_count"] = len(kids), 2
    d.objects, d.children, d.n_children = objs.ctypes.data, kids2.ctypes.data, len(kids2)
    rc, msg = create(rt, d)
    assert rc == RT_ERR_INVALID and "two parents" in msg


def test_nesting_depth_is_bounded(rt):
    b = rt.Builder(1)
    node = b.sphere([0, 0, 0], 1.0, b.empty())
    for _ in range(300):
        node = b.list([node])
    hs = b.finish(node)
    rc, msg = create(rt, copy_desc(rt, hs))
    assert rc == RT_ERR_UNSUPPORTED and "nested" in msg
    b = rt.Builder(1)
    node = b.sphere([0, 0, 0], 1.0, b.empty())
    for _ in range(200):
        node = b.list([node])
    hs = b.finish(node)  # (owns the arrays the copied description points into)
    assert create(rt, copy_desc(rt, hs))[0] in (RT_OK, RT_ERR_NO_DEVICE)

"""The C++ host mirror must build exactly what the reference constructors build
(bounding boxes, derived plane constants, camera frame, graph shape)."""
import math

import numpy as np
import pytest

KIND = dict(sphere=1, quad=2, tri=3, list=4, bvh=5, transform=6, medium=7)


def test_final_scene_graph_shape(rt):
    # src/main.rs:384-539 and SURVEY.md appendix C
    hs = rt.named_scene("book2_final", seed=7, params=[800, 1000, 40])
    d = hs.desc.contents
    objs = hs.objects()
    kinds = objs["kind"]
    assert (kinds == KIND["quad"]).sum() == 2400 + 1 + 1  # 400 boxes x 6, light, lights-list quad
    assert (kinds == KIND["sphere"]).sum() == 1000 + 6 + 2  # cloud, 6 loose (incl. glass shell), 2 medium boundaries
    assert (kinds == KIND["medium"]).sum() == 2 and d.n_media == 2
    assert (kinds == KIND["transform"]).sum() == 1
    assert (kinds == KIND["bvh"]).sum() == 2
    world = objs[d.world_root]
    assert world["kind"] == KIND["list"] and world["child_count"] == 11
    lights = objs[d.lights_root]
    assert lights["kind"] == KIND["list"] and lights["child_count"] == 1
    # children always precede their parent (the flattener emits bottom-up)
    ch = hs.children()
    for i, o in enumerate(objs):
        kids = ch[o["first_child"]:o["first_child"] + o["child_count"]]
        assert (kids < i).all()
    cam = hs.camera
    assert (cam.image_width, cam.image_height, cam.sqrt_spp, cam.max_depth) == (800, 800, 31, 40)


def test_effective_spp_and_image_height(rt):
    # src/camera.rs:205-214: sqrt_spp = floor(sqrt(spp)); height = trunc(width / aspect) >= 1
    for spp, want in [(10, 3), (1000, 31), (3000, 54), (5000, 70), (100, 10), (1, 1)]:
        b = rt.Builder(1)
        s = b.sphere([0, 0, 0], 1.0, b.empty())
        hs = b.finish(b.list([s]), spp=spp, width=10)
        assert hs.camera.sqrt_spp == want
        assert hs.camera.pixel_sample_scale == 1.0 / (want * want)
    hs = rt.named_scene("book1_final", seed=1, params=[1200, 10, 50])
    assert (hs.camera.image_width, hs.camera.image_height) == (1200, 675)
    b = rt.Builder(1)
    s = b.sphere([0, 0, 0], 1.0, b.empty())
    hs = b.finish(b.list([s]), width=3, aspect=100.0)
    assert hs.camera.image_height == 1


def test_camera_frame(rt):
    # src/camera.rs:216-245 against an independent numpy evaluation
    b = rt.Builder(1)
    s = b.sphere([0, 0, 0], 1.0, b.empty())
    hs = b.finish(b.list([s]), width=200, aspect=2.0, vfov=35.0, look_from=(3, 2, 7), look_at=(0.5, 0.2, -1), vup=(0, 1, 0),
                  defocus_angle=1.5, focus_dist=6.0)
    cam = hs.camera
    lf, la, up = np.array([3, 2, 7.0]), np.array([0.5, 0.2, -1.0]), np.array([0, 1.0, 0])
    w = (lf - la) / np.linalg.norm(lf - la)
    u = np.cross(up, w)
    u /= np.linalg.norm(u)
    v = np.cross(w, u)
    h = math.tan(math.radians(35.0) / 2)
    vh = 2 * h * 6.0
    vw = vh * (200 / 100)
    du, dv = vw * u / 200, vh * -v / 100
    p00 = lf - 6.0 * w - vw * u / 2 - vh * -v / 2 + 0.5 * (du + dv)
    assert np.allclose(list(cam.pixel_delta_u), du, rtol=1e-13)
    assert np.allclose(list(cam.pixel_delta_v), dv, rtol=1e-13)
    assert np.allclose(list(cam.pixel00_loc), p00, rtol=1e-13)
    r = 6.0 * math.tan(math.radians(0.75))
    assert np.allclose(list(cam.defocus_disk_u), u * r, rtol=1e-13)
    assert np.allclose(list(cam.defocus_disk_v), v * r, rtol=1e-13)


def test_hittables_default_bbox_contains_origin(rt):
    # src/hits.rs:10-14: #[derive(Default)] starts from the degenerate box at the origin, so a list
    # built with default()+add always covers (0,0,0); Hittables::new starts from the object's box
    b = rt.Builder(1)
    m = b.empty()
    s = b.sphere([10, 10, 10], 1.0, m)
    hs = b.finish(b.list([s]))
    o = hs.objects()
    assert list(o[-1]["bbox"]) == [0, 11, 0, 11, 0, 11]
    b = rt.Builder(1)
    s = b.sphere([10, 10, 10], 1.0, b.empty())
    hs = b.finish(b.list([s], use_new=True))
    assert list(hs.objects()[-1]["bbox"]) == [9, 11, 9, 11, 9, 11]


def test_quad_and_box_bboxes(rt):
    b = rt.Builder(1)
    m = b.empty()
    q = b.quad([1, 2, 3], [2, 0, 0], [0, 3, 0], m)  # flat in z -> padded (aabb.rs:43-51)
    bx = b.box([1, 1, 1], [2, 3, 4], m)
    hs = b.finish(b.list([q, bx]))
    o = hs.objects()
    quad = o[0]
    assert list(quad["bbox"]) == [1, 3, 2, 5, 3 - 0.00005, 3 + 0.00005]
    box_list = [x for x in o if x["kind"] == KIND["list"]][0]
    # build_box uses Hittables::default(), so the origin is included (quad.rs:128-129)
    assert box_list["child_count"] == 6
    assert box_list["bbox"][0] == 0 and box_list["bbox"][1] == pytest.approx(2 + 0.00005)


def test_degenerate_triangle_is_none(rt):
    # src/shapes/triangle.rs:29-31
    b = rt.Builder(1)
    m = b.empty()
    assert b.triangle([0, 0, 0], [1, 0, 0], [2, 0, 0], m) == rt.RT_NONE
    assert b.triangle([0, 0, 0], [1, 0, 0], [0, 1, 0], m) != rt.RT_NONE


def test_empty_bvh_is_an_error(rt):
    # src/bvh.rs:26: panic!("BVH node must contain at least one object")
    b = rt.Builder(1)
    with pytest.raises(rt.RtError):
        b.bvh([])


def test_moving_sphere_bbox_is_union(rt):
    b = rt.Builder(1)
    s = b.sphere_moving([0, 0, 0], [3, 0, 0], 1.0, b.empty())
    hs = b.finish(b.list([s], use_new=True))
    assert list(hs.objects()[0]["bbox"]) == [-1, 4, -1, 1, -1, 1]


def test_transform_bbox(rt):
    # src/shapes.rs:49-72: box of the 8 transformed corners
    b = rt.Builder(1)
    s = b.sphere([1, 0, 0], 1.0, b.empty())
    t = b.transform(s, offset=[0, 5, 0], quat=b.quat_axis_angle([0, 0, 1], 90.0), scale=[2, 2, 2])
    hs = b.finish(b.list([t], use_new=True))
    bb = hs.objects()[1]["bbox"]
    # sphere box [0,2]x[-1,1]x[-1,1] scaled by 2, rotated 90 deg about z (x->y), then +5 in y
    assert np.allclose(bb, [-2, 2, 5, 9, -2, 2], atol=1e-12)


def test_scenes_are_reproducible_from_the_seed(rt):
    a = rt.named_scene("book2_final", seed=3, params=[64, 4, 4])
    b = rt.named_scene("book2_final", seed=3, params=[64, 4, 4])
    c = rt.named_scene("book2_final", seed=4, params=[64, 4, 4])
    assert np.array_equal(a.objects()["bbox"], b.objects()["bbox"])
    assert not np.array_equal(a.objects()["bbox"], c.objects()["bbox"])

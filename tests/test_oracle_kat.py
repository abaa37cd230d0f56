"""Pin the oracle against every known-answer vector the reference's own unit tests hold for the
hot path, plus the published Philox4x32-10 vectors for the shared RNG."""
import ctypes as C
import math

import numpy as np


def D(*v):
    return (C.c_double * len(v))(*[float(x) for x in v])


def test_aabb_hit_inside_outside(orc):
    # reference src/aabb.rs:213-227
    L = orc.lib()
    assert L.orc_kat_aabb_hit(D(0, 0, 0), D(1, 1, 1), D(0.5, 0.5, -1.0), D(0, 0, 1), 0.0, 100.0) == 1
    assert L.orc_kat_aabb_hit(D(0, 0, 0), D(1, 1, 1), D(2, 2, 2), D(1, 0, 0), 0.0, 100.0) == 0


def test_aabb_longest_axis(orc):
    # src/aabb.rs:229-239
    L = orc.lib()
    assert L.orc_kat_aabb_longest_axis(D(0, 0, 0), D(2, 1, 1)) == 0
    assert L.orc_kat_aabb_longest_axis(D(0, 0, 0), D(1, 3, 1)) == 1
    assert L.orc_kat_aabb_longest_axis(D(0, 0, 0), D(1, 1, 4)) == 2
    # ties resolve toward z (aabb.rs:85-91)
    assert L.orc_kat_aabb_longest_axis(D(0, 0, 0), D(1, 1, 1)) == 2


def test_aabb_from_points_union_and_padding(orc):
    L = orc.lib()
    out = (C.c_double * 6)()
    L.orc_kat_aabb_from_points(D(1, 2, 3), D(4, 5, 6), out)  # src/aabb.rs:199-210
    assert list(out) == [1, 4, 2, 5, 3, 6]
    a, b = (C.c_double * 6)(0, 1, 0, 1, 0, 1), (C.c_double * 6)(1, 2, 1, 2, 1, 2)
    L.orc_kat_aabb_union(a, b, out)  # src/aabb.rs:241-253
    assert list(out) == [0, 2, 0, 2, 0, 2]
    # src/aabb.rs:43-51: an axis thinner than 1e-4 grows by 5e-5 on each side
    L.orc_kat_aabb_from_points(D(0, 0, 0), D(1, 0, 1), out)
    assert list(out) == [0, 1, -0.00005, 0.00005, 0, 1]


def test_sphere_uv(orc):
    # src/shapes/sphere.rs:152-169, exact equality like the reference's assert_eq!
    L = orc.lib()
    cases = [((1, 0, 0), (0.5, 0.5)), ((-1, 0, 0), (0.0, 0.5)), ((0, 1, 0), (0.5, 1.0)),
             ((0, -1, 0), (0.5, 0.0)), ((0, 0, 1), (0.25, 0.5)), ((0, 0, -1), (0.75, 0.5))]
    uv = (C.c_double * 2)()
    for p, want in cases:
        L.orc_kat_sphere_uv(D(*p), uv)
        assert (uv[0], uv[1]) == want


def test_ray_at(orc):
    out = (C.c_double * 3)()
    orc.lib().orc_kat_ray_at(D(1, 2, 3), D(4, 5, 6), 2.0, out)  # src/utils/ray.rs:57-64
    assert list(out) == [9.0, 12.0, 15.0]


def test_vec3_algebra(orc):
    # src/utils/vec3.rs:462-566
    out = (C.c_double * 20)()
    orc.lib().orc_kat_vec3(D(1, 2, 3), D(4, 5, 6), 2.0, out)
    o = list(out)
    assert o[0:3] == [5, 7, 9]
    assert o[3:6] == [-3, -3, -3]
    assert o[6:9] == [2, 4, 6]
    assert o[12] == 32.0
    assert o[13:16] == [-3, 6, -3]
    orc.lib().orc_kat_vec3(D(2, 4, 6), D(0, 0, 0), 2.0, out)
    assert list(out)[9:12] == [1, 2, 3]
    orc.lib().orc_kat_vec3(D(3, 4, 0), D(0, 0, 0), 1.0, out)
    assert out[16] == 5.0
    orc.lib().orc_kat_vec3(D(0, 5, 0), D(0, 0, 0), 1.0, out)
    assert list(out)[17:20] == [0, 1, 0]


def test_quaternion(orc):
    # src/utils/quaternion.rs:135-183
    L = orc.lib()
    out = (C.c_double * 3)()
    L.orc_kat_quat_axis_angle_rotate(D(1, 0, 0), 90.0, D(0, 1, 0), out)
    assert np.allclose(list(out), [0, 0, 1], atol=1e-10)
    # q1 * q2 rotates (0,0,1) to (1,0,0)
    def axis_angle(axis, deg):
        h = math.radians(deg) * 0.5
        n = math.sqrt(sum(a * a for a in axis))
        return [math.cos(h)] + [a / n * math.sin(h) for a in axis]
    q = (C.c_double * 4)()
    L.orc_kat_quat_mul(D(*axis_angle((0, 0, 1), 90)), D(*axis_angle((1, 0, 0), 90)), q)
    w, x, y, z = list(q)
    # rotate (0,0,1) with the product through numpy
    v = np.array([0.0, 0.0, 1.0])
    qv = np.array([x, y, z])
    rot = v + 2 * np.cross(qv, np.cross(qv, v) + w * v)
    assert np.allclose(rot, [1, 0, 0], atol=1e-10)
    # identity rotation
    L.orc_kat_quat_axis_angle_rotate(D(0, 1, 0), 0.0, D(1, 2, 3), out)
    assert np.allclose(list(out), [1, 2, 3], atol=1e-10)


def test_philox_known_answers(orc):
    # Random123 kat_vectors for philox4x32-10
    L = orc.lib()
    U = C.c_uint32
    vec = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in vec:
        out = (U * 4)()
        L.orc_kat_philox((U * 4)(*ctr), (U * 2)(*key), out)
        assert tuple(out) == want


def test_draw_layout(orc):
    # a = ((x0>>5)*2^26 + (x1>>6)) * 2^-53 on the zero-key zero-counter vector
    ab = (C.c_double * 2)()
    orc.lib().orc_kat_draw(0, 0, 0, 0, 0, ab)
    x0, x1, x2, x3 = 0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8
    assert ab[0] == ((x0 >> 5) * 2 ** 26 + (x1 >> 6)) * 2.0 ** -53
    assert ab[1] == ((x2 >> 5) * 2 ** 26 + (x3 >> 6)) * 2.0 ** -53
    assert 0.0 <= ab[0] < 1.0 and 0.0 <= ab[1] < 1.0

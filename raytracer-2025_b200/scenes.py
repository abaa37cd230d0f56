"""Client-side scene code of the reference, mirrored on the Python host API.

`obj_scene` is `obj_scene()` of the reference's src/main.rs:207-382 (BASELINE.json config 4, "assets/Final
triangle-mesh scene with assets/13.hdr environment map"), written against the same constructors
(`Wavefont::new`, `ConstantMedium::new_with_tex`, `Portal::new`, `Disney::builder`, `Transform::new`,
`build_box`, `Camera::from_json`).  It is scene CONSTRUCTION only: everything it builds is handed to the
CUDA core through rt_scene_create.

Three of the files the reference names are not shipped (.MISSING_LARGE_BLOBS): 初音未来.obj, 卒.obj and
assets/13.hdr (plus 水面_normal.png).  The reference itself would panic on the first `unwrap()`; here a mesh
that cannot be read is left out of the world ("reduced scene") and the environment is the
generated HDR image of `synthetic_hdr_environment`, flagged linear like a decoded .hdr (utils/image.rs:71-82),
so that `Environment::value`'s equirect lookup (shapes/environment.rs:14-24) runs on an image texture.
"""
import json
import os

import numpy as np

# order of the Wavefont::new calls and of world.add(..), src/main.rs:208-224 and :314-333
OBJ_FILES = [("miku", "初音未来.obj", False), ("ball", "玻璃球.obj", False), ("frame", "外框.obj", False), ("sound_box", "声匣.obj", False),
             ("mirror_door", "镜子门.obj", False), ("mirror", "镜子.obj", True), ("ring", "环.obj", False), ("portal_frame", "传送门框.obj", False),
             ("under_water", "水下.obj", False), ("water", "水面.obj", True), ("text", "文字.obj", False), ("mc", "mc.obj", False),
             ("umbralla", "伞.obj", False), ("checker", "卒.obj", False), ("forg", "雾.obj", False)]

# assets/Final/camera.json, copied value for value (Camera::from_json, camera.rs:119-160)
FINAL_CAMERA = {
    "aspect_ratio": 1.7777777777777777, "image_width": 1920, "vertical_fov_in_degrees": 23,
    "look_from": [1.842332124710083, 1.9965558052062988, 9.644098281860352],
    "look_at": [1.6544842720031738, 1.9639147520065308, 8.662442207336426],
    "vec_up": [-0.014803536236286163, 0.9994282126426697, -0.030399203300476074],
    "defocus_angle_in_degrees": 0.0, "focus_distance": 1.0000004646134415,
}


def synthetic_hdr_environment(width=512, height=256):
    """A generated equirect radiance map standing in for assets/13.hdr (not shipped): horizon-bright sky over a dim
    ground, plus a small sun above 1.0 so that the image is genuinely high dynamic range.  (H, W, 4) float32."""
    v = (np.arange(height, dtype=np.float64) + 0.5) / height          # 0 = top row
    u = (np.arange(width, dtype=np.float64) + 0.5) / width
    uu, vv = np.meshgrid(u, v)
    elev = (0.5 - vv) * np.pi                                          # +pi/2 at the top
    sky = np.stack([0.25 + 0.35 * np.cos(elev) ** 4, 0.40 + 0.35 * np.cos(elev) ** 4, 0.75 + 0.15 * np.cos(elev) ** 2], axis=-1)
    ground = np.stack([0.18 + 0 * uu, 0.16 + 0.02 * np.sin(12 * np.pi * uu), 0.14 + 0 * uu], axis=-1)
    img = np.where((elev > 0)[..., None], sky, ground)
    sun = np.exp(-(((uu - 0.31) * 2 * np.cos(elev)) ** 2 + (vv - 0.27) ** 2) / (2 * 0.012 ** 2))
    img = img + sun[..., None] * np.array([40.0, 36.0, 28.0])
    out = np.ones((height, width, 4), dtype=np.float32)
    out[..., :3] = img.astype(np.float32)
    return out


def obj_scene(rt, assets, width=None, spp=3000, depth=30, environment=None, seed=1):
    """-> (HostScene, list of the meshes that could not be loaded).  `assets` is whatever objload.Wavefont accepts:
    the asset directory, or an objload.AssetPack of it."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("objload", os.path.join(os.path.dirname(os.path.abspath(__file__)), "objload.py"))
    objload = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(objload)

    b = rt.Builder(seed)
    wf = objload.Wavefont(b, assets)
    mesh, missing = {}, []
    for name, file_name, vanilla in OBJ_FILES:
        mesh[name] = wf.new(file_name, "Final", vanilla)
        if mesh[name] is None:
            missing.append(file_name)

    forg = None
    if mesh["forg"] is not None:
        forg = b.medium(mesh["forg"], 0.05, b.solid(1.0, 0.936, 0.381))

    portal_material = b.portal([1.0, 1.0, 1.0], [0.0, -6.3, 1.1], [1.0, 0.0, 0.0, 0.0])
    anchor = np.array([-5.8035, -0.9983, -7.7198])
    portal_u = np.array([-3.8206, -0.9983, -8.3722]) - anchor
    portal_v = np.array([-5.8035, 3.1159, -7.7198]) - anchor
    portal = b.quad(anchor, portal_u, portal_v, portal_material)

    def board(material):
        return b.quad([-1.0, 0.0, -1.0], [0.0, 0.0, 2.0], [2.0, 0.0, 0.0], material)

    translucent_board = b.transform(board(b.disney(diff_trans=1.0, roughness=1.0, thin=1.0)), offset=[2.8145, -0.23603, -19.501],
                                    quat=b.quat_axis_angle([0.993, -0.082, 0.082], 90.4), scale=[2.616, 1.0, 1.0])
    light_xf = dict(offset=[-0.44579, 5.2955, 0.89889], quat=b.quat_axis_angle([0.921, 0.021, 0.389], 34.7), scale=[3.415, 3.415, 3.415])
    yellow_xf = dict(offset=[-1.0053, -1.9655, -4.242], quat=b.quat_axis_angle([0.766, 0.483, -0.423], 85.7),
                     scale=[1.0 * 1.499, 1.0 * 1.499, 1.0 * 1.499])
    light_board = b.transform(board(b.diffuse_light(b.solid(4.0, 4.0, 4.0))), **light_xf)
    yellow_board = b.transform(board(b.diffuse_light(b.solid(5.0 * 1.0, 5.0 * 0.687, 5.0 * 0.0))), **yellow_xf)
    black_box = b.transform(b.box([-1.0, -1.0, -1.0], [1.0, 1.0, 1.0], b.diffuse_light(b.solid(0.0, 0.0, 0.0))),
                            offset=[-4.9891, -6.4998, -8.3939], scale=[1.0 * 6.244, 1.0 * 6.244, 1.0 * 6.244])

    order = [mesh["miku"], light_board, mesh["ball"], mesh["frame"], mesh["sound_box"], mesh["mirror_door"], mesh["mirror"], mesh["ring"],
             mesh["portal_frame"], mesh["under_water"], mesh["water"], mesh["text"], translucent_board, mesh["mc"], portal, mesh["umbralla"],
             yellow_board, black_box, forg, mesh["checker"]]
    world = b.list([o for o in order if o is not None])
    lights = b.list([b.transform(board(b.empty()), **light_xf), b.transform(board(b.empty()), **yellow_xf)])

    env = synthetic_hdr_environment() if environment is None else environment
    background = b.image(env, raw=False, linear_format=True)  # ImageTexture::new("13.hdr"): nearest texel, no sRGB decode
    cam = FINAL_CAMERA
    hs = b.finish(world, lights, width=width or cam["image_width"], aspect=cam["aspect_ratio"], spp=spp, max_depth=depth,
                  vfov=float(cam["vertical_fov_in_degrees"]), look_from=cam["look_from"], look_at=cam["look_at"], vup=cam["vec_up"],
                  defocus_angle=cam["defocus_angle_in_degrees"], focus_dist=cam["focus_distance"], background=background)
    hs._builder = b
    return hs, missing


def camera_from_json(path):
    """Camera::from_json's field list (camera.rs:143-160)."""
    with open(path) as f:
        p = json.load(f)
    return {k: p[k] for k in FINAL_CAMERA}

"""OBJ/MTL -> host-mirror objects, following the reference's loader (src/shapes/obj.rs:83-345).

Host-side I/O outside the hot path, kept in Python: it parses the files, decodes images
with PIL and drives the `Builder` so that the flattened scene is exactly what `Wavefont::new` would
hand to `Camera::render`:

  * models are split like tobj does with GPU_LOAD_OPTIONS (triangulate, single index): a new model per
    `o`/`g` and per `usemtl` run;
  * every face becomes `Triangle::new(p1, p2-p1, p3-p1, RemappedMaterial{..})` (obj.rs:143-183), each
    model one `BVH::from_vec` (obj.rs:185-189), degenerate triangles are dropped;
  * MTL -> material mapping of load_materials (obj.rs:212-345): Disney unless `vanilla_material` and the
    material is a pure metal / pure glass; `Ke` / `map_Ke` wrap a DiffuseLight, `map_d` / `d < 1` wrap a
    Mix with Transparent, `map_Bump` becomes a raw normal map;
  * the reference zips MODELS with the MATERIALS' normal maps by index (obj.rs:129) and therefore also
    drops models beyond the number of materials — reproduced, not fixed.
"""
import os

import numpy as np

RT_NONE = 0xFFFFFFFF


def parse_mtl(path):
    """tobj's material fields + `unknown_param` (everything it does not know, as raw strings)."""
    mats = []
    cur = None
    known_tex = {"map_Kd": "diffuse_texture", "map_d": "dissolve_texture", "map_Bump": "normal_texture", "map_bump": "normal_texture",
                 "bump": "normal_texture", "map_Ka": "ambient_texture", "map_Ks": "specular_texture", "map_Ns": "shininess_texture"}
    with open(path, encoding="utf-8", errors="replace") as f:
        for line in f:
            line = line.strip()
            if not line or line.startswith("#"):
                continue
            key, _, rest = line.partition(" ")
            rest = rest.strip()
            if key == "newmtl":
                cur = {"name": rest, "unknown_param": {}}
                mats.append(cur)
            elif cur is None:
                continue
            elif key == "Kd":
                cur["diffuse"] = [float(x) for x in rest.split()[:3]]
            elif key == "Ni":
                cur["optical_density"] = float(rest.split()[0])
            elif key == "d":
                cur["dissolve"] = float(rest.split()[0])
            elif key in known_tex:
                cur[known_tex[key]] = rest
            elif key in ("Ka", "Ks", "Ns", "illum"):
                pass
            else:
                cur["unknown_param"][key] = rest
    return mats


def parse_obj(path):
    """Returns (models, mtllibs); a model = dict(name, material (name or None), faces [[(v,vt,vn) x3], ...])."""
    pos, tex, nrm = [], [], []
    models, mtllibs = [], []
    cur = None
    name = "unnamed_object"
    material = None

    def flush():
        nonlocal cur
        if cur and cur["faces"]:
            models.append(cur)
        cur = None

    def idx(tok, n):
        if tok == "":
            return None
        i = int(tok)
        return i - 1 if i > 0 else n + i

    with open(path, encoding="utf-8", errors="replace") as f:
        for line in f:
            line = line.strip()
            if not line or line.startswith("#"):
                continue
            key, _, rest = line.partition(" ")
            if key == "v":
                pos.append([float(x) for x in rest.split()[:3]])
            elif key == "vt":
                t = [float(x) for x in rest.split()[:2]]
                tex.append(t + [0.0] * (2 - len(t)))
            elif key == "vn":
                nrm.append([float(x) for x in rest.split()[:3]])
            elif key in ("o", "g"):
                flush()
                name = rest.strip() or "unnamed_object"
            elif key == "usemtl":
                flush()  # tobj starts a new model when the material changes inside an object
                material = rest.strip()
            elif key == "mtllib":
                mtllibs.append(rest.strip())
            elif key == "f":
                if cur is None:
                    cur = {"name": name, "material": material, "faces": []}
                corners = []
                for tok in rest.split():
                    parts = (tok.split("/") + ["", ""])[:3]
                    corners.append((idx(parts[0], len(pos)), idx(parts[1], len(tex)), idx(parts[2], len(nrm))))
                for k in range(1, len(corners) - 1):  # fan triangulation
                    cur["faces"].append([corners[0], corners[k], corners[k + 1]])
    flush()
    return models, mtllibs, np.array(pos, dtype=np.float64), np.array(tex, dtype=np.float64).reshape(-1, 2), np.array(nrm, dtype=np.float64)


def pil_image_loader(path):
    """-> (H, W, 4) float32 in [0,1] like `image::DynamicImage::into_rgba32f`, linear_format flag; None if missing."""
    if not os.path.isfile(path):
        return None
    from PIL import Image
    im = Image.open(path)
    linear = (im.format or "").upper() in ("HDR", "EXR", "AVIF")
    a = np.asarray(im.convert("RGBA"), dtype=np.float32) / 255.0
    return a, linear


class FileAssets:
    """The asset directory the reference reads (`assets/<prefix>/<file>`): parse on demand."""

    def __init__(self, assets_dir, image_loader=pil_image_loader):
        self.dir, self.load_image = assets_dir, image_loader

    def obj(self, prefix, file_name):
        path = os.path.join(self.dir, prefix, file_name)
        return parse_obj(path) if os.path.isfile(path) else None

    def mtl(self, prefix, lib):
        path = os.path.join(self.dir, prefix, lib)
        return parse_mtl(path) if os.path.isfile(path) else None

    def image(self, rel_path):
        return self.load_image(os.path.join(self.dir, rel_path))


class AssetPack:
    """The same three lookups served from one .npz: the parsed OBJ tables, the parsed MTL records and the decoded
    8-bit images of an asset directory, written once by `write_pack` (tests/golden/make_final_pack.py does it for
    the reference's assets/Final, which does not exist on the GPU box).  Nothing is re-derived: `obj()` returns
    exactly what `parse_obj` returned when the pack was written."""

    def __init__(self, path):
        import json
        self.z = np.load(path, allow_pickle=False)
        self.meta = json.loads(bytes(self.z["meta"]).decode("utf-8"))

    def obj(self, prefix, file_name):
        rec = self.meta["objs"].get(prefix + "/" + file_name)
        if rec is None:
            return None
        k = rec["key"]
        corners, face_model = self.z[k + "/corners"], self.z[k + "/face_model"]
        models = [{"name": m["name"], "material": m["material"], "faces": []} for m in rec["models"]]
        for f, mi in zip(corners.tolist(), face_model.tolist()):
            models[mi]["faces"].append([tuple(None if i < 0 else i for i in c) for c in f])
        return models, rec["mtllibs"], self.z[k + "/pos"], self.z[k + "/tex"], self.z[k + "/nrm"]

    def mtl(self, prefix, lib):
        return self.meta["mtls"].get(prefix + "/" + lib)

    def image(self, rel_path):
        rec = self.meta["images"].get(rel_path)
        if rec is None:
            return None
        return self.z[rec["key"]].astype(np.float32) / 255.0, bool(rec["linear"])


def write_pack(path, assets_dir, obj_files, max_image_side=None):
    """Parse `obj_files` [(prefix, file)], their MTL libraries and every image those name, and store the results.
    Images are kept as decoded RGBA8 (optionally box-reduced so that the longer side is <= max_image_side)."""
    import json
    src = FileAssets(assets_dir)
    arrays, meta = {}, {"objs": {}, "mtls": {}, "images": {}}
    tex_keys = ("diffuse_texture", "dissolve_texture", "normal_texture", "ambient_texture", "specular_texture", "shininess_texture")
    for n, (prefix, file_name) in enumerate(obj_files):
        got = src.obj(prefix, file_name)
        if got is None:
            continue
        models, mtllibs, pos, tex, nrm = got
        k = f"obj{n}"
        corners = [[[-1 if i is None else i for i in c] for c in f] for m in models for f in m["faces"]]
        arrays[k + "/corners"] = np.array(corners, dtype=np.int32).reshape(-1, 3, 3)
        arrays[k + "/face_model"] = np.array([i for i, m in enumerate(models) for _ in m["faces"]], dtype=np.int32)
        arrays[k + "/pos"], arrays[k + "/tex"], arrays[k + "/nrm"] = pos, tex, nrm
        meta["objs"][prefix + "/" + file_name] = {"key": k, "mtllibs": mtllibs,
                                                  "models": [{"name": m["name"], "material": m["material"]} for m in models]}
        for lib in mtllibs:
            mats = src.mtl(prefix, lib)
            if mats is None:
                continue
            meta["mtls"][prefix + "/" + lib] = mats
            for m in mats:
                names = [m[t] for t in tex_keys if t in m] + [v for kk, v in m["unknown_param"].items() if kk.startswith("map_")]
                for name in names:
                    if name.startswith("-bm"):
                        name = name[3:].split()[-1]
                    rel = prefix + "/" + name
                    if rel in meta["images"]:
                        continue
                    img = src.image(rel)
                    if img is None:
                        continue
                    px, linear = img
                    a = np.clip(np.rint(px * 255.0), 0, 255).astype(np.uint8)
                    if max_image_side and max(a.shape[:2]) > max_image_side:
                        from PIL import Image
                        s = max_image_side / max(a.shape[:2])
                        a = np.asarray(Image.fromarray(a, "RGBA").resize((max(1, round(a.shape[1] * s)), max(1, round(a.shape[0] * s))), Image.BOX))
                    ik = f"img{len(meta['images'])}"
                    arrays[ik] = a
                    meta["images"][rel] = {"key": ik, "linear": bool(linear)}
    arrays["meta"] = np.frombuffer(json.dumps(meta, ensure_ascii=False).encode("utf-8"), dtype=np.uint8)
    np.savez_compressed(path, **arrays)
    return meta


class Wavefont:
    """Wavefont::new(file_name, prefix, vanilla_material) — obj.rs:117-134.  `assets` is an asset directory
    (str), a FileAssets or an AssetPack."""

    def __init__(self, builder, assets, image_loader=pil_image_loader):
        self.b = builder
        self.src = FileAssets(assets, image_loader) if isinstance(assets, (str, os.PathLike)) else assets
        self._tex_cache = {}

    def _image_texture(self, rel_path, raw=False):
        key = (rel_path, raw)
        if key not in self._tex_cache:
            got = self.src.image(rel_path)
            if got is None:
                self._tex_cache[key] = self.b.image_missing()  # renders cyan, alpha 1 (texture.rs:102-105,167-169)
            else:
                pixels, linear = got
                self._tex_cache[key] = self.b.image(pixels, raw=raw, linear_format=linear)
        return self._tex_cache[key]

    def _materials(self, mtl_mats, prefix, vanilla):
        b = self.b
        transparent = b.transparent()
        mats, normals = [], []

        def fparam(m, key, default):
            try:
                return float(m["unknown_param"][key].split()[0])
            except (KeyError, ValueError, IndexError):
                return default

        for m in mtl_mats:
            if "diffuse_texture" in m:
                base_tex, base_color = self._image_texture(prefix + "/" + m["diffuse_texture"]), None
            elif "diffuse" in m:
                base_tex, base_color = None, m["diffuse"]
            else:
                raise ValueError("The material should at least have one diffuse!")
            roughness, anisotropic = fparam(m, "Pr", 0.5), fparam(m, "aniso", 0.0)
            sheen, metallic = fparam(m, "Ps", 0.0), fparam(m, "Pm", 0.0)
            clearcoat, clearcoat_gloss = fparam(m, "Pc", 0.0), fparam(m, "Pcr", 0.0)
            ior = m.get("optical_density", 1.45)
            spec_trans = 0.0
            if "Tf" in m["unknown_param"]:
                vals = []
                for s in m["unknown_param"]["Tf"].split():
                    try:
                        vals.append(float(s))
                    except ValueError:
                        pass
                spec_trans = sum(vals) / len(vals) if vals else float("nan")
            if vanilla and metallic == 1.0:
                # Metal::new(base_color.value(0,0,ZERO), roughness): a texture would be sampled at (0,0); only the
                # constant-colour case is expressible without evaluating textures on the host
                if base_color is None:
                    raise NotImplementedError("vanilla metal with a diffuse texture")
                mat = b.metal(base_color, roughness)
            elif vanilla and spec_trans == 1.0:
                mat = b.dielectric(base_tex if base_tex is not None else b.solid(*base_color), ior)
            else:
                mat = b.disney(base_color if base_color is not None else (0.8, 0.8, 0.8), tex=base_tex if base_tex is not None else RT_NONE,
                               roughness=roughness, anisotropic=anisotropic, sheen=sheen, clearcoat=clearcoat,
                               clearcoat_gloss=clearcoat_gloss, metallic=metallic, ior=ior, spec_trans=spec_trans)
            if "Ke" in m["unknown_param"]:
                try:
                    ke = [float(s) for s in m["unknown_param"]["Ke"].split()]
                except ValueError:
                    ke = []
                if len(ke) == 3:
                    mat = b.diffuse_light(b.solid(*ke), inner=mat)
            if "map_Ke" in m["unknown_param"]:
                mat = b.diffuse_light(self._image_texture(prefix + "/" + m["unknown_param"]["map_Ke"]), inner=mat)
            if "dissolve_texture" in m:
                mat = b.mix_image(transparent, mat, self._image_texture(prefix + "/" + m["dissolve_texture"]))
            if "dissolve" in m and m["dissolve"] < 1.0:
                mat = b.mix(transparent, mat, m["dissolve"])
            mats.append(mat)
            if "normal_texture" in m:
                name = m["normal_texture"]
                if name.startswith("-bm"):
                    parts = name[3:].split()
                    name = parts[-1] if parts else name
                normals.append(self._image_texture(prefix + "/" + name, raw=True))
            else:
                normals.append(RT_NONE)
        return mats, normals

    def new(self, file_name, prefix, vanilla_material):
        """-> hittable id of `Hittables[ BVH per model ]`, or None when the OBJ cannot be read."""
        b = self.b
        got = self.src.obj(prefix, file_name)
        if got is None:
            return None
        models, mtllibs, pos, tex, nrm = got
        mtl_mats = []
        ok = True
        for lib in mtllibs:
            lib_mats = self.src.mtl(prefix, lib)
            if lib_mats is not None:
                mtl_mats += lib_mats
            else:
                ok = False
        mats, normals = self._materials(mtl_mats, prefix, vanilla_material) if ok else ([], [])
        mat_index = {m["name"]: i for i, m in enumerate(mtl_mats)}
        empty = b.empty()
        bvhs = []
        for model, normal_tex in zip(models, normals):  # obj.rs:129 — models zipped with the MATERIALS' normal maps
            mid = mat_index.get(model["material"]) if ok else None
            inner = mats[mid] if mid is not None else empty
            faces = []
            for face in model["faces"]:
                p = [pos[c[0]] for c in face]
                t = [tex[c[1]] for c in face]
                n = [nrm[c[2]] for c in face]
                f = b.obj_face(inner, p, t, n, normal_tex)
                if f != RT_NONE:
                    faces.append(f)
            if faces:
                bvhs.append(b.bvh(faces))
        return b.list(bvhs)

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


class Wavefont:
    """Wavefont::new(file_name, prefix, vanilla_material) — obj.rs:117-134."""

    def __init__(self, builder, assets_dir, image_loader=pil_image_loader):
        self.b, self.dir, self.load_image = builder, assets_dir, image_loader
        self._tex_cache = {}

    def _image_texture(self, rel_path, raw=False):
        key = (rel_path, raw)
        if key not in self._tex_cache:
            got = self.load_image(os.path.join(self.dir, rel_path))
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
        path = os.path.join(self.dir, prefix, file_name)
        if not os.path.isfile(path):
            return None
        models, mtllibs, pos, tex, nrm = parse_obj(path)
        mtl_mats = []
        ok = True
        for lib in mtllibs:
            p = os.path.join(self.dir, prefix, lib)
            if os.path.isfile(p):
                mtl_mats += parse_mtl(p)
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

#!/usr/bin/env python3
"""Builds raytracer-2025_b200/rust/reference_src.patch: the edits a maintainer applies to the reference crate
(caidj0/Raytracer-2025, `src/`) so that `Camera::render` runs on the CUDA core.

    python raytracer-2025_b200/rust/make_patch.py [/root/reference]      (needs the reference checkout)

It copies `src/` to a scratch directory, inserts the code below after anchor lines that exist in the reference,
adds the three new modules of this directory (ffi.rs, flatten.rs, camera_render.rs) and diffs with one line of
context.  What the patch does, type by type:

  * `Hittable`, `Material`, `Texture` get a doc-hidden `flatten` method whose default is "unsupported";
  * every concrete type of the crate implements it by COPYING its fields into the rt_* records of include/rt2025.h -
    nothing is recomputed, so the device intersects with the very numbers the CPU code would use;
  * three closures become data next to the closure (the closures stay, the CPU path is untouched): `Mix.ratio`
    (a constant or an image alpha, material.rs:228-247), `Portal.f` (offset + quaternion, portal.rs:15-24) and
    `Disney.param_fn` (constants, optionally a base-colour texture: disney.rs:786-805, obj.rs:271-293);
  * `BVH::from_vec` consumes and re-orders its input, so the root of a `from_vec` call remembers the order it was
    given (addresses of the boxed children): rt_scene_create needs it to reproduce the reference's tie order;
  * `Camera::render` keeps `initilize()` and hands the rest to rt_scene_create / rt_render_rgb8.

NOT COMPILED: there is no rustc in this environment.  The C++ host mirror
(raytracer-2025_b200/host/rt2025.hpp) is the same logic, compiled and tested; tests/test_abi.py pins the struct layouts.
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"

# (file, anchor regex, text inserted AFTER the anchor line)
INSERT = [
    # ---------------------------------------------------------------- traits
    ("hit.rs", r"^pub trait Hittable: Send \+ Sync \{", '''    #[doc(hidden)]
    fn flatten(&self, _f: &mut crate::flatten::Flattener) -> Result<u32, crate::flatten::Unsupported> {
        Err(crate::flatten::Unsupported("a user-defined Hittable cannot be flattened for the GPU core"))
    }
'''),
    ("material.rs", r"^pub trait Material: Send \+ Sync \{", '''    #[doc(hidden)]
    fn flatten(&self, _f: &mut crate::flatten::Flattener) -> Result<crate::ffi::rt_material, crate::flatten::Unsupported> {
        Err(crate::flatten::Unsupported("a user-defined Material cannot be flattened for the GPU core"))
    }
'''),
    ("texture.rs", r"^pub trait Texture: Send \+ Sync \+ Debug \{", '''    #[doc(hidden)]
    fn flatten(&self, _f: &mut crate::flatten::Flattener) -> Result<crate::ffi::rt_texture, crate::flatten::Unsupported> {
        Err(crate::flatten::Unsupported("a user-defined Texture cannot be flattened for the GPU core"))
    }
'''),
    ("lib.rs", r"^pub mod camera;", "pub mod ffi;\npub mod flatten;\n"),
    ("camera.rs", r"^impl Camera \{", "    // `render` lives in camera/render_gpu.rs now; the rayon loop below stays available as `render_cpu`\n"),
    # ---------------------------------------------------------------- accessors the flattener needs
    ("utils/vec3.rs", r"^    pub fn x\(&self\) -> f64 \{", None),  # placeholder: see PREPEND below
    # ---------------------------------------------------------------- shapes
    ("shapes/sphere.rs", r"^impl Hittable for Sphere \{", '''    fn flatten(&self, f: &mut crate::flatten::Flattener) -> Result<u32, crate::flatten::Unsupported> {
        use crate::ffi::*;
        let mat = f.material(&self.mat)?;
        // center is a Ray: origin = centre at time 0, direction = centre(1) - centre(0) (sphere.rs:35-51)
        f.spheres.push(rt_sphere { center: self.center.origin().e(), center_vec: self.center.direction().e(), radius: self.radius, reserved: 0.0 });
        Ok(f.push_object(RT_OBJ_SPHERE, mat, f.spheres.len() as u32 - 1, &self.bbox, &[]))
    }
'''),
    ("shapes/quad.rs", r"^impl Hittable for Quad \{", '''    fn flatten(&self, f: &mut crate::flatten::Flattener) -> Result<u32, crate::flatten::Unsupported> {
        use crate::ffi::*;
        let mat = f.material(&self.mat)?;
        f.planars.push(rt_planar { anchor: self.anchor.e(), u: self.u.e(), v: self.v.e(), normal: self.normal.as_inner().e(),
                                   parm_d: self.parm_d, w: self.w.e(), area: self.area, reserved: 0.0 });
        Ok(f.push_object(RT_OBJ_QUAD, mat, f.planars.len() as u32 - 1, &self.bbox, &[]))
    }
'''),
    ("shapes/triangle.rs", r"^impl Hittable for Triangle \{", '''    fn flatten(&self, f: &mut crate::flatten::Flattener) -> Result<u32, crate::flatten::Unsupported> {
        use crate::ffi::*;
        let mat = f.material(&self.mat)?;
        f.planars.push(rt_planar { anchor: self.anchor.e(), u: self.u.e(), v: self.v.e(), normal: self.normal.as_inner().e(),
                                   parm_d: self.parm_d, w: self.w.e(), area: self.area, reserved: 0.0 });
        Ok(f.push_object(RT_OBJ_TRIANGLE, mat, f.planars.len() as u32 - 1, &self.bbox, &[]))
    }
'''),
    ("hits.rs", r"^impl Hittable for Hittables \{", '''    fn flatten(&self, f: &mut crate::flatten::Flattener) -> Result<u32, crate::flatten::Unsupported> {
        // children in insertion order: Iterator::min_by keeps the first of equal minima (hits.rs:42)
        let kids = self.objects.iter().map(|o| o.flatten(f)).collect::<Result<Vec<u32>, _>>()?;
        Ok(f.push_object(crate::ffi::RT_OBJ_LIST, crate::ffi::RT_NONE, crate::ffi::RT_NONE, &self.bbox, &kids))
    }
'''),
    ("bvh.rs", r"^impl Hittable for BVH \{", '''    fn flatten(&self, f: &mut crate::flatten::Flattener) -> Result<u32, crate::flatten::Unsupported> {
        // One RT_OBJ_BVH per BVH::from_vec call, children in the order the call received them; the library recomputes the
        // median splits from the children's bounding boxes (bvh.rs:21-43) to reproduce which child wins an exact tie.
        let Some(order) = &self.input_order else {
            return Err(crate::flatten::Unsupported("inner BVH node reached outside its from_vec root"));
        };
        let mut by_addr = std::collections::HashMap::new();
        self.flatten_leaves(f, &mut by_addr)?;
        let kids: Vec<u32> = order.iter().map(|a| by_addr[a]).collect();
        Ok(f.push_object(crate::ffi::RT_OBJ_BVH, crate::ffi::RT_NONE, crate::ffi::RT_NONE, &self.bbox, &kids))
    }
'''),
    ("bvh.rs", r"^impl BVH \{", '''    /// leaves of this subtree -> flattened ids, keyed by the address of the boxed object
    fn flatten_leaves(&self, f: &mut crate::flatten::Flattener, out: &mut std::collections::HashMap<usize, u32>) -> Result<(), crate::flatten::Unsupported> {
        for child in [&self.left, &self.right].into_iter().flatten() {
            match child.as_bvh_inner() {
                Some(inner) => inner.flatten_leaves(f, out)?,
                None => {
                    out.insert(child.as_ref() as *const dyn Hittable as *const () as usize, child.flatten(f)?);
                }
            }
        }
        Ok(())
    }
'''),
    ("shapes.rs", r"^impl Hittable for Transform \{", '''    fn flatten(&self, f: &mut crate::flatten::Flattener) -> Result<u32, crate::flatten::Unsupported> {
        use crate::ffi::*;
        let kid = self.object.flatten(f)?;
        f.transforms.push(rt_transform { offset: self.offset.e(), quat: self.quaternion.wxyz(), scale: self.scale.e() });
        Ok(f.push_object(RT_OBJ_TRANSFORM, RT_NONE, f.transforms.len() as u32 - 1, &self.bbox, &[kid]))
    }
'''),
    ("volume.rs", r"^impl Hittable for ConstantMedium \{", '''    fn flatten(&self, f: &mut crate::flatten::Flattener) -> Result<u32, crate::flatten::Unsupported> {
        use crate::ffi::*;
        let kid = self.boundary.flatten(f)?;
        let phase = crate::material::Material::flatten(self.phase_function.as_ref(), f)?;  // Box<Isotropic>: one material record per medium
        f.materials.push(phase);
        let mat = f.materials.len() as u32 - 1;
        f.media.push(rt_medium { neg_inv_density: self.neg_inv_density, reserved: 0.0 });
        Ok(f.push_object(RT_OBJ_MEDIUM, mat, f.media.len() as u32 - 1, self.boundary.bounding_box(), &[kid]))
    }
'''),
    ("shapes/obj.rs", r"^impl Hittable for Wavefont \{", '''    fn flatten(&self, f: &mut crate::flatten::Flattener) -> Result<u32, crate::flatten::Unsupported> {
        self.objects.flatten(f)  // Hittables[ BVH per model ] (obj.rs:117-134)
    }
'''),
    # ---------------------------------------------------------------- materials
    ("material.rs", r"^impl Material for EmptyMaterial \{", '''    fn flatten(&self, _f: &mut crate::flatten::Flattener) -> Result<crate::ffi::rt_material, crate::flatten::Unsupported> {
        Ok(crate::ffi::rt_material { kind: crate::ffi::RT_MAT_EMPTY, tex: crate::ffi::RT_NONE, inner: crate::ffi::RT_NONE, inner2: crate::ffi::RT_NONE, ..Default::default() })
    }
'''),
    ("material.rs", r"^impl Material for Lambertian \{", '''    fn flatten(&self, f: &mut crate::flatten::Flattener) -> Result<crate::ffi::rt_material, crate::flatten::Unsupported> {
        use crate::ffi::*;
        Ok(rt_material { kind: RT_MAT_LAMBERTIAN, tex: f.texture(&self.texture)?, inner: RT_NONE, inner2: RT_NONE, ..Default::default() })
    }
'''),
    ("material.rs", r"^impl Material for Metal \{", '''    fn flatten(&self, _f: &mut crate::flatten::Flattener) -> Result<crate::ffi::rt_material, crate::flatten::Unsupported> {
        use crate::ffi::*;
        Ok(rt_material { kind: RT_MAT_METAL, tex: RT_NONE, inner: RT_NONE, inner2: RT_NONE, color: self.albedo.e(), param: self.fuzz, ..Default::default() })
    }
'''),
    ("material.rs", r"^impl Material for Dielectric \{", '''    fn flatten(&self, f: &mut crate::flatten::Flattener) -> Result<crate::ffi::rt_material, crate::flatten::Unsupported> {
        use crate::ffi::*;
        Ok(rt_material { kind: RT_MAT_DIELECTRIC, tex: f.texture(&self.attentuation)?, inner: RT_NONE, inner2: RT_NONE, param: self.refraction_index, ..Default::default() })
    }
'''),
    ("material.rs", r"^impl Material for DiffuseLight \{", '''    fn flatten(&self, f: &mut crate::flatten::Flattener) -> Result<crate::ffi::rt_material, crate::flatten::Unsupported> {
        use crate::ffi::*;
        let inner = match &self.material { Some(m) => f.material(m)?, None => RT_NONE };  // the wrapped material precedes its wrapper
        Ok(rt_material { kind: RT_MAT_DIFFUSE_LIGHT, tex: f.texture(&self.texture)?, inner, inner2: RT_NONE, ..Default::default() })
    }
'''),
    ("material.rs", r"^impl Material for Isotropic \{", '''    fn flatten(&self, f: &mut crate::flatten::Flattener) -> Result<crate::ffi::rt_material, crate::flatten::Unsupported> {
        use crate::ffi::*;
        Ok(rt_material { kind: RT_MAT_ISOTROPIC, tex: f.texture(&self.texture)?, inner: RT_NONE, inner2: RT_NONE, ..Default::default() })
    }
'''),
    ("material.rs", r"^impl Material for Transparent \{", '''    fn flatten(&self, _f: &mut crate::flatten::Flattener) -> Result<crate::ffi::rt_material, crate::flatten::Unsupported> {
        Ok(crate::ffi::rt_material { kind: crate::ffi::RT_MAT_TRANSPARENT, tex: crate::ffi::RT_NONE, inner: crate::ffi::RT_NONE, inner2: crate::ffi::RT_NONE, ..Default::default() })
    }
'''),
    ("material.rs", r"^impl Material for Mix \{", '''    fn flatten(&self, f: &mut crate::flatten::Flattener) -> Result<crate::ffi::rt_material, crate::flatten::Unsupported> {
        use crate::ffi::*;
        let (inner, inner2) = (f.material(&self.mat1)?, f.material(&self.mat2)?);
        Ok(match &self.ratio_src {  // the closure is one of two shapes (material.rs:228-247)
            MixRatio::Constant(r) => rt_material { kind: RT_MAT_MIX, tex: RT_NONE, inner, inner2, param: *r, ..Default::default() },
            MixRatio::Alpha(tex) => {
                let as_dyn: Arc<dyn Texture> = tex.clone();
                rt_material { kind: RT_MAT_MIX, tex: f.texture(&as_dyn)?, inner, inner2, ..Default::default() }
            }
        })
    }
'''),
    ("material/portal.rs", r"^impl Material for Portal \{", '''    fn flatten(&self, _f: &mut crate::flatten::Flattener) -> Result<crate::ffi::rt_material, crate::flatten::Unsupported> {
        use crate::ffi::*;
        let mut v = [0.0; 16];
        v[0..3].copy_from_slice(&self.position_offset.e());   // v[0..3) = offset, v[3..7) = quaternion (w, x, y, z): rt2025.h
        v[3..7].copy_from_slice(&self.rotation.wxyz());
        Ok(rt_material { kind: RT_MAT_PORTAL, tex: RT_NONE, inner: RT_NONE, inner2: RT_NONE, color: self.attenuation.e(), v, ..Default::default() })
    }
'''),
    ("material/disney.rs", r"^impl Material for Disney \{", '''    fn flatten(&self, f: &mut crate::flatten::Flattener) -> Result<crate::ffi::rt_material, crate::flatten::Unsupported> {
        use crate::ffi::*;
        // param_fn is only ever built from constants (DisneyBuilder::build, Default) or constants + a base-colour texture (obj.rs:271-293)
        let Some((p, base_tex)) = &self.flat else {
            return Err(crate::flatten::Unsupported("Disney with a hand-written param_fn closure"));
        };
        let mut v = [0.0; 16];
        v[RT_DISNEY_ROUGHNESS] = p.roughness; v[RT_DISNEY_ANISOTROPIC] = p.anisotropic; v[RT_DISNEY_SHEEN] = p.sheen;
        v[RT_DISNEY_SHEEN_TINT] = p.sheen_tint; v[RT_DISNEY_CLEARCOAT] = p.clearcoat; v[RT_DISNEY_CLEARCOAT_GLOSS] = p.clearcoat_gloss;
        v[RT_DISNEY_SPECULAR_TINT] = p.specular_tint; v[RT_DISNEY_METALLIC] = p.metallic; v[RT_DISNEY_IOR] = p.ior;
        v[RT_DISNEY_FLATNESS] = p.flatness; v[RT_DISNEY_SPEC_TRANS] = p.spec_trans; v[RT_DISNEY_DIFF_TRANS] = p.diff_trans;
        v[RT_DISNEY_THIN] = if p.thin { 1.0 } else { 0.0 };
        let tex = match base_tex { Some(t) => f.texture(t)?, None => RT_NONE };
        Ok(rt_material { kind: RT_MAT_DISNEY, tex, inner: RT_NONE, inner2: RT_NONE, color: p.base_color.e(), v, ..Default::default() })
    }
'''),
    ("shapes/obj.rs", r"^impl Material for RemappedMaterial \{", '''    fn flatten(&self, f: &mut crate::flatten::Flattener) -> Result<crate::ffi::rt_material, crate::flatten::Unsupported> {
        use crate::ffi::*;
        let inner = f.material(&self.material)?;
        let normal_tex = match &self.normal_tex {
            Some(t) => { let as_dyn: Arc<dyn crate::texture::Texture> = t.clone(); f.texture(&as_dyn)? }
            None => RT_NONE,
        };
        let has_uv = self.u_vec.is_some() && self.v_vec.is_some();
        f.remaps.push(rt_remap {
            tex_ori: self.tex_ori.e(), tex_u: self.tex_u.e(), tex_v: self.tex_v.e(),
            u_vec: self.u_vec.map(|u| u.as_inner().e()).unwrap_or([0.0; 3]), v_vec: self.v_vec.map(|v| v.as_inner().e()).unwrap_or([0.0; 3]),
            normal: [self.normal[0].e(), self.normal[1].e(), self.normal[2].e()], has_uv_vecs: has_uv as u32, normal_tex,
        });
        Ok(rt_material { kind: RT_MAT_REMAPPED, tex: RT_NONE, inner, inner2: f.remaps.len() as u32 - 1, ..Default::default() })
    }
'''),
    # ---------------------------------------------------------------- textures
    ("texture.rs", r"^impl Texture for SolidColor \{", '''    fn flatten(&self, _f: &mut crate::flatten::Flattener) -> Result<crate::ffi::rt_texture, crate::flatten::Unsupported> {
        Ok(crate::ffi::rt_texture { kind: crate::ffi::RT_TEX_SOLID, a: crate::ffi::RT_NONE, b: crate::ffi::RT_NONE, color: self.albedo.e(), ..Default::default() })
    }
'''),
    ("texture.rs", r"^impl Texture for CheckerTexture \{", '''    fn flatten(&self, f: &mut crate::flatten::Flattener) -> Result<crate::ffi::rt_texture, crate::flatten::Unsupported> {
        let (a, b) = (f.texture(&self.even)?, f.texture(&self.odd)?);  // children precede the checker
        Ok(crate::ffi::rt_texture { kind: crate::ffi::RT_TEX_CHECKER, a, b, scale: self.inv_scale, ..Default::default() })
    }
'''),
    ("texture.rs", r"^impl Texture for ImageTexture \{", '''    fn flatten(&self, f: &mut crate::flatten::Flattener) -> Result<crate::ffi::rt_texture, crate::flatten::Unsupported> {
        use crate::ffi::*;
        // a missing file stays "no image": the device returns cyan / alpha 1 like texture.rs:102-105,167-169
        let a = match self.image.rgba32f() {
            None => RT_NONE,
            Some((pixels, width, height, linear)) => {
                let mut flags = if linear { RT_IMG_LINEAR } else { 0 };  // raw, Hdr, OpenExr, Avif: no sRGB decode (image.rs:71-82)
                if matches!(self.interp, ImageInterpMethod::Linear) { flags |= RT_IMG_INTERP; }
                f.images.push(rt_image { width, height, flags, reserved: 0, texel_offset: f.texels.len() as u64 });
                f.texels.extend_from_slice(pixels);
                f.images.len() as u32 - 1
            }
        };
        Ok(rt_texture { kind: RT_TEX_IMAGE, a, b: RT_NONE, ..Default::default() })
    }
'''),
    ("texture.rs", r"^impl Texture for NoiseTexture \{", '''    fn flatten(&self, f: &mut crate::flatten::Flattener) -> Result<crate::ffi::rt_texture, crate::flatten::Unsupported> {
        f.perlins.push(self.noise.tables());
        Ok(crate::ffi::rt_texture { kind: crate::ffi::RT_TEX_NOISE, a: f.perlins.len() as u32 - 1, b: crate::ffi::RT_NONE, scale: self.scale, ..Default::default() })
    }
'''),
    # ---------------------------------------------------------------- small accessors
    ("utils/quaternion.rs", r"^impl Quaternion \{", '''    /// (w, x, y, z), the order of rt_transform.quat
    pub fn wxyz(&self) -> [f64; 4] {
        [self.w, self.x, self.y, self.z]
    }
'''),
    ("utils/perlin.rs", r"^impl Perlin \{", '''    /// the tables as the device reads them (rt_perlin)
    pub fn tables(&self) -> crate::ffi::rt_perlin {
        let mut t = crate::ffi::rt_perlin { randvec: [[0.0; 3]; 256], perm_x: [0; 256], perm_y: [0; 256], perm_z: [0; 256] };
        for i in 0..Perlin::POINT_COUNT {
            t.randvec[i] = self.randvec[i].as_inner().e();
            t.perm_x[i] = self.perm_x[i] as u32;
            t.perm_y[i] = self.perm_y[i] as u32;
            t.perm_z[i] = self.perm_z[i] as u32;
        }
        t
    }
'''),
    ("utils/image.rs", r"^impl Image \{", '''    /// decoded RGBA32F pixels (row-major), width, height, "already linear" (raw or Hdr / OpenExr / Avif, image.rs:71-82)
    pub fn rgba32f(&self) -> Option<(&[f32], u32, u32, bool)> {
        let (img, fmt) = self.img.as_ref()?;
        let linear = self.raw || matches!(fmt, ImageFormat::Hdr | ImageFormat::OpenExr | ImageFormat::Avif);
        Some((img.as_raw().as_slice(), img.width(), img.height(), linear))
    }
'''),
]

# (file, regex, replacement): closures that become data, the BVH input order
REPLACE = [
    ("camera.rs", r"    pub fn render\(&mut self, world: &dyn Hittable, lights: Option<&dyn Hittable>\) -> RgbImage \{", "    pub fn render_cpu(&mut self, world: &dyn Hittable, lights: Option<&dyn Hittable>) -> RgbImage {"),
    ("camera.rs", r"(\nimpl Default for Camera \{)", "\nmod render_gpu;\n\\1"),
    ("utils/vec3.rs", r"(    pub fn x\(&self\) -> f64 \{)", "    /// the three components, for the flattener\n    pub fn e(&self) -> [f64; 3] {\n        self.e\n    }\n\n\\1"),
    ("material.rs", r"(pub struct Mix \{\n    mat1: Arc<dyn Material>,\n    mat2: Arc<dyn Material>,\n    ratio: RatioFn,\n)",
     "\\1    ratio_src: MixRatio,\n"),
    ("material.rs", r"(type RatioFn = [^\n]*\n)", "\\1\n/// what `ratio` was built from, kept for the flattener\nenum MixRatio {\n    Constant(f64),\n    Alpha(Arc<ImageTexture>),\n}\n"),
    ("material.rs", r"(            ratio: Box::new\(move \|_, _, _\| ratio\),\n)", "\\1            ratio_src: MixRatio::Constant(ratio),\n"),
    ("material.rs", r"(    \) -> Mix \{\n        Mix \{\n            mat1,\n            mat2,\n)", "\\1            ratio_src: MixRatio::Alpha(tex.clone()),\n"),
    ("material/portal.rs", r"(pub struct Portal \{\n    pub attenuation: Color,\n)", "\\1    position_offset: Vec3,\n    rotation: Quaternion,\n"),
    ("material/portal.rs", r"(        Portal \{\n            attenuation,\n)", "\\1            position_offset,\n            rotation,\n"),
    ("material/disney.rs", r"(pub struct Disney \{\n    pub param_fn: DisneyParamFn,\n)",
     "\\1    /// the constants (and base-colour texture) `param_fn` was built from; None for a hand-written closure\n    pub flat: Option<(DisneyParameters, Option<std::sync::Arc<dyn crate::texture::Texture>>)>,\n"),
    ("material/disney.rs", r"(            param_fn: Box::new\(\|_, _, _\| DisneyParameters::default\(\)\),\n)", "\\1            flat: Some((DisneyParameters::default(), None)),\n"),
    ("material/disney.rs", r"(        let params = self\.params;\n        Disney \{\n)", "\\1            flat: Some((params.clone(), None)),\n"),
    ("shapes/obj.rs", r"(            Arc::new\(Disney \{\n)",
     "\\1                flat: Some((DisneyParameters { roughness, anisotropic, sheen, clearcoat, clearcoat_gloss, metallic, ior, spec_trans, ..Default::default() },\n                            Some(base_color.clone()))),\n"),
    ("bvh.rs", r"(pub struct BVH \{\n)", "\\1    /// Some(..) on the node a `from_vec` call returns: addresses of the boxed children in the order they were given\n    input_order: Option<Vec<usize>>,\n"),
    ("bvh.rs", r"    pub fn from_vec\(mut objects: Vec<Box<dyn Hittable>>\) -> BVH \{\n",
     "    pub fn from_vec(objects: Vec<Box<dyn Hittable>>) -> BVH {\n        let order = objects.iter().map(|o| o.as_ref() as *const dyn Hittable as *const () as usize).collect();\n"
     "        let mut root = BVH::build(objects);\n        root.input_order = Some(order);\n        root\n    }\n\n    fn build(mut objects: Vec<Box<dyn Hittable>>) -> BVH {\n"),
]


def main():
    src = os.path.join(REF, "src")
    if not os.path.isdir(src):
        raise SystemExit(f"{src} not found: the patch is generated against the reference checkout")
    tmp = tempfile.mkdtemp()
    a, b = os.path.join(tmp, "a", "src"), os.path.join(tmp, "b", "src")
    shutil.copytree(src, a)
    shutil.copytree(src, b)
    for rel, rx, text in INSERT:
        if text is None:
            continue
        p = os.path.join(b, rel)
        lines = open(p, encoding="utf-8").read().split("\n")
        hits = [i for i, l in enumerate(lines) if re.search(rx, l)]
        assert len(hits) == 1, (rel, rx, hits)
        lines[hits[0] + 1:hits[0] + 1] = text.rstrip("\n").split("\n")
        open(p, "w", encoding="utf-8").write("\n".join(lines))
    for rel, rx, repl in REPLACE:
        p = os.path.join(b, rel)
        s = open(p, encoding="utf-8").read()
        s2, n = re.subn(rx, repl, s, count=1)
        assert n == 1, (rel, rx)
        open(p, "w", encoding="utf-8").write(s2)
    # recursion inside from_vec now goes through the private builder; inner nodes carry no input order
    p = os.path.join(b, "bvh.rs")
    s = open(p, encoding="utf-8").read()
    body = s.split("    fn build(mut objects", 1)
    body[1] = body[1].replace("BVH::from_vec(", "BVH::build(").replace("BVH { left, right, bbox }", "BVH { input_order: None, left, right, bbox }")
    s = "    fn build(mut objects".join(body)
    open(p, "w", encoding="utf-8").write(s)
    # `as_bvh_inner`: lets a BVH recognise its own inner nodes behind `dyn Hittable`
    p = os.path.join(b, "hit.rs")
    s = open(p, encoding="utf-8").read()
    s = s.replace("    #[doc(hidden)]\n    fn flatten(", "    #[doc(hidden)]\n    fn as_bvh_inner(&self) -> Option<&crate::bvh::BVH> {\n        None\n    }\n    #[doc(hidden)]\n    fn flatten(", 1)
    open(p, "w", encoding="utf-8").write(s)
    p = os.path.join(b, "bvh.rs")
    s = open(p, encoding="utf-8").read()
    s = s.replace("impl Hittable for BVH {\n", "impl Hittable for BVH {\n    fn as_bvh_inner(&self) -> Option<&BVH> {\n        if self.input_order.is_none() { Some(self) } else { None }\n    }\n", 1)
    open(p, "w", encoding="utf-8").write(s)
    for new in ("ffi.rs", "flatten.rs"):
        shutil.copy(os.path.join(HERE, "src", new), os.path.join(b, new))
    os.makedirs(os.path.join(b, "camera"), exist_ok=True)  # a child module of `camera`: it reads Camera's private fields
    shutil.copy(os.path.join(HERE, "src", "camera_render.rs"), os.path.join(b, "camera", "render_gpu.rs"))
    out = subprocess.run(["diff", "-U1", "-r", "-N", "a/src", "b/src"], cwd=tmp, capture_output=True, text=True).stdout
    out = re.sub(r"^(---|\+\+\+) (\S+)\t.*$", r"\1 \2", out, flags=re.M)  # no timestamps: the patch is reproducible
    dst = os.path.join(HERE, "reference_src.patch")
    open(dst, "w", encoding="utf-8").write(out)
    print(f"{dst}: {len(out.splitlines())} lines, {out.count('+++ ')} files")
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()

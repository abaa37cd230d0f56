//! Flattener: accumulates the arrays of `rt_scene_desc`.  Shared `Arc<dyn Material>` /
//! `Arc<dyn Texture>` are emitted once (keyed by pointer).  Children are always pushed before their
//! parent, which the library's validator requires.
use crate::ffi::*;
use std::collections::HashMap;

#[derive(Debug)]
pub struct Unsupported(pub &'static str);

#[derive(Default)]
pub struct Flattener {
    pub objects: Vec<rt_object>, pub children: Vec<u32>, pub spheres: Vec<rt_sphere>, pub planars: Vec<rt_planar>,
    pub transforms: Vec<rt_transform>, pub media: Vec<rt_medium>, pub materials: Vec<rt_material>,
    pub textures: Vec<rt_texture>, pub images: Vec<rt_image>, pub texels: Vec<f32>, pub perlins: Vec<rt_perlin>, pub remaps: Vec<rt_remap>,
    tex_ids: HashMap<usize, u32>, mat_ids: HashMap<usize, u32>,
}

impl Flattener {
    pub fn push_object(&mut self, kind: u32, material: u32, data: u32, bbox: &crate::aabb::AABB, kids: &[u32]) -> u32 {
        let o = rt_object {
            kind, material, data, reserved: 0,
            first_child: self.children.len() as u32, child_count: kids.len() as u32,
            bbox: [*bbox.x().min(), *bbox.x().max(), *bbox.y().min(), *bbox.y().max(), *bbox.z().min(), *bbox.z().max()],
        };
        self.children.extend_from_slice(kids);
        self.objects.push(o);
        self.objects.len() as u32 - 1
    }
    /// `Arc` identity -> one table entry
    pub fn texture(&mut self, t: &std::sync::Arc<dyn crate::texture::Texture>) -> Result<u32, Unsupported> {
        let key = std::sync::Arc::as_ptr(t) as *const () as usize;
        if let Some(&id) = self.tex_ids.get(&key) { return Ok(id); }
        let rec = t.flatten(self)?;
        self.textures.push(rec);
        let id = self.textures.len() as u32 - 1;
        self.tex_ids.insert(key, id);
        Ok(id)
    }
    pub fn material(&mut self, m: &std::sync::Arc<dyn crate::material::Material>) -> Result<u32, Unsupported> {
        let key = std::sync::Arc::as_ptr(m) as *const () as usize;
        if let Some(&id) = self.mat_ids.get(&key) { return Ok(id); }
        let rec = m.flatten(self)?;
        self.materials.push(rec);
        let id = self.materials.len() as u32 - 1;
        self.mat_ids.insert(key, id);
        Ok(id)
    }
    pub fn desc(&self, world_root: u32, lights_root: u32) -> rt_scene_desc {
        rt_scene_desc {
            version: RT_ABI_VERSION, struct_size: std::mem::size_of::<rt_scene_desc>() as u32, world_root, lights_root,
            n_objects: self.objects.len() as u32, n_children: self.children.len() as u32, n_spheres: self.spheres.len() as u32,
            n_planars: self.planars.len() as u32, n_transforms: self.transforms.len() as u32, n_media: self.media.len() as u32,
            n_materials: self.materials.len() as u32, n_textures: self.textures.len() as u32, n_images: self.images.len() as u32,
            n_perlins: self.perlins.len() as u32, n_texels: self.texels.len() as u64, n_remaps: self.remaps.len() as u32, reserved0: 0,
            objects: self.objects.as_ptr(), children: self.children.as_ptr(), spheres: self.spheres.as_ptr(), planars: self.planars.as_ptr(),
            transforms: self.transforms.as_ptr(), media: self.media.as_ptr(), materials: self.materials.as_ptr(),
            textures: self.textures.as_ptr(), images: self.images.as_ptr(), texels: self.texels.as_ptr(), perlins: self.perlins.as_ptr(), remaps: self.remaps.as_ptr(),
        }
    }
}

// ---- bodies to paste into the per-type impl blocks (fields are module-private) -------------------
//
// impl Hittable for Sphere (src/shapes/sphere.rs):
//     fn flatten(&self, f: &mut Flattener) -> Result<u32, Unsupported> {
//         let mat = f.material(&self.mat)?;
//         f.spheres.push(rt_sphere { center: self.center.origin().e(), center_vec: self.center.direction().e(),
//                                    radius: self.radius, reserved: 0.0 });
//         Ok(f.push_object(RT_OBJ_SPHERE, mat, f.spheres.len() as u32 - 1, &self.bbox, &[]))
//     }
// impl Hittable for Quad / Triangle (src/shapes/quad.rs, triangle.rs): see INTEGRATION.md §3 (RT_OBJ_QUAD / RT_OBJ_TRIANGLE).
// impl Hittable for Hittables (src/hits.rs):
//         let kids = self.objects.iter().map(|o| o.flatten(f)).collect::<Result<Vec<_>, _>>()?;
//         Ok(f.push_object(RT_OBJ_LIST, RT_NONE, RT_NONE, &self.bbox, &kids))
// impl Hittable for BVH (src/bvh.rs): BVH::from_vec consumes its input, so BVH keeps a copy of the flattened ids of
//     its original children (`Vec<Box<dyn Hittable>>` order) to emit RT_OBJ_BVH; the library recomputes the split.
// impl Hittable for Transform (src/shapes.rs):
//         let kid = self.object.flatten(f)?;
//         f.transforms.push(rt_transform { offset: self.offset.e(), quat: [q.w, q.x, q.y, q.z], scale: self.scale.e() });
//         Ok(f.push_object(RT_OBJ_TRANSFORM, RT_NONE, f.transforms.len() as u32 - 1, &self.bbox, &[kid]))
// impl Hittable for ConstantMedium (src/volume.rs):
//         let kid = self.boundary.flatten(f)?;
//         let mat = /* Isotropic */ self.phase_function.flatten_boxed(f)?;
//         f.media.push(rt_medium { neg_inv_density: self.neg_inv_density, reserved: 0.0 });
//         Ok(f.push_object(RT_OBJ_MEDIUM, mat, f.media.len() as u32 - 1, self.boundary.bounding_box(), &[kid]))
// impl Material for Lambertian / Metal / Dielectric / DiffuseLight / Isotropic / Transparent / Mix / Portal:
//         one rt_material each (kind, tex, inner, inner2, color, param, v) exactly as host/rt2025.hpp does.
// impl Material for Disney: the `param_fn` closure is only ever built from constants (DisneyBuilder::build) or constants +
//         a base-colour texture (obj.rs:271-293), so Disney stores those next to the closure and emits RT_MAT_DISNEY.
// impl Material for RemappedMaterial (obj.rs): one rt_remap (tex_ori, tex_u, tex_v, u_vec, v_vec, normal[3], normal_tex) + RT_MAT_REMAPPED.
// impl Texture for SolidColor / CheckerTexture / ImageTexture / NoiseTexture: one rt_texture each; ImageTexture copies
//         its Rgba32F pixels into f.texels and sets RT_IMG_LINEAR for raw / Hdr / OpenExr / Avif, RT_IMG_INTERP for
//         ImageInterpMethod::Linear; NoiseTexture copies its Perlin tables into f.perlins.

// host_capi.cpp — flat C entry points over the C++ host mirror so that Python (ctypes) tests
// and bench.py can build scenes the way src/main.rs does: named benchmark scenes, plus a
// builder whose calls map 1:1 onto the reference constructors.
// Nothing here computes radiance or intersections; it only produces rt_scene_desc/rt_camera.
#include <cstdio>
#include <cstring>
#include <string>

#include "scenes.hpp"

using namespace rt2025;

static Vec3 V(const double* p) { return Vec3(p[0], p[1], p[2]); }
template <class T, class U>
static uint32_t push(std::vector<T>& v, U p) {
    v.push_back(std::move(p));
    return (uint32_t)v.size() - 1;
}

extern "C" {

struct rth_scene {
    BuiltScene s;
};

struct rth_builder {
    std::vector<TexturePtr> tex;
    std::vector<MaterialPtr> mat;
    std::vector<HittablePtr> obj;
    std::string error;
};

static thread_local std::string g_err;
const char* rth_last_error() { return g_err.c_str(); }

// params: book2_final/cornell/book1: [image_width, spp, max_depth]; soups: [n]
rth_scene* rth_scene_named(const char* name, uint64_t seed, const double* params, int n_params) {
    try {
        auto* sc = new rth_scene();
        Random::seed(seed);
        std::string n(name);
        auto P = [&](int i, double dflt) { return i < n_params ? params[i] : dflt; };
        if (n == "book2_final")
            final_scene(sc->s, (uint32_t)P(0, 800), (size_t)P(1, 1000), (uint32_t)P(2, 40));
        else if (n == "cornell_glass")
            cornell_box(sc->s, false, (uint32_t)P(0, 600), (size_t)P(1, 1000), (uint32_t)P(2, 50));
        else if (n == "cornell_shipped")
            cornell_box(sc->s, true, (uint32_t)P(0, 1080), (size_t)P(1, 100), (uint32_t)P(2, 10));
        else if (n == "book1_final")
            book1_final(sc->s, (uint32_t)P(0, 1200), (size_t)P(1, 10), (uint32_t)P(2, 50));
        else if (n == "tri_soup")
            soup(sc->s, seed, (uint64_t)P(0, 1000000), false);
        else if (n == "sphere_soup")
            soup(sc->s, seed, (uint64_t)P(0, 1000000), true);
        else {
            delete sc;
            g_err = "unknown scene: " + n;
            return nullptr;
        }
        return sc;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
const rt_scene_desc* rth_scene_desc(const rth_scene* s) { return &s->s.desc; }
const rt_camera* rth_scene_camera(const rth_scene* s) { return &s->s.camera; }
void rth_scene_free(rth_scene* s) { delete s; }

// ---- builder -----------------------------------------------------------------------------
rth_builder* rth_builder_new(uint64_t seed) {
    Random::seed(seed);
    return new rth_builder();
}
void rth_builder_free(rth_builder* b) { delete b; }

#define GUARD(expr)                        \
    try {                                  \
        expr                               \
    } catch (const std::exception& e) {    \
        g_err = e.what();                  \
        return RT_NONE;                    \
    }

uint32_t rth_tex_solid(rth_builder* b, double r, double g, double bl) {
    return push(b->tex, std::make_shared<SolidColor>(Color(r, g, bl)));
}
uint32_t rth_tex_checker(rth_builder* b, double scale, uint32_t even, uint32_t odd) {
    return push(b->tex, std::make_shared<CheckerTexture>(scale, b->tex.at(even), b->tex.at(odd)));
}
uint32_t rth_tex_noise(rth_builder* b, double scale) { return push(b->tex, std::make_shared<NoiseTexture>(scale)); }
uint32_t rth_tex_image_missing(rth_builder* b) { return push(b->tex, std::make_shared<ImageTexture>("missing")); }
// rgba: width*height*4 floats in [0,1] as decoded; raw != 0 -> ImageTexture::new_raw_image
uint32_t rth_tex_image(rth_builder* b, uint32_t width, uint32_t height, const float* rgba, int raw, int linear_format) {
    Image im;
    im.width = width;
    im.height = height;
    im.rgba.assign(rgba, rgba + (size_t)width * height * 4);
    im.linear_format = linear_format != 0;
    return push(b->tex, ImageTexture::from_pixels(std::move(im), raw != 0));
}
uint32_t rth_tex_gradient(rth_builder* b, const double* bottom, const double* top) {
    return push(b->tex, std::make_shared<GradientTexture>(V(bottom), V(top)));
}

uint32_t rth_mat_empty(rth_builder* b) { return push(b->mat, std::make_shared<EmptyMaterial>()); }
uint32_t rth_mat_lambertian(rth_builder* b, uint32_t tex) { return push(b->mat, std::make_shared<Lambertian>(b->tex.at(tex))); }
uint32_t rth_mat_metal(rth_builder* b, const double* albedo, double fuzz) {
    return push(b->mat, std::make_shared<Metal>(V(albedo), fuzz));
}
uint32_t rth_mat_dielectric(rth_builder* b, uint32_t tex, double ri) {
    return push(b->mat, std::make_shared<Dielectric>(b->tex.at(tex), ri));
}
uint32_t rth_mat_diffuse_light(rth_builder* b, uint32_t tex, uint32_t inner) {
    if (inner == RT_NONE) return push(b->mat, std::make_shared<DiffuseLight>(b->tex.at(tex)));
    return push(b->mat, DiffuseLight::new_with_material(b->tex.at(tex), b->mat.at(inner)));
}
uint32_t rth_mat_isotropic(rth_builder* b, uint32_t tex) { return push(b->mat, std::make_shared<Isotropic>(b->tex.at(tex))); }
uint32_t rth_mat_transparent(rth_builder* b) { return push(b->mat, std::make_shared<Transparent>()); }
uint32_t rth_mat_mix(rth_builder* b, uint32_t m1, uint32_t m2, double ratio) {
    return push(b->mat, std::make_shared<Mix>(b->mat.at(m1), b->mat.at(m2), ratio));
}
// Mix::from_image(mat1, mat2, tex): the ratio is the image's alpha (material.rs:236-247)
uint32_t rth_mat_mix_image(rth_builder* b, uint32_t m1, uint32_t m2, uint32_t tex) {
    auto it = std::dynamic_pointer_cast<const ImageTexture>(b->tex.at(tex));
    if (!it) {
        g_err = "Mix::from_image needs an ImageTexture";
        return RT_NONE;
    }
    return push(b->mat, Mix::from_image(b->mat.at(m1), b->mat.at(m2), it));
}
uint32_t rth_mat_portal(rth_builder* b, const double* att, const double* offset, const double* quat_wxyz) {
    Quaternion q{quat_wxyz[0], quat_wxyz[1], quat_wxyz[2], quat_wxyz[3]};
    return push(b->mat, std::make_shared<Portal>(V(att), V(offset), q));
}

// params: the 13 DisneyParameters scalars in RT_DISNEY_* order; tex = base-colour texture or RT_NONE
uint32_t rth_mat_disney(rth_builder* b, const double* base_color, uint32_t tex, const double* params) {
    DisneyParameters p;
    p.base_color = V(base_color);
    p.roughness = params[RT_DISNEY_ROUGHNESS], p.anisotropic = params[RT_DISNEY_ANISOTROPIC], p.sheen = params[RT_DISNEY_SHEEN];
    p.sheen_tint = params[RT_DISNEY_SHEEN_TINT], p.clearcoat = params[RT_DISNEY_CLEARCOAT], p.clearcoat_gloss = params[RT_DISNEY_CLEARCOAT_GLOSS];
    p.specular_tint = params[RT_DISNEY_SPECULAR_TINT], p.metallic = params[RT_DISNEY_METALLIC], p.ior = params[RT_DISNEY_IOR];
    p.flatness = params[RT_DISNEY_FLATNESS], p.spec_trans = params[RT_DISNEY_SPEC_TRANS], p.diff_trans = params[RT_DISNEY_DIFF_TRANS];
    p.thin = params[RT_DISNEY_THIN] != 0.0;
    return push(b->mat, std::make_shared<Disney>(p, tex == RT_NONE ? nullptr : b->tex.at(tex)));
}
// one face of an OBJ model as load_object builds it (obj.rs:143-183): pos/uv/nrm are 3x3 doubles
// (uv rows are (u, v, 0)); returns the RemappedMaterial handle
uint32_t rth_mat_remapped(rth_builder* b, uint32_t inner, const double* pos, const double* uv, const double* nrm, uint32_t normal_tex) {
    std::shared_ptr<const ImageTexture> nt;
    if (normal_tex != RT_NONE) nt = std::dynamic_pointer_cast<const ImageTexture>(b->tex.at(normal_tex));
    return push(b->mat, std::make_shared<RemappedMaterial>(b->mat.at(inner), V(pos), V(pos + 3), V(pos + 6), V(uv), V(uv + 3), V(uv + 6),
                                                           V(nrm), V(nrm + 3), V(nrm + 6), nt));
}

uint32_t rth_sphere(rth_builder* b, const double* c, double r, uint32_t mat) {
    return push(b->obj, std::make_shared<Sphere>(V(c), r, b->mat.at(mat)));
}
uint32_t rth_sphere_moving(rth_builder* b, const double* c1, const double* c2, double r, uint32_t mat) {
    return push(b->obj, Sphere::new_with_motion(V(c1), V(c2), r, b->mat.at(mat)));
}
uint32_t rth_quad(rth_builder* b, const double* anchor, const double* u, const double* v, uint32_t mat) {
    GUARD(return push(b->obj, std::make_shared<Quad>(V(anchor), V(u), V(v), b->mat.at(mat)));)
}
uint32_t rth_triangle(rth_builder* b, const double* anchor, const double* u, const double* v, uint32_t mat) {
    auto t = Triangle::create(V(anchor), V(u), V(v), b->mat.at(mat));
    if (!t) return RT_NONE;  // Triangle::new -> None
    return push(b->obj, t);
}
uint32_t rth_box(rth_builder* b, const double* p0, const double* p1, uint32_t mat) {
    GUARD(return push(b->obj, build_box(V(p0), V(p1), b->mat.at(mat)));)
}
// use_new != 0: Hittables::new(first) then add(rest); else Hittables::default() then add(all)
uint32_t rth_list(rth_builder* b, const uint32_t* ids, uint32_t n, int use_new) {
    std::shared_ptr<Hittables> l;
    uint32_t i = 0;
    if (use_new && n > 0) {
        l = std::make_shared<Hittables>(b->obj.at(ids[0]));
        i = 1;
    } else
        l = std::make_shared<Hittables>();
    for (; i < n; i++) l->add(b->obj.at(ids[i]));
    return push(b->obj, l);
}
uint32_t rth_bvh(rth_builder* b, const uint32_t* ids, uint32_t n) {
    std::vector<HittablePtr> v;
    for (uint32_t i = 0; i < n; i++) v.push_back(b->obj.at(ids[i]));
    GUARD(return push(b->obj, std::make_shared<BVH>(std::move(v)));)
}
uint32_t rth_transform(rth_builder* b, uint32_t child, const double* offset, const double* quat_wxyz, const double* scale) {
    std::optional<Vec3> o, s;
    std::optional<Quaternion> q;
    if (offset) o = V(offset);
    if (scale) s = V(scale);
    if (quat_wxyz) q = Quaternion{quat_wxyz[0], quat_wxyz[1], quat_wxyz[2], quat_wxyz[3]};
    return push(b->obj, std::make_shared<Transform>(b->obj.at(child), o, q, s));
}
void rth_quat_from_axis_angle(const double* axis, double degrees, double* out_wxyz) {
    Quaternion q = Quaternion::from_axis_angle(V(axis), degrees);
    out_wxyz[0] = q.w, out_wxyz[1] = q.x, out_wxyz[2] = q.y, out_wxyz[3] = q.z;
}
uint32_t rth_medium(rth_builder* b, uint32_t boundary, double density, uint32_t tex) {
    return push(b->obj, ConstantMedium::new_with_tex(b->obj.at(boundary), density, b->tex.at(tex)));
}

struct rth_camera_params {
    double aspect_ratio;
    uint32_t image_width;
    uint32_t samples_per_pixel;
    uint32_t max_depth;
    uint32_t background_tex;  // builder texture handle or RT_NONE for the default black
    double vertical_fov_in_degrees;
    double look_from[3], look_at[3], vec_up[3];
    double defocus_angle_in_degrees;
    double focus_distance;
    uint32_t toon_map;
    uint32_t reserved;
};

rth_scene* rth_builder_finish(rth_builder* b, uint32_t world, uint32_t lights, const rth_camera_params* cp) {
    try {
        Camera cam;
        cam.aspect_ratio = cp->aspect_ratio;
        cam.image_width = cp->image_width;
        cam.samples_per_pixel = cp->samples_per_pixel;
        cam.max_depth = cp->max_depth;
        if (cp->background_tex != RT_NONE) cam.background.texture = b->tex.at(cp->background_tex);
        cam.vertical_fov_in_degrees = cp->vertical_fov_in_degrees;
        cam.look_from = V(cp->look_from);
        cam.look_at = V(cp->look_at);
        cam.vec_up = V(cp->vec_up);
        cam.defocus_angle_in_degrees = cp->defocus_angle_in_degrees;
        cam.focus_distance = cp->focus_distance;
        cam.toon_map = cp->toon_map ? ToonMap::ACES : ToonMap::None;
        auto* sc = new rth_scene();
        sc->s.finish(cam, *b->obj.at(world), lights == RT_NONE ? nullptr : b->obj.at(lights).get());
        return sc;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}


}  // extern "C"

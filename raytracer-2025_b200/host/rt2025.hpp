// rt2025.hpp — C++ host-side mirror of the reference crate's public API.
//
// The reference is a Rust crate; no Rust toolchain exists in this image, so the host side
// above the C ABI (include/rt2025.h) is written in C++ with the reference's names and
// argument meaning: Hittable / Material / Texture objects are built exactly like in
// src/main.rs and `Camera::render(world, lights)` is the entry point (camera.rs:161).
// There is deliberately NO hit()/scatter() here: the product has no CPU path.  Every object
// only knows how to (a) report the bounding box the reference would compute and (b) flatten
// itself into the plain-old-data object graph of rt2025.h.
//
// All arithmetic that feeds the flat description (bounding boxes, plane constants, camera
// frame) follows the reference operation by operation and this file must be compiled with
// -ffp-contract=off so the numbers are the ones the Rust code would produce.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "rt2025.h"

namespace rt2025 {

constexpr double PI = 3.14159265358979323846264338327950288;
constexpr double INF = std::numeric_limits<double>::infinity();

// ---- utils/vec3.rs ---------------------------------------------------------------------
struct Vec3 {
    double e[3]{0, 0, 0};
    constexpr Vec3() = default;
    constexpr Vec3(double x, double y, double z) : e{x, y, z} {}
    double x() const { return e[0]; }
    double y() const { return e[1]; }
    double z() const { return e[2]; }
    double operator[](int i) const { return e[i]; }
    double& operator[](int i) { return e[i]; }
    // vec3.rs:96-98
    double length_squared() const { return e[0] * e[0] + e[1] * e[1] + e[2] * e[2]; }
    double length() const { return std::sqrt(length_squared()); }
    // vec3.rs:107-109
    double dot(const Vec3& r) const { return e[0] * r[0] + e[1] * r[1] + e[2] * r[2]; }
    // vec3.rs:111-117
    Vec3 cross(const Vec3& r) const {
        return Vec3(e[1] * r[2] - e[2] * r[1], e[2] * r[0] - e[0] * r[2], e[0] * r[1] - e[1] * r[0]);
    }
    static Vec3 ZERO() { return Vec3(0, 0, 0); }
};
using Point3 = Vec3;
using Color = Vec3;
inline Vec3 operator+(const Vec3& a, const Vec3& b) { return Vec3(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline Vec3 operator-(const Vec3& a, const Vec3& b) { return Vec3(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
inline Vec3 operator*(const Vec3& a, const Vec3& b) { return Vec3(a[0] * b[0], a[1] * b[1], a[2] * b[2]); }
inline Vec3 operator/(const Vec3& a, const Vec3& b) { return Vec3(a[0] / b[0], a[1] / b[1], a[2] / b[2]); }
inline Vec3 operator-(const Vec3& a) { return Vec3(-a[0], -a[1], -a[2]); }
inline Vec3 operator*(double s, const Vec3& a) { return Vec3(s * a[0], s * a[1], s * a[2]); }
inline Vec3 operator*(const Vec3& a, double s) { return Vec3(a[0] * s, a[1] * s, a[2] * s); }
// vec3.rs:222-236: Vec3 / f64 multiplies by the reciprocal
inline Vec3 operator/(const Vec3& a, double s) { return (1.0 / s) * a; }

// UnitVec3::from_vec3 (vec3.rs:303-310): None when the normalised vector is not finite
inline std::optional<Vec3> unit_vector(const Vec3& v) {
    Vec3 r = v / v.length();
    if (std::isfinite(r[0]) && std::isfinite(r[1]) && std::isfinite(r[2])) return r;
    return std::nullopt;
}

// Rust f64::min / f64::max ignore a NaN operand
inline double rmin(double a, double b) { return std::isnan(a) ? b : (std::isnan(b) ? a : (a < b ? a : b)); }
inline double rmax(double a, double b) { return std::isnan(a) ? b : (std::isnan(b) ? a : (a > b ? a : b)); }

// ---- utils/interval.rs -----------------------------------------------------------------
struct Interval {
    double min = 0.0, max = 0.0;  // #[derive(Default)] -> {0,0}
    Interval() = default;
    Interval(double a, double b) : min(rmin(a, b)), max(rmax(a, b)) {}  // interval.rs:10-15
    static Interval raw(double mn, double mx) {
        Interval i;
        i.min = mn;
        i.max = mx;
        return i;
    }
    static Interval EMPTY() { return raw(INF, -INF); }
    double size() const { return rmax(max - min, 0.0); }  // :42-44
    Interval expand(double delta) const {                 // :29-35
        double padding = delta / 2.0;
        return raw(min - padding, max + padding);
    }
    static Interval union_(const Interval& a, const Interval& b) {  // :58-63
        return raw(rmin(a.min, b.min), rmax(a.max, b.max));
    }
};

// ---- aabb.rs ---------------------------------------------------------------------------
struct AABB {
    Interval x, y, z;  // #[derive(Default)] -> the degenerate box at the origin
    static AABB EMPTY() { return AABB{Interval::EMPTY(), Interval::EMPTY(), Interval::EMPTY()}; }
    AABB pad_to_minimums() const {  // aabb.rs:43-51
        const double DELTA = 0.0001;
        auto f = [&](const Interval& t) { return t.size() < DELTA ? t.expand(DELTA) : t; };
        return AABB{f(x), f(y), f(z)};
    }
    static AABB from_points(const Point3& a, const Point3& b) {  // :21-28
        return AABB{Interval(a[0], b[0]), Interval(a[1], b[1]), Interval(a[2], b[2])}.pad_to_minimums();
    }
    AABB union_(const AABB& r) const {  // :94-100
        return AABB{Interval::union_(x, r.x), Interval::union_(y, r.y), Interval::union_(z, r.z)};
    }
    const Interval& axis_interval(int n) const { return n == 0 ? x : (n == 1 ? y : z); }
    void all_points(Point3 out[8]) const {  // :30-41
        out[0] = Point3(x.min, y.min, z.min);
        out[1] = Point3(x.min, y.min, z.max);
        out[2] = Point3(x.min, y.max, z.min);
        out[3] = Point3(x.min, y.max, z.max);
        out[4] = Point3(x.max, y.min, z.min);
        out[5] = Point3(x.max, y.min, z.max);
        out[6] = Point3(x.max, y.max, z.min);
        out[7] = Point3(x.max, y.max, z.max);
    }
};

// ---- utils/quaternion.rs ---------------------------------------------------------------
struct Quaternion {
    double w = 1, x = 0, y = 0, z = 0;
    static Quaternion identity() { return Quaternion{1, 0, 0, 0}; }
    static Quaternion from_euler(double yaw, double pitch, double roll) {  // :23-37
        double cy = std::cos(0.5 * yaw), sy = std::sin(0.5 * yaw);
        double cp = std::cos(0.5 * pitch), sp = std::sin(0.5 * pitch);
        double cr = std::cos(0.5 * roll), sr = std::sin(0.5 * roll);
        return Quaternion{cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy,
                          cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy};
    }
    static Quaternion from_axis_angle(const Vec3& axis, double angle_in_degrees) {  // :39-51
        double half = (angle_in_degrees * (PI / 180.0)) * 0.5;
        double s = std::sin(half), c = std::cos(half);
        auto a = unit_vector(axis);
        if (!a) throw std::runtime_error("Quaternion::from_axis_angle: axis not normalisable");
        return Quaternion{c, (*a)[0] * s, (*a)[1] * s, (*a)[2] * s};
    }
    Quaternion conjugate() const { return Quaternion{w, -x, -y, -z}; }
    Quaternion mul(const Quaternion& r) const {  // :94-104
        return Quaternion{w * r.w - x * r.x - y * r.y - z * r.z, w * r.x + x * r.w + y * r.z - z * r.y,
                          w * r.y - x * r.z + y * r.w + z * r.x, w * r.z + x * r.y - y * r.x + z * r.w};
    }
    Vec3 rotate_vector(const Vec3& v) const {  // :72-82
        Quaternion qv{0.0, v[0], v[1], v[2]};
        Quaternion r = this->mul(qv).mul(conjugate());
        return Vec3(r.x, r.y, r.z);
    }
};

// ---- seeded stand-in for utils/random.rs (scene construction only) -----------------------
// The reference draws scene content from the unseeded thread RNG (main.rs:397,498;
// perlin.rs:18,104), so no two reference runs build the same scene.  Scenes here come from
// one stated 64-bit seed: splitmix64 -> xoshiro256++.
class Random {
  public:
    static void seed(uint64_t s) {
        for (auto& v : state()) {
            s += 0x9E3779B97F4A7C15ull;
            uint64_t z = s;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            v = z ^ (z >> 31);
        }
    }
    static uint64_t u64() {
        auto& s = state();
        auto rotl = [](uint64_t x, int k) { return (x << k) | (x >> (64 - k)); };
        uint64_t result = rotl(s[0] + s[3], 23) + s[0];
        uint64_t t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return result;
    }
    static double f64() { return (double)(u64() >> 11) * 0x1.0p-53; }
    static double random_range(double lo, double hi) { return lo + (hi - lo) * f64(); }
    static size_t usize_inclusive(size_t lo, size_t hi) { return lo + (size_t)(u64() % (uint64_t)(hi - lo + 1)); }

  private:
    static uint64_t (&state())[4] {
        static uint64_t s[4] = {1, 2, 3, 4};
        return s;
    }
};
inline Vec3 random_vec3_range(double lo, double hi) {  // vec3.rs:53-61
    double a = Random::random_range(lo, hi), b = Random::random_range(lo, hi), c = Random::random_range(lo, hi);
    return Vec3(a, b, c);
}

// ---- flattening ------------------------------------------------------------------------
class Texture;
class Material;
class Hittable;

// Accumulates the arrays of rt_scene_desc; shared textures / materials (Arc in the
// reference) are emitted once.
class Flattener {
  public:
    std::vector<rt_object> objects;
    std::vector<uint32_t> children;
    std::vector<rt_sphere> spheres;
    std::vector<rt_planar> planars;
    std::vector<rt_transform> transforms;
    std::vector<rt_medium> media;
    std::vector<rt_material> materials;
    std::vector<rt_texture> textures;
    std::vector<rt_image> images;
    std::vector<float> texels;
    std::vector<rt_perlin> perlins;
    std::vector<rt_remap> remaps;
    uint32_t world_root = RT_NONE, lights_root = RT_NONE;

    uint32_t texture(const std::shared_ptr<const Texture>& t);
    uint32_t material(const std::shared_ptr<const Material>& m);
    uint32_t object(const Hittable& h);

    uint32_t push_object(uint32_t kind, uint32_t material, uint32_t data, const AABB& box,
                         const std::vector<uint32_t>& kids) {
        rt_object o{};
        o.kind = kind;
        o.material = material;
        o.data = data;
        o.first_child = (uint32_t)children.size();
        o.child_count = (uint32_t)kids.size();
        children.insert(children.end(), kids.begin(), kids.end());
        o.bbox[0] = box.x.min, o.bbox[1] = box.x.max;
        o.bbox[2] = box.y.min, o.bbox[3] = box.y.max;
        o.bbox[4] = box.z.min, o.bbox[5] = box.z.max;
        objects.push_back(o);
        return (uint32_t)objects.size() - 1;
    }

    rt_scene_desc desc() const {
        rt_scene_desc d{};
        d.version = RT_ABI_VERSION;
        d.struct_size = sizeof(rt_scene_desc);
        d.world_root = world_root;
        d.lights_root = lights_root;
        d.n_objects = (uint32_t)objects.size();
        d.n_children = (uint32_t)children.size();
        d.n_spheres = (uint32_t)spheres.size();
        d.n_planars = (uint32_t)planars.size();
        d.n_transforms = (uint32_t)transforms.size();
        d.n_media = (uint32_t)media.size();
        d.n_materials = (uint32_t)materials.size();
        d.n_textures = (uint32_t)textures.size();
        d.n_images = (uint32_t)images.size();
        d.n_perlins = (uint32_t)perlins.size();
        d.n_texels = texels.size();
        d.n_remaps = (uint32_t)remaps.size();
        d.remaps = remaps.data();
        d.objects = objects.data();
        d.children = children.data();
        d.spheres = spheres.data();
        d.planars = planars.data();
        d.transforms = transforms.data();
        d.media = media.data();
        d.materials = materials.data();
        d.textures = textures.data();
        d.images = images.data();
        d.texels = texels.data();
        d.perlins = perlins.data();
        return d;
    }

  private:
    std::map<const void*, uint32_t> tex_ids_, mat_ids_;
};

// ---- texture.rs ------------------------------------------------------------------------
class Texture {
  public:
    virtual ~Texture() = default;
    virtual rt_texture flatten(Flattener& f) const = 0;
};
using TexturePtr = std::shared_ptr<const Texture>;

class SolidColor : public Texture {  // texture.rs:9-36
  public:
    explicit SolidColor(const Color& albedo) : albedo_(albedo) {}
    static std::shared_ptr<SolidColor> from_rgb(double r, double g, double b) {
        return std::make_shared<SolidColor>(Color(r, g, b));
    }
    rt_texture flatten(Flattener&) const override {
        rt_texture t{};
        t.kind = RT_TEX_SOLID;
        t.a = t.b = RT_NONE;
        for (int i = 0; i < 3; i++) t.color[i] = albedo_[i];
        return t;
    }

  private:
    Color albedo_;
};

class CheckerTexture : public Texture {  // texture.rs:38-73
  public:
    CheckerTexture(double scale, TexturePtr even, TexturePtr odd)
        : inv_scale_(1.0 / scale), even_(std::move(even)), odd_(std::move(odd)) {}
    rt_texture flatten(Flattener& f) const override {
        rt_texture t{};
        t.kind = RT_TEX_CHECKER;
        t.a = f.texture(even_);
        t.b = f.texture(odd_);
        t.scale = inv_scale_;
        return t;
    }

  private:
    double inv_scale_;
    TexturePtr even_, odd_;
};

// Decoded picture: RGBA32F, row-major, as `image::DynamicImage::into_rgba32f` would give
// (utils/image.rs:46-53).  Decoding files is host I/O outside the hot path; callers that have
// pixels (e.g. decoded by Python) hand them in with from_pixels.
struct Image {
    uint32_t width = 0, height = 0;
    std::vector<float> rgba;
    bool linear_format = false;  // Hdr / OpenExr / Avif: no sRGB decode on fetch (image.rs:75-81)
};

class ImageTexture : public Texture {  // texture.rs:81-174
  public:
    // ImageTexture::new(file): nearest lookup, sRGB-decoded.  A file that cannot be loaded
    // renders cyan (texture.rs:167-169).  This build has no image decoders linked in, so
    // every path is "missing" unless pixels are supplied with from_pixels.
    explicit ImageTexture(const std::string& file_name) : file_(file_name) {}
    static std::shared_ptr<ImageTexture> from_pixels(Image img, bool raw_and_linear_interp) {
        auto t = std::make_shared<ImageTexture>(std::string());
        t->img_ = std::move(img);
        t->raw_ = raw_and_linear_interp;  // new_raw_image: raw = true, interp = Linear (:94-100)
        return t;
    }
    rt_texture flatten(Flattener& f) const override {
        rt_texture t{};
        t.kind = RT_TEX_IMAGE;
        t.a = t.b = RT_NONE;
        if (img_.height != 0 && img_.width != 0) {
            rt_image im{};
            im.width = img_.width;
            im.height = img_.height;
            im.flags = ((raw_ || img_.linear_format) ? RT_IMG_LINEAR : 0u) | (raw_ ? RT_IMG_INTERP : 0u);
            im.texel_offset = f.texels.size();
            f.texels.insert(f.texels.end(), img_.rgba.begin(), img_.rgba.end());
            f.images.push_back(im);
            t.a = (uint32_t)f.images.size() - 1;
        }
        return t;
    }

  private:
    std::string file_;
    Image img_;
    bool raw_ = false;
};

// utils/perlin.rs:16-37,91-107 with the seeded Random
inline rt_perlin make_perlin() {
    rt_perlin p{};
    for (int i = 0; i < 256; i++) {  // UnitVec3::random_unit_vector, vec3.rs:313-322
        double r1 = Random::f64(), r2 = Random::f64();
        p.randvec[i][0] = std::cos(2.0 * PI * r1) * 2.0 * std::sqrt(r2 * (1.0 - r2));
        p.randvec[i][1] = std::sin(2.0 * PI * r1) * 2.0 * std::sqrt(r2 * (1.0 - r2));
        p.randvec[i][2] = 1.0 - 2.0 * r2;
    }
    auto perm = [](uint32_t* a) {
        for (uint32_t i = 0; i < 256; i++) a[i] = i;
        for (size_t i = 255; i >= 1; i--) {  // perlin.rs:102-107: Fisher-Yates, target in 0..=i
            size_t target = Random::usize_inclusive(0, i);
            std::swap(a[i], a[target]);
        }
    };
    perm(p.perm_x);
    perm(p.perm_y);
    perm(p.perm_z);
    return p;
}

class NoiseTexture : public Texture {  // texture.rs:176-196
  public:
    explicit NoiseTexture(double scale) : noise_(make_perlin()), scale_(scale) {}
    rt_texture flatten(Flattener& f) const override {
        rt_texture t{};
        t.kind = RT_TEX_NOISE;
        f.perlins.push_back(noise_);
        t.a = (uint32_t)f.perlins.size() - 1;
        t.b = RT_NONE;
        t.scale = scale_;
        return t;
    }

  private:
    rt_perlin noise_;
    double scale_;
};

// Not part of the crate: the book-1 sky as a Texture evaluated on the unit direction that
// Environment::value passes as `p` (environment.rs:14-24).  See rt2025.h RT_TEX_GRADIENT_Y.
class GradientTexture : public Texture {
  public:
    GradientTexture(const Color& bottom, const Color& top) : c0_(bottom), c1_(top) {}
    rt_texture flatten(Flattener&) const override {
        rt_texture t{};
        t.kind = RT_TEX_GRADIENT_Y;
        t.a = t.b = RT_NONE;
        for (int i = 0; i < 3; i++) t.color[i] = c0_[i], t.color2[i] = c1_[i];
        return t;
    }

  private:
    Color c0_, c1_;
};

// ---- material.rs -----------------------------------------------------------------------
class Material {
  public:
    virtual ~Material() = default;
    virtual rt_material flatten(Flattener& f) const = 0;

  protected:
    static rt_material blank(uint32_t kind) {
        rt_material m{};
        m.kind = kind;
        m.tex = m.inner = m.inner2 = RT_NONE;
        return m;
    }
};
using MaterialPtr = std::shared_ptr<const Material>;

class EmptyMaterial : public Material {  // material.rs:36-47
  public:
    rt_material flatten(Flattener&) const override { return blank(RT_MAT_EMPTY); }
};
class Lambertian : public Material {  // :49-66
  public:
    explicit Lambertian(TexturePtr t) : tex_(std::move(t)) {}
    rt_material flatten(Flattener& f) const override {
        auto m = blank(RT_MAT_LAMBERTIAN);
        m.tex = f.texture(tex_);
        return m;
    }

  private:
    TexturePtr tex_;
};
class Metal : public Material {  // :68-95
  public:
    Metal(const Color& albedo, double fuzz) : albedo_(albedo), fuzz_(fuzz < 0.0 ? 0.0 : (fuzz > 1.0 ? 1.0 : fuzz)) {}
    rt_material flatten(Flattener&) const override {
        auto m = blank(RT_MAT_METAL);
        for (int i = 0; i < 3; i++) m.color[i] = albedo_[i];
        m.param = fuzz_;
        return m;
    }

  private:
    Color albedo_;
    double fuzz_;
};
class Dielectric : public Material {  // :97-144
  public:
    Dielectric(TexturePtr attenuation, double refraction_index) : tex_(std::move(attenuation)), ri_(refraction_index) {}
    rt_material flatten(Flattener& f) const override {
        auto m = blank(RT_MAT_DIELECTRIC);
        m.tex = f.texture(tex_);
        m.param = ri_;
        return m;
    }

  private:
    TexturePtr tex_;
    double ri_;
};
class DiffuseLight : public Material {  // :146-186
  public:
    explicit DiffuseLight(TexturePtr t) : tex_(std::move(t)) {}
    static std::shared_ptr<DiffuseLight> new_with_material(TexturePtr t, MaterialPtr inner) {
        auto d = std::make_shared<DiffuseLight>(std::move(t));
        d->inner_ = std::move(inner);
        return d;
    }
    rt_material flatten(Flattener& f) const override {
        auto m = blank(RT_MAT_DIFFUSE_LIGHT);
        m.tex = f.texture(tex_);
        if (inner_) m.inner = f.material(inner_);
        return m;
    }

  private:
    TexturePtr tex_;
    MaterialPtr inner_;
};
class Isotropic : public Material {  // :188-207
  public:
    explicit Isotropic(TexturePtr t) : tex_(std::move(t)) {}
    rt_material flatten(Flattener& f) const override {
        auto m = blank(RT_MAT_ISOTROPIC);
        m.tex = f.texture(tex_);
        return m;
    }

  private:
    TexturePtr tex_;
};
class Transparent : public Material {  // :209-218
  public:
    rt_material flatten(Flattener&) const override { return blank(RT_MAT_TRANSPARENT); }
};
class Mix : public Material {  // :220-268 — the ratio closure becomes data
  public:
    Mix(MaterialPtr m1, MaterialPtr m2, double ratio) : m1_(std::move(m1)), m2_(std::move(m2)), ratio_(ratio) {}
    static std::shared_ptr<Mix> from_image(MaterialPtr m1, MaterialPtr m2, std::shared_ptr<const ImageTexture> tex) {
        auto m = std::make_shared<Mix>(std::move(m1), std::move(m2), 0.0);
        m->alpha_ = std::move(tex);
        return m;
    }
    rt_material flatten(Flattener& f) const override {
        auto m = blank(RT_MAT_MIX);
        m.inner = f.material(m1_);
        m.inner2 = f.material(m2_);
        m.param = ratio_;
        if (alpha_) m.tex = f.texture(alpha_);
        return m;
    }

  private:
    MaterialPtr m1_, m2_;
    double ratio_;
    std::shared_ptr<const ImageTexture> alpha_;
};
class Portal : public Material {  // material/portal.rs:9-31 — the closure is (offset, rotation)
  public:
    Portal(const Color& attenuation, const Vec3& offset, const Quaternion& q) : att_(attenuation), off_(offset), q_(q) {}
    rt_material flatten(Flattener&) const override {
        auto m = blank(RT_MAT_PORTAL);
        for (int i = 0; i < 3; i++) m.color[i] = att_[i], m.v[i] = off_[i];
        m.v[3] = q_.w, m.v[4] = q_.x, m.v[5] = q_.y, m.v[6] = q_.z;
        return m;
    }

  private:
    Color att_;
    Vec3 off_;
    Quaternion q_;
};

// material/disney.rs:18-116,718-805.  `param_fn` closures become data: the parameters are either
// constants (DisneyBuilder) or constants + a base-colour texture (the OBJ loader, obj.rs:271-293).
struct DisneyParameters {
    Color base_color{0.8, 0.8, 0.8};
    double roughness = 0.5, anisotropic = 0.0, sheen = 0.0, sheen_tint = 0.0, clearcoat = 0.0, clearcoat_gloss = 0.0;
    double specular_tint = 0.0, metallic = 0.0, ior = 1.45, flatness = 0.0, spec_trans = 0.0, diff_trans = 0.0;
    bool thin = false;
};
class Disney : public Material {
  public:
    Disney() = default;
    explicit Disney(const DisneyParameters& p, TexturePtr base_color_tex = nullptr) : p_(p), tex_(std::move(base_color_tex)) {}
    rt_material flatten(Flattener& f) const override {
        auto m = blank(RT_MAT_DISNEY);
        for (int i = 0; i < 3; i++) m.color[i] = p_.base_color[i];
        if (tex_) m.tex = f.texture(tex_);
        m.v[RT_DISNEY_ROUGHNESS] = p_.roughness, m.v[RT_DISNEY_ANISOTROPIC] = p_.anisotropic, m.v[RT_DISNEY_SHEEN] = p_.sheen;
        m.v[RT_DISNEY_SHEEN_TINT] = p_.sheen_tint, m.v[RT_DISNEY_CLEARCOAT] = p_.clearcoat, m.v[RT_DISNEY_CLEARCOAT_GLOSS] = p_.clearcoat_gloss;
        m.v[RT_DISNEY_SPECULAR_TINT] = p_.specular_tint, m.v[RT_DISNEY_METALLIC] = p_.metallic, m.v[RT_DISNEY_IOR] = p_.ior;
        m.v[RT_DISNEY_FLATNESS] = p_.flatness, m.v[RT_DISNEY_SPEC_TRANS] = p_.spec_trans, m.v[RT_DISNEY_DIFF_TRANS] = p_.diff_trans;
        m.v[RT_DISNEY_THIN] = p_.thin ? 1.0 : 0.0;
        return m;
    }

  private:
    DisneyParameters p_;
    TexturePtr tex_;
};
// the fluent DisneyBuilder of disney.rs:718-805
class DisneyBuilder {
  public:
    DisneyBuilder& base_color(const Color& c) { p_.base_color = c; return *this; }
    DisneyBuilder& roughness(double v) { p_.roughness = v; return *this; }
    DisneyBuilder& anisotropic(double v) { p_.anisotropic = v; return *this; }
    DisneyBuilder& sheen(double v) { p_.sheen = v; return *this; }
    DisneyBuilder& sheen_tint(double v) { p_.sheen_tint = v; return *this; }
    DisneyBuilder& clearcoat(double v) { p_.clearcoat = v; return *this; }
    DisneyBuilder& clearcoat_gloss(double v) { p_.clearcoat_gloss = v; return *this; }
    DisneyBuilder& specular_tint(double v) { p_.specular_tint = v; return *this; }
    DisneyBuilder& metallic(double v) { p_.metallic = v; return *this; }
    DisneyBuilder& ior(double v) { p_.ior = v; return *this; }
    DisneyBuilder& flatness(double v) { p_.flatness = v; return *this; }
    DisneyBuilder& spec_trans(double v) { p_.spec_trans = v; return *this; }
    DisneyBuilder& diff_trans(double v) { p_.diff_trans = v; return *this; }
    DisneyBuilder& thin(bool v) { p_.thin = v; return *this; }
    std::shared_ptr<Disney> build() const { return std::make_shared<Disney>(p_); }

  private:
    DisneyParameters p_;
};

// shapes/obj.rs:20-81,150-176,196-211: the per-face material the OBJ loader wraps around a mesh material
class RemappedMaterial : public Material {
  public:
    // what load_object computes for one face: positions p1..p3, texture coordinates t1..t3 (z = 0),
    // vertex normals n1..n3, the model's raw normal map (or none)
    RemappedMaterial(MaterialPtr material, const Point3& p1, const Point3& p2, const Point3& p3, const Vec3& t1, const Vec3& t2,
                     const Vec3& t3, const Vec3& n1, const Vec3& n2, const Vec3& n3, std::shared_ptr<const ImageTexture> normal_tex)
        : material_(std::move(material)), normal_tex_(std::move(normal_tex)) {
        tex_ori_ = t1;
        tex_u_ = t2 - t1;
        tex_v_ = t3 - t1;
        Vec3 world_u = p2 - p1, world_v = p3 - p1;
        // uv_local_to_world, obj.rs:196-211
        double ua = tex_v_.y() / (-tex_u_.y() * tex_v_.x() + tex_u_.x() * tex_v_.y());
        double ub = tex_u_.y() / (tex_u_.y() * tex_v_.x() - tex_u_.x() * tex_v_.y());
        double va = tex_v_.x() / (tex_u_.y() * tex_v_.x() - tex_u_.x() * tex_v_.y());
        double vb = tex_u_.x() / (-tex_u_.y() * tex_v_.x() + tex_u_.x() * tex_v_.y());
        auto uv = unit_vector(world_u * ua + world_v * ub), vv = unit_vector(world_u * va + world_v * vb);
        has_uv_ = uv.has_value() && vv.has_value();
        if (has_uv_) u_vec_ = *uv, v_vec_ = *vv;
        normal_[0] = n1, normal_[1] = n2, normal_[2] = n3;
    }
    rt_material flatten(Flattener& f) const override {
        auto m = blank(RT_MAT_REMAPPED);
        m.inner = f.material(material_);
        rt_remap r{};
        for (int i = 0; i < 3; i++) {
            r.tex_ori[i] = tex_ori_[i], r.tex_u[i] = tex_u_[i], r.tex_v[i] = tex_v_[i];
            r.u_vec[i] = u_vec_[i], r.v_vec[i] = v_vec_[i];
            for (int k = 0; k < 3; k++) r.normal[k][i] = normal_[k][i];
        }
        r.has_uv_vecs = has_uv_ ? 1u : 0u;
        r.normal_tex = normal_tex_ ? f.texture(normal_tex_) : RT_NONE;
        f.remaps.push_back(r);
        m.inner2 = (uint32_t)f.remaps.size() - 1;
        return m;
    }

  private:
    MaterialPtr material_;
    Vec3 tex_ori_, tex_u_, tex_v_, u_vec_, v_vec_, normal_[3];
    bool has_uv_ = false;
    std::shared_ptr<const ImageTexture> normal_tex_;
};

// ---- hit.rs: Hittable ------------------------------------------------------------------
class Hittable {
  public:
    virtual ~Hittable() = default;
    virtual const AABB& bounding_box() const = 0;
    virtual uint32_t flatten(Flattener& f) const = 0;
};
using HittablePtr = std::shared_ptr<const Hittable>;  // Box<dyn Hittable> in the reference

class Sphere : public Hittable {  // shapes/sphere.rs:17-51
  public:
    Sphere(const Point3& static_center, double radius, MaterialPtr mat) : mat_(std::move(mat)) {
        Vec3 rvec(radius, radius, radius);
        center_ = static_center;
        center_vec_ = Vec3::ZERO();
        radius_ = rmax(0.0, radius);
        bbox_ = AABB::from_points(static_center - rvec, static_center + rvec);
    }
    static std::shared_ptr<Sphere> new_with_motion(const Point3& c1, const Point3& c2, double radius, MaterialPtr mat) {
        auto s = std::make_shared<Sphere>(c1, radius, std::move(mat));
        Vec3 rvec(radius, radius, radius);
        s->center_vec_ = c2 - c1;
        Point3 at0 = c1 + 0.0 * s->center_vec_, at1 = c1 + 1.0 * s->center_vec_;  // Ray::at
        AABB box1 = AABB::from_points(at0 - rvec, at0 + rvec);
        AABB box2 = AABB::from_points(at1 - rvec, at1 + rvec);
        s->bbox_ = box1.union_(box2);
        return s;
    }
    const AABB& bounding_box() const override { return bbox_; }
    uint32_t flatten(Flattener& f) const override {
        rt_sphere s{};
        for (int i = 0; i < 3; i++) s.center[i] = center_[i], s.center_vec[i] = center_vec_[i];
        s.radius = radius_;
        f.spheres.push_back(s);
        return f.push_object(RT_OBJ_SPHERE, f.material(mat_), (uint32_t)f.spheres.size() - 1, bbox_, {});
    }

  private:
    Point3 center_;
    Vec3 center_vec_;
    double radius_;
    MaterialPtr mat_;
    AABB bbox_;
};

// Quad and Triangle: shapes/quad.rs:31-49, shapes/triangle.rs:29-46
class Planar : public Hittable {
  public:
    const AABB& bounding_box() const override { return bbox_; }
    uint32_t flatten(Flattener& f) const override {
        rt_planar p{};
        for (int i = 0; i < 3; i++) {
            p.anchor[i] = anchor_[i], p.u[i] = u_[i], p.v[i] = v_[i];
            p.normal[i] = normal_[i], p.w[i] = w_[i];
        }
        p.parm_d = parm_d_;
        p.area = area_;
        f.planars.push_back(p);
        return f.push_object(kind_, f.material(mat_), (uint32_t)f.planars.size() - 1, bbox_, {});
    }

  protected:
    Planar(uint32_t kind, const Point3& anchor, const Vec3& u, const Vec3& v, MaterialPtr mat)
        : kind_(kind), anchor_(anchor), u_(u), v_(v), mat_(std::move(mat)) {}
    bool derive(bool triangle) {
        Vec3 n = u_.cross(v_);
        auto normal = unit_vector(n);
        if (!normal) return false;
        normal_ = *normal;
        parm_d_ = normal_.dot(anchor_);
        w_ = n / n.length_squared();
        area_ = triangle ? n.length() / 2.0 : n.length();
        return true;
    }
    uint32_t kind_;
    Point3 anchor_;
    Vec3 u_, v_, w_, normal_;
    double parm_d_ = 0, area_ = 0;
    MaterialPtr mat_;
    AABB bbox_;
};

class Quad : public Planar {
  public:
    Quad(const Point3& anchor, const Vec3& u, const Vec3& v, MaterialPtr mat)
        : Planar(RT_OBJ_QUAD, anchor, u, v, std::move(mat)) {
        if (!derive(false)) throw std::runtime_error("The length of normal should be normalizable!");
        // quad.rs:52-58
        AABB d1 = AABB::from_points(anchor, anchor + u + v);
        AABB d2 = AABB::from_points(anchor + u, anchor + v);
        bbox_ = d1.union_(d2);
    }
};

class Triangle : public Planar {
  public:
    // Triangle::new returns None for a degenerate triangle (triangle.rs:29-31)
    static std::shared_ptr<Triangle> create(const Point3& anchor, const Vec3& u, const Vec3& v, MaterialPtr mat) {
        std::shared_ptr<Triangle> t(new Triangle(anchor, u, v, std::move(mat)));
        if (!t->derive(true)) return nullptr;
        // triangle.rs:49-54
        AABB b1 = AABB::from_points(anchor, anchor + u);
        AABB b2 = AABB::from_points(anchor, anchor + v);
        t->bbox_ = b1.union_(b2);
        return t;
    }

  private:
    Triangle(const Point3& anchor, const Vec3& u, const Vec3& v, MaterialPtr mat)
        : Planar(RT_OBJ_TRIANGLE, anchor, u, v, std::move(mat)) {}
};

class Hittables : public Hittable {  // hits.rs:10-31
  public:
    Hittables() = default;  // Default: empty list, bbox = degenerate box at the origin (!)
    explicit Hittables(HittablePtr object) : bbox_(object->bounding_box()) { objects.push_back(std::move(object)); }
    void clear() { objects.clear(); }
    void add(HittablePtr object) {
        bbox_ = bbox_.union_(object->bounding_box());
        objects.push_back(std::move(object));
    }
    const AABB& bounding_box() const override { return bbox_; }
    uint32_t flatten(Flattener& f) const override {
        std::vector<uint32_t> kids;
        kids.reserve(objects.size());
        for (auto& o : objects) kids.push_back(o->flatten(f));
        return f.push_object(RT_OBJ_LIST, RT_NONE, RT_NONE, bbox_, kids);
    }
    std::vector<HittablePtr> objects;

  private:
    AABB bbox_;
};

class BVH : public Hittable {  // bvh.rs:5-46 — only the root is kept; consumers rebuild the split
  public:
    explicit BVH(Hittables world) : BVH(std::move(world.objects)) {}
    explicit BVH(std::vector<HittablePtr> objects) : objects_(std::move(objects)) {
        if (objects_.empty()) throw std::runtime_error("BVH node must contain at least one object");
        bbox_ = AABB::EMPTY();
        for (auto& o : objects_) bbox_ = bbox_.union_(o->bounding_box());
    }
    const AABB& bounding_box() const override { return bbox_; }
    uint32_t flatten(Flattener& f) const override {
        std::vector<uint32_t> kids;
        kids.reserve(objects_.size());
        for (auto& o : objects_) kids.push_back(o->flatten(f));
        return f.push_object(RT_OBJ_BVH, RT_NONE, RT_NONE, bbox_, kids);
    }

  private:
    std::vector<HittablePtr> objects_;
    AABB bbox_;
};

class Transform : public Hittable {  // shapes.rs:23-86
  public:
    Transform(HittablePtr object, std::optional<Vec3> offset, std::optional<Quaternion> q, std::optional<Vec3> scale)
        : object_(std::move(object)),
          offset_(offset.value_or(Vec3::ZERO())),
          q_(q.value_or(Quaternion::identity())),
          scale_(scale.value_or(Vec3(1.0, 1.0, 1.0))) {
        Point3 pts[8];
        object_->bounding_box().all_points(pts);
        Vec3 mn(INF, INF, INF), mx(-INF, -INF, -INF);
        for (auto& p : pts) {
            Vec3 t = transform(p);
            for (int i = 0; i < 3; i++) mn[i] = rmin(mn[i], t[i]), mx[i] = rmax(mx[i], t[i]);
        }
        bbox_ = AABB::from_points(mn, mx);
    }
    Vec3 transform(const Vec3& v) const { return q_.rotate_vector(v * scale_) + offset_; }  // :74-78
    const AABB& bounding_box() const override { return bbox_; }
    uint32_t flatten(Flattener& f) const override {
        uint32_t kid = object_->flatten(f);
        rt_transform t{};
        for (int i = 0; i < 3; i++) t.offset[i] = offset_[i], t.scale[i] = scale_[i];
        t.quat[0] = q_.w, t.quat[1] = q_.x, t.quat[2] = q_.y, t.quat[3] = q_.z;
        f.transforms.push_back(t);
        return f.push_object(RT_OBJ_TRANSFORM, RT_NONE, (uint32_t)f.transforms.size() - 1, bbox_, {kid});
    }

  private:
    HittablePtr object_;
    Vec3 offset_;
    Quaternion q_;
    Vec3 scale_;
    AABB bbox_;
};

class ConstantMedium : public Hittable {  // volume.rs:16-35,75-77
  public:
    static std::shared_ptr<ConstantMedium> new_with_tex(HittablePtr boundary, double density, TexturePtr tex) {
        auto m = std::make_shared<ConstantMedium>();
        m->boundary_ = std::move(boundary);
        m->neg_inv_density_ = -1.0 / density;
        m->phase_ = std::make_shared<Isotropic>(std::move(tex));
        return m;
    }
    const AABB& bounding_box() const override { return boundary_->bounding_box(); }
    uint32_t flatten(Flattener& f) const override {
        uint32_t kid = boundary_->flatten(f);
        rt_medium m{};
        m.neg_inv_density = neg_inv_density_;
        f.media.push_back(m);
        return f.push_object(RT_OBJ_MEDIUM, f.material(phase_), (uint32_t)f.media.size() - 1, bounding_box(), {kid});
    }

  private:
    HittablePtr boundary_;
    double neg_inv_density_ = 0;
    MaterialPtr phase_;
};

// quad.rs:128-189
inline std::shared_ptr<Hittables> build_box(const Point3& a, const Point3& b, MaterialPtr mat) {
    auto sides = std::make_shared<Hittables>();
    Point3 mn(rmin(a[0], b[0]), rmin(a[1], b[1]), rmin(a[2], b[2]));
    Point3 mx(rmax(a[0], b[0]), rmax(a[1], b[1]), rmax(a[2], b[2]));
    Vec3 dx(mx.x() - mn.x(), 0.0, 0.0), dy(0.0, mx.y() - mn.y(), 0.0), dz(0.0, 0.0, mx.z() - mn.z());
    sides->add(std::make_shared<Quad>(Point3(mn.x(), mn.y(), mx.z()), dx, dy, mat));
    sides->add(std::make_shared<Quad>(Point3(mx.x(), mn.y(), mx.z()), -dz, dy, mat));
    sides->add(std::make_shared<Quad>(Point3(mx.x(), mn.y(), mn.z()), -dx, dy, mat));
    sides->add(std::make_shared<Quad>(Point3(mn.x(), mn.y(), mn.z()), dz, dy, mat));
    sides->add(std::make_shared<Quad>(Point3(mn.x(), mx.y(), mx.z()), dx, -dz, mat));
    sides->add(std::make_shared<Quad>(Point3(mn.x(), mn.y(), mn.z()), dx, dz, mat));
    return sides;
}

inline uint32_t Flattener::texture(const TexturePtr& t) {
    auto it = tex_ids_.find(t.get());
    if (it != tex_ids_.end()) return it->second;
    rt_texture r = t->flatten(*this);  // children first, so indices always point backwards
    textures.push_back(r);
    uint32_t id = (uint32_t)textures.size() - 1;
    tex_ids_[t.get()] = id;
    return id;
}
inline uint32_t Flattener::material(const MaterialPtr& m) {
    auto it = mat_ids_.find(m.get());
    if (it != mat_ids_.end()) return it->second;
    rt_material r = m->flatten(*this);
    materials.push_back(r);
    uint32_t id = (uint32_t)materials.size() - 1;
    mat_ids_[m.get()] = id;
    return id;
}
inline uint32_t Flattener::object(const Hittable& h) { return h.flatten(*this); }

// ---- utils/color.rs, shapes/environment.rs, camera.rs ----------------------------------
enum class ToonMap { None = 0, ACES = 1 };

struct Environment {  // shapes/environment.rs:9-11
    TexturePtr texture;
};

struct RgbImage {
    uint32_t width = 0, height = 0;
    std::vector<uint8_t> data;  // row-major RGB8, like image::RgbImage
};

class Camera {  // camera.rs:46-75
  public:
    double aspect_ratio = 1.0;
    uint32_t image_width = 100;
    size_t samples_per_pixel = 10;
    uint32_t max_depth = 10;
    Environment background{std::make_shared<SolidColor>(Color(0, 0, 0))};
    double vertical_fov_in_degrees = 90.0;
    Point3 look_from{0, 0, 0};
    Point3 look_at{0, 0, -1};
    Vec3 vec_up{0, 1, 0};
    double defocus_angle_in_degrees = 0.0;
    double focus_distance = 10.0;
    ToonMap toon_map = ToonMap::None;

    // additions that have no counterpart in the reference (its RNG is unseeded and it has one
    // device): the Philox seed and the last call's statistics
    uint64_t seed = 0x2025;
    rt_stats last_stats{};

    Camera() = default;
    Camera(double aspect, uint32_t width) : aspect_ratio(aspect), image_width(width) {}

    // camera.rs:204-245, operation for operation
    rt_camera initialize() const {
        rt_camera c{};
        double hf = (double)image_width / aspect_ratio;
        uint32_t image_height = hf >= 4294967295.0 ? 0xFFFFFFFFu : (hf > 0.0 ? (uint32_t)hf : 0u);  // `as u32` saturates
        if (image_height < 1) image_height = 1;
        double sq = std::sqrt((double)samples_per_pixel);
        uint32_t sqrt_spp = (uint32_t)sq;
        c.image_width = image_width;
        c.image_height = image_height;
        c.sqrt_spp = sqrt_spp;
        c.max_depth = max_depth;
        c.toon_map = (uint32_t)toon_map;
        c.pixel_sample_scale = 1.0 / (double)(sqrt_spp * sqrt_spp);
        c.recip_sqrt_spp = 1.0 / (double)sqrt_spp;
        Point3 center = look_from;
        double theta = vertical_fov_in_degrees * (PI / 180.0);
        double h = std::tan(theta / 2.0);
        double viewport_height = 2.0 * h * focus_distance;
        double viewport_width = viewport_height * ((double)image_width / (double)image_height);
        auto w = unit_vector(look_from - look_at);
        if (!w) throw std::runtime_error("Camera axis w should be normalizable!");
        auto u = unit_vector(vec_up.cross(*w));
        if (!u) throw std::runtime_error("Camera axis u should be normalizable!");
        Vec3 v = w->cross(*u);
        Vec3 viewport_u = viewport_width * (*u);
        Vec3 viewport_v = viewport_height * (-v);
        Vec3 du = viewport_u / (double)image_width;
        Vec3 dv = viewport_v / (double)image_height;
        Vec3 upper_left = center - focus_distance * (*w) - viewport_u / 2.0 - viewport_v / 2.0;
        Vec3 p00 = upper_left + 0.5 * (du + dv);
        double defocus_radius = focus_distance * std::tan((defocus_angle_in_degrees / 2.0) * (PI / 180.0));
        Vec3 disk_u = (*u) * defocus_radius, disk_v = v * defocus_radius;
        c.defocus_angle_in_degrees = defocus_angle_in_degrees;
        for (int i = 0; i < 3; i++) {
            c.center[i] = center[i];
            c.pixel00_loc[i] = p00[i];
            c.pixel_delta_u[i] = du[i];
            c.pixel_delta_v[i] = dv[i];
            c.defocus_disk_u[i] = disk_u[i];
            c.defocus_disk_v[i] = disk_v[i];
        }
        return c;
    }

    // Flatten world, lights and the background texture; fills cam.background_tex.
    void flatten(Flattener& f, rt_camera& cam, const Hittable& world, const Hittable* lights) const {
        f.world_root = world.flatten(f);
        f.lights_root = lights ? lights->flatten(f) : RT_NONE;
        cam.background_tex = f.texture(background.texture);
    }

    // camera.rs:161-202.  Everything after initilize() runs behind the C ABI on the GPU.
    RgbImage render(const Hittable& world, const Hittable* lights) {
        rt_camera cam = initialize();
        Flattener f;
        flatten(f, cam, world, lights);
        rt_scene_desc d = f.desc();
        rt_scene* scene = nullptr;
        if (rt_scene_create(&d, nullptr, &scene) != RT_OK) throw std::runtime_error(rt_last_error());
        rt_render_opts o{};
        o.struct_size = sizeof(o);
        o.seed = seed;
        RgbImage img;
        img.width = cam.image_width;
        img.height = cam.image_height;
        img.data.resize((size_t)cam.image_width * cam.image_height * 3);
        // render + Color::to_rgb on the device: only the 8-bit pixels come back
        int rc = rt_render_rgb8(scene, &cam, &o, img.data.data(), &last_stats);
        std::string err = rc == RT_OK ? "" : rt_last_error();
        rt_scene_destroy(scene);
        if (rc != RT_OK) throw std::runtime_error(err);
        return img;
    }
};

}  // namespace rt2025

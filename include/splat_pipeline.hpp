// splat_pipeline.hpp -- the host side above the C ABI, in compiled code.
//
// The reference is a Rust crate and its toolchain is absent from this image, so the interface a
// caller of the reference programs against is mirrored here in C++17 (header only, needs nothing
// but splat.h and libsplat_b200.so): same type names, field names, method names, argument meaning
// and error behaviour as
//     struct Camera                         src/camera.rs:4-127
//     struct Gaussian / GaussianList        src/gaussians.rs:31-38, :408-445
//     naive_gaussians / load_from_ply       src/gaussians.rs:319-374, :375-405 (+ set_property :258-282)
//     GaussianSplatPipeline01 / 02          src/pipelines.rs:54-57 / :172-175
//     ::render_to_buffer(&mut Buffer2d)     src/pipelines.rs:66-86 / :260-280      <- the boundary
// What sits BELOW render_to_buffer in the reference (sort, vertex / fragment / blend callbacks, the
// euc rasteriser) is not here: it runs in the CUDA library, and there is no CPU path -- creating a
// pipeline without a usable B200 throws.  The Rust shim of INTEGRATION.md has the same body as
// render_to_buffer below: marshal the camera, upload the scene once, one splat_render call.
//
// Where Rust panics (a PLY with an element other than `vertex`, an unreadable file) this throws
// splat_b200::Error; library errors (negative codes of splat.h) are thrown with the library's own
// message.  Nothing is caught and retried on a CPU.
#ifndef SPLAT_B200_PIPELINE_HPP
#define SPLAT_B200_PIPELINE_HPP

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "splat.h"

namespace splat_b200 {

struct Error : std::runtime_error {
  int code;   // a SPLAT_ERR_* value, or 0 for host-side failures (file format, I/O)
  Error(int c, const std::string &what) : std::runtime_error(what), code(c) {}
};

// ------------------------------------------------------------------------------------------------
// euc::Buffer<T, 2> as the viewer uses it (main.rs:28, :73, :79): [w, h] extent, row-major storage,
// fill(), raw().
template <class T>
class Buffer2d {
 public:
  static Buffer2d fill(std::array<size_t, 2> size, T value) { return Buffer2d(size[0], size[1], value); }
  Buffer2d(size_t w, size_t h, T value = T()) : w_(w), h_(h), data_(w * h, value) {}
  std::array<size_t, 2> size() const { return {w_, h_}; }
  void fill(T value) { std::fill(data_.begin(), data_.end(), value); }
  T *raw_mut() { return data_.data(); }
  const T *raw() const { return data_.data(); }
  T &at(size_t x, size_t y) { return data_[y * w_ + x]; }
  const T &at(size_t x, size_t y) const { return data_[y * w_ + x]; }

 private:
  size_t w_, h_;
  std::vector<T> data_;
};

// ------------------------------------------------------------------------------------------------
// Small column-major 4x4 / 3-vector helpers: nalgebra-glm 0.18.0 restated in f32, one rounding per
// operation, in the order splat_b200/camera.py uses (tests compare the two).
using Vec3 = std::array<float, 3>;
using Vec4 = std::array<float, 4>;
struct Mat4 {
  float m[16];   // column-major: element (row r, column c) = m[4*c + r], exactly nalgebra's as_slice()
  float &operator()(int r, int c) { return m[4 * c + r]; }
  float operator()(int r, int c) const { return m[4 * c + r]; }
  static Mat4 identity() {
    Mat4 a{};
    for (int i = 0; i < 4; ++i) a(i, i) = 1.0f;
    return a;
  }
  const float *as_slice() const { return m; }
};

namespace glm {
inline Vec3 sub(const Vec3 &a, const Vec3 &b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
inline float dot(const Vec3 &a, const Vec3 &b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
inline Vec3 cross(const Vec3 &a, const Vec3 &b) {
  return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
}
inline Vec3 normalize(const Vec3 &v) {
  const float n = std::sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
  return {v[0] / n, v[1] / n, v[2] / n};
}
inline Vec4 mul(const Mat4 &a, const Vec4 &v) {
  Vec4 r;
  for (int i = 0; i < 4; ++i) r[i] = ((a(i, 0) * v[0] + a(i, 1) * v[1]) + a(i, 2) * v[2]) + a(i, 3) * v[3];
  return r;
}
// glm::look_at: right-handed (camera.rs:65)
inline Mat4 look_at(const Vec3 &eye, const Vec3 &center, const Vec3 &up) {
  const Vec3 z = normalize(sub(eye, center));   // the camera looks down -z
  const Vec3 x = normalize(cross(up, z));
  const Vec3 y = normalize(cross(z, x));
  const Vec3 neg = {-eye[0], -eye[1], -eye[2]};
  Mat4 a = Mat4::identity();
  const Vec3 *axes[3] = {&x, &y, &z};
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) a(r, c) = (*axes[r])[c];
    a(r, 3) = dot(*axes[r], neg);
  }
  return a;
}
// glm::perspective(aspect, fovy, near, far) = perspective_rh_no, NDC z in [-1, 1] (camera.rs:67)
inline Mat4 perspective(float aspect, float fovy, float near, float far) {
  const float t = std::tan(fovy / 2.0f);
  Mat4 a{};
  a(0, 0) = 1.0f / (aspect * t);
  a(1, 1) = 1.0f / t;
  a(2, 2) = -(far + near) / (far - near);
  a(2, 3) = -((2.0f * far) * near) / (far - near);
  a(3, 2) = -1.0f;
  return a;
}
// glm::rotation(angle, axis): axis-angle as a homogeneous 4x4 (camera.rs:57, :62); zero axis -> identity
inline Mat4 rotation(float angle, const Vec3 &axis) {
  Mat4 a = Mat4::identity();
  const float n = std::sqrt((axis[0] * axis[0] + axis[1] * axis[1]) + axis[2] * axis[2]);
  if (n == 0.0f) return a;
  const float ux = axis[0] / n, uy = axis[1] / n, uz = axis[2] / n;
  const float s = std::sin(angle), c = std::cos(angle), k = 1.0f - c;
  a(0, 0) = ux * ux * k + c;      a(0, 1) = ux * uy * k - uz * s; a(0, 2) = ux * uz * k + uy * s;
  a(1, 0) = ux * uy * k + uz * s; a(1, 1) = uy * uy * k + c;      a(1, 2) = uy * uz * k - ux * s;
  a(2, 0) = ux * uz * k - uy * s; a(2, 1) = uy * uz * k + ux * s; a(2, 2) = uz * uz * k + c;
  return a;
}
}  // namespace glm

// ------------------------------------------------------------------------------------------------
// struct Camera, camera.rs:4-19, impl :21-127.  Public fields keep the reference's names.
struct Camera {
  float znear = 0.01f, zfar = 100.0f;
  float h, w;
  float fovy = 3.14159265358979323846f / 2.0f;
  Vec3 position{0.0f, 0.0f, 3.0f};
  Vec3 target{0.0f, 0.0f, 0.0f};
  Vec3 up{0.0f, -1.0f, 0.0f};
  float yaw = 0.0f, pitch = 0.0f;
  bool is_pose_dirty = true, is_intrin_dirty = true;
  Mat4 view_matrix = Mat4::identity();
  Mat4 projection_matrix = Mat4::identity();

  // Camera::new(h, w, start_position), camera.rs:22-39
  Camera(float h_, float w_) : h(h_), w(w_) {}
  Camera(float h_, float w_, const Vec3 &start_position) : h(h_), w(w_), position(start_position) {}

  // camera.rs:41-68
  void compute_matrices() {
    Vec4 pos4{position[0], position[1], position[2], 1.0f};
    const Vec4 pivot{target[0], target[1], target[2], 1.0f};
    const Vec3 viewdir = glm::normalize(glm::sub(position, target));
    const float cos_angle = glm::dot(viewdir, up);
    const float sgn = pitch > 0.0f ? 1.0f : (pitch < 0.0f ? -1.0f : 0.0f);
    if (cos_angle * sgn > 0.99f) pitch = 0.0f;
    auto about_pivot = [&](const Mat4 &R, const Vec4 &p) {
      const Vec4 d{p[0] - pivot[0], p[1] - pivot[1], p[2] - pivot[2], p[3] - pivot[3]};
      const Vec4 r = glm::mul(R, d);
      return Vec4{r[0] + pivot[0], r[1] + pivot[1], r[2] + pivot[2], r[3] + pivot[3]};
    };
    pos4 = about_pivot(glm::rotation(yaw, up), pos4);
    const Vec3 right = glm::cross(up, position);   // camera.rs:61 uses the UNROTATED self.position
    const Vec4 fin = about_pivot(glm::rotation(pitch, right), pos4);
    view_matrix = glm::look_at({fin[0], fin[1], fin[2]}, target, up);
    projection_matrix = glm::perspective(w / h, fovy, znear, zfar);
  }
  const Mat4 &get_view_matrix() const { return view_matrix; }          // camera.rs:70
  const Mat4 &get_project_matrix() const { return projection_matrix; } // camera.rs:80
  void update_resolution(float height, float width) { h = height; w = width; is_intrin_dirty = true; }
  // camera.rs:84-89: (htanx, htany, focal)
  Vec3 get_htanfovxy_focal() const {
    const float htany = std::tan(fovy / 2.0f);
    const float htanx = (htany / h) * w;
    const float focal = h / (2.0f * htany);
    return {htanx, htany, focal};
  }
  float get_focal() const { return get_htanfovxy_focal()[2]; }
  void update_pitch_angle(float delta) { pitch = pitch + delta; is_pose_dirty = true; }
  void update_yaw_angle(float delta) { yaw = yaw + delta; is_pose_dirty = true; }
  // camera.rs:103-126; self.position is never updated by orbiting, and the SH view
  // direction keeps using it (pipelines.rs:99)
  void update_camera_pose() { compute_matrices(); is_pose_dirty = false; }
};

// what the kernels need from a Camera, marshalled exactly like the Rust shim does
inline splat_camera camera_struct(const Camera &c) {
  splat_camera s;
  std::memcpy(s.view, c.get_view_matrix().as_slice(), sizeof s.view);
  std::memcpy(s.proj, c.get_project_matrix().as_slice(), sizeof s.proj);
  for (int i = 0; i < 3; ++i) s.position[i] = c.position[i];
  s.w = c.w;
  s.h = c.h;
  const Vec3 hf = c.get_htanfovxy_focal();
  s.htanx = hf[0]; s.htany = hf[1]; s.focal = hf[2];
  return s;
}

// ------------------------------------------------------------------------------------------------
// struct Gaussian, gaussians.rs:31-38 (cov3d is derived: the device recomputes it at upload).
// Laid out as the 59 consecutive floats splat_upload_aos takes.
struct Gaussian {
  float position[3];
  float scale[3];      // already exp()'d
  float opacity;       // already sigmoid()'d
  float rotation[4];   // nalgebra coords order (i, j, k, w)
  float sh[48];        // f_dc_0..2 then f_rest_0..44, as stored

  // PropertyAccess::new, gaussians.rs:247-256
  static Gaussian make() {
    Gaussian g;
    std::memset(&g, 0, sizeof g);
    g.rotation[3] = 1.0f;
    return g;
  }
};
static_assert(sizeof(Gaussian) == 59 * sizeof(float), "Gaussian must be 59 packed floats (splat_upload_aos)");

// struct GaussianList, gaussians.rs:408-416: every attribute is one contiguous column-major nalgebra
// matrix (k x N), i.e. k consecutive floats per Gaussian.
struct GaussianList {
  std::vector<float> positions;   // 4 x N  (x, y, z, 1)
  std::vector<float> scales;      // 3 x N
  std::vector<float> opacities;   // N
  std::vector<float> rotations;   // 4 x N  (i, j, k, w)
  std::vector<float> sh;          // 48 x N
  size_t num_gaussians = 0;

  // gaussians.rs:419-440
  static GaussianList from_vec(const std::vector<Gaussian> &gs) {
    GaussianList l;
    const size_t n = gs.size();
    l.num_gaussians = n;
    l.positions.resize(4 * n); l.scales.resize(3 * n); l.opacities.resize(n); l.rotations.resize(4 * n); l.sh.resize(48 * n);
    for (size_t i = 0; i < n; ++i) {
      const Gaussian &g = gs[i];
      for (int k = 0; k < 3; ++k) { l.positions[4 * i + k] = g.position[k]; l.scales[3 * i + k] = g.scale[k]; }
      l.positions[4 * i + 3] = 1.0f;
      l.opacities[i] = g.opacity;
      for (int k = 0; k < 4; ++k) l.rotations[4 * i + k] = g.rotation[k];
      std::memcpy(&l.sh[48 * i], g.sh, sizeof g.sh);
    }
    return l;
  }
  static GaussianList naive_gaussians();   // gaussians.rs:442-445
};

// The 4-Gaussian test scene, gaussians.rs:319-374 (the 0.28209 literal at :330 is kept)
inline std::vector<Gaussian> naive_gaussians() {
  struct Spec { float pos[3], scale[3], color[3]; };
  const Spec specs[4] = {{{0, 0, 0}, {0.03f, 0.03f, 0.03f}, {1, 0, 1}},
                         {{1, 0, 0}, {0.2f, 0.03f, 0.03f}, {1, 0, 0}},
                         {{0, 1, 0}, {0.03f, 0.2f, 0.03f}, {0, 1, 0}},
                         {{0, 0, 1}, {0.03f, 0.03f, 0.2f}, {0, 0, 1}}};
  std::vector<Gaussian> out;
  for (const Spec &s : specs) {
    Gaussian g = Gaussian::make();
    for (int k = 0; k < 3; ++k) {
      g.position[k] = s.pos[k];
      g.scale[k] = s.scale[k];
      g.sh[k] = (s.color[k] - 0.5f) / 0.28209f;
    }
    g.opacity = 1.0f;
    out.push_back(g);
  }
  return out;
}
inline GaussianList GaussianList::naive_gaussians() { return from_vec(splat_b200::naive_gaussians()); }

// ------------------------------------------------------------------------------------------------
// load_from_ply, gaussians.rs:375-405 with set_property :258-282: scale_i -> exp, opacity ->
// 1/(1+exp(-v)), rot_0 -> w and rot_1..3 -> i, j, k, f_dc_i -> sh[i], f_rest_i -> sh[3+i] (no channel
// transpose), unknown properties ignored; then the mean position -- accumulated sequentially in f32
// in file order, one division (:394-399) -- is subtracted.  Any element other than `vertex` is the
// reference's panic "Unexpected element!".  binary_little_endian (what 3DGS trainers write) and ascii.
namespace detail {
struct PlyProp { std::string name; int size; char kind; };   // kind: f float, d double, u/i integers
inline bool ply_type(const std::string &t, PlyProp *p) {
  static const std::map<std::string, std::pair<int, char>> types = {
      {"float", {4, 'f'}}, {"float32", {4, 'f'}}, {"double", {8, 'd'}}, {"float64", {8, 'd'}}, {"uchar", {1, 'u'}}, {"uint8", {1, 'u'}},
      {"char", {1, 'i'}},  {"int8", {1, 'i'}},    {"short", {2, 'i'}},  {"int16", {2, 'i'}},   {"ushort", {2, 'u'}}, {"uint16", {2, 'u'}},
      {"int", {4, 'i'}},   {"int32", {4, 'i'}},   {"uint", {4, 'u'}},   {"uint32", {4, 'u'}}};
  auto it = types.find(t);
  if (it == types.end()) return false;
  p->size = it->second.first;
  p->kind = it->second.second;
  return true;
}
inline float ply_scalar(const unsigned char *p, const PlyProp &pr) {   // little-endian hosts only (x86-64, aarch64)
  switch (pr.kind) {
    case 'f': { float v; std::memcpy(&v, p, 4); return v; }
    case 'd': { double v; std::memcpy(&v, p, 8); return (float)v; }
    case 'u': { uint32_t v = 0; std::memcpy(&v, p, pr.size); return (float)v; }
    default: {
      if (pr.size == 1) { int8_t v; std::memcpy(&v, p, 1); return (float)v; }
      if (pr.size == 2) { int16_t v; std::memcpy(&v, p, 2); return (float)v; }
      int32_t v; std::memcpy(&v, p, 4); return (float)v;
    }
  }
}
// set_property, gaussians.rs:258-282
inline void set_property(Gaussian &g, const std::string &key, float v) {
  if (key == "x") g.position[0] = v;
  else if (key == "y") g.position[1] = v;
  else if (key == "z") g.position[2] = v;
  else if (key == "scale_0") g.scale[0] = std::exp(v);
  else if (key == "scale_1") g.scale[1] = std::exp(v);
  else if (key == "scale_2") g.scale[2] = std::exp(v);
  else if (key == "opacity") g.opacity = 1.0f / (1.0f + std::exp(-v));
  else if (key == "rot_0") g.rotation[3] = v;
  else if (key == "rot_1") g.rotation[0] = v;
  else if (key == "rot_2") g.rotation[1] = v;
  else if (key == "rot_3") g.rotation[2] = v;
  else if (key.compare(0, 5, "f_dc_") == 0) {
    const int i = std::atoi(key.c_str() + 5);
    if (i >= 0 && i < 3) g.sh[i] = v;
  } else if (key.compare(0, 7, "f_rest_") == 0) {
    const int i = std::atoi(key.c_str() + 7);
    if (i >= 0 && i < 45) g.sh[3 + i] = v;
  }
}
}  // namespace detail

inline std::vector<Gaussian> load_from_ply(const std::string &filename) {
  std::ifstream f(filename, std::ios::binary);
  if (!f) throw Error(0, "cannot open " + filename);
  std::string data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  const std::string marker = "end_header\n";
  const size_t hpos = data.find(marker);
  if (data.compare(0, 3, "ply") != 0 || hpos == std::string::npos) throw Error(0, filename + ": not a PLY file");
  const size_t body = hpos + marker.size();
  std::istringstream header(data.substr(0, body));
  std::string line, format;
  std::vector<detail::PlyProp> props;
  size_t n = 0;
  bool in_vertex = false;
  while (std::getline(header, line)) {
    std::istringstream ls(line);
    std::string tok;
    if (!(ls >> tok)) continue;
    if (tok == "format") {
      ls >> format;
    } else if (tok == "element") {
      std::string name;
      ls >> name >> n;
      if (name != "vertex") throw Error(0, "Unexpected element!");   // gaussians.rs:390
      in_vertex = true;
    } else if (tok == "property" && in_vertex) {
      std::string type, name;
      ls >> type >> name;
      detail::PlyProp p;
      if (!detail::ply_type(type, &p)) throw Error(0, filename + ": unsupported property type " + type);
      p.name = name;
      props.push_back(p);
    }
  }
  std::vector<Gaussian> out(n, Gaussian::make());
  if (format == "binary_little_endian") {
    size_t stride = 0;
    for (const auto &p : props) stride += (size_t)p.size;
    if (data.size() < body + n * stride) throw Error(0, filename + ": truncated vertex payload");
    const unsigned char *row = reinterpret_cast<const unsigned char *>(data.data()) + body;
    for (size_t i = 0; i < n; ++i, row += stride) {
      const unsigned char *p = row;
      for (const auto &pr : props) {
        detail::set_property(out[i], pr.name, detail::ply_scalar(p, pr));
        p += pr.size;
      }
    }
  } else if (format == "ascii") {
    std::istringstream vals(data.substr(body));
    for (size_t i = 0; i < n; ++i)
      for (const auto &pr : props) {
        double v;
        if (!(vals >> v)) throw Error(0, filename + ": truncated vertex payload");
        detail::set_property(out[i], pr.name, (float)v);
      }
  } else {
    throw Error(0, filename + ": unsupported PLY format " + format);
  }
  if (n) {
    // gaussians.rs:394-399: running f32 sums in file order, one division, then the subtraction
    float avg[3] = {0.0f, 0.0f, 0.0f};
    for (const Gaussian &g : out)
      for (int k = 0; k < 3; ++k) avg[k] += g.position[k];
    for (int k = 0; k < 3; ++k) avg[k] /= (float)n;
    for (Gaussian &g : out)
      for (int k = 0; k < 3; ++k) g.position[k] -= avg[k];
  }
  return out;
}

// `trim` (src/bin/00_ply_load.rs:9-63): copy the first `count` vertices of a binary_little_endian PLY into a
// new file whose header differs only in the vertex count (tiny test scenes).  Returns the vertices written.
inline size_t trim_ply(const std::string &src, const std::string &dst, size_t count = 3) {
  std::ifstream f(src, std::ios::binary);
  if (!f) throw Error(0, "cannot open " + src);
  std::string data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  const std::string marker = "end_header\n";
  const size_t hpos = data.find(marker);
  if (data.compare(0, 3, "ply") != 0 || hpos == std::string::npos) throw Error(0, src + ": not a PLY file");
  const size_t body = hpos + marker.size();
  std::string out_header, line;
  std::istringstream header(data.substr(0, body));
  size_t n = 0, stride = 0;
  while (std::getline(header, line)) {
    std::istringstream ls(line);
    std::string tok, a, b;
    ls >> tok >> a >> b;
    if (tok == "element") {
      if (a != "vertex") throw Error(0, "Unexpected element!");
      n = (size_t)std::strtoull(b.c_str(), nullptr, 10);
      line = "element vertex " + std::to_string(std::min(n, count));
    } else if (tok == "property") {
      detail::PlyProp p;
      if (!detail::ply_type(a, &p)) throw Error(0, src + ": unsupported property type " + a);
      stride += (size_t)p.size;
    } else if (tok == "format" && a != "binary_little_endian") {
      throw Error(0, "trim_ply handles binary_little_endian files");
    }
    out_header += line + "\n";
  }
  const size_t m = std::min(n, count);
  if (data.size() < body + m * stride) throw Error(0, src + ": truncated vertex payload");
  std::ofstream o(dst, std::ios::binary);
  if (!o) throw Error(0, "cannot write " + dst);
  o.write(out_header.data(), (std::streamsize)out_header.size());
  o.write(data.data() + body, (std::streamsize)(m * stride));
  return m;
}

// ------------------------------------------------------------------------------------------------
// The device side of a pipeline: one library context and the identity of the scene it holds.  Kept
// OUT of the pipeline structs' public fields, like the shim of INTEGRATION.md keeps its handle.
class Device {
 public:
  explicit Device(float lowpass, int device = 0) : Device(lowpass, std::vector<int32_t>{device}) {}
  // several devices: a group context -- the frame is sharded by screen-tile stripes, the scene broadcast at
  // upload, the stripes gathered over NVLink inside splat_render (splat.h, "Multi-GPU"); same pixels
  Device(float lowpass, const std::vector<int32_t> &devices) {
    if (devices.empty()) throw Error(SPLAT_ERR_INVALID, "no device listed");
    splat_config cfg;
    splat_config_default(&cfg);
    cfg.device = devices[0];
    cfg.lowpass = lowpass;
    const int rc = devices.size() == 1 ? splat_create(&ctx_, &cfg)
                                       : splat_create_multi(&ctx_, &cfg, devices.data(), (int32_t)devices.size());
    if (rc != SPLAT_OK) throw Error(rc, std::string("splat_create: ") + splat_create_error());   // no CPU fallback
  }
  Device(const Device &) = delete;
  Device &operator=(const Device &) = delete;
  ~Device() {
    if (pinned_) splat_unpin_host(pinned_);
    if (ctx_) splat_destroy(ctx_);
  }
  // pin the caller's colour buffer once (cudaHostRegister) so that the per-frame copies of splat_render run at
  // full PCIe rate; pinned again when the buffer was re-allocated.  Failure to pin is not an error.
  void pin(void *p, size_t bytes) {
    if (p == pinned_ && bytes == pinned_bytes_) return;
    if (pinned_) splat_unpin_host(pinned_);
    pinned_ = splat_pin_host(p, bytes) == SPLAT_OK ? p : nullptr;
    pinned_bytes_ = pinned_ ? bytes : 0;
  }

  splat_ctx *ctx() { return ctx_; }
  void check(int rc, const char *what) {
    if (rc != SPLAT_OK) throw Error(rc, std::string(what) + ": " + splat_last_error(ctx_));
  }
  // the scene is uploaded when (pointer, length) of the caller's storage changes, or after invalidate()
  bool holds(const void *p, size_t n) const { return valid_ && p == scene_ptr_ && n == scene_len_; }
  void now_holds(const void *p, size_t n) { scene_ptr_ = p; scene_len_ = n; valid_ = true; }
  void invalidate() { valid_ = false; }
  splat_timings timings() {
    splat_timings t;
    check(splat_get_timings(ctx_, &t), "splat_get_timings");
    return t;
  }

 private:
  splat_ctx *ctx_ = nullptr;
  const void *scene_ptr_ = nullptr;
  size_t scene_len_ = 0;
  bool valid_ = false;
  void *pinned_ = nullptr;
  size_t pinned_bytes_ = 0;
};

// GaussianSplatPipeline01, pipelines.rs:54-169: `gaussians: Vec<Gaussian>`, low-pass +0.01 (gaussians.rs:156-157)
class GaussianSplatPipeline01 {
 public:
  std::vector<Gaussian> gaussians;
  Camera camera;
  GaussianSplatPipeline01(std::vector<Gaussian> g, Camera c, int device = 0)
      : gaussians(std::move(g)), camera(std::move(c)), dev_(0.01f, device) {}
  GaussianSplatPipeline01(std::vector<Gaussian> g, Camera c, const std::vector<int32_t> &devices)
      : gaussians(std::move(g)), camera(std::move(c)), dev_(0.01f, devices) {}
  // pipelines.rs:66-86.  `color` is blended onto and overwritten.
  void render_to_buffer(Buffer2d<uint32_t> &color) {
    if (!dev_.holds(gaussians.data(), gaussians.size())) {
      dev_.check(splat_upload_aos(dev_.ctx(), reinterpret_cast<const float *>(gaussians.data()), gaussians.size()), "splat_upload_aos");
      dev_.now_holds(gaussians.data(), gaussians.size());
    }
    const splat_camera cam = camera_struct(camera);
    const auto sz = color.size();
    dev_.pin(color.raw_mut(), sz[0] * sz[1] * sizeof(uint32_t));
    dev_.check(splat_render(dev_.ctx(), &cam, color.raw_mut(), (uint32_t)sz[0], (uint32_t)sz[1]), "splat_render");
  }
  void scene_changed() { dev_.invalidate(); }   // after mutating `gaussians` in place
  splat_timings timings() { return dev_.timings(); }

 private:
  Device dev_;
};

// GaussianSplatPipeline02, pipelines.rs:172-281: `gaussians: GaussianList`, low-pass +0.3 (gaussians.rs:517-518)
class GaussianSplatPipeline02 {
 public:
  GaussianList gaussians;
  Camera camera;
  GaussianSplatPipeline02(GaussianList g, Camera c, int device = 0) : gaussians(std::move(g)), camera(std::move(c)), dev_(0.3f, device) {}
  GaussianSplatPipeline02(GaussianList g, Camera c, const std::vector<int32_t> &devices)
      : gaussians(std::move(g)), camera(std::move(c)), dev_(0.3f, devices) {}
  // pipelines.rs:260-280
  void render_to_buffer(Buffer2d<uint32_t> &color) {
    upload_if_needed();
    const splat_camera cam = camera_struct(camera);
    const auto sz = color.size();
    dev_.pin(color.raw_mut(), sz[0] * sz[1] * sizeof(uint32_t));
    dev_.check(splat_render(dev_.ctx(), &cam, color.raw_mut(), (uint32_t)sz[0], (uint32_t)sz[1]), "splat_render");
  }
  // `color.fill(clear); render_to_buffer(&mut color)` (main.rs:73-74) in one call: no host fill, no upload of it
  void render_cleared_to_buffer(Buffer2d<uint32_t> &color, uint32_t clear = 0) {
    upload_if_needed();
    const splat_camera cam = camera_struct(camera);
    const auto sz = color.size();
    dev_.pin(color.raw_mut(), sz[0] * sz[1] * sizeof(uint32_t));
    dev_.check(splat_render_cleared(dev_.ctx(), &cam, color.raw_mut(), (uint32_t)sz[0], (uint32_t)sz[1], clear), "splat_render_cleared");
  }
  void scene_changed() { dev_.invalidate(); }
  splat_timings timings() { return dev_.timings(); }

 private:
  void upload_if_needed() {
    const GaussianList &g = gaussians;
    if (dev_.holds(g.positions.data(), g.num_gaussians)) return;
    if (g.positions.size() != 4 * g.num_gaussians || g.scales.size() != 3 * g.num_gaussians || g.opacities.size() != g.num_gaussians ||
        g.rotations.size() != 4 * g.num_gaussians || g.sh.size() != 48 * g.num_gaussians)
      throw Error(SPLAT_ERR_INVALID, "GaussianList: array sizes do not match num_gaussians");
    dev_.check(splat_upload_soa(dev_.ctx(), g.positions.data(), g.scales.data(), g.opacities.data(), g.rotations.data(), g.sh.data(),
                                g.num_gaussians), "splat_upload_soa");
    dev_.now_holds(g.positions.data(), g.num_gaussians);
  }
  Device dev_;
};

}  // namespace splat_b200
#endif  // SPLAT_B200_PIPELINE_HPP

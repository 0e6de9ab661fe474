// my_function.cpp -- host side of the segmentation calls (see my_function.h) + the CLI helpers
// Split / analyse_path (reference tmc3/my_function.cpp:147-178: argv[1] and argv[2], text after '=').
// extracted_contour (my_function.cpp:8-145, OpenCV) is outside the hot path and not provided.
#include "my_function.h"

#include <cstring>
#include <fstream>
#include <iostream>

#include "png_read.h"

#include <cstring>

namespace bseg_host {
namespace {
bseg_ctx* g_ctx = nullptr;
bseg_params g_params;
bool g_params_init = false;
const int32_t* g_cloud_ptr = nullptr;
size_t g_cloud_n = 0;
uint64_t g_cloud_sum = 0;       // checksum of the positions the device holds (the caller may edit the cloud in place)
bool g_cloud_shifted = false;   // the host cloud carries the shifted coordinates (buildingSeg ran on it)
int32_t g_cloud_min[3] = {0, 0, 0};

uint64_t checksum(const int32_t* p, size_t n)
{
  uint64_t a = 1469598103934665603ull, b = 0;
  for (size_t i = 0; i < 3 * n; ++i) {
    a += (uint32_t)p[i];
    b += a;
  }
  return a ^ (b << 1);
}
}  // namespace

void check(int rc, const char* what)
{
  if (rc != 0)
    throw std::runtime_error(std::string(what) + ": " + bseg_last_error(g_ctx));
}

bseg_params& params()
{
  if (!g_params_init) {
    bseg_default_params(&g_params);
    g_params_init = true;
  }
  return g_params;
}

bseg_ctx* context()
{
  if (!g_ctx) {
    int dev = 0;
    if (const char* e = std::getenv("BSEG_DEVICE")) dev = std::atoi(e);
    int rc = bseg_create(&g_ctx, dev);
    if (rc != 0)
      throw std::runtime_error(std::string("bseg_create: ") + bseg_last_error(nullptr));
  }
  return g_ctx;
}

bool ensure_cloud(PCCPointSet3& cloud, bool shift_caller_cloud)
{
  const size_t n = cloud.getPointCount();
  const uint64_t sum = checksum(cloud.positionData(), n);
  if (!shift_caller_cloud && g_cloud_ptr == cloud.positionData() && g_cloud_n == n && g_cloud_sum == sum && n > 0)
    return false;
  int32_t mx[3];
  check(bseg_set_points(context(), cloud.positionData(), (int64_t)n, g_cloud_min, mx,
                        shift_caller_cloud ? cloud.positionData() : nullptr),
        "bseg_set_points");
  g_cloud_ptr = cloud.positionData();
  g_cloud_n = n;
  g_cloud_shifted = shift_caller_cloud;
  g_cloud_sum = shift_caller_cloud ? checksum(cloud.positionData(), n) : sum;
  return true;
}

void note_shifted(const PCCPointSet3& cloud)
{
  g_cloud_ptr = cloud.positionData();
  g_cloud_n = cloud.getPointCount();
  g_cloud_sum = checksum(cloud.positionData(), g_cloud_n);
  g_cloud_shifted = true;
}

// seg_plane works in the coordinates of the cloud it is given (my_function.cpp:190,227,242-250): when that cloud is
// not the shifted one, the device adds the subtracted minimum back for the grower's int32 arithmetic
const int32_t* grow_offset() { return g_cloud_shifted ? nullptr : g_cloud_min; }

}  // namespace bseg_host

std::vector<plane> seg_plane::get_planes()
{
  using namespace bseg_host;
  const size_t n = Cloud.getPointCount();
  ensure_cloud(Cloud, false);
  bseg_params p = params();
  p.K = K;
  p.th_thickness = th_thickness;
  p.th_point_count = th_pointCount;
  // the caller's vectors are authoritative (the reference's constructor accepts arbitrary ones)
  std::vector<int32_t> flat(n * (size_t)K, -1);
  for (size_t i = 0; i < n; ++i)
    for (size_t k = 0; k < Neigh[i].size() && k < (size_t)K; ++k) flat[i * K + k] = Neigh[i][k];
  check(bseg_override_neigh_normals(context(), &p, flat.data(), n ? &Normal[0][0] : nullptr),
        "bseg_override_neigh_normals");
  check(bseg_set_grow_offset(context(), grow_offset()), "bseg_set_grow_offset");
  int32_t np = 0;
  std::vector<int32_t> label(n);
  Cloud.planeIdx.resize(n);
  check(bseg_grow_planes(context(), &p, n ? Cloud.planeIdx.data() : nullptr, label.data(), &np), "bseg_grow_planes");
  std::vector<int32_t> seeds(np);
  std::vector<double> normals((size_t)np * 3);
  std::vector<int32_t> centers((size_t)np * 3);
  std::vector<int64_t> off((size_t)np + 1);
  check(bseg_get_planes(context(), seeds.data(), normals.data(), centers.data(), off.data(), nullptr), "bseg_get_planes");
  std::vector<int32_t> idx((size_t)off[np]);
  check(bseg_get_planes(context(), seeds.data(), normals.data(), centers.data(), off.data(), idx.data()),
        "bseg_get_planes");
  std::vector<plane> planes((size_t)np);
  for (int q = 0; q < np; ++q) {
    plane& pl = planes[q];
    pl.id = q + 1;
    pl.normal = Vec3<double>(normals[3 * q], normals[3 * q + 1], normals[3 * q + 2]);
    pl.center = Vec3<int>(centers[3 * q], centers[3 * q + 1], centers[3 * q + 2]);
    pl.pointIdx.assign(idx.begin() + off[q], idx.begin() + off[q + 1]);
  }
  return planes;
}

bool seg_plane::Broad(int, int)
{
  throw std::runtime_error("seg_plane::Broad runs on the device inside get_planes(); it has no host implementation");
}

namespace bseg_host {
static std::vector<uint16_t> g_last_rgb;
const std::vector<uint16_t>& last_plane_rgb() { return g_last_rgb; }
}  // namespace bseg_host

void seg_plane::set_plane_color(std::vector<plane>& planes)
{
  using namespace bseg_host;
  if (!Cloud.hasColors())
    Cloud.addColors();  // the reference requires colours (PCCPointSet.h:289-293 asserts)
  // my_function.cpp:260-275 paints the pointIdx of exactly the planes it is handed, in their order: the caller may
  // have filtered or reordered the vector.  The device holds the lists of the last get_planes(); a plane is named
  // by its id, and a plane whose pointIdx was edited on the host is refused (nothing here paints on the CPU).
  const int32_t device_planes = bseg_plane_count(context());
  std::vector<int64_t> off((size_t)(device_planes > 0 ? device_planes : 0) + 1, 0);
  if (device_planes > 0)
    check(bseg_get_planes(context(), nullptr, nullptr, nullptr, off.data(), nullptr), "bseg_get_planes");
  std::vector<int32_t> ids(planes.size());
  std::vector<uint16_t> rgb(planes.size() * 3);
  bool canonical = (int64_t)planes.size() == (int64_t)device_planes;
  for (size_t q = 0; q < planes.size(); ++q) {
    const int id = planes[q].id;
    if (id < 1 || id > device_planes)
      throw std::runtime_error("set_plane_color: plane id " + std::to_string(id) + " is not a plane of the last get_planes()");
    if ((int64_t)planes[q].pointIdx.size() != off[id] - off[id - 1])
      throw std::runtime_error("set_plane_color: pointIdx of plane " + std::to_string(id) + " was modified on the host");
    ids[q] = id;
    canonical = canonical && id == (int)q + 1;
    for (int k = 0; k < 3; ++k) rgb[3 * q + k] = uint16_t(55 + rand() % 200);  // braced-init order, :268
  }
  check(bseg_paint(context(), canonical ? nullptr : ids.data(), (int32_t)planes.size(), rgb.data(), Cloud.colorData()),
        "bseg_paint");
  // colour table by plane id for the label image (black for planes that were not listed)
  g_last_rgb.assign((size_t)(device_planes > 0 ? device_planes : 0) * 3, 0);
  for (size_t q = 0; q < planes.size(); ++q)
    for (int k = 0; k < 3; ++k) g_last_rgb[3 * (size_t)(ids[q] - 1) + k] = rgb[3 * q + k];
}

vector<string> Split(const string& s, const string& seperator)
{
  // my_function.cpp:147-160: cut at every occurrence of the whole separator STRING; empty tokens are kept and
  // the remainder (possibly empty) is always the last token
  vector<string> out;
  size_t pos = 0;
  for (;;) {
    const size_t next = seperator.empty() ? string::npos : s.find(seperator, pos);
    if (next == string::npos)
      break;
    out.push_back(s.substr(pos, next - pos));
    pos = next + seperator.length();
  }
  out.push_back(s.substr(pos));
  return out;
}

param analyse_path(char* argv[])
{
  param p;
  p.frame = 0;
  // my_function.cpp:163-178: flag names are ignored, the path is token [1] of Split(arg, "=") -- the text between
  // the first and the second '='.  The reference indexes [1] unchecked; an argument without '=' is an error here.
  auto value = [](const char* a) {
    const vector<string> t = Split(string(a ? a : ""), "=");
    if (t.size() < 2)
      throw std::runtime_error("tmc3: expected -<flag>=<path>, got '" + string(a ? a : "") + "'");
    return t[1];
  };
  p.readPath = value(argv[1]);
  p.savePath = value(argv[2]);
  return p;
}

// DetectedPlane records (my_function.h:41-46) of the planes get_planes returned, classes from bseg_plane_classes
std::vector<DetectedPlane> detect_planes(const std::vector<plane>& planes, double ground_z, double facade_max_nz,
                                         double roof_min_nz, std::vector<uint8_t>* point_class)
{
  bseg_ctx* ctx = bseg_host::context();
  const int32_t P = bseg_plane_count(ctx);
  if (P < 0 || (size_t)P != planes.size())
    throw std::runtime_error("detect_planes: the plane vector is not the one the last get_planes returned");
  std::vector<double> eq((size_t)P * 4);
  std::vector<uint8_t> cls((size_t)P);
  size_t n = 0;
  if (point_class) {
    n = (size_t)bseg_point_count(ctx);
    point_class->assign(n, 0);
  }
  bseg_host::check(bseg_plane_classes(ctx, facade_max_nz, roof_min_nz, ground_z, eq.data(), cls.data(),
                                      point_class && n ? point_class->data() : nullptr),
                   "bseg_plane_classes");
  std::vector<DetectedPlane> out((size_t)P);
  for (size_t i = 0; i < out.size(); ++i) {
    DetectedPlane& d = out[i];
    d.indices.assign(planes[i].pointIdx.begin(), planes[i].pointIdx.end());
    for (int k = 0; k < 4; ++k) d.equation[k] = eq[4 * i + k];
    d.normal = planes[i].normal;
    d.d = eq[4 * i + 3];
    d.cls = cls[i];
  }
  return out;
}

// ---- extracted_contour (my_function.cpp:8-145) ---------------------------------------------------------------------
void extracted_contour(string read_path, string save_path, string flip)
{
  std::vector<uint8_t> src;  // R,G,B (cv::imread keeps B,G,R; channel 1 is green either way)
  int cols = 0, rows = 0;
  std::string why;
  if (!bseg_png::read_rgb(read_path, src, cols, rows, &why))
    throw std::runtime_error("extracted_contour: " + read_path + ": " + why);  // (:10-12 prints and then crashes on the empty Mat)
  bseg_ctx* ctx = bseg_host::context();
  // :17-26 extractChannel(src, 1), threshold(10, 255, THRESH_BINARY), morphologyEx(MORPH_CLOSE, ellipse 5x5, iterations 2)
  std::vector<uint8_t> morphed((size_t)cols * rows);
  bseg_host::check(bseg_contour_mask(ctx, src.data(), cols, rows, 3, 1, 10, 2, morphed.data()), "bseg_contour_mask");
  // :30-33 findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE)
  int64_t nc = 0, np = 0;
  bseg_host::check(bseg_find_contours(morphed.data(), cols, rows, 1, nullptr, 0, nullptr, 0, &nc, &np), "bseg_find_contours");
  std::vector<int32_t> pts((size_t)(np > 0 ? np : 1) * 2);
  std::vector<int64_t> off((size_t)nc + 1);
  bseg_host::check(bseg_find_contours(morphed.data(), cols, rows, 1, pts.data(), np, off.data(), nc, &nc, &np), "bseg_find_contours");
  // :36-59 the contours that look like buildings, drawn into a copy of the image
  std::vector<uint8_t> result(src);
  const uint8_t yellow_bgr_as_rgb[3] = {0, 255, 255};  // Scalar(255, 255, 0) is B, G, R
  for (int64_t k = 0; k < nc; ++k) {
    double area = 0, perimeter = 0;
    bseg_host::check(bseg_contour_measure(pts.data() + 2 * off[k], off[k + 1] - off[k], &area, &perimeter), "bseg_contour_measure");
    if (area > 500 && perimeter > 100)
      bseg_host::check(bseg_draw_contour(result.data(), cols, rows, 3, pts.data() + 2 * off[k], off[k + 1] - off[k], yellow_bgr_as_rgb),
                       "bseg_draw_contour");
  }
  // :64-128 csa.obj: every contour (not only the drawn ones), normalised to [0,1], y flipped, extruded to z = 0 and 1.
  // The three comment lines are the reference's, GBK bytes as they stand in its source file.
  {
    std::ofstream obj("csa.obj");
    if (!obj.is_open()) {
      std::cerr << "extracted_contour: cannot create csa.obj" << std::endl;
      return;  // (:66-69)
    }
    obj << "# \xb4\xd3\xc2\xd6\xc0\xaa\xc9\xfa\xb3\xc9\xb5\xc4" "3D\xc4\xa3\xd0\xcd" << std::endl;
    obj << "# \xc2\xd6\xc0\xaa\xca\xfd\xc1\xbf: " << (size_t)nc << std::endl;
    obj << "# \xb6\xa5\xb5\xe3\xb9\xe9\xd2\xbb\xbb\xaf\xb5\xbd\xb7\xb6\xce\xa7 [0,1] (x,y)" << std::endl << std::endl;
    int vertexIndex = 1;
    std::vector<std::vector<int>> groups;
    for (int64_t k = 0; k < nc; ++k) {
      std::vector<int> group;
      for (int64_t e = off[k]; e < off[k + 1]; ++e) {
        const float x = static_cast<float>(pts[2 * e]) / cols;
        const float y = 1.0f - static_cast<float>(pts[2 * e + 1]) / rows;
        obj << "v " << x << " " << y << " 0.0" << std::endl;
        group.push_back(vertexIndex++);
        obj << "v " << x << " " << y << " " << 1 << std::endl;
        group.push_back(vertexIndex++);
      }
      groups.push_back(group);
    }
    obj << std::endl << "# \xb2\xe0\xc3\xe6 (\xcb\xc4\xb1\xdf\xd0\xce\xc3\xe6)" << std::endl;
    for (const std::vector<int>& v : groups) {
      const int n = (int)v.size() / 2;
      for (int i = 0; i < n; ++i) {
        const int next = (i + 1) % n;
        obj << "f " << v[i * 2] << " " << v[next * 2] << " " << v[next * 2 + 1] << " " << v[i * 2 + 1] << std::endl;
      }
    }
  }
  // :141-144 imwrite(save_path, result); flip(result, 0); imwrite(flip, flipped)
  bseg_host::check(bseg_png_write(save_path.c_str(), result.data(), cols, rows, 3, 0), "bseg_png_write");
  std::vector<uint8_t> flipped(result.size());
  for (int y = 0; y < rows; ++y)
    memcpy(&flipped[(size_t)y * cols * 3], &result[(size_t)(rows - 1 - y) * cols * 3], (size_t)cols * 3);
  bseg_host::check(bseg_png_write(flip.c_str(), flipped.data(), cols, rows, 3, 0), "bseg_png_write");
}

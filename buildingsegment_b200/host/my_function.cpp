// my_function.cpp -- host side of the segmentation calls (see my_function.h) + the CLI helpers
// Split / analyse_path (reference tmc3/my_function.cpp:147-178: argv[1] and argv[2], text after '=').
// extracted_contour (my_function.cpp:8-145, OpenCV) is outside the hot path and not provided.
#include "my_function.h"

#include <cstring>

namespace bseg_host {
namespace {
bseg_ctx* g_ctx = nullptr;
bseg_params g_params;
bool g_params_init = false;
const int32_t* g_cloud_ptr = nullptr;
size_t g_cloud_n = 0;
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
  if (!shift_caller_cloud && g_cloud_ptr == cloud.positionData() && g_cloud_n == n && n > 0)
    return false;
  int32_t mn[3], mx[3];
  check(bseg_set_points(context(), cloud.positionData(), (int64_t)n, mn, mx,
                        shift_caller_cloud ? cloud.positionData() : nullptr),
        "bseg_set_points");
  g_cloud_ptr = cloud.positionData();
  g_cloud_n = n;
  return true;
}

void note_shifted(const PCCPointSet3& cloud)
{
  g_cloud_ptr = cloud.positionData();
  g_cloud_n = cloud.getPointCount();
}

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
  const size_t n = Cloud.getPointCount();
  if (!Cloud.hasColors())
    Cloud.addColors();  // the reference requires colours (PCCPointSet.h:289-293 asserts)
  std::vector<uint16_t> rgb(planes.size() * 3);
  for (size_t q = 0; q < planes.size(); ++q)
    for (int k = 0; k < 3; ++k) rgb[3 * q + k] = uint16_t(55 + rand() % 200);  // braced-init order, :268
  check(bseg_paint(context(), rgb.data(), Cloud.colorData()), "bseg_paint");
  g_last_rgb = rgb;
  (void)n;
}

vector<string> Split(const string& s, const string& seperator)
{
  vector<string> out;
  size_t pos = 0;
  while (pos <= s.size()) {
    size_t next = s.find_first_of(seperator, pos);
    if (next == string::npos) next = s.size();
    if (next > pos) out.push_back(s.substr(pos, next - pos));
    pos = next + 1;
  }
  return out;
}

param analyse_path(char* argv[])
{
  param p;
  p.frame = 0;
  // flag names are ignored: only the text after '=' of argv[1] (input) and argv[2] (output) counts
  auto value = [](const char* a) {
    string s(a ? a : "");
    size_t eq = s.find('=');
    return eq == string::npos ? s : s.substr(eq + 1);
  };
  p.readPath = value(argv[1]);
  p.savePath = value(argv[2]);
  return p;
}

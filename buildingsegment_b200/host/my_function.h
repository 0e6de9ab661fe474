// my_function.h -- the reference's segmentation calls (tmc3/my_function.h:16-123) as thin host
// shims over the C ABI of libbseg.so (include/bseg.h).  Same names, same argument meaning:
//
//   get_Normal_and_K_neighbor<K>(cloud, normal, neigh)   my_function.h:48-85   -> bseg_knn_normals
//   seg_plane(cloud, normal, neigh, K)                   my_function.h:98-104
//   seg_plane::get_planes()                              my_function.cpp:180-217 -> bseg_grow_planes
//   seg_plane::set_plane_color(planes)                   my_function.cpp:260-275 -> bseg_paint
//
// No Open3D, no OpenCV, no CPU segmentation code: every call lands on the GPU through the ABI and
// throws std::runtime_error (carrying bseg_last_error) when the library reports a failure.
#pragma once
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "bseg.h"
#include "ply.h"

using namespace pcc;
using namespace std;

class param {
public:
  string dataType;
  string readPath;
  string savePath;
  string num_sizePath;
  int frame;
};

struct plane {
  int id;  // > 0
  Vec3<double> normal;
  Vec3<int> center;
  std::vector<int> pointIdx;
};

// my_function.h:41-46: the plane record the reference declares and never fills (Eigen types there; plain arrays here,
// this build has no Eigen).  detect_planes() fills it from the planes of get_planes: equation ax + by + cz + d = 0 with
// (a, b, c) = plane::normal and d = -(normal . center); `cls` is the roof / facade / ground class (bseg_plane_classes).
struct DetectedPlane {
  std::vector<size_t> indices;  // points of the plane (plane::pointIdx)
  double equation[4];           // ax + by + cz + d = 0
  Vec3<double> normal;
  double d;
  int cls;                      // BSEG_CLASS_*
};

// The planes of the last seg_plane::get_planes as DetectedPlane records, classified on the device;
// ground_z: e.g. buildingSeg::groundTH() (TMC3.cpp:181-198).  point_class (optional): class per point.
std::vector<DetectedPlane> detect_planes(const std::vector<plane>& planes, double ground_z, double facade_max_nz = 0.3,
                                         double roof_min_nz = 0.7, std::vector<uint8_t>* point_class = nullptr);

// my_function.cpp:8-145: building outlines from the count image ("像素数量.png" of save_image).  Green channel > 10,
// closed with the 5x5 ellipse twice (on the device: bseg_contour_mask), external contours (bseg_find_contours), those with
// area > 500 and perimeter > 100 drawn 2 pixels thick in Scalar(255,255,0) into a copy of the image that is written to
// save_path and, flipped vertically, to `flip`; EVERY contour is extruded to z in {0, 1} in "csa.obj" (current
// directory, as in the reference).  No OpenCV: PNG in through host/png_read.cpp, PNG out through bseg_png_write.
void extracted_contour(string read_path, string save_path, string flip);

param analyse_path(char* argv[]);
vector<string> Split(const string& s, const string& seperator);

namespace bseg_host {

// One process-wide libbseg context, as the reference is one process / one thread (TMC3.cpp:202-229).
bseg_ctx* context();
bseg_params& params();  // defaults = the reference's literals; change before the first call
void check(int rc, const char* what);

// Uploads `cloud` unless it is the cloud the device already holds (same storage, same size, and the
// positions were last written by us).  Returns true when an upload happened.
bool ensure_cloud(PCCPointSet3& cloud, bool shift_caller_cloud);
void note_shifted(const PCCPointSet3& cloud);
const int32_t* grow_offset();  // minimum the device subtracted from an unshifted cloud, NULL for a shifted one

// the colour sequence the last set_plane_color drew (55 + rand() % 200, three per plane): the label image uses it
const std::vector<uint16_t>& last_plane_rgb();

// device results -> the reference's host containers
void fetch_knn(size_t n, int K, std::vector<Vec3<double>>& normal, std::vector<std::vector<int>>& neigh);

}  // namespace bseg_host

template <int K>
void get_Normal_and_K_neighbor(PCCPointSet3& pointCloud, std::vector<Vec3<double>>& normal,
                               std::vector<vector<int>>& neigh)
{
  bseg_host::ensure_cloud(pointCloud, false);
  bseg_params p = bseg_host::params();
  p.K = K;
  bseg_host::params().K = K;
  // hybrid-radius normals oriented to +z, then the K nearest of every point (my_function.h:63-78)
  const size_t n = pointCloud.getPointCount();
  normal.resize(n);
  neigh.resize(n);
  std::vector<int32_t> flat(n * K);
  static_assert(sizeof(Vec3<double>) == 24, "Vec3<double> must be 3 packed doubles");
  bseg_host::check(bseg_knn_normals(bseg_host::context(), &p, flat.data(), n ? &normal[0][0] : nullptr, nullptr),
                   "bseg_knn_normals");
  for (size_t i = 0; i < n; ++i) {
    // SearchKNN returns min(K, N) indices
    size_t m = K;
    while (m > 0 && flat[i * K + m - 1] < 0) --m;
    neigh[i].assign(flat.begin() + i * K, flat.begin() + i * K + m);
  }
  // the reference also dumps a debug PLY to a hard-coded Windows path (my_function.h:81): dropped
}

class seg_plane {
public:
  seg_plane(PCCPointSet3& pointCloud, std::vector<Vec3<double>>& normal, std::vector<std::vector<int>>& neigh,
            int num_neigh)
    : Cloud(pointCloud), Normal(normal), Neigh(neigh), K(num_neigh)
  {
    Cloud.planeIdx.resize(Cloud.getPointCount(), -1);
  }

  // my_function.cpp:180-217: the committed planes, ids 1..P in seed order; Cloud.planeIdx is left as
  // the reference leaves it (orphan marks included)
  std::vector<plane> get_planes();

  // my_function.cpp:220-258 runs inside get_planes on the device; kept for source compatibility
  bool Broad(int Idx, int depth);

  // my_function.cpp:260-275: black everything, then 55 + rand() % 200 three times per plane
  void set_plane_color(std::vector<plane>& planes);

private:
  PCCPointSet3& Cloud;
  std::vector<Vec3<double>>& Normal;
  std::vector<std::vector<int>>& Neigh;
  int K;
  int th_thickness = 300;   // my_function.h:117
  int th_pointCount = 400;  // my_function.h:118
};

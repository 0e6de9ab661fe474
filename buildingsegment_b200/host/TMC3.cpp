// TMC3.cpp -- the tmc3 driver with the reference's entry points (tmc3/TMC3.cpp:44-229):
// struct Box, class buildingSeg (constructor, compute_gird_picture, save_image, pixel, groundTH)
// and main.  Same CLI: `tmc3 -a=<in.ply> -s=<out.ply>` (flag names ignored, readme.txt:12).
// Every computation runs on the GPU through libbseg's C ABI; this file only marshals.
#include <cstdio>
#include <cstring>
#include <iostream>
#include <memory>
#include <vector>

#include <algorithm>

#include "my_function.h"

struct Box {
  Vec3<int> min = std::numeric_limits<int32_t>::max();
  Vec3<int> max = std::numeric_limits<int32_t>::lowest();
};

namespace {

// The reference hands the three byte images to stbi_write_png (TMC3.cpp:98,108,119).  libbseg's writer produces
// the same bytes (include/bseg.h bseg_png_*) and encodes on worker threads: save_image returns once the pixels are
// copied, the files are complete after bseg_png_wait() (end of main), off the segmentation's critical path.
void write_png_rgb(const std::string& path, int w, int h, const uint8_t* rgb, int stride)
{
  bseg_host::check(bseg_png_write_async(path.c_str(), rgb, w, h, 3, stride), "bseg_png_write_async");
}

}  // namespace

//============================================================================
class buildingSeg {
public:
  PCCPointSet3 pointcloud;

  // TMC3.cpp:55-79: copy the cloud, bounding box, shift BOTH the copy and the caller's cloud to
  // min = 0 (:71), size the image.  The box and the shift come from the device (bseg_set_points).
  buildingSeg(PCCPointSet3& pointcloud)
  {
    int32_t mn[3], mx[3];
    const size_t n = pointcloud.getPointCount();
    bseg_host::check(bseg_set_points(bseg_host::context(), pointcloud.positionData(), (int64_t)n, mn, mx,
                                     pointcloud.positionData()),
                     "bseg_set_points");
    bseg_host::note_shifted(pointcloud);
    this->pointcloud = pointcloud;  // the shifted copy
    for (int k = 0; k < 3; ++k) {
      box.min[k] = mn[k];
      box.max[k] = mx[k];
    }
    width = (box.max[0] - box.min[0]) / bin + 2;
    height = (box.max[1] - box.min[1]) / bin + 2;
    image.resize(size_t(width) * height * channels, 0);
  }

  // TMC3.cpp:81-121: three PNGs; the uint8 conversion (255.0 * v / max, truncated) runs on the device
  void save_image(std::string path)
  {
    const size_t npx = size_t(width) * height;
    if (png.size() != 3 * npx * 3)
      compute_gird_picture();
    // the reference's file names are the GBK bytes of these three Chinese names (:98,:108,:119)
    static const char nameA[] = "\xc6\xbd\xbe\xf9\xb8\xdf\xb6\xc8.png";              // mean height
    static const char nameB[] = "\xcf\xf1\xcb\xd8\xca\xfd\xc1\xbf.png";              // pixel count
    static const char nameC[] = "\xcf\xf1\xcb\xd8\xca\xfd\xc1\xbf+\xb8\xdf\xb6\xc8.png";  // count + height (always black)
    write_png_rgb(path + nameA, width, height, png.data(), width * 3);
    write_png_rgb(path + nameB, width, height, png.data() + 3 * npx, width * 3);
    write_png_rgb(path + nameC, width, height, png.data() + 6 * npx, width * 3);
  }

  // extension (include/bseg.h bseg_label_raster): the plane label of the highest point per pixel, in the colours
  // set_plane_color gave the planes -- "labels.png" on the grid of the three images above
  // (uses the segmentation the device holds: call it after get_planes and before compute_gird_picture, which
  // uploads this object's own copy of the cloud again)
  void save_label_image(std::string path)
  {
    bseg_params p = bseg_host::params();
    p.bin = bin;
    p.bin_height = bin_height;
    const size_t npx = size_t(width) * height;
    std::vector<uint8_t> rgb(3 * npx);
    const std::vector<uint16_t>& prgb = bseg_host::last_plane_rgb();
    bseg_host::check(bseg_label_raster(bseg_host::context(), &p, prgb.empty() ? nullptr : prgb.data(), nullptr, rgb.data()),
                     "bseg_label_raster");
    write_png_rgb(path + "labels.png", width, height, rgb.data(), width * 3);
  }

  double& pixel(int x, int y, int ch) { return image[(size_t(y) * width + x) * channels + ch]; }  // :123-125

  // TMC3.cpp:127-172 (+ groundTH :181-198): median-height cut, bilinear splat in point order, mean
  // height, log count -- on the device, bit-exact with the reference's in-order fp64 sums
  void compute_gird_picture()
  {
    bseg_host::ensure_cloud(pointcloud, false);
    bseg_params p = bseg_host::params();
    p.bin = bin;
    p.bin_height = bin_height;
    const size_t npx = size_t(width) * height;
    png.resize(3 * npx * 3);
    bseg_host::check(bseg_raster(bseg_host::context(), &p, image.data(), png.data(), png.data() + 3 * npx,
                                 png.data() + 6 * npx, &ground_th),
                     "bseg_raster");
  }

  double groundTH()
  {
    compute_gird_picture();
    return ground_th;
  }

  Box box;
  int width = 0, height = 0;
  std::vector<double> image;
  std::vector<uint8_t> png;
  double ground_th = 0.0;
  int bin = 100, bin_height = 1000;  // TMC3.cpp:177
  int channels = 3;                  // TMC3.cpp:178
};

#ifndef TMC3_NO_MAIN
int main(int argc, char* argv[])
{
  if (argc < 3) {
    std::cerr << "usage: tmc3 -a=<in.ply> -s=<out.ply> [--raster=<dir/>] [--contours=<dir/>] [--classes]" << std::endl;
    return 2;
  }
  try {
    param path = analyse_path(argv);
    PCCPointSet3 pointCloud;
    double positionScale = 1000;  // metres -> millimetres, TMC3.cpp:207
    if (!ply::read(path.readPath, {"x", "y", "z"}, positionScale, pointCloud)) {
      std::cerr << "tmc3: cannot read " << path.readPath << std::endl;
      return 1;
    }

    buildingSeg seg = buildingSeg(pointCloud);

    std::vector<Vec3<double>> normal;
    std::vector<vector<int>> neigh;
    get_Normal_and_K_neighbor<15>(pointCloud, normal, neigh);
    seg_plane h = seg_plane(pointCloud, normal, neigh, 15);
    vector<plane> plances = h.get_planes();
    h.set_plane_color(plances);

    ply::write(pointCloud, {"x", "y", "z"}, 1.00, {0, 0, 0}, path.savePath, false);

    // the plane classes the data structures hint at (my_function.h:41-46), while the device still holds the planes
    // (the raster below uploads buildingSeg's own copy of the cloud); ground level = groundTH's rule (TMC3.cpp:181-198)
    for (int a = 3; a < argc; ++a)
      if (std::strcmp(argv[a], "--classes") == 0) {
        const size_t n = pointCloud.getPointCount();
        int zmax = 0;
        for (size_t i = 0; i < n; ++i) zmax = std::max(zmax, (int)pointCloud[i][2]);
        std::vector<size_t> hist((size_t)zmax / 1000 + 1, 0);
        for (size_t i = 0; i < n; ++i) hist[(size_t)pointCloud[i][2] / 1000]++;
        size_t total = 0, bin = 0;
        for (; bin < hist.size(); ++bin) {
          total += hist[bin];
          if (total > n / 2)
            break;
        }
        std::vector<uint8_t> cls;
        const std::vector<DetectedPlane> dp = detect_planes(plances, 1000.0 * (double)bin, 0.3, 0.7, &cls);
        size_t per[5] = {0, 0, 0, 0, 0};
        for (uint8_t k : cls) per[k < 5 ? k : 4]++;
        std::cout << "tmc3: " << dp.size() << " planes; points: roof " << per[BSEG_CLASS_ROOF] << ", facade " << per[BSEG_CLASS_FACADE]
                  << ", ground " << per[BSEG_CLASS_GROUND] << ", other " << per[BSEG_CLASS_OTHER] << ", unlabelled " << per[0]
                  << std::endl;
      }

    // the raster path is compiled but commented out of the reference's main (TMC3.cpp:223-226);
    // here it is opt-in
    for (int a = 3; a < argc; ++a)
      if (std::strncmp(argv[a], "--raster=", 9) == 0) {
        seg.save_label_image(std::string(argv[a] + 9));
        seg.compute_gird_picture();
        seg.save_image(std::string(argv[a] + 9));
      }
    bseg_host::check(bseg_png_wait(), "bseg_png_wait");  // the images encoded on worker threads are on disk now
    // TMC3.cpp:226 (commented out there): outlines from the count image save_image has just written; and the plane
    // classes the data structures hint at (my_function.h:41-46)
    for (int a = 3; a < argc; ++a) {
      if (std::strncmp(argv[a], "--contours=", 11) == 0) {
        const std::string base(argv[a] + 11);
        extracted_contour(base + "\xcf\xf1\xcb\xd8\xca\xfd\xc1\xbf.png", base + "extracted_contours.png",
                          base + "extracted_contours_flip.png");
      }
    }
    std::cout << "tmc3: " << pointCloud.getPointCount() << " points, " << plances.size() << " planes -> "
              << path.savePath << std::endl;
  } catch (const std::exception& e) {
    std::cerr << "tmc3: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
#endif

// TMC3.cpp -- the tmc3 driver with the reference's entry points (tmc3/TMC3.cpp:44-229):
// struct Box, class buildingSeg (constructor, compute_gird_picture, save_image, pixel, groundTH)
// and main.  Same CLI: `tmc3 -a=<in.ply> -s=<out.ply>` (flag names ignored, readme.txt:12).
// Every computation runs on the GPU through libbseg's C ABI; this file only marshals.
#include <zlib.h>

#include <cstdio>
#include <cstring>
#include <iostream>
#include <memory>
#include <vector>

#include "my_function.h"

struct Box {
  Vec3<int> min = std::numeric_limits<int32_t>::max();
  Vec3<int> max = std::numeric_limits<int32_t>::lowest();
};

namespace {

// minimal PNG writer (8-bit RGB, zlib deflate): the reference hands the same pixel buffers to
// stbi_write_png (TMC3.cpp:98,108,119); the decoded images are identical, the file bytes need not be
void put32(std::vector<uint8_t>& v, uint32_t x)
{
  for (int s = 24; s >= 0; s -= 8) v.push_back(uint8_t(x >> s));
}

void chunk(std::vector<uint8_t>& out, const char tag[4], const std::vector<uint8_t>& data)
{
  put32(out, uint32_t(data.size()));
  std::vector<uint8_t> body(tag, tag + 4);
  body.insert(body.end(), data.begin(), data.end());
  out.insert(out.end(), body.begin(), body.end());
  put32(out, uint32_t(crc32(0L, body.data(), uInt(body.size()))));
}

bool write_png_rgb(const std::string& path, int w, int h, const uint8_t* rgb, int stride)
{
  std::vector<uint8_t> raw;
  raw.reserve(size_t(h) * (size_t(w) * 3 + 1));
  for (int y = 0; y < h; ++y) {
    raw.push_back(0);  // filter: none
    raw.insert(raw.end(), rgb + size_t(y) * stride, rgb + size_t(y) * stride + size_t(w) * 3);
  }
  uLongf clen = compressBound(uLong(raw.size()));
  std::vector<uint8_t> comp(clen);
  if (compress2(comp.data(), &clen, raw.data(), uLong(raw.size()), 6) != Z_OK)
    return false;
  comp.resize(clen);
  std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  std::vector<uint8_t> ihdr;
  put32(ihdr, uint32_t(w));
  put32(ihdr, uint32_t(h));
  const uint8_t tail[5] = {8, 2, 0, 0, 0};  // 8 bit, colour type 2 (RGB)
  ihdr.insert(ihdr.end(), tail, tail + 5);
  chunk(out, "IHDR", ihdr);
  chunk(out, "IDAT", comp);
  chunk(out, "IEND", {});
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f)
    return false;
  const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
  return (std::fclose(f) == 0) && ok;
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
    std::cerr << "usage: tmc3 -a=<in.ply> -s=<out.ply> [--raster=<dir/>]" << std::endl;
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

    // the raster path is compiled but commented out of the reference's main (TMC3.cpp:223-226);
    // here it is opt-in
    for (int a = 3; a < argc; ++a)
      if (std::strncmp(argv[a], "--raster=", 9) == 0) {
        seg.save_label_image(std::string(argv[a] + 9));
        seg.compute_gird_picture();
        seg.save_image(std::string(argv[a] + 9));
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

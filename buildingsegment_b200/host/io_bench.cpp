// io_bench.cpp -- times the host I/O either side of the segmentation path, which bench.py reports in its `io`
// block next to (never inside) the points/s figures ("PLY/PNG I/O reported separately"):
//   ply::read   (reference tmc3/ply.cpp:190-504; here a bulk block decoder behind the same signature)
//   ply::write  (ply.cpp:88-186: binary float64 xyz + uchar g,b,r)
//   the PNG encode of one W x H x 3 byte image (stbi_write_png, TMC3.cpp:98 -> bseg_png_encode, same bytes)
// usage: io_bench <in.ply> <out.ply> <png_w> <png_h>     -> one JSON line
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "bseg.h"
#include "ply.h"

using namespace pcc;

static double now()
{
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv)
{
  if (argc < 5) {
    std::fprintf(stderr, "usage: io_bench <in.ply> <out.ply> <png_w> <png_h>\n");
    return 2;
  }
  PCCPointSet3 cloud;
  double t0 = now();
  if (!ply::read(argv[1], {"x", "y", "z"}, 1000.0, cloud)) {
    std::fprintf(stderr, "io_bench: cannot read %s\n", argv[1]);
    return 1;
  }
  const double read_s = now() - t0;
  t0 = now();
  if (!ply::write(cloud, {"x", "y", "z"}, 1.0, {0, 0, 0}, argv[2], false)) {
    std::fprintf(stderr, "io_bench: cannot write %s\n", argv[2]);
    return 1;
  }
  const double write_s = now() - t0;
  const int w = std::atoi(argv[3]), h = std::atoi(argv[4]);
  std::vector<uint8_t> img((size_t)w * h * 3, 0);
  uint32_t s = 12345;
  for (size_t px = 0; px < (size_t)w * h; ++px) {  // like image B of save_image: ~1/3 occupied, bright
    s = s * 1664525u + 1013904223u;
    if ((s >> 24) < 85) img[3 * px + 1] = (uint8_t)(232 + ((s >> 16) & 15));
  }
  int64_t len = 0;
  t0 = now();
  if (bseg_png_encode(img.data(), w, h, 3, 0, nullptr, 0, &len) != 0)
    return 1;
  const double png_s = now() - t0;
  std::printf("{\"points\": %zu, \"ply_read_s\": %.4f, \"ply_write_s\": %.4f, \"png_w\": %d, \"png_h\": %d, \"png_encode_s\": %.4f, "
              "\"png_bytes\": %lld}\n",
              cloud.getPointCount(), read_s, write_s, w, h, png_s, (long long)len);
  return 0;
}

// ply_check.cpp -- CPU test harness for the host-side ply::read / ply::write (no GPU, no libbseg):
//   ply_check <in.ply> <scale> <out.ply> <dump.bin>
// reads with {"x","y","z"}, dumps positions (int32 x3) and colours (uint16 x3) raw, writes the cloud
// back as binary PLY with scale 1 / offset 0 exactly like the reference's main (TMC3.cpp:221).
#include <cstdio>
#include <cstdlib>

#include "ply.h"

int main(int argc, char** argv)
{
  if (argc < 5) return 2;
  pcc::PCCPointSet3 cloud;
  if (!pcc::ply::read(argv[1], {"x", "y", "z"}, std::atof(argv[2]), cloud)) return 1;
  FILE* f = std::fopen(argv[4], "wb");
  const long long n = (long long)cloud.getPointCount();
  const int hc = cloud.hasColors() ? 1 : 0;
  std::fwrite(&n, 8, 1, f);
  std::fwrite(&hc, 4, 1, f);
  for (long long i = 0; i < n; ++i) std::fwrite(&cloud[i][0], 4, 3, f);
  if (hc)
    for (long long i = 0; i < n; ++i) std::fwrite(&cloud.getColor(i)[0], 2, 3, f);
  std::fclose(f);
  bool ascii = argc > 5 && argv[5][0] == 'a';
  return pcc::ply::write(cloud, {"x", "y", "z"}, 1.0, {0, 0, 0}, argv[3], ascii) ? 0 : 3;
}

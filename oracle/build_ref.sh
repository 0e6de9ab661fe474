#!/bin/sh
# build_ref.sh -- compile the reference's OWN lines (region grower + raster) into
# oracle/_ref/libbseg_ref.so, straight from /root/reference/tmc3 where they lie.
# TEST INFRASTRUCTURE: pins oracle/bseg_oracle.c (orc_grow, orc_raster, ...) bit-for-bit.
# Nothing is copied into the repository: the generated TU lives in the git-ignored oracle/_ref/.
#
# What is extracted (line ranges of the GBK-encoded sources, compiled as raw bytes):
#   my_function.h:25-30    struct plane
#   my_function.h:89-123   class seg_plane
#   my_function.cpp:180-275  get_planes / Broad / set_plane_color
#   TMC3.cpp:44-200        struct Box + class buildingSeg (ctor, raster, save_image, groundTH)
# What cannot be compiled here: everything that needs Open3D / OpenCV (my_function.h:48-85,
# my_function.cpp:8-145, TMC3.cpp main) -- headers absent, no network.
set -e
REF=${1:-/root/reference/tmc3}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT="$HERE/_ref"
[ -d "$REF" ] || { echo "build_ref: $REF not present (GPU box) -- keeping prebuilt files"; exit 0; }
mkdir -p "$OUT"
TU="$OUT/ref_tu.cpp"
{
  echo '#include <cstdint>'
  echo '#include <cstdlib>'
  echo '#include <cmath>'
  echo '#include <string>'
  echo '#include <vector>'
  echo '#include <limits>'
  echo '#include <iostream>'
  echo '#include "ply.h"'
  echo 'using namespace pcc;'
  echo 'using namespace std;'
  sed -n '25,30p;89,123p' "$REF/my_function.h"
  sed -n '180,275p' "$REF/my_function.cpp"
  echo
  echo '#define STB_IMAGE_WRITE_IMPLEMENTATION'
  echo '#include "stb_image_write.h"'
  echo '#define private public'
  sed -n '44,200p' "$REF/TMC3.cpp"
  echo '#undef private'
  cat "$HERE/ref_wrap.inc"
} > "$TU"
/usr/bin/g++ -std=c++17 -O2 -ffp-contract=off -fwrapv -w -fPIC -shared -fvisibility=hidden \
    -I"$REF" "$TU" -o "$OUT/libbseg_ref.so" -lpthread
echo "build_ref: built $OUT/libbseg_ref.so"

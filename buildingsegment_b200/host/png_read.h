// png_read.h -- the part of cv::imread(path, IMREAD_COLOR) that extracted_contour needs (my_function.cpp:9): 8-bit
// PNG files, non-interlaced, greyscale / RGB / RGBA / palette, decoded to 3 interleaved bytes per pixel in R,G,B order
// (OpenCV keeps B,G,R: the green channel the reference thresholds is index 1 either way).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace bseg_png {
// false when the file is missing or not a PNG this reader covers (the reference prints a message and goes on with an
// empty image, my_function.cpp:10-12; the shim throws instead)
bool read_rgb(const std::string& path, std::vector<uint8_t>& rgb, int& width, int& height, std::string* why = nullptr);
}  // namespace bseg_png

// contour.cu -- SURVEY 8(f)-3: extracted_contour (my_function.cpp:8-145) without OpenCV.
//
// The reference thresholds the green channel of the count image ("像素数量.png", TMC3.cpp:101-108: every occupied
// pixel is >= 232 there) at 10 (:19-20), closes it with a 5x5 elliptic element, two iterations (:24-26), takes the
// external contours with CHAIN_APPROX_SIMPLE (:33), keeps those with area > 500 and perimeter > 100 (:37-43), draws
// them 2 pixels thick in (255,255,0) (:57-59), writes every contour extruded to z in {0,1} as csa.obj (:64-128) and
// saves the overlay and its vertical flip (:141-144).
//
// What runs where:
//   device  threshold + the four morphology passes (dilate, dilate, erode, erode): 84 taps per pixel over a raster that
//           is 4 * 10^8 pixels for a 2 km tile -- the part that is worth a GPU; tiles of the image in shared memory;
//   host    border following (Suzuki-Abe as OpenCV's findContours runs it: raster scan, marks on the border pixels, the
//           "last border seen in this row" rule that makes RETR_EXTERNAL skip components nested in holes): inherently
//           sequential, touches border pixels only; contourArea, arcLength, the overlay and the OBJ text.
// cv2 is the test oracle (tests/test_contour.py): masks, contours (points, order, start points), areas and perimeters
// are identical; the 2-pixel overlay strokes follow the definition "pixels within one pixel of the segment", which is
// not OpenCV's fixed-point polygon fill bit for bit (documented in DESIGN.md).
#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace {

constexpr int CT = 32;  // tile edge of the morphology kernels (+ 2 pixels of apron on every side)

// 5x5 MORPH_ELLIPSE of OpenCV's getStructuringElement: rows 0 and 4 hold only the centre column
__device__ __forceinline__ bool in_ellipse5(int dy, int dx) { return (dy != 0 && dy != 4) || dx == 2; }

__global__ void contour_threshold_kernel(const uint8_t* __restrict__ rgb, int64_t npx, int comp, int channel, int thresh,
                                         uint8_t* __restrict__ out)
{
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < npx)
    out[k] = rgb[k * comp + channel] > thresh ? 255 : 0;  // THRESH_BINARY (:20)
}

// one pass of dilate (MAXOP) or erode; pixels outside the image do not take part (OpenCV's morphologyDefaultBorderValue)
template <bool MAXOP>
__global__ void __launch_bounds__(CT* CT) contour_morph_kernel(const uint8_t* __restrict__ in, int W, int H,
                                                              uint8_t* __restrict__ out)
{
  __shared__ uint8_t t[CT + 4][CT + 4];
  const int x0 = blockIdx.x * CT - 2, y0 = blockIdx.y * CT - 2;
  const uint8_t neutral = MAXOP ? 0 : 255;
  for (int k = threadIdx.y * CT + threadIdx.x; k < (CT + 4) * (CT + 4); k += CT * CT) {
    const int ly = k / (CT + 4), lx = k - ly * (CT + 4);
    const int x = x0 + lx, y = y0 + ly;
    t[ly][lx] = (x >= 0 && x < W && y >= 0 && y < H) ? in[(int64_t)y * W + x] : neutral;
  }
  __syncthreads();
  const int x = blockIdx.x * CT + threadIdx.x, y = blockIdx.y * CT + threadIdx.y;
  if (x >= W || y >= H)
    return;
  uint8_t v = neutral;
#pragma unroll
  for (int dy = 0; dy < 5; ++dy)
#pragma unroll
    for (int dx = 0; dx < 5; ++dx)
      if (in_ellipse5(dy, dx)) {
        const uint8_t u = t[threadIdx.y + dy][threadIdx.x + dx];
        v = MAXOP ? (u > v ? u : v) : (u < v ? u : v);
      }
  out[(int64_t)y * W + x] = v;
}

}  // namespace

// threshold(channel) > thresh, then morphological close with the 5x5 ellipse, `iterations` times (dilate x it, erode x it)
int stage_contour_mask(bseg_ctx* c, const uint8_t* h_pixels, int32_t W, int32_t H, int32_t comp, int32_t channel, int32_t thresh,
                       int32_t iterations, uint8_t* h_mask)
{
  const int64_t npx = (int64_t)W * H;
  if (npx == 0)
    return 0;
  RC_CHECK(dev_ensure(c, c->out_tmp, (size_t)npx * comp + 2 * (size_t)npx + 256));
  uint8_t* d_px = dptr<uint8_t>(c->out_tmp);
  uint8_t* d_a = d_px + (((size_t)npx * comp + 127) & ~(size_t)127);
  uint8_t* d_b = d_a + (((size_t)npx + 127) & ~(size_t)127);
  RC_CHECK(dev_ensure(c, c->out_tmp, (size_t)(d_b - d_px) + (size_t)npx + 128));
  d_px = dptr<uint8_t>(c->out_tmp);
  d_a = d_px + (((size_t)npx * comp + 127) & ~(size_t)127);
  d_b = d_a + (((size_t)npx + 127) & ~(size_t)127);
  CU_CHECK(c, cudaMemcpyAsync(d_px, h_pixels, (size_t)npx * comp, cudaMemcpyHostToDevice, c->stream));
  contour_threshold_kernel<<<(unsigned)ceil_div64(npx, 256), 256, 0, c->stream>>>(d_px, npx, comp, channel, thresh, d_a);
  KLAUNCH_CHECK(c);
  const dim3 grid((unsigned)((W + CT - 1) / CT), (unsigned)((H + CT - 1) / CT)), block(CT, CT);
  uint8_t* src = d_a;
  uint8_t* dst = d_b;
  for (int pass = 0; pass < 2 * iterations; ++pass) {
    if (pass < iterations) contour_morph_kernel<true><<<grid, block, 0, c->stream>>>(src, W, H, dst);
    else contour_morph_kernel<false><<<grid, block, 0, c->stream>>>(src, W, H, dst);
    KLAUNCH_CHECK(c);
    uint8_t* t = src;
    src = dst;
    dst = t;
  }
  CU_CHECK(c, cudaMemcpyAsync(h_mask, src, (size_t)npx, cudaMemcpyDeviceToHost, c->stream));
  CU_CHECK(c, cudaStreamSynchronize(c->stream));
  return 0;
}

// ---- host: external contours as cv::findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE / NONE) returns them ------------
// Suzuki-Abe border following the way OpenCV runs it: the image gets a one-pixel frame of zeros, is scanned in raster
// order, an outer border starts at a 0 -> 1 transition whose pixel is still unmarked, and is skipped in RETR_EXTERNAL
// mode when the last border pixel crossed in this row carries a positive mark (we are inside another contour, i.e. in
// one of its holes).  A traced border pixel is marked 2, or -126 where the border has the outside on its right.
// Directions: code 0 = +x, counting counter-clockwise on the screen (y down): 1 = (+1,-1), 2 = (0,-1), ...
namespace {
const int CDX[16] = {1, 1, 0, -1, -1, -1, 0, 1, 1, 1, 0, -1, -1, -1, 0, 1};
const int CDY[16] = {0, -1, -1, -1, 0, 1, 1, 1, 0, -1, -1, -1, 0, 1, 1, 1};

void trace_border(int8_t* img, int64_t step, int x0, int y0, bool simple, std::vector<int32_t>& pts)
{
  const int8_t nbd = 2, nbd_right = (int8_t)(2 | -128);
  int8_t* i0 = img + (int64_t)y0 * step + x0;
  int s = 4, s_end = 4;
  int8_t* i1;
  do {
    s = (s - 1) & 7;
    i1 = i0 + CDY[s] * step + CDX[s];
  } while (*i1 == 0 && s != s_end);
  if (s == s_end) {  // a single pixel
    *i0 = nbd_right;
    pts.push_back(x0 - 1);
    pts.push_back(y0 - 1);
    return;
  }
  int8_t* i3 = i0;
  int prev_s = s ^ 4;
  int px = x0, py = y0;
  for (;;) {
    s_end = s;
    int8_t* i4;
    for (;;) {
      ++s;
      i4 = i3 + CDY[s & 15] * step + CDX[s & 15];
      if (*i4 != 0)
        break;
    }
    s &= 7;
    if ((unsigned)(s - 1) < (unsigned)s_end) *i3 = nbd_right;  // the "right bound" check
    else if (*i3 == 1) *i3 = nbd;
    if (s != prev_s || !simple) {
      pts.push_back(px - 1);
      pts.push_back(py - 1);
    }
    prev_s = s;
    px += CDX[s];
    py += CDY[s];
    if (i4 == i0 && i3 == i1)
      break;
    i3 = i4;
    s = (s + 4) & 7;
  }
}
}  // namespace

// contours in the order cv::findContours returns them (the last one found first); pts = x, y pairs
void contour_find_host(const uint8_t* mask, int32_t W, int32_t H, bool simple, std::vector<int32_t>& pts,
                             std::vector<int64_t>& offsets)
{
  const int64_t step = (int64_t)W + 2;
  std::vector<int8_t> img((size_t)step * (H + 2), 0);
  for (int y = 0; y < H; ++y) {
    int8_t* row = img.data() + (int64_t)(y + 1) * step + 1;
    const uint8_t* m = mask + (int64_t)y * W;
    for (int x = 0; x < W; ++x) row[x] = m[x] ? 1 : 0;
  }
  std::vector<std::vector<int32_t>> found;
  for (int y = 1; y <= H; ++y) {
    int8_t* row = img.data() + (int64_t)y * step;
    int lnbd_x = 0;
    int prev = 0;
    for (int x = 1; x <= W + 1; ++x) {
      int p = row[x];
      if (p == prev)
        continue;
      bool skip = false;
      if (!(prev == 0 && p == 1)) skip = true;  // a hole border, or a pixel that is marked already: not an external start
      if (!skip && row[lnbd_x] > 0) skip = true;  // inside another contour
      if (!skip) {
        found.emplace_back();
        trace_border(img.data(), step, x, y, simple, found.back());
        p = row[x];
      }
      prev = p;
      if (prev & -2) lnbd_x = x;
    }
  }
  pts.clear();
  offsets.assign(1, 0);
  for (size_t k = found.size(); k-- > 0;) {
    pts.insert(pts.end(), found[k].begin(), found[k].end());
    offsets.push_back((int64_t)pts.size() / 2);
  }
}

// cv::contourArea (Green's formula, doubles) and cv::arcLength(closed) (float differences, float sqrt, double sum)
double contour_area(const int32_t* xy, int64_t n)
{
  if (n == 0)
    return 0.0;
  double a00 = 0.0;
  double xp = (double)xy[2 * (n - 1)], yp = (double)xy[2 * (n - 1) + 1];
  for (int64_t i = 0; i < n; ++i) {
    const double x = (double)xy[2 * i], y = (double)xy[2 * i + 1];
    a00 += xp * y - x * yp;
    xp = x;
    yp = y;
  }
  return std::fabs(a00 * 0.5);
}

double contour_perimeter(const int32_t* xy, int64_t n)
{
  if (n <= 1)
    return 0.0;
  double per = 0.0;
  float xp = (float)xy[2 * (n - 1)], yp = (float)xy[2 * (n - 1) + 1];
  for (int64_t i = 0; i < n; ++i) {
    const float x = (float)xy[2 * i], y = (float)xy[2 * i + 1];
    const float dx = x - xp, dy = y - yp;
    per += std::sqrt(dx * dx + dy * dy);
    xp = x;
    yp = y;
  }
  return per;
}

// closed polyline drawn like cv::drawContours(thickness 2): OpenCV fills, per segment, the polygon one pixel to either
// side of it -- every pixel that polygon touches -- and a disc of radius 1 at the joints.  Here: a pixel is painted
// when its SQUARE meets the strip (|distance to the segment's line| <= 1 + half the square's extent along the normal,
// inside the segment's span) or its centre is within 1 of an end point.  comp bytes per pixel.
void contour_draw(uint8_t* img, int32_t W, int32_t H, int32_t comp, const int32_t* xy, int64_t n, const uint8_t* color)
{
  for (int64_t i = 0; i < n; ++i) {
    const double ax = xy[2 * i], ay = xy[2 * i + 1];
    const int64_t j = (i + 1) % n;
    const double bx = xy[2 * j], by = xy[2 * j + 1];
    const int xlo = (int)std::fmin(ax, bx) - 2, xhi = (int)std::fmax(ax, bx) + 2;
    const int ylo = (int)std::fmin(ay, by) - 2, yhi = (int)std::fmax(ay, by) + 2;
    const double dx = bx - ax, dy = by - ay, l2 = dx * dx + dy * dy, l = std::sqrt(l2);
    const double reach = l > 0 ? 1.0 + 0.5 * (std::fabs(dx) + std::fabs(dy)) / l : 1.0;
    for (int y = ylo < 0 ? 0 : ylo; y <= yhi && y < H; ++y)
      for (int x = xlo < 0 ? 0 : xlo; x <= xhi && x < W; ++x) {
        bool on = false;
        const double ea = (x - ax) * (x - ax) + (y - ay) * (y - ay), eb = (x - bx) * (x - bx) + (y - by) * (y - by);
        if (ea <= 1.0 || eb <= 1.0) on = true;
        else if (l > 0) {
          const double t = ((x - ax) * dx + (y - ay) * dy) / l2;
          if (t >= 0.0 && t <= 1.0) on = std::fabs((x - ax) * dy - (y - ay) * dx) / l < reach;
        }
        if (on)
          for (int k = 0; k < comp && k < 3; ++k) img[((int64_t)y * W + x) * comp + k] = color[k];
      }
  }
}

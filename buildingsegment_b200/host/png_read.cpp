// png_read.cpp -- PNG decoder for extracted_contour's input image (see png_read.h): chunk walk, zlib inflate (stored,
// fixed and dynamic Huffman blocks, RFC 1950/1951) and the five scanline filters (PNG spec 9.2).
#include "png_read.h"

#include <cstdio>
#include <cstring>

namespace bseg_png {
namespace {

struct Bits {
  const uint8_t* p;
  size_t n, pos = 0;
  uint32_t buf = 0;
  int cnt = 0;
  bool ok = true;
  int get(int k)
  {
    while (cnt < k) {
      if (pos >= n) {
        ok = false;
        return 0;
      }
      buf |= (uint32_t)p[pos++] << cnt;
      cnt += 8;
    }
    const int v = (int)(buf & ((1u << k) - 1));
    buf >>= k;
    cnt -= k;
    return v;
  }
};

struct Huff {
  uint16_t count[16], symbol[320];
  void build(const uint8_t* len, int n)
  {
    memset(count, 0, sizeof(count));
    for (int i = 0; i < n; ++i) count[len[i]]++;
    uint16_t offs[16];
    offs[1] = 0;
    for (int l = 1; l < 15; ++l) offs[l + 1] = offs[l] + count[l];
    for (int i = 0; i < n; ++i)
      if (len[i]) symbol[offs[len[i]]++] = (uint16_t)i;
    count[0] = 0;
  }
  int decode(Bits& b) const
  {
    int code = 0, first = 0, index = 0;
    for (int l = 1; l <= 15; ++l) {
      code |= b.get(1);
      if (!b.ok)
        return -1;
      const int c = count[l];
      if (code - c < first)
        return symbol[index + (code - first)];
      index += c;
      first += c;
      first <<= 1;
      code <<= 1;
    }
    return -1;
  }
};

bool inflate(const uint8_t* src, size_t n, std::vector<uint8_t>& out)
{
  static const uint16_t LBASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
  static const uint8_t LEXT[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
  static const uint16_t DBASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
  static const uint8_t DEXT[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
  if (n < 2)
    return false;
  Bits b{src + 2, n - 2};  // zlib header: CMF, FLG
  for (;;) {
    const int last = b.get(1), type = b.get(2);
    if (!b.ok)
      return false;
    if (type == 0) {
      b.buf = 0;
      b.cnt = 0;
      if (b.pos + 4 > b.n)
        return false;
      const unsigned len = b.p[b.pos] | (b.p[b.pos + 1] << 8);
      b.pos += 4;
      if (b.pos + len > b.n)
        return false;
      out.insert(out.end(), b.p + b.pos, b.p + b.pos + len);
      b.pos += len;
    } else if (type == 1 || type == 2) {
      Huff hl, hd;
      uint8_t lens[320];
      if (type == 1) {
        int i = 0;
        for (; i < 144; ++i) lens[i] = 8;
        for (; i < 256; ++i) lens[i] = 9;
        for (; i < 280; ++i) lens[i] = 7;
        for (; i < 288; ++i) lens[i] = 8;
        hl.build(lens, 288);
        for (i = 0; i < 30; ++i) lens[i] = 5;
        hd.build(lens, 30);
      } else {
        static const uint8_t ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        const int nlen = b.get(5) + 257, ndist = b.get(5) + 1, ncode = b.get(4) + 4;
        if (!b.ok || nlen > 286 || ndist > 30)
          return false;
        uint8_t cl[19] = {0};
        for (int i = 0; i < ncode; ++i) cl[ORDER[i]] = (uint8_t)b.get(3);
        Huff hc;
        hc.build(cl, 19);
        int i = 0;
        while (i < nlen + ndist) {
          const int sym = hc.decode(b);
          if (sym < 0)
            return false;
          if (sym < 16) {
            lens[i++] = (uint8_t)sym;
          } else {
            int rep, val = 0;
            if (sym == 16) {
              if (i == 0)
                return false;
              val = lens[i - 1];
              rep = 3 + b.get(2);
            } else if (sym == 17) {
              rep = 3 + b.get(3);
            } else {
              rep = 11 + b.get(7);
            }
            if (i + rep > nlen + ndist)
              return false;
            while (rep--) lens[i++] = (uint8_t)val;
          }
        }
        hl.build(lens, nlen);
        hd.build(lens + nlen, ndist);
      }
      for (;;) {
        const int sym = hl.decode(b);
        if (sym < 0 || !b.ok)
          return false;
        if (sym < 256) {
          out.push_back((uint8_t)sym);
        } else if (sym == 256) {
          break;
        } else {
          const int ls = sym - 257;
          if (ls >= 29)
            return false;
          const int len = LBASE[ls] + b.get(LEXT[ls]);
          const int ds = hd.decode(b);
          if (ds < 0 || ds >= 30)
            return false;
          const size_t dist = (size_t)DBASE[ds] + (size_t)b.get(DEXT[ds]);
          if (!b.ok || dist > out.size())
            return false;
          const size_t from = out.size() - dist;
          for (int k = 0; k < len; ++k) out.push_back(out[from + k]);
        }
      }
    } else {
      return false;
    }
    if (last)
      break;
  }
  return true;
}

int paeth(int a, int b, int c)
{
  const int p = a + b - c, pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

}  // namespace

bool read_rgb(const std::string& path, std::vector<uint8_t>& rgb, int& width, int& height, std::string* why)
{
  auto fail = [&](const char* m) {
    if (why) *why = m;
    return false;
  };
  FILE* f = fopen(path.c_str(), "rb");
  if (!f)
    return fail("cannot open the file");
  std::vector<uint8_t> file;
  uint8_t tmp[1 << 16];
  size_t got;
  while ((got = fread(tmp, 1, sizeof(tmp), f)) > 0) file.insert(file.end(), tmp, tmp + got);
  fclose(f);
  static const uint8_t SIG[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (file.size() < 8 || memcmp(file.data(), SIG, 8) != 0)
    return fail("not a PNG file");
  size_t pos = 8;
  int depth = 0, ctype = 0, interlace = 0;
  width = height = 0;
  std::vector<uint8_t> idat, plte;
  while (pos + 12 <= file.size()) {
    const uint32_t len = ((uint32_t)file[pos] << 24) | (file[pos + 1] << 16) | (file[pos + 2] << 8) | file[pos + 3];
    const char* type = reinterpret_cast<const char*>(&file[pos + 4]);
    const uint8_t* data = &file[pos + 8];
    if (pos + 12 + len > file.size())
      return fail("truncated chunk");
    if (!memcmp(type, "IHDR", 4) && len >= 13) {
      width = (int)(((uint32_t)data[0] << 24) | (data[1] << 16) | (data[2] << 8) | data[3]);
      height = (int)(((uint32_t)data[4] << 24) | (data[5] << 16) | (data[6] << 8) | data[7]);
      depth = data[8];
      ctype = data[9];
      interlace = data[12];
    } else if (!memcmp(type, "PLTE", 4)) {
      plte.assign(data, data + len);
    } else if (!memcmp(type, "IDAT", 4)) {
      idat.insert(idat.end(), data, data + len);
    } else if (!memcmp(type, "IEND", 4)) {
      break;
    }
    pos += 12 + (size_t)len;
  }
  if (width <= 0 || height <= 0 || depth != 8 || interlace != 0)
    return fail("only 8-bit non-interlaced PNG files are read");
  const int comp = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
  if (!comp)
    return fail("unknown colour type");
  std::vector<uint8_t> raw;
  raw.reserve((size_t)height * ((size_t)width * comp + 1));
  if (!inflate(idat.data(), idat.size(), raw))
    return fail("corrupt zlib stream");
  const size_t stride = (size_t)width * comp;
  if (raw.size() < (size_t)height * (stride + 1))
    return fail("image data too short");
  std::vector<uint8_t> prev(stride, 0), cur(stride);
  rgb.assign((size_t)width * height * 3, 0);
  for (int y = 0; y < height; ++y) {
    const uint8_t* in = &raw[(size_t)y * (stride + 1)];
    const int ft = in[0];
    for (size_t k = 0; k < stride; ++k) {
      const int a = k >= (size_t)comp ? cur[k - comp] : 0, b = prev[k], c = k >= (size_t)comp ? prev[k - comp] : 0;
      int v = in[1 + k];
      switch (ft) {
        case 0: break;
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) >> 1; break;
        case 4: v += paeth(a, b, c); break;
        default: return fail("unknown scanline filter");
      }
      cur[k] = (uint8_t)v;
    }
    uint8_t* out = &rgb[(size_t)y * width * 3];
    for (int x = 0; x < width; ++x) {
      const uint8_t* px = &cur[(size_t)x * comp];
      if (ctype == 2 || ctype == 6) {
        out[3 * x] = px[0];
        out[3 * x + 1] = px[1];
        out[3 * x + 2] = px[2];
      } else if (ctype == 3) {
        const size_t e = (size_t)px[0] * 3;
        for (int k = 0; k < 3; ++k) out[3 * x + k] = e + k < plte.size() ? plte[e + k] : 0;
      } else {
        out[3 * x] = out[3 * x + 1] = out[3 * x + 2] = px[0];
      }
    }
    prev.swap(cur);
  }
  return true;
}

}  // namespace bseg_png

// png.cu -- host-only: the PNG container the raster's byte images are written in.
//
// The reference hands its three W*H*3 byte images to stbi_write_png of the vendored stb_image_write v1.16
// (TMC3.cpp:98,108,119 -> stb_image_write.h:1215).  A file written here is BYTE-identical to what that call produces
// with its default settings (compression level 8, filter chosen per row), so downstream tools that hash or diff the
// images see no change.  That pins three published behaviours, restated here from the format's point of view:
//
//   rows      every row is encoded with the five PNG filter types (None, Sub, Up, Average, Paeth) and the one with the
//             smallest sum of |signed byte| wins, the first on ties.  On the first row "Up" degenerates to "None",
//             "Average" to half of the left neighbour and "Paeth" to the left neighbour -- while the filter byte
//             written is still the nominal type.
//   deflate   one fixed-Huffman block.  Matches come from a 16384-bucket hash of 3-byte prefixes; a bucket keeps at
//             most 16 positions and drops its older half when full; only positions where a token STARTS are entered.
//             A match (>= 3 bytes, <= 258, distance < 32768, the LAST candidate of the bucket among equally long ones)
//             is abandoned for a literal when the next byte has a strictly longer match in its own bucket (one-step
//             lazy matching).  If the result is larger than stored blocks would be, stored blocks (<= 32767 bytes)
//             are written instead.  zlib header 0x78 0x5e, Adler-32 trailer.
//   chunks    IHDR (8 bit; colour type by component count: 1 grey, 2 grey+alpha, 3 RGB, 4 RGBA), one IDAT, IEND.
//
// Encoding is sequential by nature (the match finder's state runs through the whole image), so the three images of
// save_image are encoded on three worker threads, off the caller's critical path (bseg_png_write_async / _wait).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

// ---- row filters -----------------------------------------------------------------------------------------------
inline int paeth(int a, int b, int c)
{
  const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
  if (pa <= pb && pa <= pc) return a & 0xff;
  if (pb <= pc) return b & 0xff;
  return c & 0xff;
}

// filter `type` (0..4) applied to row y; prev == nullptr on the first row, where the stb encoder's degenerate forms apply
void filter_row(const uint8_t* cur, const uint8_t* prev, int bytes, int bpp, int type, uint8_t* out)
{
  enum { NONE = 0, SUB = 1, UP = 2, AVG = 3, PAETH = 4, AVG_FIRST = 5, PAETH_FIRST = 6 };
  int eff = type;
  if (!prev) eff = type == UP ? NONE : (type == AVG ? AVG_FIRST : (type == PAETH ? PAETH_FIRST : type));
  if (eff == NONE) {
    memcpy(out, cur, (size_t)bytes);
    return;
  }
  for (int i = 0; i < bpp && i < bytes; ++i) {  // the first pixel has no left neighbour
    switch (eff) {
      case SUB: case AVG_FIRST: case PAETH_FIRST: out[i] = cur[i]; break;
      case UP: out[i] = (uint8_t)(cur[i] - prev[i]); break;
      case AVG: out[i] = (uint8_t)(cur[i] - (prev[i] >> 1)); break;
      case PAETH: out[i] = (uint8_t)(cur[i] - paeth(0, prev[i], 0)); break;
    }
  }
  for (int i = bpp; i < bytes; ++i) {
    int pred = 0;
    switch (eff) {
      case SUB: pred = cur[i - bpp]; break;
      case UP: pred = prev[i]; break;
      case AVG: pred = (cur[i - bpp] + prev[i]) >> 1; break;
      case PAETH: pred = paeth(cur[i - bpp], prev[i], prev[i - bpp]); break;
      case AVG_FIRST: pred = cur[i - bpp] >> 1; break;
      case PAETH_FIRST: pred = paeth(cur[i - bpp], 0, 0); break;
    }
    out[i] = (uint8_t)(cur[i] - pred);
  }
}

// ---- deflate ---------------------------------------------------------------------------------------------------
struct BitSink {
  std::vector<uint8_t>& out;
  uint32_t buf = 0;
  int n = 0;
  explicit BitSink(std::vector<uint8_t>& o) : out(o) {}
  void put(uint32_t code, int bits)
  {
    buf |= code << n;
    n += bits;
    while (n >= 8) {
      out.push_back((uint8_t)buf);
      buf >>= 8;
      n -= 8;
    }
  }
  static uint32_t reversed(uint32_t code, int bits)
  {
    uint32_t r = 0;
    while (bits--) {
      r = (r << 1) | (code & 1u);
      code >>= 1;
    }
    return r;
  }
  void huff(uint32_t code, int bits) { put(reversed(code, bits), bits); }
  // fixed Huffman literal/length alphabet (RFC 1951 3.2.6)
  void symbol(int s)
  {
    if (s <= 143) huff(0x30 + s, 8);
    else if (s <= 255) huff(0x190 + s - 144, 9);
    else if (s <= 279) huff(s - 256, 7);
    else huff(0xc0 + s - 280, 8);
  }
};

inline uint32_t prefix_hash(const uint8_t* d)
{
  uint32_t h = d[0] + ((uint32_t)d[1] << 8) + ((uint32_t)d[2] << 16);
  h ^= h << 3;
  h += h >> 5;
  h ^= h << 4;
  h += h >> 17;
  h ^= h << 25;
  h += h >> 6;
  return h;
}

inline int match_len(const uint8_t* a, const uint8_t* b, int limit)
{
  int i = 0;
  while (i < limit && i < 258 && a[i] == b[i]) ++i;
  return i;
}

void zlib_stream(const uint8_t* data, int len, std::vector<uint8_t>& out)
{
  static const uint16_t len_base[] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258, 259};
  static const uint8_t len_extra[] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
  static const uint16_t dist_base[] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577, 32768};
  static const uint8_t dist_extra[] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
  constexpr int BUCKETS = 16384, KEEP = 8;  // compression level 8: a bucket holds at most 2 * 8 positions
  out.clear();
  out.push_back(0x78);
  out.push_back(0x5e);
  {
    BitSink bits(out);
    bits.put(1, 1);  // last block
    bits.put(1, 2);  // fixed Huffman
    // bucket b: positions pos[b * 2 * KEEP ..], cnt[b] of them, oldest first
    std::vector<int32_t> pos((size_t)BUCKETS * 2 * KEEP);
    std::vector<uint8_t> cnt(BUCKETS, 0);
    int i = 0;
    while (i < len - 3) {
      const uint32_t h = prefix_hash(data + i) & (BUCKETS - 1);
      int32_t* bucket = &pos[(size_t)h * 2 * KEEP];
      int best = 3, best_at = -1;
      for (int j = 0; j < cnt[h]; ++j)
        if (bucket[j] > i - 32768) {
          const int d = match_len(data + bucket[j], data + i, len - i);
          if (d >= best) {
            best = d;
            best_at = bucket[j];
          }
        }
      if (cnt[h] == 2 * KEEP) {  // full: the older half goes
        memmove(bucket, bucket + KEEP, sizeof(int32_t) * KEEP);
        cnt[h] = KEEP;
      }
      bucket[cnt[h]++] = i;
      if (best_at >= 0) {  // a strictly longer match one byte later wins: this byte becomes a literal
        const uint32_t h2 = prefix_hash(data + i + 1) & (BUCKETS - 1);
        const int32_t* b2 = &pos[(size_t)h2 * 2 * KEEP];
        for (int j = 0; j < cnt[h2]; ++j)
          if (b2[j] > i - 32767 && match_len(data + b2[j], data + i + 1, len - i - 1) > best) {
            best_at = -1;
            break;
          }
      }
      if (best_at >= 0) {
        const int dist = i - best_at;
        int j = 0;
        while (best > len_base[j + 1] - 1) ++j;
        bits.symbol(j + 257);
        if (len_extra[j]) bits.put((uint32_t)(best - len_base[j]), len_extra[j]);
        j = 0;
        while (dist > dist_base[j + 1] - 1) ++j;
        bits.put(BitSink::reversed((uint32_t)j, 5), 5);
        if (dist_extra[j]) bits.put((uint32_t)(dist - dist_base[j]), dist_extra[j]);
        i += best;
      } else {
        bits.symbol(data[i]);
        ++i;
      }
    }
    for (; i < len; ++i) bits.symbol(data[i]);
    bits.symbol(256);
    while (bits.n) bits.put(0, 1);
  }
  // stored blocks when compression lost
  const size_t stored = (size_t)len + 2 + (size_t)((len + 32766) / 32767) * 5;
  if (out.size() > stored) {
    out.resize(2);
    for (int j = 0; j < len;) {
      int block = len - j;
      if (block > 32767) block = 32767;
      out.push_back(len - j == block ? 1 : 0);
      out.push_back((uint8_t)block);
      out.push_back((uint8_t)(block >> 8));
      out.push_back((uint8_t)~block);
      out.push_back((uint8_t)(~block >> 8));
      out.insert(out.end(), data + j, data + j + block);
      j += block;
    }
  }
  uint32_t s1 = 1, s2 = 0;  // Adler-32
  int block = len % 5552;
  for (int j = 0; j < len;) {
    for (int k = 0; k < block; ++k) {
      s1 += data[j + k];
      s2 += s1;
    }
    s1 %= 65521;
    s2 %= 65521;
    j += block;
    block = 5552;
  }
  out.push_back((uint8_t)(s2 >> 8));
  out.push_back((uint8_t)s2);
  out.push_back((uint8_t)(s1 >> 8));
  out.push_back((uint8_t)s1);
}

// ---- container ---------------------------------------------------------------------------------------------------
uint32_t crc32_of(const uint8_t* p, size_t n)
{
  static uint32_t table[256];
  static std::once_flag once;
  std::call_once(once, [] {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xedb88320u ^ (c >> 1) : c >> 1;
      table[i] = c;
    }
  });
  uint32_t c = ~0u;
  for (size_t i = 0; i < n; ++i) c = (c >> 8) ^ table[p[i] ^ (c & 0xff)];
  return ~c;
}

void be32(std::vector<uint8_t>& v, uint32_t x)
{
  for (int s = 24; s >= 0; s -= 8) v.push_back((uint8_t)(x >> s));
}

void chunk(std::vector<uint8_t>& v, const char tag[4], const uint8_t* body, size_t n)
{
  be32(v, (uint32_t)n);
  const size_t at = v.size();
  v.insert(v.end(), tag, tag + 4);
  if (n) v.insert(v.end(), body, body + n);
  be32(v, crc32_of(v.data() + at, n + 4));
}

bool encode(const uint8_t* pixels, int stride, int w, int h, int comp, std::vector<uint8_t>& png)
{
  if (!pixels || w <= 0 || h <= 0 || comp < 1 || comp > 4)
    return false;
  if (stride == 0) stride = w * comp;
  const int bytes = w * comp;
  if ((int64_t)(bytes + 1) * h >= ((int64_t)1 << 31))
    return false;  // the stream is indexed with int (as in the encoder this one mirrors)
  std::vector<uint8_t> filtered((size_t)(bytes + 1) * h), trial((size_t)bytes), keep((size_t)bytes);
  for (int y = 0; y < h; ++y) {
    const uint8_t* cur = pixels + (size_t)stride * y;
    const uint8_t* prev = y ? cur - stride : nullptr;
    int best = 0, best_cost = 0x7fffffff;
    for (int f = 0; f < 5; ++f) {
      filter_row(cur, prev, bytes, comp, f, trial.data());
      int cost = 0;
      for (int i = 0; i < bytes; ++i) cost += std::abs((int)(int8_t)trial[i]);
      if (cost < best_cost) {
        best_cost = cost;
        best = f;
        keep.swap(trial);
      }
    }
    uint8_t* dst = filtered.data() + (size_t)(bytes + 1) * y;
    dst[0] = (uint8_t)best;
    memcpy(dst + 1, keep.data(), (size_t)bytes);
  }
  std::vector<uint8_t> z;
  zlib_stream(filtered.data(), (int)filtered.size(), z);
  static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
  static const int ctype[5] = {-1, 0, 4, 2, 6};
  png.assign(sig, sig + 8);
  std::vector<uint8_t> ihdr;
  be32(ihdr, (uint32_t)w);
  be32(ihdr, (uint32_t)h);
  const uint8_t tail[5] = {8, (uint8_t)ctype[comp], 0, 0, 0};
  ihdr.insert(ihdr.end(), tail, tail + 5);
  chunk(png, "IHDR", ihdr.data(), ihdr.size());
  chunk(png, "IDAT", z.data(), z.size());
  chunk(png, "IEND", nullptr, 0);
  return true;
}

bool write_file(const std::string& path, const std::vector<uint8_t>& png)
{
  FILE* f = fopen(path.c_str(), "wb");
  if (!f)
    return false;
  const bool ok = fwrite(png.data(), 1, png.size(), f) == png.size();
  return (fclose(f) == 0) && ok;
}

// ---- asynchronous writes -------------------------------------------------------------------------------------------
struct Pending {
  std::thread th;
  bool ok = false;
};
std::mutex g_mu;
std::vector<Pending*> g_pending;

}  // namespace

extern "C" {

BSEG_API int bseg_png_encode(const uint8_t* pixels, int32_t w, int32_t h, int32_t comp, int32_t stride_bytes, uint8_t* out,
                             int64_t cap, int64_t* len)
{
  std::vector<uint8_t> png;
  if (!len || !encode(pixels, stride_bytes, w, h, comp, png))
    return bseg_fail(nullptr, BSEG_E_ARG, "bseg_png_encode: bad image (%d x %d x %d)", w, h, comp);
  *len = (int64_t)png.size();
  if (out) {
    if (cap < *len)
      return bseg_fail(nullptr, BSEG_E_CAPACITY, "bseg_png_encode: %lld bytes needed, %lld given", (long long)*len, (long long)cap);
    memcpy(out, png.data(), png.size());
  }
  return 0;
}

BSEG_API int bseg_png_write(const char* path, const uint8_t* pixels, int32_t w, int32_t h, int32_t comp, int32_t stride_bytes)
{
  std::vector<uint8_t> png;
  if (!path || !encode(pixels, stride_bytes, w, h, comp, png))
    return bseg_fail(nullptr, BSEG_E_ARG, "bseg_png_write: bad image (%d x %d x %d)", w, h, comp);
  if (!write_file(path, png))
    return bseg_fail(nullptr, BSEG_E_ARG, "bseg_png_write: cannot write %s", path);
  return 0;
}

BSEG_API int bseg_png_write_async(const char* path, const uint8_t* pixels, int32_t w, int32_t h, int32_t comp,
                                  int32_t stride_bytes)
{
  if (!path || !pixels || w <= 0 || h <= 0 || comp < 1 || comp > 4)
    return bseg_fail(nullptr, BSEG_E_ARG, "bseg_png_write_async: bad image (%d x %d x %d)", w, h, comp);
  const int stride = stride_bytes ? stride_bytes : w * comp;
  // the caller's buffer may change or vanish: the worker owns a packed copy
  auto* copy = new std::vector<uint8_t>((size_t)w * comp * h);
  for (int y = 0; y < h; ++y) memcpy(copy->data() + (size_t)w * comp * y, pixels + (size_t)stride * y, (size_t)w * comp);
  Pending* p = new Pending();
  const std::string file(path);
  p->th = std::thread([p, copy, file, w, h, comp] {
    std::vector<uint8_t> png;
    p->ok = encode(copy->data(), w * comp, w, h, comp, png) && write_file(file, png);
    delete copy;
  });
  std::lock_guard<std::mutex> lock(g_mu);
  g_pending.push_back(p);
  return 0;
}

BSEG_API int bseg_png_wait(void)
{
  std::vector<Pending*> mine;
  {
    std::lock_guard<std::mutex> lock(g_mu);
    mine.swap(g_pending);
  }
  int bad = 0;
  for (Pending* p : mine) {
    p->th.join();
    if (!p->ok) ++bad;
    delete p;
  }
  return bad ? bseg_fail(nullptr, BSEG_E_ARG, "bseg_png_wait: %d image(s) could not be written", bad) : 0;
}

}  // extern "C"

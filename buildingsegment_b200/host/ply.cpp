// ply.cpp -- bulk PLY I/O behind the reference's ply::read / ply::write signatures.
//
// Same file format and value semantics as the reference (tmc3/ply.cpp:88-504; every rule cited in
// ply.h), different mechanics: the reference issues 3-6 iostream calls per point; here the vertex
// block is decoded from / encoded into large buffers with a precomputed per-record layout, which is
// what SURVEY.md 8(f)-1 asks for once the GPU path makes I/O the dominant cost.
#include "ply.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <vector>

namespace pcc {
namespace ply {
namespace {

enum class Ty { F64, F32, U64, U32, U16, U8, I64, I32, I16, I8, Unknown };

struct Prop {
  std::string name;
  Ty type = Ty::Unknown;
  size_t bytes = 0;
  size_t offset = 0;
};

Ty parseType(const std::string& t, size_t& bytes)
{
  struct { const char* n; Ty t; size_t b; } tab[] = {
    {"float64", Ty::F64, 8}, {"double", Ty::F64, 8}, {"float", Ty::F32, 4}, {"float32", Ty::F32, 4},
    {"uint64", Ty::U64, 8}, {"uint32", Ty::U32, 4}, {"uint", Ty::U32, 4}, {"uint16", Ty::U16, 2},
    {"ushort", Ty::U16, 2}, {"uchar", Ty::U8, 1}, {"uint8", Ty::U8, 1}, {"int64", Ty::I64, 8},
    {"int32", Ty::I32, 4}, {"int", Ty::I32, 4}, {"int16", Ty::I16, 2}, {"short", Ty::I16, 2},
    {"char", Ty::I8, 1}, {"int8", Ty::I8, 1}};
  for (const auto& e : tab)
    if (t == e.n) {
      bytes = e.b;
      return e.t;
    }
  bytes = 0;
  return Ty::Unknown;
}

std::vector<std::string> tokens(const std::string& line)
{
  std::vector<std::string> out;
  std::istringstream ss(line);
  std::string t;
  while (ss >> t) out.push_back(t);
  return out;
}

bool hostIsLittleEndian()
{
  const uint16_t v = 1;
  return *reinterpret_cast<const uint8_t*>(&v) == 1;
}

template <typename T>
T loadAs(const uint8_t* p, bool swap)
{
  uint8_t b[sizeof(T)];
  if (swap)
    for (size_t i = 0; i < sizeof(T); ++i) b[i] = p[sizeof(T) - 1 - i];
  else
    std::memcpy(b, p, sizeof(T));
  T v;
  std::memcpy(&v, b, sizeof(T));
  return v;
}

double loadReal(const Prop& pr, const uint8_t* rec, bool swap)
{
  // by width, as the reference does (ply.cpp:437-465): a 4-byte position property is read as a float whatever its
  // declared type (int / uint32 bit patterns are reinterpreted, not converted), an 8-byte one as a double
  return pr.bytes == 4 ? double(loadAs<float>(rec + pr.offset, swap)) : loadAs<double>(rec + pr.offset, swap);
}

}  // namespace

bool write(const PCCPointSet3& cloud, const PropertyNameMap& names, double positionScale, Vec3<double> positionOffset,
           const std::string& fileName, bool asAscii)
{
  FILE* f = std::fopen(fileName.c_str(), "wb");
  if (!f)
    return false;
  const size_t n = cloud.getPointCount();
  std::string h = "ply\n";
  h += asAscii ? "format ascii 1.0\n"
               : (hostIsLittleEndian() ? "format binary_little_endian 1.0\n" : "format binary_big_endian 1.0\n");
  h += "element vertex " + std::to_string(n) + "\n";
  const char* ptype = asAscii ? "property float " : "property float64 ";
  for (int k = 0; k < 3; ++k) h += std::string(ptype) + names.position[k] + "\n";
  if (cloud.hasColors()) h += "property uchar green\nproperty uchar blue\nproperty uchar red\n";
  if (cloud.hasReflectances()) h += "property uint16 refc\n";
  if (cloud.hasFrameIndex()) h += "property uint8 frameindex\n";
  h += "element face 0\nproperty list uint8 int32 vertex_index\nend_header\n";
  bool ok = std::fwrite(h.data(), 1, h.size(), f) == h.size();

  if (asAscii) {
    std::string line;
    char buf[256];
    for (size_t i = 0; i < n && ok; ++i) {
      const Vec3<double> p = cloud[i] * positionScale + positionOffset;
      int len = std::snprintf(buf, sizeof(buf), "%.5f %.5f %.5f", p[0], p[1], p[2]);
      line.assign(buf, len);
      if (cloud.hasColors()) {
        const auto& c = cloud.getColor(i);
        len = std::snprintf(buf, sizeof(buf), " %d %d %d", int(c[0]), int(c[1]), int(c[2]));
        line.append(buf, len);
      }
      if (cloud.hasReflectances()) line += " " + std::to_string(int(cloud.getReflectance(i)));
      if (cloud.hasFrameIndex()) line += " " + std::to_string(int(cloud.getFrameIndex(i)));
      line += "\n";
      ok = std::fwrite(line.data(), 1, line.size(), f) == line.size();
    }
  } else {
    // record: 3 doubles [+ 3 uint8 colours] [+ uint16 reflectance] [+ uint16 frame index]
    const size_t rec = 24 + (cloud.hasColors() ? 3 : 0) + (cloud.hasReflectances() ? 2 : 0) + (cloud.hasFrameIndex() ? 2 : 0);
    const size_t chunk = 1 << 16;
    std::vector<uint8_t> buf(rec * chunk);
    for (size_t i0 = 0; i0 < n && ok; i0 += chunk) {
      const size_t m = n - i0 < chunk ? n - i0 : chunk;
      uint8_t* o = buf.data();
      for (size_t i = i0; i < i0 + m; ++i) {
        const Vec3<double> p = cloud[i] * positionScale + positionOffset;
        std::memcpy(o, p.data, 24);
        o += 24;
        if (cloud.hasColors()) {
          const auto& c = cloud.getColor(i);
          o[0] = uint8_t(c[0]); o[1] = uint8_t(c[1]); o[2] = uint8_t(c[2]);
          o += 3;
        }
        if (cloud.hasReflectances()) {
          const uint16_t r = cloud.getReflectance(i);
          std::memcpy(o, &r, 2);
          o += 2;
        }
        if (cloud.hasFrameIndex()) {
          const uint16_t fi = cloud.getFrameIndex(i);
          std::memcpy(o, &fi, 2);
          o += 2;
        }
      }
      ok = std::fwrite(buf.data(), rec, m, f) == m;
    }
  }
  ok = (std::fclose(f) == 0) && ok;
  return ok;
}

bool read(const std::string& fileName, const PropertyNameMap& names, double positionScale, PCCPointSet3& cloud)
{
  FILE* f = std::fopen(fileName.c_str(), "rb");
  if (!f)
    return false;
  auto fail = [&](const char* msg) {
    std::cout << "Error: " << msg << std::endl;
    std::fclose(f);
    return false;
  };
  char line[4096];
  if (!std::fgets(line, sizeof(line), f) || tokens(line).empty() || tokens(line)[0] != "ply")
    return fail("corrupted file!");
  bool isAscii = false, bigEndian = false, inVertex = true;
  double version = 1.0;
  size_t pointCount = 0;
  std::vector<Prop> props;
  for (;;) {
    if (!std::fgets(line, sizeof(line), f))
      return fail("corrupted header!");
    const auto tk = tokens(line);
    if (tk.empty() || tk[0] == "comment")
      continue;
    if (tk[0] == "format") {
      if (tk.size() != 3) return fail("corrupted format info!");
      isAscii = tk[1] == "ascii";
      bigEndian = tk[1] == "binary_big_endian";
      version = std::atof(tk[2].c_str());
    } else if (tk[0] == "element") {
      if (tk.size() != 3) return fail("corrupted element info!");
      if (tk[1] == "vertex") pointCount = size_t(std::atoll(tk[2].c_str()));
      else inVertex = false;
    } else if (tk[0] == "property" && inVertex) {
      if (tk.size() != 3) return fail("corrupted property info!");
      Prop p;
      p.name = tk[2];
      p.type = parseType(tk[1], p.bytes);
      props.push_back(p);
    } else if (tk[0] == "end_header") {
      break;
    }
  }
  if (version != 1.0)
    return fail("non-supported version!");
  const size_t none = size_t(-1);
  size_t iX = none, iY = none, iZ = none, iR = none, iG = none, iB = none, iRefl = none, iFrame = none;
  size_t rec = 0;
  for (size_t a = 0; a < props.size(); ++a) {
    Prop& p = props[a];
    p.offset = rec;
    rec += p.bytes;
    const bool real = p.bytes == 8 || p.bytes == 4;
    if (p.name == names.position[0] && real) iX = a;
    else if (p.name == names.position[1] && real) iY = a;
    else if (p.name == names.position[2] && real) iZ = a;
    else if (p.name == "red" && p.bytes == 1) iR = a;
    else if (p.name == "green" && p.bytes == 1) iG = a;
    else if (p.name == "blue" && p.bytes == 1) iB = a;
    else if ((p.name == "reflectance" || p.name == "refc") && p.bytes <= 2) iRefl = a;
    else if (p.name == "frameindex" && p.bytes <= 2) iFrame = a;
  }
  if (iX == none || iY == none || iZ == none)
    return fail("missing coordinates!");
  const bool withColors = iR != none && iG != none && iB != none;
  cloud.addRemoveAttributes(withColors, iRefl != none);
  if (iFrame != none) cloud.addFrameIndex(); else cloud.removeFrameIndex();
  cloud.resize(pointCount);

  if (isAscii) {
    size_t i = 0;
    while (i < pointCount && std::fgets(line, sizeof(line), f)) {
      const auto tk = tokens(line);
      if (tk.empty())
        continue;
      if (tk.size() < props.size()) {
        std::fclose(f);
        return false;
      }
      auto& pos = cloud[i];
      pos[0] = int32_t(std::atof(tk[iX].c_str()) * positionScale);
      pos[1] = int32_t(std::atof(tk[iY].c_str()) * positionScale);
      pos[2] = int32_t(std::atof(tk[iZ].c_str()) * positionScale);
      if (withColors) {
        auto& c = cloud.getColor(i);
        c[0] = attr_t(std::atoi(tk[iG].c_str()));
        c[1] = attr_t(std::atoi(tk[iB].c_str()));
        c[2] = attr_t(std::atoi(tk[iR].c_str()));
      }
      if (iRefl != none) cloud.getReflectance(i) = uint16_t(std::atoi(tk[iRefl].c_str()));
      if (iFrame != none) cloud.getFrameIndex(i) = uint8_t(std::atoi(tk[iFrame].c_str()));
      ++i;
    }
  } else {
    const bool swap = bigEndian == hostIsLittleEndian();
    const size_t chunk = 1 << 16;
    std::vector<uint8_t> buf(rec * chunk);
    // fast path: the layout every generator here and the reference writer produce
    const bool xyzF32 = props[iX].type == Ty::F32 && props[iY].type == Ty::F32 && props[iZ].type == Ty::F32;
    size_t i = 0;
    while (i < pointCount) {
      const size_t want = pointCount - i < chunk ? pointCount - i : chunk;
      const size_t got = rec ? std::fread(buf.data(), rec, want, f) : 0;
      if (got == 0)
        break;
      for (size_t r = 0; r < got; ++r) {
        const uint8_t* p = buf.data() + r * rec;
        auto& pos = cloud[i + r];
        if (xyzF32 && !swap) {
          float x, y, z;
          std::memcpy(&x, p + props[iX].offset, 4);
          std::memcpy(&y, p + props[iY].offset, 4);
          std::memcpy(&z, p + props[iZ].offset, 4);
          pos[0] = int32_t(x * positionScale);
          pos[1] = int32_t(y * positionScale);
          pos[2] = int32_t(z * positionScale);
        } else {
          pos[0] = int32_t(loadReal(props[iX], p, swap) * positionScale);
          pos[1] = int32_t(loadReal(props[iY], p, swap) * positionScale);
          pos[2] = int32_t(loadReal(props[iZ], p, swap) * positionScale);
        }
        if (withColors) {
          auto& c = cloud.getColor(i + r);
          c[0] = p[props[iG].offset];
          c[1] = p[props[iB].offset];
          c[2] = p[props[iR].offset];
        }
        if (iRefl != none)
          cloud.getReflectance(i + r) = props[iRefl].bytes == 1 ? attr_t(p[props[iRefl].offset])
                                                                 : loadAs<uint16_t>(p + props[iRefl].offset, swap);
        if (iFrame != none)
          cloud.getFrameIndex(i + r) = props[iFrame].bytes == 1 ? p[props[iFrame].offset]
                                                                 : uint8_t(loadAs<uint16_t>(p + props[iFrame].offset, swap));
      }
      i += got;
      if (got < want)
        break;
    }
  }
  std::fclose(f);
  return true;
}

}  // namespace ply
}  // namespace pcc

// PCCPointSet.h -- the PCCPointSet3 container of the segmentation path.
// Mirrors the part of the reference's tmc3/PCCPointSet.h:64-614 the path touches: AoS int32
// positions (operator[] returns a reference, :271-275), optional uint16 [G,B,R] colours
// (:289-293), optional reflectances / frame index, and the authors' public `planeIdx` (:67).
// Positions are contiguous N x 3 int32, which is what the C ABI (include/bseg.h) consumes.
#pragma once
#include <cassert>
#include <cstddef>
#include <cstdint>
#include <vector>

#include "PCCMath.h"

namespace pcc {

class PCCPointSet3 {
public:
  typedef point_t PointType;

  std::vector<int> planeIdx;  // my_function.h:103 resizes it to N, -1 = unassigned

  PCCPointSet3() = default;

  PointType& operator[](const size_t i) { assert(i < positions.size()); return positions[i]; }
  const PointType& operator[](const size_t i) const { assert(i < positions.size()); return positions[i]; }

  Vec3<attr_t>& getColor(const size_t i) { assert(withColors && i < colors.size()); return colors[i]; }
  const Vec3<attr_t>& getColor(const size_t i) const { assert(withColors && i < colors.size()); return colors[i]; }
  void setColor(const size_t i, const Vec3<attr_t>& c) { assert(withColors && i < colors.size()); colors[i] = c; }

  attr_t& getReflectance(const size_t i) { return reflectances[i]; }
  const attr_t& getReflectance(const size_t i) const { return reflectances[i]; }
  void setReflectance(const size_t i, attr_t r) { reflectances[i] = r; }
  uint8_t& getFrameIndex(const size_t i) { return frameidx[i]; }
  const uint8_t& getFrameIndex(const size_t i) const { return frameidx[i]; }

  bool hasColors() const { return withColors; }
  bool hasReflectances() const { return withReflectances; }
  bool hasFrameIndex() const { return withFrameIndex; }
  void addColors() { withColors = true; colors.resize(positions.size()); }
  void removeColors() { withColors = false; colors.clear(); }
  void addReflectances() { withReflectances = true; reflectances.resize(positions.size()); }
  void removeReflectances() { withReflectances = false; reflectances.clear(); }
  void addFrameIndex() { withFrameIndex = true; frameidx.resize(positions.size()); }
  void removeFrameIndex() { withFrameIndex = false; frameidx.clear(); }
  void addRemoveAttributes(bool wantColors, bool wantReflectances)
  {
    if (wantColors) addColors(); else removeColors();
    if (wantReflectances) addReflectances(); else removeReflectances();
  }

  size_t getPointCount() const { return positions.size(); }
  void resize(const size_t n)
  {
    positions.resize(n);
    if (withColors) colors.resize(n);
    if (withReflectances) reflectances.resize(n);
    if (withFrameIndex) frameidx.resize(n);
  }
  void reserve(const size_t n) { positions.reserve(n); }
  void clear()
  {
    positions.clear(); colors.clear(); reflectances.clear(); frameidx.clear(); planeIdx.clear();
  }
  void swapPoints(size_t a, size_t b)
  {
    std::swap(positions[a], positions[b]);
    if (withColors) std::swap(colors[a], colors[b]);
    if (withReflectances) std::swap(reflectances[a], reflectances[b]);
    if (withFrameIndex) std::swap(frameidx[a], frameidx[b]);
  }
  Box3<int32_t> computeBoundingBox() const
  {
    Box3<int32_t> b(std::numeric_limits<int32_t>::max(), std::numeric_limits<int32_t>::lowest());
    for (const auto& p : positions) b.insert(p);
    return b;
  }

  // contiguous views handed to the C ABI
  int32_t* positionData() { return positions.empty() ? nullptr : &positions[0][0]; }
  const int32_t* positionData() const { return positions.empty() ? nullptr : &positions[0][0]; }
  attr_t* colorData() { return colors.empty() ? nullptr : &colors[0][0]; }

private:
  std::vector<PointType> positions;
  std::vector<Vec3<attr_t>> colors;
  std::vector<attr_t> reflectances;
  std::vector<uint8_t> frameidx;
  bool withColors = false;
  bool withReflectances = false;
  bool withFrameIndex = false;
};

}  // namespace pcc

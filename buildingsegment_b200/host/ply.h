// ply.h -- PLY reader / writer with the reference's signatures (tmc3/ply.h:51-85).
#pragma once
#include <array>
#include <string>

#include "PCCMath.h"
#include "PCCPointSet.h"

namespace pcc {
namespace ply {

struct PropertyNameMap {
  // names of the position properties, e.g. {"x", "y", "z"}
  std::array<const char*, 3> position;
};

// ply.cpp:88-186 -- header: format, `element vertex N`, float64 x/y/z (float in ascii), uchar
// green/blue/red when the cloud has colours, uint16 refc, uint8 frameindex, `element face 0` +
// list property; body: position * scale + offset as 3 doubles, then the attributes.
bool write(const PCCPointSet3& cloud, const PropertyNameMap& attributeNames, double positionScale,
           Vec3<double> positionOffset, const std::string& fileName, bool asAscii);

// ply.cpp:190-504 -- positions are int32(value * scale) (C++ truncation toward zero, :407-409,
// :437-465); colours are detected by property name (red/green/blue, 1 byte) and stored [G, B, R].
bool read(const std::string& fileName, const PropertyNameMap& attributeNames, double positionScale,
          PCCPointSet3& cloud);

}  // namespace ply
}  // namespace pcc

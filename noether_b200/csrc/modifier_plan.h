// Host side of the tool-path modifier chain: from the structure of the input (offsets only — the poses stay on the
// device) and the chain, the sequence of device passes with their launch tables and constants.
#pragma once
#include "../../include/noether_b200.h"
#include "modifier_ops.h"

#include <cstdint>
#include <vector>

namespace nb2
{
namespace mod
{
/** tool paths -> segments -> waypoints, as offsets */
struct Structure
{
  std::vector<uint32_t> path_off;  // n_paths + 1, indexes segments
  std::vector<uint32_t> seg_off;   // n_segments + 1, indexes waypoints
  std::vector<int64_t> path_plane; // n_paths
  uint64_t paths() const { return path_off.empty() ? 0 : path_off.size() - 1; }
  uint64_t segments() const { return seg_off.empty() ? 0 : seg_off.size() - 1; }
  uint64_t waypoints() const { return seg_off.empty() ? 0 : seg_off.back(); }
};

/** One modifier of the chain, ready to launch */
struct Step
{
  int kind = 0;
  bool device = false;       // false: bookkeeping only (Concatenate; empty inputs)
  bool restructure = false;  // true: out-of-place gather through `map`; false: waypoint i -> waypoint i
  Structure out;             // structure after the step
  std::vector<SegMap> map;   // one entry per output segment (restructuring steps)
  double R[2][9] = {};       // rotation for sign > 0 / sign < 0 (tool drag), R[0] for UniformOrientation (row-major)
  double v[16] = {};         // constants of the per-waypoint operation (see modifiers.cu)
  int n_points = 0;
  std::vector<double> rot;   // circular leads: rotation of step i for sign > 0 at [9 i], for sign < 0 at [9 (n_points + i)]
};

/** What a data-dependent modifier needs to see of the poses (they live on the device): implemented by the executor */
struct Probe
{
  virtual ~Probe() = default;
  /** translation of the first and of the last waypoint of every segment (6 doubles per segment; zeros for an empty
   *  segment) and estimateToolPathDirection of tool path 0 (utils.cpp:7-29) */
  virtual void rasterInputs(const Structure& cur, std::vector<double>& seg_ends, double dir0[3]) = 0;
  /** polyline length of every segment: the chord lengths added up in waypoint order (computeLength, utils.cpp:213-234) */
  virtual void segmentLengths(const Structure& cur, std::vector<double>& lengths) = 0;
};

/** Eigen::AngleAxisd(angle, axis).toRotationMatrix() for a unit axis, row-major (Eigen 3.4 Geometry/AngleAxis.h) */
void angleAxisMatrix(double angle, const double axis[3], double R[9]);

/** The pass for modifier `m` on tool paths of structure `cur`.  Throws nb2::Error (via fail) where the reference would
 *  throw or read out of bounds.  RasterOrganization decides on the poses themselves: it asks `probe`. */
Step buildStep(const Structure& cur, const nb2_modifier& m, Probe* probe);
void checkStructure(const Structure& s);
}  // namespace mod
}  // namespace nb2

// TEST INFRASTRUCTURE ONLY — see oracle.h.
#pragma once
#include "oracle.h"
#include "oracle_math.h"

#include <array>
#include <string>
#include <vector>

namespace orc
{
extern std::string g_last_error;
void setError(const std::string& msg);

/** noether::ToolPathWaypoint (Eigen::Isometry3d, 4x4 column-major) + provenance of the cut point */
struct Waypoint
{
  double T[16];          // column-major 4x4
  float pos[3];          // float32 cut point as stored by vtkCutter's output vtkPoints
  float nrm[3];          // float32 interpolated (un-normalised) point normal as stored by vtkCutter
  uint32_t edge_lo;      // mesh vertex with the lower plane scalar on the generating edge
  uint32_t edge_hi;      // mesh vertex with the higher plane scalar
};
using Segment = std::vector<Waypoint>;
struct Path
{
  int64_t plane = -1;  // lattice index
  std::vector<Segment> segments;
};
}  // namespace orc

struct orc_paths
{
  std::vector<orc::Path> paths;
};

namespace orc
{
void eigenSolver3f(const float A[9], float evals[3], float evecs[9]);
int pca(const float* xyz, uint64_t V, int mode, orc_pca_result* out);
int centroid(const float* xyz, uint64_t V, int mode, double out[3]);
int principalAxisDirection(const orc_pca_result* pca, double rotation_offset, double out[3]);
int pcaRotatedDirection(const orc_pca_result* pca, double rotation_angle, const double nominal[3], double out[3]);
int makeLattice(const orc_pca_result* pca,
                const double cut_direction[3],
                const double cut_origin[3],
                double line_spacing,
                int bidirectional,
                orc_lattice* out);
int normalsFromMeshFaces(const float* xyz, uint64_t V, const uint32_t* tri, uint64_t F, float* out);

int slice(const float* xyz,
          const float* nrm,
          uint64_t V,
          const uint32_t* tri,
          uint64_t F,
          const orc_lattice& L,
          int slice_mode,
          int scan_mode,
          uint64_t plane_begin,
          uint64_t plane_end,
          orc_paths& out);

int rasterOrganization(orc_paths& p);
int directionOfTravel(orc_paths& p);
int joinCloseSegments(orc_paths& p, double distance);
int minimumSegmentLength(orc_paths& p, double min_length);
int uniformSpacing(orc_paths& p, double point_spacing);
int snakeOrganization(orc_paths& p);
int linearApproachDeparture(orc_paths& p, const double offset[3], int n_points, bool departure);
int offsetModifier(orc_paths& p, const double offset16[16]);
int fixedOrientation(orc_paths& p, const double ref_x[3]);
int uniformOrientation(orc_paths& p);
int concatenate(orc_paths& p);
int toolDrag(orc_paths& p, double angle_offset, double tool_radius, bool biased);
int movingAverageOrientationSmoothing(orc_paths& p, int window_size);
int circularLead(orc_paths& p, double arc_angle, double arc_radius, int n_points, bool lead_out);

inline V3d translation(const Waypoint& w) { return { w.T[12], w.T[13], w.T[14] }; }
}  // namespace orc

// Host entry points of the frame arithmetic (frame_math.h): used by the explicit C-ABI calls (nb2_make_lattice,
// nb2_principal_axis_direction).  nb2_plan evaluates the same functions on the device (kernels_frame.cu).
#include "frame_math.h"

#include <cmath>

namespace nb2
{
void eigenSolver3f(const float A[9], float evals[3], float evecs[9]) { frame::eigenSolver3f(A, evals, evecs); }

void principalAxisDirection(const nb2_pca& pca, double rotation_offset, double out[3])
{
  frame::principalAxisDirection(pca, std::sin(rotation_offset), std::cos(rotation_offset), out);
}

void pcaRotatedDirection(const nb2_pca& pca, double rotation_angle, const double nominal[3], double out[3])
{
  frame::pcaRotatedDirection(pca, std::sin(rotation_angle), std::cos(rotation_angle), nominal, out);
}

const char* latticeErrorMessage(int code)
{
  switch (code)
  {
    case frame::kLatticeBadSpacing:
      return "line_spacing must be a positive finite number";
    case frame::kLatticeNoNormal:
      return "cut direction is parallel to the mesh normal (or not finite): no cutting plane normal";
    case frame::kLatticeTooManyPlanes:
      return "line_spacing yields more than 2^31 cutting planes";
    default:
      return "";
  }
}

int makeLattice(const nb2_pca& pca, const double dir[3], const double origin[3], double spacing, int bidirectional,
                nb2_lattice& L)
{
  const int code = frame::makeLattice(pca, dir, origin, spacing, bidirectional, L);
  if (code != frame::kLatticeOk)
  {
    setLastError(latticeErrorMessage(code));
    return NB2_ERR_INVALID_ARGUMENT;
  }
  return NB2_OK;
}

}  // namespace nb2

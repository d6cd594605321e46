// TEST INFRASTRUCTURE ONLY — see oracle.h.
//
// PCA / centroid / lattice / vertex-normal restatements.
//   pcl::PCA::initCompute            -> orc_pca                (called at plane_slicer_raster_planner.cpp:540-541)
//   Eigen::SelfAdjointEigenSolver    -> orc_eigen_solver_3f    (3P, Eigen 3.4 iterative path, float)
//   pcl::PCA::project + getMinMax3D  -> orc_pca (scales)       (plane_slicer_raster_planner.cpp:544-548)
//   PrincipalAxisDirectionGenerator  -> orc_principal_axis_direction (principal_axis_direction_generator.cpp:14-26)
//   CentroidOriginGenerator          -> orc_centroid           (centroid_origin_generator.cpp:8-17)
//   getDistancesToMinMaxCuts + lattice -> orc_make_lattice     (plane_slicer_raster_planner.cpp:460-501,550-586)
//   NormalsFromMeshFacesMeshModifier -> orc_normals_from_mesh_faces (normals_from_mesh_faces_modifier.cpp:15-46)
#include "oracle.h"
#include "oracle_internal.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <limits>
#include <vector>

namespace orc
{
std::string g_last_error;
void setError(const std::string& msg) { g_last_error = msg; }

// ---------------------------------------------------------------------------------------------------------------
// Eigen 3.4: JacobiRotation<float>::makeGivens(p, q) (real case, r not requested)
// ---------------------------------------------------------------------------------------------------------------
static void makeGivens(float p, float q, float& c, float& s)
{
  if (q == 0.f)
  {
    c = p < 0.f ? -1.f : 1.f;
    s = 0.f;
  }
  else if (p == 0.f)
  {
    c = 0.f;
    s = q < 0.f ? 1.f : -1.f;
  }
  else if (std::abs(p) > std::abs(q))
  {
    float t = q / p;
    float u = std::sqrt(1.f + t * t);
    if (p < 0.f)
      u = -u;
    c = 1.f / u;
    s = -t * c;
  }
  else
  {
    float t = p / q;
    float u = std::sqrt(1.f + t * t);
    if (q < 0.f)
      u = -u;
    s = -1.f / u;
    c = -t * s;
  }
}

// Eigen 3.4: internal::tridiagonal_qr_step<ColMajor>(diag, subdiag, start, end, matrixQ, n) for n == 3
static void tridiagonalQrStep(float* diag, float* subdiag, int start, int end, float* Q /*col-major 3x3*/)
{
  float td = (diag[end - 1] - diag[end]) * 0.5f;
  float e = subdiag[end - 1];
  float mu = diag[end];
  if (td == 0.f)
  {
    mu -= std::abs(e);
  }
  else if (e != 0.f)
  {
    const float e2 = e * e;
    // numext::hypot -> internal::positive_real_hypot(|x|, |y|) = p * sqrt(1 + (q/p)^2) (Eigen/src/Core/MathFunctionsImpl.h)
    const float ax = std::abs(td), ay = std::abs(e);
    const float pmax = std::max(ax, ay);
    const float qp = pmax == 0.f ? 0.f : std::min(ay, ax) / pmax;
    const float h = pmax == 0.f ? 0.f : pmax * std::sqrt(1.f + qp * qp);
    if (e2 == 0.f)
      mu -= e / ((td + (td > 0.f ? h : -h)) / e);
    else
      mu -= e2 / (td + (td > 0.f ? h : -h));
  }

  float x = diag[start] - mu;
  float z = subdiag[start];
  for (int k = start; k < end && z != 0.f; ++k)
  {
    float c, s;
    makeGivens(x, z, c, s);

    // T = G' T G
    const float sdk = s * diag[k] + c * subdiag[k];
    const float dkp1 = s * subdiag[k] + c * diag[k + 1];

    diag[k] = c * (c * diag[k] - s * subdiag[k]) - s * (c * subdiag[k] - s * diag[k + 1]);
    diag[k + 1] = s * sdk + c * dkp1;
    subdiag[k] = c * sdk - s * dkp1;

    if (k > start)
      subdiag[k - 1] = c * subdiag[k - 1] - s * z;

    x = subdiag[k];
    if (k < end - 1)
    {
      z = -s * subdiag[k + 1];
      subdiag[k + 1] = c * subdiag[k + 1];
    }

    // Q = Q * G : q.applyOnTheRight(k, k+1, rot)  ==  apply_rotation_in_the_plane(col k, col k+1, rot.transpose())
    for (int i = 0; i < 3; ++i)
    {
      const float xi = Q[k * 3 + i];
      const float yi = Q[(k + 1) * 3 + i];
      Q[k * 3 + i] = c * xi - s * yi;
      Q[(k + 1) * 3 + i] = s * xi + c * yi;
    }
  }
}

/** Eigen 3.4 SelfAdjointEigenSolver<Matrix3f>::compute(matrix, ComputeEigenvectors) — iterative (non-direct) path.
 *  Reads the lower triangle only.  Output eigenvalues ascending, eigenvectors as columns (column-major). */
void eigenSolver3f(const float A[9], float evals[3], float evecs[9])
{
  auto a = [&](int r, int c) { return A[c * 3 + r]; };
  // map the coefficients to [-1:1]: scale = max |lower-triangular coefficient|
  float scale = 0.f;
  for (int c = 0; c < 3; ++c)
    for (int r = c; r < 3; ++r)
      scale = std::max(scale, std::abs(a(r, c)));
  if (scale == 0.f)
    scale = 1.f;
  float m[3][3];
  for (int c = 0; c < 3; ++c)
    for (int r = c; r < 3; ++r)
      m[r][c] = a(r, c) / scale;

  // internal::tridiagonalization_inplace_selector<MatrixType, 3, false>::run
  float diag[3], subdiag[2];
  float Q[9];
  {
    const float tol = std::numeric_limits<float>::min();
    diag[0] = m[0][0];
    const float v1norm2 = m[2][0] * m[2][0];
    if (v1norm2 <= tol)
    {
      diag[1] = m[1][1];
      diag[2] = m[2][2];
      subdiag[0] = m[1][0];
      subdiag[1] = m[2][1];
      const float I[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
      std::copy(I, I + 9, Q);
    }
    else
    {
      const float beta = std::sqrt(m[1][0] * m[1][0] + v1norm2);
      const float invBeta = 1.f / beta;
      const float m01 = m[1][0] * invBeta;
      const float m02 = m[2][0] * invBeta;
      const float q = 2.f * m01 * m[2][1] + m02 * (m[2][2] - m[1][1]);
      diag[1] = m[1][1] + m02 * q;
      diag[2] = m[2][2] - m02 * q;
      subdiag[0] = beta;
      subdiag[1] = m[2][1] - m01 * q;
      // mat << 1, 0, 0,  0, m01, m02,  0, m02, -m01   (row-wise comma initialiser) stored column-major here
      const float Qrow[9] = { 1, 0, 0, 0, m01, m02, 0, m02, -m01 };
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
          Q[c * 3 + r] = Qrow[r * 3 + c];
    }
  }

  // internal::computeFromTridiagonal_impl
  {
    const int n = 3;
    const int maxIterations = 30;
    int end = n - 1;
    int start = 0;
    int iter = 0;
    const float considerAsZero = std::numeric_limits<float>::min();
    const float precision_inv = 1.f / std::numeric_limits<float>::epsilon();
    while (end > 0)
    {
      for (int i = start; i < end; ++i)
      {
        if (std::abs(subdiag[i]) < considerAsZero)
        {
          subdiag[i] = 0.f;
        }
        else
        {
          const float scaled_subdiag = precision_inv * subdiag[i];
          if (scaled_subdiag * scaled_subdiag <= (std::abs(diag[i]) + std::abs(diag[i + 1])))
            subdiag[i] = 0.f;
        }
      }
      while (end > 0 && subdiag[end - 1] == 0.f)
        end--;
      if (end <= 0)
        break;
      iter++;
      if (iter > maxIterations * n)
        break;
      start = end - 1;
      while (start > 0 && subdiag[start - 1] != 0.f)
        start--;
      tridiagonalQrStep(diag, subdiag, start, end, Q);
    }
    // sort eigenvalues (ascending) and the corresponding vectors: selection by minCoeff of the tail
    for (int i = 0; i < n - 1; ++i)
    {
      int k = 0;
      for (int j = 1; j < n - i; ++j)
        if (diag[i + j] < diag[i + k])
          k = j;
      if (k > 0)
      {
        std::swap(diag[i], diag[k + i]);
        for (int r = 0; r < 3; ++r)
          std::swap(Q[i * 3 + r], Q[(k + i) * 3 + r]);
      }
    }
  }

  for (int i = 0; i < 3; ++i)
    evals[i] = diag[i] * scale;
  std::copy(Q, Q + 9, evecs);
}

// ---------------------------------------------------------------------------------------------------------------
// pcl::PCA<PointXYZ>::initCompute + project + getMinMax3D
// ---------------------------------------------------------------------------------------------------------------
int pca(const float* xyz, uint64_t V, int mode, orc_pca_result* out)
{
  if (V < 3)
  {
    setError("[pcl::PCA::initCompute] number of points < 3");
    return ORC_ERR_TOO_FEW_POINTS;
  }

  float mean[3];
  float cov[9];
  if (mode == ORC_MOMENTS_F32_SEQUENTIAL)
  {
    // pcl::compute3DCentroid: float accumulators in vertex order, then centroid /= N (P1)
    float acc[3] = { 0.f, 0.f, 0.f };
    for (uint64_t i = 0; i < V; ++i)
    {
      acc[0] += xyz[3 * i + 0];
      acc[1] += xyz[3 * i + 1];
      acc[2] += xyz[3 * i + 2];
    }
    const float n = static_cast<float>(V);
    for (int c = 0; c < 3; ++c)
      mean[c] = acc[c] / n;

    // alpha = (1 / (N - 1)) * X * X^T with X the demeaned cloud, float.  Eigen evaluates this as a GEMM whose
    // blocked/vectorised summation order is not restated here: sequential float accumulation stands in (documented
    // tolerance-level parity).
    float acc2[6] = { 0, 0, 0, 0, 0, 0 };  // xx xy xz yy yz zz
    for (uint64_t i = 0; i < V; ++i)
    {
      const float dx = xyz[3 * i + 0] - mean[0];
      const float dy = xyz[3 * i + 1] - mean[1];
      const float dz = xyz[3 * i + 2] - mean[2];
      acc2[0] += dx * dx;
      acc2[1] += dx * dy;
      acc2[2] += dx * dz;
      acc2[3] += dy * dy;
      acc2[4] += dy * dz;
      acc2[5] += dz * dz;
    }
    const float f = 1.f / (static_cast<float>(V) - 1.f);
    const float xx = f * acc2[0], xy = f * acc2[1], xz = f * acc2[2], yy = f * acc2[3], yz = f * acc2[4],
                zz = f * acc2[5];
    const float C[9] = { xx, xy, xz, xy, yy, yz, xz, yz, zz };
    std::copy(C, C + 9, cov);
  }
  else
  {
    // double accumulation about the first vertex (pivot), cast to float at the end
    const double px = xyz[0], py = xyz[1], pz = xyz[2];
    long double s[3] = { 0, 0, 0 };
    long double q[6] = { 0, 0, 0, 0, 0, 0 };
    for (uint64_t i = 0; i < V; ++i)
    {
      const double dx = static_cast<double>(xyz[3 * i + 0]) - px;
      const double dy = static_cast<double>(xyz[3 * i + 1]) - py;
      const double dz = static_cast<double>(xyz[3 * i + 2]) - pz;
      s[0] += dx;
      s[1] += dy;
      s[2] += dz;
      q[0] += dx * dx;
      q[1] += dx * dy;
      q[2] += dx * dz;
      q[3] += dy * dy;
      q[4] += dy * dz;
      q[5] += dz * dz;
    }
    const long double n = static_cast<long double>(V);
    const long double m[3] = { s[0] / n, s[1] / n, s[2] / n };
    mean[0] = static_cast<float>(static_cast<double>(px + m[0]));
    mean[1] = static_cast<float>(static_cast<double>(py + m[1]));
    mean[2] = static_cast<float>(static_cast<double>(pz + m[2]));
    const long double d = n - 1.0L;
    const float xx = static_cast<float>(static_cast<double>((q[0] - n * m[0] * m[0]) / d));
    const float xy = static_cast<float>(static_cast<double>((q[1] - n * m[0] * m[1]) / d));
    const float xz = static_cast<float>(static_cast<double>((q[2] - n * m[0] * m[2]) / d));
    const float yy = static_cast<float>(static_cast<double>((q[3] - n * m[1] * m[1]) / d));
    const float yz = static_cast<float>(static_cast<double>((q[4] - n * m[1] * m[2]) / d));
    const float zz = static_cast<float>(static_cast<double>((q[5] - n * m[2] * m[2]) / d));
    const float C[9] = { xx, xy, xz, xy, yy, yz, xz, yz, zz };
    std::copy(C, C + 9, cov);
  }

  float ev_asc[3], evec_asc[9];
  eigenSolver3f(cov, ev_asc, evec_asc);
  // pcl::PCA: "organize eigenvectors and eigenvalues in ascendent order": column i <- solver column 2 - i
  for (int i = 0; i < 3; ++i)
  {
    out->evals[i] = ev_asc[2 - i];
    for (int r = 0; r < 3; ++r)
      out->evecs[i * 3 + r] = evec_asc[(2 - i) * 3 + r];
  }
  std::copy(mean, mean + 3, out->mean);
  std::copy(cov, cov + 9, out->cov);

  // pca.project(cloud, proj): proj_i = E^T (x_i - mean), float; getMinMax3D(proj): float min/max (P1, P3)
  float mn[3] = { FLT_MAX, FLT_MAX, FLT_MAX };
  float mx[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
  for (uint64_t i = 0; i < V; ++i)
  {
    const float d[3] = { xyz[3 * i + 0] - mean[0], xyz[3 * i + 1] - mean[1], xyz[3 * i + 2] - mean[2] };
    for (int c = 0; c < 3; ++c)
    {
      const float* e = &out->evecs[c * 3];
      const float p = eigenSum3(e[0] * d[0], e[1] * d[1], e[2] * d[2]);
      mn[c] = std::min(mn[c], p);
      mx[c] = std::max(mx[c], p);
    }
  }
  for (int c = 0; c < 3; ++c)
    out->scales[c] = mx[c] - mn[c];
  return ORC_OK;
}

int centroid(const float* xyz, uint64_t V, int mode, double out[3])
{
  if (V < 1)
  {
    setError("Failed to compute mesh centroid");  // centroid_origin_generator.cpp:13-14
    return ORC_ERR_TOO_FEW_POINTS;
  }
  if (mode == ORC_MOMENTS_F32_SEQUENTIAL)
  {
    // pcl::computeCentroid -> CentroidPoint<PointXYZ>: Eigen::Vector3f sum in vertex order, divided by n (P2)
    float acc[3] = { 0.f, 0.f, 0.f };
    for (uint64_t i = 0; i < V; ++i)
      for (int c = 0; c < 3; ++c)
        acc[c] += xyz[3 * i + c];
    const float n = static_cast<float>(V);
    for (int c = 0; c < 3; ++c)
      out[c] = static_cast<double>(acc[c] / n);
  }
  else
  {
    const double p[3] = { xyz[0], xyz[1], xyz[2] };
    long double s[3] = { 0, 0, 0 };
    for (uint64_t i = 0; i < V; ++i)
      for (int c = 0; c < 3; ++c)
        s[c] += static_cast<double>(xyz[3 * i + c]) - p[c];
    for (int c = 0; c < 3; ++c)
      out[c] = static_cast<double>(static_cast<float>(static_cast<double>(p[c] + s[c] / static_cast<long double>(V))));
  }
  return ORC_OK;
}

/** Eigen::AngleAxisd(angle, axis).toRotationMatrix() * v */
static V3d angleAxisRotate(double angle, const V3d& axis, const V3d& v)
{
  const V3d sin_axis = std::sin(angle) * axis;
  const double c = std::cos(angle);
  const V3d cos1_axis = (1.0 - c) * axis;
  double R[3][3];
  double tmp;
  tmp = cos1_axis.x * axis.y;
  R[0][1] = tmp - sin_axis.z;
  R[1][0] = tmp + sin_axis.z;
  tmp = cos1_axis.x * axis.z;
  R[0][2] = tmp + sin_axis.y;
  R[2][0] = tmp - sin_axis.y;
  tmp = cos1_axis.y * axis.z;
  R[1][2] = tmp - sin_axis.x;
  R[2][1] = tmp + sin_axis.x;
  R[0][0] = cos1_axis.x * axis.x + c;
  R[1][1] = cos1_axis.y * axis.y + c;
  R[2][2] = cos1_axis.z * axis.z + c;
  V3d r;
  for (int i = 0; i < 3; ++i)
    r[i] = eigenSum3(R[i][0] * v.x, R[i][1] * v.y, R[i][2] * v.z);
  return r;
}

int principalAxisDirection(const orc_pca_result* pca, double rotation_offset, double out[3])
{
  // principal_axis_direction_generator.cpp:20-25
  const V3d e1 = { pca->evecs[3], pca->evecs[4], pca->evecs[5] };
  const V3d e2 = { pca->evecs[6], pca->evecs[7], pca->evecs[8] };
  const V3d r = angleAxisRotate(rotation_offset, eigenNormalized(e2), eigenNormalized(e1));
  out[0] = r.x;
  out[1] = r.y;
  out[2] = r.z;
  return ORC_OK;
}

int pcaRotatedDirection(const orc_pca_result* pca, double rotation_angle, const double nominal[3], double out[3])
{
  // pca_rotated_direction_generator.cpp:24: AngleAxisd(rotation_angle_, pca_vecs.col(2).normalized()) * nominal_dir
  // (the nominal direction is NOT normalised, unlike PrincipalAxis' e1)
  const V3d e2 = { pca->evecs[6], pca->evecs[7], pca->evecs[8] };
  const V3d v = { nominal[0], nominal[1], nominal[2] };
  const V3d r = angleAxisRotate(rotation_angle, eigenNormalized(e2), v);
  out[0] = r.x;
  out[1] = r.y;
  out[2] = r.z;
  return ORC_OK;
}

int makeLattice(const orc_pca_result* pca,
                const double cut_direction_[3],
                const double cut_origin_[3],
                double line_spacing,
                int bidirectional,
                orc_lattice* L)
{
  if (!(line_spacing > 0.0) || !std::isfinite(line_spacing))
  {
    setError("line_spacing must be a positive finite number");
    return ORC_ERR_INVALID;
  }
  // plane_slicer_raster_planner.cpp:548-552
  const V3d e2f = { pca->evecs[6], pca->evecs[7], pca->evecs[8] };
  const V3d mesh_normal = eigenNormalized(e2f);
  const V3d centroid = { pca->mean[0], pca->mean[1], pca->mean[2] };
  // pca_vecs = (eigenvectors.array().rowwise() * scales.transpose()).cast<double>(): column j scaled by scales[j] (float)
  double pca_vecs[3][3];  // [row][col]
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r)
      pca_vecs[r][c] = static_cast<double>(pca->evecs[c * 3 + r] * pca->scales[c]);

  const V3d cut_direction = { cut_direction_[0], cut_direction_[1], cut_direction_[2] };
  const V3d cut_origin = { cut_origin_[0], cut_origin_[1], cut_origin_[2] };
  // :557
  const V3d cut_normal = eigenNormalized(eigenCross(eigenNormalized(cut_direction), mesh_normal));

  // getDistancesToMinMaxCuts (:460-501)
  static const int mask[3][8] = { { 1, -1, 1, -1, 1, -1, 1, -1 }, { 1, 1, -1, -1, 1, 1, -1, -1 }, { 1, 1, 1, 1, -1, -1, -1, -1 } };
  double d_max = -std::numeric_limits<double>::max();
  double d_min = std::numeric_limits<double>::max();
  for (int col = 0; col < 8; ++col)
  {
    V3d corner;
    for (int r = 0; r < 3; ++r)
    {
      // (obb_size / 2.0) * corner_mask.cast<double>() : row r of a 3x3 times column of +-1, then += centroid
      const double h0 = pca_vecs[r][0] / 2.0, h1 = pca_vecs[r][1] / 2.0, h2 = pca_vecs[r][2] / 2.0;
      corner[r] = eigenSum3(h0 * mask[0][col], h1 * mask[1][col], h2 * mask[2][col]) + centroid[r];
    }
    // Hyperplane(n, e): offset = -n.dot(e); signedDistance(p) = n.dot(p) + offset
    const double offset = -eigenDot(cut_normal, corner);
    double d = eigenDot(cut_normal, cut_origin) + offset;
    d *= -1;
    if (d > d_max)
      d_max = d;
    else if (d < d_min)
      d_min = d;
  }

  L->no_cuts = 0;
  if (!bidirectional)
  {
    if (d_max < 0.0)
      L->no_cuts = 1;
    if (d_min < 0.0)
      d_min = 0.0;
  }

  const double cut_span = std::abs(d_max - d_min);
  const auto num_planes = static_cast<std::size_t>(std::ceil(cut_span / line_spacing));
  const auto n_planes_start_offset = static_cast<std::size_t>(std::floor(std::abs(d_min) / line_spacing));
  const double start_offset = n_planes_start_offset * line_spacing;
  const V3d start_loc = cut_origin + cut_normal * std::copysign(start_offset, d_min);

  for (int c = 0; c < 3; ++c)
  {
    L->cut_direction[c] = cut_direction[c];
    L->cut_normal[c] = cut_normal[c];
    L->cut_origin[c] = cut_origin[c];
    L->start_loc[c] = start_loc[c];
  }
  L->line_spacing = line_spacing;
  L->d_min = d_min;
  L->d_max = d_max;
  L->num_planes = L->no_cuts ? 0 : num_planes;
  L->pad_ = 0;
  return ORC_OK;
}

int normalsFromMeshFaces(const float* xyz, uint64_t V, const uint32_t* tri, uint64_t F, float* out)
{
  std::fill(out, out + 3 * V, 0.f);
  for (uint64_t f = 0; f < F; ++f)
  {
    const uint32_t i0 = tri[3 * f + 0], i1 = tri[3 * f + 1], i2 = tri[3 * f + 2];
    if (i0 >= V || i1 >= V || i2 >= V)
    {
      setError("triangle references a vertex out of range");
      return ORC_ERR_INVALID;
    }
    // getFaceNormal (utils.cpp:154-161): (p1 - p0).cross(p_last - p0).normalized(), float
    const V3f p0 = { xyz[3 * i0], xyz[3 * i0 + 1], xyz[3 * i0 + 2] };
    const V3f p1 = { xyz[3 * i1], xyz[3 * i1 + 1], xyz[3 * i1 + 2] };
    const V3f pn = { xyz[3 * i2], xyz[3 * i2 + 1], xyz[3 * i2 + 2] };
    const V3f n = eigenNormalized(eigenCross(p1 - p0, pn - p0));
    // normals_map({0,1,2}, p.vertices).colwise() += n   (:25) — polygon order, float adds
    const uint32_t idx[3] = { i0, i1, i2 };
    for (int k = 0; k < 3; ++k)
    {
      out[3 * idx[k] + 0] += n.x;
      out[3 * idx[k] + 1] += n.y;
      out[3 * idx[k] + 2] += n.z;
    }
  }
  // colwise().normalize() (:29): Eigen 3.4 normalize() leaves a zero vector untouched
  for (uint64_t v = 0; v < V; ++v)
  {
    V3f n = { out[3 * v], out[3 * v + 1], out[3 * v + 2] };
    const float z = eigenSquaredNorm(n);
    if (z > 0.f)
    {
      const float s = std::sqrt(z);
      out[3 * v + 0] = n.x / s;
      out[3 * v + 1] = n.y / s;
      out[3 * v + 2] = n.z / s;
    }
  }
  return ORC_OK;
}

}  // namespace orc

// AABBCenterOriginGenerator::generate (aabb_center_origin_generator.cpp:9-17): pcl::getMinMax3D (float min / max over the
// points), Array3f difference, cast to double, / 2.0.  TEST INFRASTRUCTURE ONLY.
extern "C" int orc_aabb_center_origin(const float* xyz, uint64_t n, double out[3])
{
  if (!xyz || !out || n == 0)
  {
    orc::setError("invalid argument");
    return ORC_ERR_INVALID;
  }
  for (int i = 0; i < 3; ++i)
  {
    float mn = xyz[i], mx = xyz[i];
    for (uint64_t v = 1; v < n; ++v)
    {
      mn = std::min(mn, xyz[3 * v + i]);
      mx = std::max(mx, xyz[3 * v + i]);
    }
    const float range = mx - mn;
    out[i] = static_cast<double>(range) / 2.0;
  }
  return ORC_OK;
}

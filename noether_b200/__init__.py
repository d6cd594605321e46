"""noether_b200 — B200-native raster-slicing path of ros-industrial/noether (noether_tpp 0.16.0).

The product is libnoether_b200.so (hand-written sm_100a CUDA behind the C ABI in include/noether_b200.h) plus the
boost_plugin_loader shim in plugin/.  This Python package is the host-side mirror of the reference's component
interface used by tests and the benchmark; it only ever calls the C ABI.
"""
from . import _capi, meshes  # noqa: F401
from .planner import (AABBCenterOriginGenerator, CentroidOriginGenerator, Context, Factory, FixedDirectionGenerator,  # noqa: F401
                      FixedOriginGenerator, NormalsFromMeshFacesMeshModifier, NormalEstimationPCLMeshModifier, FaceMidpointSubdivisionMeshModifier,
                      FaceSubdivisionByAreaMeshModifier, EuclideanClusteringMeshModifier, cluster_order, cluster_submeshes, MultiToolPathPlanner, PlaneSlicerLegacyRasterPlanner,
                      PCARotatedDirectionGenerator, PlaneSlicerRasterPlanner, PrincipalAxisDirectionGenerator, ToolPaths,
                      default_context, make_lattice,
                      ToolPathModifier, SnakeOrganizationModifier, RasterOrganizationModifier,
                      JoinCloseSegmentsToolPathModifier, MinimumSegmentLengthToolPathModifier, LinearApproachModifier, LinearDepartureModifier,
                      OffsetModifier, FixedOrientationModifier, UniformOrientationModifier, ConcatenateModifier,
                      DirectionOfTravelOrientationModifier, MovingAverageOrientationSmoothingModifier,
                      CircularLeadInModifier, CircularLeadOutModifier,
                      ToolDragOrientationToolPathModifier,
                      BiasedToolDragOrientationToolPathModifier, CompoundModifier)

__version__ = "0.1.0"

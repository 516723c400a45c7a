"""tc2li_slam_b200 — B200-native ORB feature-extraction front-end, drop-in for TC2LI_SLAM::ORBextractor.

Hand-written sm_100a CUDA (csrc/) behind the C ABI of include/orbx.h; this package is the Python mirror of the
reference's class interface used by the parity tests and the benchmark.  No CPU fallback exists.
"""
from .orbx import (AREA_QUERY_DTYPE, KP_DTYPE, PROJ_QUERY_DTYPE, BatchExtractor, ExtractorPool, ORBVocabulary, ORBextractor, OrbxError, PinnedArray, StereoORBextractor, level_geometry, lib, tables)

__all__ = ["AREA_QUERY_DTYPE", "KP_DTYPE", "PROJ_QUERY_DTYPE", "BatchExtractor", "ExtractorPool", "ORBVocabulary", "ORBextractor", "OrbxError", "PinnedArray", "StereoORBextractor", "level_geometry", "lib", "tables"]

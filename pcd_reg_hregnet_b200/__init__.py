"""B200-native (sm_100a) HRegNet registration forward path behind the reference's API.

    from pcd_reg_hregnet_b200 import HRegNet, furthest_point_sample, knn_points, ...

The CUDA kernels live in libhregnet_b200.so (C ABI: include/hregnet_b200.h); nothing here falls back to the CPU."""
from .layers import (CoarseReg, DescExtractor, FineReg, KeypointDetector, WeightedSVDHead,  # noqa: F401
                     calc_cosine_similarity, knn_group)
from .model_v2 import FineReg1, FineReg2, Model_V2  # noqa: F401
from .model_v4 import Model_V4  # noqa: F401
from .models import HierFeatureExtraction, HRegNet  # noqa: F401
from .ops import (furthest_point_sample, gather_operation, knn_gather, knn_points,  # noqa: F401
                  weighted_furthest_point_sample)
from . import runner  # noqa: F401,E402  (Registrar: host-buffer serving API)

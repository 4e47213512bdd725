"""Mirror of learning3d/utils/__init__.py for the hot-path symbols (utils/__init__.py:1-23).

As in the reference, the names re-exported here are the model_common_utils variants (the later
import at utils/__init__.py:5-14 shadows the ppfnet_util ones); the pointconv_util / ppfnet_util
variants stay reachable under their own module paths.
"""
from .svd import SVDHead
from .transformer import Transformer, Identity
from .ppfnet_util import angle_difference, sample_and_group, sample_and_group_multi
from .model_common_utils import (
    knn,
    pc_normalize,
    square_distance,
    index_points,
    farthest_point_sample,
    knn_point,
    query_ball_point,
    get_graph_feature,
)
from . import pointconv_util, ppfnet_util
from .lib import pointnet2_utils

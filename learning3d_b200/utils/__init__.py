"""Mirror of learning3d/utils/__init__.py for the hot-path symbols (utils/__init__.py:1-23)."""
from .model_common_utils import knn, get_graph_feature, knn_point

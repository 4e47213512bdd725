"""Mirror of learning3d/losses/__init__.py for the hot-path losses (losses/__init__.py:5-12).
Unlike the reference there is no try/except apology: a missing CUDA library is an error."""
from .emd import EMDLoss
from .chamfer_distance import ChamferDistanceLoss, chamfer_distance

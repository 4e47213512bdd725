"""Drop-in for the reference's `_emd_ext` extension package (losses/cuda/emd_torch/setup.py builds
`_emd_ext._emd`; losses/cuda/emd_torch/pkg/layer/emd_loss_layer.py:4 imports it).  See _emd.py."""
from . import _emd  # noqa: F401

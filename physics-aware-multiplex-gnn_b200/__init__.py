"""pamnet_b200 -- B200-native PAMNet message-passing hot path behind the reference's module API.

    from pamnet_b200 import Config, PAMNet, PAMNet_s          # == reference models.py
    from pamnet_b200 import ops                               # radius / knn / scatter / bases on CUDA
    from pamnet_b200 import FusedAdamEMA                      # clip + Adam + EMA of main_qm9.py in two launches
    from pamnet_b200 import DeviceDataset                     # dataset resident in HBM, one-launch collate
"""
from .data import Batch, DeviceDataset, molecules_of, synthetic_qm9_batch, synthetic_rna_batch  # noqa: F401
from .models import Config, PAMNet, PAMNet_s  # noqa: F401
from . import layers, ops, optim  # noqa: F401
from .optim import FusedAdamEMA  # noqa: F401
from ._lib import PamnetError  # noqa: F401

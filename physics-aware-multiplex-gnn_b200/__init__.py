"""pamnet_b200 -- B200-native PAMNet message-passing hot path behind the reference's API."""
from .data import Batch, synthetic_qm9_batch, synthetic_rna_batch  # noqa: F401

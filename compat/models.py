"""Drop-in for the reference's ``models.py``: put this directory first on PYTHONPATH (together with the repo
root) and ``from models import PAMNet, PAMNet_s, Config`` (main_qm9.py:13, main_rna_puzzles.py, main_pdbbind.py,
inference_rna_puzzles.py:10) resolves to the B200 implementation."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pamnet_b200  # noqa: E402,F401
from pamnet_b200 import Config, PAMNet, PAMNet_s  # noqa: E402,F401

"""Puts the product source root on sys.path for the drop-in CLI scripts."""
import sys
from pathlib import Path

_ROOT = Path(__file__).resolve().parents[1] / "robust-multimodal-pd_b200"
if str(_ROOT) not in sys.path:
    sys.path.insert(0, str(_ROOT))

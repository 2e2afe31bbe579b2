"""Drop-in module name for the reference's Main.py (`from models import *`, Main.py:14)."""
from igmc_b200.models import IGMC, DGCNN_RS, RGCNConv, FusedAdam  # noqa: F401

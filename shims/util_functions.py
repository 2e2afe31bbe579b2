"""Drop-in module name for the reference's Main.py (`from util_functions import *`, Main.py:12)."""
from igmc_b200.util_functions import *  # noqa: F401,F403
from igmc_b200.util_functions import Batch, Data, MyDataset, MyDynamicDataset, RatingGraph, SubgraphExtractor  # noqa: F401

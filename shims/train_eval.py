"""Drop-in module name for the reference's Main.py (`from train_eval import *`, Main.py:15)."""
from igmc_b200.train_eval import (train_multiple_epochs, test_once, train, eval_loss, eval_rmse,  # noqa: F401
                                  eval_loss_ensemble, eval_rmse_ensemble, visualize, TrainEngine)

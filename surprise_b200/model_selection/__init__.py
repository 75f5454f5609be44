from .split import KFold, PredefinedKFold, RepeatedKFold, ShuffleSplit, train_test_split
from .validation import cross_validate, fit_and_score

__all__ = ["KFold", "PredefinedKFold", "RepeatedKFold", "ShuffleSplit", "train_test_split", "cross_validate",
           "fit_and_score"]

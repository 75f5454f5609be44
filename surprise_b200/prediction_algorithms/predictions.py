"""Prediction tuple and PredictionImpossible (reference: prediction_algorithms/predictions.py:13-50)."""
from collections import namedtuple


class PredictionImpossible(Exception):
    """Raised by estimate() when no sensible estimate exists; predict() falls back to the global mean."""
    pass


class Prediction(namedtuple("Prediction", ["uid", "iid", "r_ui", "est", "details"])):
    __slots__ = ()

    def __str__(self):
        s = "user: {uid:<10} ".format(uid=self.uid)
        s += "item: {iid:<10} ".format(iid=self.iid)
        s += "r_ui = {r_ui:1.2f}   ".format(r_ui=self.r_ui) if self.r_ui is not None else "r_ui = None   "
        s += "est = {est:1.2f}   ".format(est=self.est)
        s += str(self.details)
        return s

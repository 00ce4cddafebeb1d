"""load_estimator (reference: crl_kino/estimator/load_estimator.py:7-14)."""
import numpy as np
import torch

from .network import ClassifierModel, EstimatorModel, _Pair


def _load(path_or_sd):
    if isinstance(path_or_sd, dict):
        return path_or_sd
    if str(path_or_sd).endswith(".npz"):
        z = np.load(path_or_sd)
        return {k: z[k] for k in z.files}
    return torch.load(path_or_sd, map_location=torch.device("cpu"))


def load_estimator(estimator_path, classifier_path):
    """Paths to the reference's *_statedict.pth files (or dicts of arrays)."""
    pair = _Pair()
    estimator = EstimatorModel(pair)
    estimator.load_state_dict(_load(estimator_path))
    classifier = ClassifierModel(pair)
    classifier.load_state_dict(_load(classifier_path))
    return estimator, classifier

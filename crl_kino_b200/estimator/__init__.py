from .load_estimator import load_estimator  # noqa: F401
from .network import ClassifierModel, EstimatorModel  # noqa: F401

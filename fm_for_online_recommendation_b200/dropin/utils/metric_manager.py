"""import shim: `from utils.metric_manager import regression_metric, classfication_metric` resolves to the device versions"""
from fm_for_online_recommendation_b200.metrics import classfication_metric, regression_metric  # noqa: F401

"""Drop-in for the reference's models/models_online_deep/nfm_adam.py: same import path, same class name.
Put fm_for_online_recommendation_b200/dropin first on sys.path and main_experiment*.py runs unchanged."""
from fm_for_online_recommendation_b200.deep import NFMAdam  # noqa: F401

"""Drop-in for the reference's models/models_online/FM_FTRL.py: same import path, same class name."""
from fm_for_online_recommendation_b200.classical import FM_FTRL  # noqa: F401

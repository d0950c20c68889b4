"""Drop-in for the reference's models/models_online/SFTRL_Vanila.py: same import path, same class name."""
from fm_for_online_recommendation_b200.classical import SFTRL_Vanila  # noqa: F401

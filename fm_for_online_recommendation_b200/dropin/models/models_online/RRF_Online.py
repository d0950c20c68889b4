from fm_for_online_recommendation_b200.classical import RRF_Online  # noqa: F401

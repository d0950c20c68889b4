from fm_for_online_recommendation_b200.deep import AFMAdam  # noqa: F401

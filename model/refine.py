from pvraft_b200.refine import FlotRefine  # noqa: F401  (reference: model/refine.py:6)

from pvraft_b200.extractor import FlotEncoder  # noqa: F401  (reference: model/extractor.py:7)

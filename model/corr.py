from pvraft_b200.corr import CorrBlock  # noqa: F401  (reference: model/corr.py:8)

from pvraft_b200.gconv import SetConv  # noqa: F401  (reference: model/flot/gconv.py:4)

from pvraft_b200.update import ConvGRU, ConvRNN, FlowHead, MotionEncoder, UpdateBlock  # noqa: F401  (model/update.py)

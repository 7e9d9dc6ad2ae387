from pvraft_b200.raft import RSF  # noqa: F401  (reference: model/RAFTSceneFlow.py:10)

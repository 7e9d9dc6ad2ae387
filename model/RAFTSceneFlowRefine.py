from pvraft_b200.raft import RSF_refine  # noqa: F401  (reference: model/RAFTSceneFlowRefine.py:10)

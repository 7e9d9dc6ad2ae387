from pvraft_b200.graph import Graph  # noqa: F401  (reference: model/flot/graph.py:4)

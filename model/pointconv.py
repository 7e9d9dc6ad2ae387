from pvraft_b200.pointconv import knn_point, square_distance  # noqa: F401  (reference: model/pointconv.py)

"""Drop-in for the reference's tools/metric.py (same names and signatures), see pvraft_b200/loss.py."""
from pvraft_b200.loss import compute_epe, compute_epe_train  # noqa: F401

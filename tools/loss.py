"""Drop-in for the reference's tools/loss.py (same names and signatures) on the device-side kernels: with this repository
ahead of the reference on PYTHONPATH, `from tools.loss import sequence_loss` (tools/engine.py:19) resolves here."""
from pvraft_b200.loss import compute_loss, sequence_loss  # noqa: F401

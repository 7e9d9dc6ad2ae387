"""Drop-in `model` package: the reference's import paths (`from model.RAFTSceneFlow import RSF`,
tools/engine.py:17, test.py:14-15) resolved to the B200-native implementation in `pvraft_b200`.
Put this repository ahead of the reference on sys.path and train.py / test.py run unchanged."""

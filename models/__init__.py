"""Drop-in alias: the reference's callers do ``from models.mvsnet import CascadeMVSNet``
(train.py:9, eval.py:11, test.ipynb:34).  See INTEGRATION.md."""

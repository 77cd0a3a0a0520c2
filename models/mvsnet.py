from casmvsnet_pl_b200.models.mvsnet import *  # noqa: F401,F403
from casmvsnet_pl_b200.models.mvsnet import CascadeMVSNet, CostRegNet, FeatureNet  # noqa: F401

from casmvsnet_pl_b200.models.modules import *  # noqa: F401,F403

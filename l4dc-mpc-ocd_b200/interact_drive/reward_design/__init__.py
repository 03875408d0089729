"""Cost design and reward inference: mirror of interact_drive/reward_design/ of the reference
(mpc_ord.py; first_order_ioc.py on the engine's feature Jacobian; second_order_ioc.py -- LocalCIOC -- on the feature
Hessians)."""
from .first_order_ioc import InverseLocallyOptimalControl, LinearInverseLocallyOptimalControl  # noqa: F401
from .second_order_ioc import LocalCIOC  # noqa: F401

"""Cost design and reward inference: mirror of interact_drive/reward_design/ of the reference
(mpc_ord.py; first_order_ioc.py on the engine's feature Jacobian).  LocalCIOC (second_order_ioc.py) needs
the Hessian of the reward and is out of scope."""
from .first_order_ioc import InverseLocallyOptimalControl, LinearInverseLocallyOptimalControl  # noqa: F401

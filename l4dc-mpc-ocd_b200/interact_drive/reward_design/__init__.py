"""Cost design: mirror of interact_drive/reward_design/mpc_ord.py of the reference."""

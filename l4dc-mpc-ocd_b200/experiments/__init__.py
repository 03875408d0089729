"""Mirror of the reference's ``experiments`` package for the MPC hot path: the reward features car,
the three scenario constructors and the run_mpc_ord driver."""

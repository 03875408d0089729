"""Drop-in mirror of the reference's ``interact_drive`` package for the MPC hot path: same class
and function names, argument meaning and error behaviour; the arithmetic runs in the CUDA engine."""

"""B200-native batched MPC engine for the planner hot path of avikj/L4DC-MPC-OCD.

``engine``          torch-tensor front-end of the C ABI (libocd_b200.so; CUDA only, no CPU fallback)
``interact_drive``  drop-in mirror of the reference's planner / car / world / MPC_ORD interface
``experiments``     the reference's scenario constructors and the run_mpc_ord driver
"""
import importlib as _importlib

__all__ = ["Engine", "HostContext", "PlannerParams", "Scenario", "MATH_FAST", "MATH_PRECISE", "OPT_SGD", "OPT_LBFGS",
           "device_count", "kernel_form"]           # + OcdError, OcdCudaError (from _native, resolved below)


def __getattr__(name):
    """The engine (and with it ``libocd_b200.so``) is loaded on first use, not at package import: helpers that are
    plain numpy -- ``synthetic``, ``cmaes`` -- stay importable in a process that must not map the CUDA library
    (bench.py's ``--impl reference`` arm).  Any engine name still raises ImportError when the library has not been
    built: there is no CPU fallback."""
    if name in __all__:
        return getattr(_importlib.import_module(__name__ + ".engine"), name)
    if name in ("_native", "engine"):
        return _importlib.import_module(__name__ + "." + name)
    if name in ("OcdError", "OcdCudaError"):
        return getattr(_importlib.import_module(__name__ + "._native"), name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


def install_as_reference() -> None:
    """Make `import interact_drive...` / `import experiments...` resolve to this package's drop-in
    mirror, so that scripts written against the reference run on the engine unchanged:

        import l4dc_mpc_ocd_b200; l4dc_mpc_ocd_b200.install_as_reference()
        from interact_drive.reward_design.mpc_ord import MPC_ORD, finite_horizon_env
    """
    import importlib
    import sys
    pkg = __name__
    for name in ("interact_drive", "interact_drive.simulation_utils", "interact_drive.math_utils",
                 "interact_drive.world", "interact_drive.car", "interact_drive.car.car",
                 "interact_drive.car.fixed_control_car", "interact_drive.car.fixed_velocity_car",
                 "interact_drive.car.fixed_plan_car", "interact_drive.car.planner_car",
                 "interact_drive.car.linear_reward_car", "interact_drive.planner",
                 "interact_drive.planner.car_planner", "interact_drive.planner.naive_planner",
                 "interact_drive.reward_design", "interact_drive.reward_design.mpc_ord",
                 "interact_drive.reward_design.first_order_ioc", "interact_drive.reward_design.second_order_ioc",
                 "experiments",
                 "experiments.merging", "experiments.local_opt_scenario", "experiments.replanning_world",
                 "experiments.run_mpc_ord"):
        sys.modules[name] = importlib.import_module(pkg + "." + name)


__all__.append("install_as_reference")

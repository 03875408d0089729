"""Import-time stand-in for pycma (absent here; un-pinned in the reference, setup.py:6).
TEST INFRASTRUCTURE ONLY: lets interact_drive/reward_design/mpc_ord.py import; any use raises."""


class _Missing:
    def __getattr__(self, name):
        raise NotImplementedError("cma is not installed; the golden generator never runs CMA-ES")


evolution_strategy = _Missing()

"""Mirror of interact_drive/planner/car_planner.py:7-17 of the reference."""


class CarPlanner(object):
    """Base class of the trajectory planners of one car."""

    def __init__(self, world, car):
        self.world = world
        self.car = car

    def generate_plan(self):
        raise NotImplementedError

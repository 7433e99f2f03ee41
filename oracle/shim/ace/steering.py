"""Stand-in for `ace.steering.SteeringGeometry` (third-party ac-extras, absent here).

Documented synthetic vehicle constants: the asset
data/vehicles/audi_r8_lms_2016 is not in the reference tree."""
from types import SimpleNamespace

WHEELBASE = 2.65
WIDTH = 1.99
DELTA_MAX = 0.30


class SteeringGeometry:
    def __init__(self, data_path=None):
        self.vehicle_data = SimpleNamespace(wheelbase=WHEELBASE, width=WIDTH)

    def max_steering_angle(self):
        return DELTA_MAX

from .controller import build_mpc
from .dynamics import SpatialBicycleModel
from .paths import ReferencePath
from .spatial_mpc import SpatialMPC

__all__ = ["build_mpc", "SpatialBicycleModel", "ReferencePath", "SpatialMPC"]

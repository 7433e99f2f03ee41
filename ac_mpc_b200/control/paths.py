"""Reference-path container with the reference's row layout
(/root/reference/src/acmpc/control/paths.py:4-72): a (7, n) float64 structure-of-arrays whose rows
are xs, ys, psis, kappas, distances, widths, velocities.  On the B200 path this object is only a
host-side *view* of the `waypoints` output of the kernel (include/acmpc_b200.h)."""
from __future__ import annotations

import numpy as np

ROWS = ("xs", "ys", "psis", "kappas", "distances", "widths", "velocities")


def _row_property(index: int):
    def getter(self):
        return self._reference_path[index, :]

    def setter(self, values):
        self._reference_path[index, :] = values

    return property(getter, setter)


class ReferencePath:
    def __init__(self, n_positions: int, data: np.ndarray | None = None):
        self._n_positions = int(n_positions)
        if data is None:
            self._reference_path = np.zeros((len(ROWS), self._n_positions))
        else:
            data = np.asarray(data, dtype=np.float64)
            if data.shape != (len(ROWS), self._n_positions):
                raise ValueError(f"expected {(len(ROWS), self._n_positions)}, got {data.shape}")
            self._reference_path = data

    def __len__(self) -> int:
        return self._n_positions

    def get_state(self, index: int) -> np.ndarray:
        """[x, y, psi] of waypoint `index`."""
        return self._reference_path[:3, index]

    def as_array(self) -> np.ndarray:
        return self._reference_path


for _i, _name in enumerate(ROWS):
    setattr(ReferencePath, _name, _row_property(_i))

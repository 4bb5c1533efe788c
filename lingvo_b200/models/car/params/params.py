"""Car model registrations (ref `lingvo/tasks/car/params/params.py`)."""

from lingvo_b200.models.car.params import kitti  # noqa: F401
from lingvo_b200.models.car.params import waymo  # noqa: F401
from lingvo_b200.models.car.params import waymo_deepfusion  # noqa: F401

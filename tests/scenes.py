"""Scene builders used by the tests: the product's workload module (yulio_raytracer_b200/workloads.py) re-exported."""
from yulio_raytracer_b200.workloads import *  # noqa: F401,F403
from yulio_raytracer_b200.workloads import _CORNELL_GROUPS, _CORNELL_KD, _bundle, _regression_image, _regression_material, _regression_shape  # noqa: F401

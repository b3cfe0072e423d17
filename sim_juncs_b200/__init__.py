"""sim_juncs_b200 -- B200-native FDTD engine for the hot path of sim_juncs."""
from .engine import Sim, SjError  # noqa: F401

from .differential_env import DifferentialDriveEnv, normalize_angle  # noqa: F401
from .differential_gym import DifferentialDriveGym  # noqa: F401
from .robot_env import RobotEnv  # noqa: F401

__all__ = ["DifferentialDriveEnv", "DifferentialDriveGym", "RobotEnv"]

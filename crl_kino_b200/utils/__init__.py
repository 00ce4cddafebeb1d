from .rigid import RectObs, RectRobot, Obstacle, Rigid, load_obstacles  # noqa: F401

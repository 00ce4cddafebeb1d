from .rrt import RRT  # noqa: F401

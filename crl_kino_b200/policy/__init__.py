from .dwa import DWA  # noqa: F401

__all__ = ["DWA"]

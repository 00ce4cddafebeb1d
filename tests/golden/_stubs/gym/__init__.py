"""Minimal gym stub (golden-vector generation only): the reference subclasses
gym.Env / gym.Wrapper at import time and builds spaces.Box action bounds."""


class Env:
    pass


class Wrapper:
    pass


from . import spaces  # noqa: E402,F401

"""Abstract env interface (reference: crl_kino/env/robot_env.py:11-22)."""
from abc import abstractmethod


class RobotEnv(object):
    @abstractmethod
    def motion(self):
        pass

    @abstractmethod
    def valid_state_check(self):
        pass

    @abstractmethod
    def get_bounds(self):
        pass

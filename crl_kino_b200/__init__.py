"""crl_kino_b200: B200 (sm_100a) implementation of crl_kino's kinodynamic-RRT
extend path behind the reference's own plugin API.

Sub-packages mirror the reference layout for the hot path only:
env/ (DifferentialDriveEnv, DifferentialDriveGym), policy/ (DWA, policy_forward),
estimator/, planner/ (RRT, RRT_RL, RRT_RL_Estimator, BatchedRRT), utils/rigid.
All numerics run in hand-written CUDA kernels reached through the C ABI of
include/crlk.h; there is no CPU fallback.
"""
__version__ = "0.1.0"

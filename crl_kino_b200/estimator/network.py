"""EstimatorModel / ClassifierModel callables (reference:
crl_kino/estimator/network.py:13-127).  Both CNNs are evaluated together by
crlk_estimator_forward; each wrapper returns its own output column."""
import numpy as np
import torch

from .. import ops


class _Pair:
    """Shared device handle of an (estimator, classifier) state-dict pair."""

    def __init__(self):
        self.sd = {}
        self._handle = None

    @property
    def handle(self):
        if self._handle is None:
            assert "est" in self.sd and "cls" in self.sd, "load both state dicts first"
            self._handle = ops.EstimatorHandle(self.sd["est"], self.sd["cls"])
        return self._handle


class _CnnBase:
    kind = None

    def __init__(self, pair=None):
        self.pair = pair if pair is not None else _Pair()

    def load_state_dict(self, sd):
        self.pair.sd[self.kind] = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
                                   for k, v in sd.items()}
        if "est" not in self.pair.sd:
            self.pair.sd["est"] = self.pair.sd[self.kind]
        if "cls" not in self.pair.sd:
            self.pair.sd["cls"] = self.pair.sd[self.kind]
        self.pair._handle = None

    def forward_obs(self, obs32):
        e, p = ops.estimator_forward(self.pair.handle, obs32)
        return e if self.kind == "est" else p

    def __call__(self, spatial, ext):
        """(spatial[B,1,27,27], ext[B,3]) -> [B,1] like the reference modules.
        ext[:, 0] must be the goal distance (rrt_rl_estimator.py:59-62)."""
        spatial = ops.as_dev(spatial, torch.float32).reshape(-1, 729)
        ext = ops.as_dev(ext, torch.float32).reshape(-1, 3)
        B = spatial.shape[0]
        obs = torch.zeros(B, ops.OBS_LD, dtype=torch.float32, device=spatial.device)
        obs[:, 0] = ext[:, 0]          # kernel recomputes sqrt(gx^2 + gy^2): (d, 0) -> d
        obs[:, 2:4] = ext[:, 1:3]
        obs[:, 4:733] = spatial
        return self.forward_obs(obs).reshape(B, 1)


class EstimatorModel(_CnnBase):
    kind = "est"


class ClassifierModel(_CnnBase):
    kind = "cls"

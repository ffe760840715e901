"""Drop-in for the hot path of the reference's ``mfg_synthetic.actor_critic`` (mfg_synthetic.py:24-925):
the forward actor-critic with the synthetic reward  r = -1/2 sum_i pi_i ||P_i||^2  (:249-265).

Everything else is inherited from the mfg_ac2 drop-in (same policy, critic, updates).  ``sweep`` is the
reference's ``__main__`` experiment (:902-925) -- one independent learner per (shift, theta_initial) pair,
1000 episodes each with constant step sizes -- run as ONE launch of independent serial learners instead of
200 sequential runs.  The analytic check of the learned policy against the MFG backward equation
(:726-899) is evaluation code and out of scope.
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine
from ._lib import num_features
from .mfg_ac2 import actor_critic as _actor_critic


class actor_critic(_actor_critic):
    reward_kind = "synthetic"

    def __init__(self, theta=10, shift=0, alpha_scale=100, d=21, mat_pi0=None, path_to_dir=None, device=None,
                 dtype="float32", seed=None):
        """Reference signature actor_critic(theta=10, shift=0, alpha_scale=100, d=21) (mfg_synthetic.py:26)."""
        super().__init__(theta=theta, shift=shift, alpha_scale=alpha_scale, d=d, mat_pi0=mat_pi0,
                         path_to_dir=path_to_dir, device=device, dtype=dtype, seed=seed)

    def sweep(self, shifts, thetas, num_episodes=1000, gamma=1, constant=1, lr_critic=0.1, lr_actor=0.001, T=15):
        """All (shift, theta_initial) pairs of mfg_synthetic.py:907-914 as independent learners in one launch.
        Returns an array [len(shifts) * len(thetas), 3] of (shift, theta_initial, theta_final)."""
        grid = np.array([(s, t) for s in shifts for t in thetas], dtype=np.float64)
        L = grid.shape[0]
        d = self.d
        if d not in (4, 15, 16):
            raise NotImplementedError("the learners kernel is built for d in {4, 15, 16}")
        theta = torch.as_tensor(grid[:, 1].copy(), device=self.device)
        shift = torch.as_tensor(grid[:, 0].copy(), device=self.device)
        w = torch.as_tensor(np.random.rand(L, num_features(d)), device=self.device)
        engine.learners(theta, w, self._dev(self.mat_pi0), int(num_episodes), T, shift=shift,
                        alpha_scale=self.alpha_scale, episode0=self.first_episode, gamma=gamma, lr_critic=lr_critic,
                        lr_actor=lr_actor, constant=bool(constant), reward=self.reward_kind,
                        discount=self.discount_kind, seed=self.seed, want_total_reward=False)
        return np.column_stack([grid, theta.cpu().numpy()])

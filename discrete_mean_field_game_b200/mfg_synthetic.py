"""Drop-in for the hot path of the reference's ``mfg_synthetic.actor_critic`` (mfg_synthetic.py:24-925):
the forward actor-critic with the synthetic reward  r = -1/2 sum_i pi_i ||P_i||^2  (:249-265).

Everything else is inherited from the mfg_ac2 drop-in (same policy, critic, updates).  ``sweep`` is the
reference's ``__main__`` experiment (:902-925) -- one independent learner per (shift, theta_initial) pair,
1000 episodes each with constant step sizes -- run as ONE launch of independent serial learners instead of
200 sequential runs.  The analytic check of the learned policy against the MFG backward equation
(``calc_reward_vector``, ``evaluate_synthetic``, ``evaluate_synthetic_JSD``, :726-899) rolls
all start rows out in ONE launch and evaluates the backward equation on the recorded actions on the device
(``dmfg_synthetic_check``).
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine
from ._lib import num_features
from .mfg_ac2 import HELPER_POP_OFFSET, actor_critic as _actor_critic


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
        if d not in (4, 15, 16) and not (d == 21 and self.dtype == torch.float32):
            raise NotImplementedError("the learners kernel is built for d in {4, 15, 16} (and 21 in float32)")
        theta = torch.as_tensor(grid[:, 1].copy(), device=self.device)
        shift = torch.as_tensor(grid[:, 0].copy(), device=self.device)
        w = torch.as_tensor(np.random.rand(L, num_features(d)), device=self.device)
        engine.learners(theta, w, self._dev(self.mat_pi0), int(num_episodes), T, shift=shift,
                        alpha_scale=self.alpha_scale, episode0=self.first_episode, gamma=gamma, lr_critic=lr_critic,
                        lr_actor=lr_actor, constant=bool(constant), reward=self.reward_kind,
                        discount=self.discount_kind, seed=self.seed, want_total_reward=False)
        return np.column_stack([grid, theta.cpu().numpy()])

    # ------------------------------------------------ consumer of a9: check against the MFG backward equation
    def generate_trajectory(self, pi0, total_hours, y=None):
        """(mat_trajectory [total_hours,d], array_actions [total_hours-1,d,d]) -- this variant also returns the
        actions (mfg_synthetic.py:549-578)."""
        pi0 = np.asarray(pi0, dtype=np.float64).reshape(1, self.d)
        T = int(total_hours) - 1
        noise = None if y is None else self._dev(np.asarray(y).reshape(T, 1, self.d, self.d))
        out = engine.rollout(self._dev(pi0), self.theta, self.shift, self.alpha_scale, T, reward="none",
                             noise_y=noise, seed=self.seed, pop_offset=HELPER_POP_OFFSET, step_offset=self._draws, outputs=("states", "actions"))
        if y is None:
            self._draws += T
        return out["states"][:, 0].double().cpu().numpy(), out["actions"][:, 0].double().cpu().numpy()

    def calc_reward_vector(self, P):
        """v_i = -1/2 ||P_i||^2 (mfg_synthetic.py:726-738)."""
        P = np.asarray(P, dtype=np.float64)
        return -0.5 * np.sum(P * P, axis=1)

    def _synthetic_check(self, day_first, day_last, actions, want_jsd):
        """l1 / jsd [days, 15] of the rows day_first..day_last (1-based, inclusive).  ``actions`` [days,15,d,d]
        replaces the rollout (parity against the reference's own sampled actions)."""
        if actions is None:
            pi0 = self._dev(self.mat_pi0[day_first - 1:day_last])
            out = engine.rollout(pi0, self.theta, self.shift, self.alpha_scale, 15, reward="none", seed=self.seed,
                                 pop_offset=HELPER_POP_OFFSET, step_offset=self._draws, outputs=("actions",))
            self._draws += 15
            acts = out["actions"]
        else:
            a = np.asarray(actions, dtype=np.float64)
            acts = self._dev(np.ascontiguousarray(a.transpose(1, 0, 2, 3)), torch.float64)
        return engine.synthetic_check(acts, want_jsd=want_jsd)

    def evaluate_synthetic(self, day_first=1, day_last=26, verbose=0, actions=None):
        """Mean and standard deviation over all (day, hour) of sum_ij |P_ij - A_ij|, A built from the value
        function of the backward equation (mfg_synthetic.py:741-812)."""
        l1, _ = self._synthetic_check(day_first, day_last, actions, False)
        diff_mean, diff_std = float(l1.mean()), float(l1.std(unbiased=False))
        if verbose:
            print("Mean over all hours", diff_mean)
            print("Standard deviation", diff_std)
        return diff_mean, diff_std

    def evaluate_synthetic_JSD(self, day_first=1, day_last=26, write_file=0, filename='synthetic_log.csv', verbose=0,
                               actions=None):
        """Same with sum_i JSD(P_i, A_i) (mfg_synthetic.py:815-899).  write_file is not supported: the log it
        writes is a debugging dump of every row pair."""
        if write_file:
            raise NotImplementedError("write_file: the row-by-row dump of mfg_synthetic.py:879-882 is not provided")
        _, js = self._synthetic_check(day_first, day_last, actions, True)
        diff_mean, diff_std = float(js.mean()), float(js.std(unbiased=False))
        if verbose:
            print("Mean over all hours", diff_mean)
            print("Standard deviation", diff_std)
        return diff_mean, diff_std

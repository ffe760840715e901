"""Drop-in for the reference's ``mfg_ac2.actor_critic`` (mfg_ac2.py:23-836) on B200.

Same constructor, method names, argument order and attributes as the
reference's forward actor-critic solver; every numeric method runs on the GPU
through libdmfg (include/dmfg.h).  Host arrays go in and come out as NumPy,
like the reference -- that is the end-to-end path ``bench.py`` times.

What is different, on purpose:
  * noise is explicit and reproducible: a Philox key (``seed``) instead of the
    process-global NumPy state; ``sample_action(pi, y=...)`` accepts the Gamma
    variates themselves for parity runs;
  * start states can be given in memory (``mat_pi0=``) instead of being read
    from ``./train_normalized_round2``;
  * inputs are never mutated (the reference overwrites P==0, mfg_ac2.py:369);
  * ``train_batch`` / ``rollout_batch`` expose the batched path the reference
    does not have: B independent populations per launch, ``per_episode`` or
    ``per_step`` synchronous updates, optional multi-GPU gradient all-reduce.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch

from . import engine
from ._lib import num_features

_FAST_D = (4, 15, 16)
# sample_action / generate_trajectory / evaluate draw from their OWN Philox populations: train() uses learner ids
# 0.. and train_batch population ids pop_offset + b at the same (seed, step) positions, and the reference draws
# fresh noise for every call (AC_IRL.generate_batch sits at 1 << 32)
HELPER_POP_OFFSET = 1 << 33


def _torch_dtype(dtype):
    if dtype in ("float32", np.float32, torch.float32):
        return torch.float32
    if dtype in ("float64", np.float64, torch.float64):
        return torch.float64
    raise TypeError("dtype must be float32 or float64")


class actor_critic:
    """Forward actor-critic with the pre-specified reward (mfg_ac2.py:23)."""

    reward_kind = "ac2"          # mfg_ac2.py:257-287; the synthetic variant overrides this
    discount_kind = "step"       # mfg_ac2.py:505
    first_episode = 0            # mfg_ac2.py:460

    def __init__(self, theta=8.86349, shift=0.16, alpha_scale=12000, d=21, mat_pi0=None,
                 path_to_dir=None, device=None, dtype="float32", seed=None):
        engine.require_cuda()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.dtype = _torch_dtype(dtype)
        self.theta = theta
        self.shift = shift
        self.alpha_scale = alpha_scale
        self.d = d
        self.w = self.init_w(d)
        if mat_pi0 is not None:
            self.mat_pi0 = np.array(mat_pi0, dtype=np.float64)[:, :d].copy()
        else:
            self.init_pi0(path_to_dir or os.path.join(os.getcwd(), "train_normalized_round2"))
        self.num_start_samples = self.mat_pi0.shape[0]
        self.mat_alpha = np.zeros([d, d])
        self.mat_alpha_deriv = np.zeros([d, d])
        self.seed = int(np.random.randint(2 ** 31 - 1)) if seed is None else int(seed)
        self._draws = 0          # Philox step counter of the single-population methods
        self._episodes = 0       # episodes consumed by train() so far (Philox stream position)
        self._batch_episodes = 0  # episodes consumed by train_batch() so far: repeated calls never replay noise

    # ------------------------------------------------------------------ set-up
    def init_w(self, d):
        """U[0,1) critic weights, column vector [F,1] (mfg_ac2.py:165-176)."""
        return np.random.rand(num_features(d), 1)

    def init_pi0(self, path_to_dir, verbose=0):
        """Start-state table from trend_distribution_day<k>.csv files: first line of each file,
        space separated, first d columns, NOT renormalised (mfg_ac2.py:179-208)."""
        rows = []
        for k in range(1, 1 + len(os.listdir(path_to_dir))):
            name = "trend_distribution_day%d.csv" % k
            with open(os.path.join(path_to_dir, name)) as f:
                rows.append([float(v) for v in f.readline().split()][: self.d])
            if verbose:
                print(name)
        self.mat_pi0 = np.array(rows, dtype=np.float64)

    # ------------------------------------------------------------ device helpers
    def _dev(self, x, dtype=None):
        return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype or self.dtype, device=self.device)

    def _w_dev(self, w=None):
        w = self.w if w is None else w
        return self._dev(np.asarray(w, dtype=np.float64).reshape(-1), torch.float64)

    # -------------------------------------------------------------------- a1
    def sample_action(self, pi, y=None):
        """P ~ prod_i Dirichlet(alpha_i * alpha_scale)  (mfg_ac2.py:211-254).

        Sets ``mat_alpha`` / ``mat_alpha_deriv`` like the reference.  ``y`` [d,d] injects the Gamma
        variates (parity); otherwise each call consumes the next step of this object's Philox stream.
        """
        pi = np.asarray(pi, dtype=np.float64).reshape(1, self.d)
        noise = None if y is None else self._dev(np.asarray(y).reshape(1, 1, self.d, self.d))
        out = engine.rollout(self._dev(pi), self.theta, self.shift, self.alpha_scale, 1, reward="none",
                             noise_y=noise, seed=self.seed, pop_offset=HELPER_POP_OFFSET, step_offset=self._draws,
                             outputs=("actions", "alpha", "alpha_deriv"))
        if y is None:
            self._draws += 1
        self.mat_alpha = out["alpha"][0, 0].double().cpu().numpy()
        self.mat_alpha_deriv = out["alpha_deriv"][0, 0].double().cpu().numpy()
        return out["actions"][0, 0].double().cpu().numpy()

    # --------------------------------------------------------------- a2/a3/a6
    def _evaluate(self, P, pi, d, outputs, reward=None):
        """Evaluate GIVEN (pi, P) pairs on the device (no sampling, no normalisation)."""
        P = np.asarray(P, dtype=np.float64).reshape(1, 1, d, d)
        pi = np.asarray(pi, dtype=np.float64).reshape(1, d)
        return engine.rollout(self._dev(pi), self.theta, self.shift, self.alpha_scale, 1,
                              reward=reward or self.reward_kind, actions_in=self._dev(P), outputs=outputs)

    def calc_reward(self, P, pi, d):
        """<pi, (P*P) pi - ((P*P) 1) * pi>  (mfg_ac2.py:257-287); returns shape (1,) like the reference."""
        out = self._evaluate(P, pi, d, ("rewards",))
        return out["rewards"].double().cpu().numpy().reshape(1)

    def step(self, P, pi):
        """pi' = P^T pi (mfg_ac2.py:497)."""
        return self._evaluate(P, pi, self.d, ("pi_final",))["pi_final"][0].double().cpu().numpy()

    def calc_gradient_vectorized(self, P, pi):
        """d/dtheta log prod_i Dir(P_i; alpha_i)  (mfg_ac2.py:347-381).

        alpha, alpha' are recomputed from (pi, theta) on the device -- identical to the reference's
        side state whenever sample_action(pi) was the previous call, which is how it is always used.
        """
        out = self._evaluate(P, pi, self.d, ("grads",))
        return float(out["grads"][0, 0])

    calc_gradient = calc_gradient_vectorized        # mfg_ac2.py:402 ("do not use") -- same value
    calc_gradient_basic = calc_gradient_vectorized  # mfg_ac2.py:384

    # -------------------------------------------------------------------- a4
    def calc_features(self, pi):
        """phi(pi) = [pi_i pi_j (i<=j), pi, 1]  (mfg_ac2.py:325-344)."""
        pi = np.asarray(pi, dtype=np.float64).reshape(1, -1)
        out = engine.critic_eval(self._dev(pi), want_features=True)
        return out["features"][0].double().cpu().numpy()

    def calc_value(self, pi):
        """V(pi; w) = phi(pi) . w  (mfg_ac2.py:290-311); shape follows self.w like the reference."""
        pi = np.asarray(pi, dtype=np.float64).reshape(1, -1)
        w = np.asarray(self.w, dtype=np.float64)
        out = engine.critic_eval(self._dev(pi), self._w_dev(w), want_features=False, want_values=True)
        v = out["values"].double().cpu().numpy()
        return v.reshape(1) if w.ndim == 2 else float(v[0])

    # -------------------------------------------------------------------- a8
    def train_log(self, vector, filename, str_format):
        """Append one CSV line (mfg_ac2.py:441-445)."""
        os.makedirs(os.path.dirname(filename) or ".", exist_ok=True)
        with open(filename, "a") as f:
            np.asarray(vector).tofile(f, sep=",", format=str_format)
            f.write("\n")

    def train(self, num_episodes=4000, gamma=1, constant=0, lr_critic=0.1, lr_actor=0.001, consecutive=100,
              file_theta="results/theta.csv", file_pi="results/pi.csv", file_reward="results/reward.csv",
              write_file=0, write_all=0, start_rows=None, noise_y=None, verbose=True):
        """One learner, per-step online updates of w then theta: mfg_ac2.py:448-539.

        Runs as ONE serial learner on the GPU (dmfg_ac_learners) in chunks of ``consecutive`` episodes,
        reporting at the same episodes as the reference.  ``start_rows`` [E] / ``noise_y`` [E,15,d,d]
        replay recorded draws (parity); otherwise the object's Philox stream is used.
        """
        T = 15                                       # mfg_ac2.py:478
        d = self.d
        theta = torch.tensor([float(self.theta)], dtype=torch.float64, device=self.device)
        w = self._w_dev().reshape(1, -1).clone()
        mat = self._dev(self.mat_pi0)
        list_reward = []
        e = 0
        while e < num_episodes:
            # the reference reports after every episode whose index is a multiple of `consecutive`
            k = -(-e // consecutive) * consecutive
            n = min(k + 1, num_episodes) - e
            kw = {}
            if noise_y is not None:
                kw["noise_y"] = self._dev(np.asarray(noise_y)[e:e + n].reshape(1, n, T, d, d))
            if start_rows is not None:
                kw["start_rows"] = self._dev(np.asarray(start_rows)[e:e + n].reshape(1, n), torch.int32)
            if write_all:      # mfg_ac2.py:461-463,488-494: every (pi, P) goes to temp.csv -- the host-driven per-step path
                res = self._run_learner_generic(theta, w, mat, n, T, self.first_episode + e, self._episodes,
                                                gamma, constant, lr_critic, lr_actor, kw, dump="temp.csv")
            else:
                res = self._run_learner(theta, w, mat, n, T, self.first_episode + e, self._episodes,
                                        gamma, constant, lr_critic, lr_actor, kw)
            list_reward += res["total_reward"][0].cpu().tolist()
            e += n
            if (e - 1) % consecutive == 0:
                self.theta = float(theta[0])
                pi = res["pi_final"][0].double().cpu().numpy()
                reward_avg = sum(list_reward) / consecutive          # quirk A.16 kept: divides by `consecutive`
                if verbose:
                    print("Theta\n", self.theta)
                    print("pi\n", pi)
                    print("Average reward during previous %d episodes: " % consecutive, str(reward_avg))
                list_reward = []
                if write_file:
                    self.train_log(np.array([self.theta]), file_theta, "%.5e")
                    self.train_log(pi, file_pi, "%.3e")
                    self.train_log(np.array([reward_avg]), file_reward, "%.3e")
        self.theta = float(theta[0])
        self.w = w[0].cpu().numpy().reshape(-1, 1)
        self._episodes += num_episodes

    def _run_learner(self, theta, w, mat, n, T, episode0, noise_ep, gamma, constant, lr_critic, lr_actor, kw):
        """episode0 drives the step-size schedule (restarts at every train() call like the reference);
        noise_ep positions the Philox stream (never restarts)."""
        if self.d in _FAST_D or (self.d == 21 and self.dtype == torch.float32):    # d = 21: the reference's default
            return engine.learners(theta, w, mat, n, T, shift=self.shift, alpha_scale=self.alpha_scale,
                                   episode0=episode0, gamma=gamma, lr_critic=lr_critic, lr_actor=lr_actor,
                                   constant=bool(constant), reward=self.reward_kind, discount=self.discount_kind,
                                   seed=self.seed, noise_episode_offset=noise_ep, **kw)
        return self._run_learner_generic(theta, w, mat, n, T, episode0, noise_ep, gamma, constant, lr_critic,
                                         lr_actor, kw)

    @staticmethod
    def write_all_step(path, num_steps, pi, P):
        """The per-step block the reference appends to temp.csv with write_all=1 (mfg_ac2.py:488-494)."""
        with open(path, 'ab') as f:
            np.savetxt(f, np.array(['num_steps = %d' % num_steps]), fmt='%s')
            np.savetxt(f, np.array(['distribution']), fmt='%s')
            np.savetxt(f, np.asarray(pi).reshape(1, -1), delimiter=',', fmt='%.6f')
            np.savetxt(f, np.array(['Action']), fmt='%s')
            np.savetxt(f, np.asarray(P), delimiter=',', fmt='%.3f')

    def _run_learner_generic(self, theta, w, mat, n, T, episode0, noise_ep, gamma, constant, lr_critic, lr_actor,
                             kw, dump=None):
        """Any d: the same per-step semantics driven from the host -- one transition launch (generic
        kernel), then the device-side update; parameters never leave the GPU."""
        d = self.d
        total = torch.zeros((1, n), dtype=torch.float64, device=self.device)
        w1 = w[0]
        pi = None
        for e in range(n):
            episode = episode0 + e
            if "start_rows" in kw:
                row = int(kw["start_rows"][0, e])
            else:
                row = engine.philox((0, 0, (episode + noise_ep) & 0xFFFFFFFF, 0xC0000000),
                                    (self.seed & 0xFFFFFFFF, self.seed >> 32))[0]
                row = (row * mat.shape[0]) >> 32
            pi = mat[row:row + 1].contiguous()
            lr_c = lr_critic if constant else lr_critic / (episode + 1.0)
            lr_a = lr_actor if constant else lr_actor / ((episode + 1.0) * math.log(math.log(episode + 20.0)))
            disc = 1.0
            if dump:
                with open(dump, 'a') as f:
                    f.write('Episode %d \n\n' % episode)
            for t in range(T):
                noise = kw["noise_y"][0, e, t].reshape(1, 1, d, d) if "noise_y" in kw else None
                g_next = gamma if self.discount_kind == "step" else disc
                out = engine.rollout(pi, 0.0, self.shift, self.alpha_scale, 1, w=w1, theta_dev=theta,
                                     gamma=g_next, reward=self.reward_kind, noise_y=noise, seed=self.seed,
                                     step_offset=(episode + noise_ep) * T + t,
                                     outputs=("pi_final", "actions") if dump else ("pi_final",), want_acc=True)
                if dump:
                    self.write_all_step(dump, t + 1, pi[0].double().cpu().numpy(), out["actions"][0, 0].double().cpu().numpy())
                engine.apply_update(d, theta, w1, out["acc"], lr_c, lr_a, 1.0)
                total[0, e] += out["acc"][-1]
                disc *= gamma
                pi = out["pi_final"]
        return dict(total_reward=total, pi_final=pi)

    # -------------------------------------------------------------------- a9
    def generate_trajectory(self, pi0, total_hours, y=None):
        """[total_hours, d] states pi^0..pi^N, rollout only (mfg_ac2.py:566-592)."""
        pi0 = np.asarray(pi0, dtype=np.float64).reshape(1, self.d)
        T = int(total_hours) - 1
        noise = None if y is None else self._dev(np.asarray(y).reshape(T, 1, self.d, self.d))
        out = engine.rollout(self._dev(pi0), self.theta, self.shift, self.alpha_scale, T, reward="none",
                             noise_y=noise, seed=self.seed, pop_offset=HELPER_POP_OFFSET, step_offset=self._draws, outputs=("states",))
        if y is None:
            self._draws += T
        return out["states"][:, 0].double().cpu().numpy()

    # ------------------------------------------------ consumer of a9: evaluation against measured days
    def JSD(self, P, Q):
        """Jensen-Shannon divergence (mfg_ac2.py:546-563); inputs are not mutated."""
        P = np.asarray(P, dtype=np.float64).reshape(1, 1, -1)
        Q = np.asarray(Q, dtype=np.float64).reshape(1, 1, -1)
        _, js = engine.traj_metrics(self._dev(P, torch.float64), self._dev(Q, torch.float64), time_major=False)
        return float(js[0, 0])

    def evaluate(self, theta=8.86349, shift=0.5, alpha_scale=1e4, d=21, episode_length=16,
                 indir='test_normalized_round2', outfile='eval_mfg_round2/test_eval_fixed_reward.csv',
                 write_header=0, empirical=None, y=None):
        """Roll the fixed policy forward from the first row of every test day and compare with the measured
        trajectory: mean L1 / Jensen-Shannon of the final distribution and over the day (mfg_ac2.py:595-670).
        All test days run as ONE batched rollout + one metrics launch.  Files are visited in SORTED order
        (the reference uses directory order).  ``empirical`` [n,episode_length,>=d] replaces reading ``indir``;
        ``y`` [n,episode_length-1,d,d] injects the Gamma variates (parity)."""
        self.theta, self.shift, self.alpha_scale, self.d = theta, shift, alpha_scale, d
        if empirical is None:
            path = os.path.join(os.getcwd(), indir)
            empirical = np.stack([np.loadtxt(os.path.join(path, f), delimiter=' ', ndmin=2) for f in sorted(os.listdir(path))])
        emp = np.ascontiguousarray(np.asarray(empirical, dtype=np.float64)[:, :episode_length, :d])
        n, T = emp.shape[0], episode_length - 1
        noise = None if y is None else self._dev(np.ascontiguousarray(np.transpose(np.asarray(y), (1, 0, 2, 3))))
        out = engine.rollout(self._dev(emp[:, 0]), theta, shift, alpha_scale, T, reward="none", noise_y=noise,
                             seed=self.seed, pop_offset=HELPER_POP_OFFSET, step_offset=self._draws, outputs=("states",))
        if y is None:
            self._draws += T
        l1, js = engine.traj_metrics(out["states"], self._dev(emp))
        l1, js = l1.cpu().numpy(), js.cpu().numpy()
        arrays = (l1[:, -1], l1.mean(1), js[:, -1], js.mean(1))
        stats = [(float(np.mean(a)), float(np.std(a))) for a in arrays]
        if outfile:
            os.makedirs(os.path.dirname(outfile) or ".", exist_ok=True)
            with open(outfile, 'a') as f:
                if write_header:
                    f.write('theta,shift,alpha_scale,mean_l1_final,std_l1_final,mean_l1_mean,std_l1_mean,'
                            'mean_JSD_final,std_JSD_final,mean_JSD_mean,std_JSD_mean\n')
                f.write("%f,%f,%f,%.3e,%.3e,%.3e,%.3e,%.3e,%.3e,%.3e,%.3e\n" % (
                    (theta, shift, alpha_scale) + tuple(v for ms in stats for v in ms)))
        return tuple(m for m, _ in stats)

    def gridsearch(self, theta_range, shift_range, alpha_range, indir, outfile, empirical=None, verbose=True):
        """Best (theta, shift, alpha_scale) per metric over the grid (mfg_ac2.py:673-689); returns the list of
        [value, theta, shift, alpha_scale] the reference prints."""
        best = [[100, 0, 0, 0] for _ in range(4)]
        if empirical is None:
            path = os.path.join(os.getcwd(), indir)
            empirical = np.stack([np.loadtxt(os.path.join(path, f), delimiter=' ', ndmin=2) for f in sorted(os.listdir(path))])
        for theta in theta_range:
            for shift in shift_range:
                for alpha_scale in alpha_range:
                    if verbose:
                        print("Theta %f, shift %f, alpha %d" % (theta, shift, alpha_scale))
                    result = self.evaluate(theta, shift, alpha_scale, d=self.d, indir=indir, outfile=outfile,
                                           write_header=0, empirical=empirical)
                    for idx in range(4):
                        if result[idx] <= best[idx][0]:
                            best[idx] = [result[idx], theta, shift, alpha_scale]
        if verbose:
            print(best)
        return best

    # ------------------------------------------------ batched path (extension)
    def rollout_batch(self, pi0, T=15, record=False, seed=None, pop_offset=0, noise_y=None, with_td=True):
        """B independent populations, frozen (theta, w): host arrays in, host arrays out.

        pi0 [B,d] NumPy (or a pinned / CUDA tensor).  Returns rewards/deltas/grads [T,B], pi_final [B,d],
        'acc' = [sum delta*g, sum delta*phi, sum r] and, with record=True, states [T+1,B,d] and
        actions [T,B,d,d] (time-major, the layout the IRL sampler consumes).
        """
        pi_dev = self._to_device(pi0)
        outs = ("rewards", "grads", "pi_final") + (("deltas",) if with_td else ()) + \
               (("states", "actions") if record else ())
        noise = None if noise_y is None else self._to_device(noise_y)
        res = engine.rollout(pi_dev, self.theta, self.shift, self.alpha_scale, T,
                             w=self._w_dev() if with_td else None, reward=self.reward_kind,
                             discount=self.discount_kind, noise_y=noise,
                             seed=self.seed if seed is None else seed, pop_offset=pop_offset,
                             outputs=outs, want_acc=with_td)
        return {k: v.cpu().numpy() for k, v in res.items()}

    def _to_device(self, x):
        if isinstance(x, torch.Tensor):
            return x.to(device=self.device, dtype=self.dtype, non_blocking=True).contiguous()
        return self._dev(x)

    def train_batch(self, pi0, num_episodes=1, T=15, gamma=1, constant=0, lr_critic=0.1, lr_actor=0.001,
                    update="per_episode", seed=None, pop_offset=0, group=None, first_episode=None, history=False,
                    fuse_step=True):
        """Batched actor-critic: B populations share (theta, w).

        update="per_episode": parameters frozen within an episode, one batch-mean update
            theta += lr_a(e)/B * sum_{b,t} delta*g,  w += lr_c(e)/B * sum_{b,t} delta*phi  per episode
            (one all-reduce per episode when ``group`` is a torch.distributed process group);
        update="per_step": a synchronous batch-mean update after EVERY transition -- reduces to the
            reference's train() exactly at B = 1.
        pi0 [B,d]: host array / pinned tensor (copied in every episode) or a CUDA tensor; a list / tuple of
        ``num_episodes`` such arrays, or a callable ``episode_index -> array``, gives every episode its own start
        states.  Host inputs are double-buffered: the copy of episode e+1 runs on a side stream under the kernel
        of episode e, and with ``history=True`` (theta, w, mean reward) of EVERY episode are read back into pinned
        host memory asynchronously (returned as ``theta_history`` / ``w_history``).
        Returns dict(theta, mean_reward [num_episodes]).
        """
        from . import parallel
        d = self.d
        seed = self.seed if seed is None else seed
        theta = torch.tensor([float(self.theta)], dtype=torch.float64, device=self.device)
        w = self._w_dev().clone()
        _, world = parallel.world_info(group)
        first = self.first_episode if first_episode is None else first_episode      # step-size schedule only
        # Philox position: a persistent per-object episode counter (like train()'s _episodes), NOT first_episode --
        # a loop of train_batch calls never replays noise, and resuming the schedule does not move the noise
        noise0 = self.first_episode + self._batch_episodes
        mean_rewards = []
        F = w.numel()
        B_local = -1
        total_pops = None                  # populations over ALL ranks (shards may differ by one): set at episode 0

        def source(e):
            if callable(pi0):
                return pi0(e)
            if isinstance(pi0, (list, tuple)):
                return pi0[e]
            return pi0

        # ---- input pipeline: pinned host -> device on a side stream, two device buffers ------------------
        on_device = isinstance(source(0), torch.Tensor) and source(0).is_cuda if num_episodes > 0 else True
        compute = torch.cuda.current_stream(self.device)
        bufs, ev_ready, ev_free = [None, None], [None, None], [None, None]
        copy_stream = None
        if not on_device and num_episodes > 0:
            copy_stream = getattr(self, "_copy_stream", None)
            if copy_stream is None:
                copy_stream = self._copy_stream = torch.cuda.Stream(self.device)

        def stage(e):
            """issue the host->device copy of episode e into buffer e % 2 (side stream)"""
            src = source(e)
            if not isinstance(src, torch.Tensor):
                src = torch.as_tensor(np.ascontiguousarray(src, dtype=np.float32 if self.dtype == torch.float32 else np.float64))
            k = e & 1
            with torch.cuda.stream(copy_stream):
                if ev_free[k] is not None:
                    copy_stream.wait_event(ev_free[k])            # the kernel that read this buffer is done
                if bufs[k] is None or bufs[k].shape != src.shape:
                    bufs[k] = torch.empty(src.shape, dtype=self.dtype, device=self.device)
                bufs[k].copy_(src, non_blocking=True)
                ev_ready[k] = torch.cuda.Event()
                ev_ready[k].record(copy_stream)

        hist = None
        if history and num_episodes > 0:
            hist = torch.empty((num_episodes, F + 2), dtype=torch.float64).pin_memory()
        if copy_stream is not None:
            copy_stream.wait_stream(compute)
            stage(0)
        for e in range(num_episodes):
            episode = first + e
            if copy_stream is not None:
                compute.wait_event(ev_ready[e & 1])
                pi = bufs[e & 1]
                if e + 1 < num_episodes:
                    stage(e + 1)
            else:
                pi = self._to_device(source(e))
            B = pi.shape[0]
            if total_pops is None or B != B_local:
                B_local, total_pops = B, parallel.total_count(B, group, self.device)
            lr_c = lr_critic if constant else lr_critic / (episode + 1.0)
            lr_a = lr_actor if constant else lr_actor / ((episode + 1.0) * math.log(math.log(episode + 20.0)))
            if update == "per_episode":
                out = engine.rollout(pi, 0.0, self.shift, self.alpha_scale, T, w=w, theta_dev=theta, gamma=gamma,
                                     reward=self.reward_kind, discount=self.discount_kind, seed=seed,
                                     pop_offset=pop_offset, step_offset=(noise0 + e) * T, outputs=(), want_acc=True)
                if copy_stream is not None:
                    ev_free[e & 1] = torch.cuda.Event()
                    ev_free[e & 1].record(compute)
                acc = parallel.allreduce_sum_(out["acc"], group)
                engine.apply_update(d, theta, w, acc, lr_c, lr_a, 1.0 / total_pops)
                mean_rewards.append(acc[-1] / total_pops)
            elif update == "per_step":
                disc, tot = 1.0, torch.zeros((), dtype=torch.float64, device=self.device)
                # one rank, float streams, d in {15, 16, 21}: ONE launch per transition (dmfg_ac_step: sampling, TD sums
                # and the batch-mean update fused); otherwise rollout(T=1) -> (all-reduce) -> apply_update
                fused = fuse_step and world == 1 and self.dtype == torch.float32 and d in engine.AC_STEP_D
                for t in range(T):
                    g_next = gamma if self.discount_kind == "step" else disc
                    if fused:
                        out = engine.ac_step(pi, theta, w, self.shift, self.alpha_scale, lr_c, lr_a, 1.0 / total_pops,
                                             gamma=g_next, reward=self.reward_kind, seed=seed, pop_offset=pop_offset,
                                             step_offset=(noise0 + e) * T + t)
                        if t == 0 and copy_stream is not None:
                            ev_free[e & 1] = torch.cuda.Event()
                            ev_free[e & 1].record(compute)
                        tot = tot + out["acc"][-1] / total_pops
                        disc *= gamma
                        pi = out["pi_final"]
                        continue
                    out = engine.rollout(pi, 0.0, self.shift, self.alpha_scale, 1, w=w, theta_dev=theta,
                                         gamma=g_next, reward=self.reward_kind, seed=seed, pop_offset=pop_offset,
                                         step_offset=(noise0 + e) * T + t, outputs=("pi_final",), want_acc=True)
                    if t == 0 and copy_stream is not None:
                        ev_free[e & 1] = torch.cuda.Event()
                        ev_free[e & 1].record(compute)
                    acc = parallel.allreduce_sum_(out["acc"], group)
                    engine.apply_update(d, theta, w, acc, lr_c, lr_a, 1.0 / total_pops)
                    tot = tot + acc[-1] / total_pops
                    disc *= gamma
                    pi = out["pi_final"]
                mean_rewards.append(tot)
            else:
                raise ValueError("update must be 'per_episode' or 'per_step'")
            if hist is not None:                               # the step's result, device -> pinned host, async
                row = torch.cat([theta, w, mean_rewards[-1].reshape(1)])
                hist[e].copy_(row, non_blocking=True)
        extra = {}
        if hist is not None:
            torch.cuda.current_stream(self.device).synchronize()
            extra = dict(theta_history=hist[:, 0].numpy().copy(), w_history=hist[:, 1:1 + F].numpy().copy())
        self._batch_episodes += num_episodes
        self.theta = float(theta[0])                       # the device->host read of the step's result
        self.w = w.cpu().numpy().reshape(-1, 1)
        return dict(theta=self.theta, mean_reward=torch.stack(mean_rewards).cpu().numpy()
                    if mean_rewards else np.zeros(0), **extra)

"""Drop-in for the reference's ``ac_irl.AC_IRL`` (ac_irl.py:31-954) on B200: maximum-entropy IRL
(guided-cost-learning style) with the forward actor-critic in the loop.

Same constructor, method names, argument order and attributes as the reference.  The numeric work
runs on the GPU through libdmfg (include/dmfg.h): ``dmfg_rollout`` (sample_action, P^T pi,
gradient), ``dmfg_rnet_forward/backward`` (networks.r_net*), ``dmfg_irl_loss_grad`` (the loss of
create_training_method), ``dmfg_adam_tf`` (AdamOptimizer.minimize), ``dmfg_td_accumulate`` +
``dmfg_ac_apply_update`` (TD error and the w / theta updates), ``dmfg_dirichlet_logq`` (calc_z).

Differences, on purpose:
  * data can be given in memory (``mat_pi0=``, ``mat_pi0_test=``, ``demonstrations=``, ...) instead of
    being read from ``./train_normalized_round2`` etc.; the file readers are kept (and work on Linux:
    upstream needs pandas behind a Windows-only import, ac_irl.py:16-22,186);
  * noise is an explicit Philox key (``seed``), not the global NumPy / TF state;
  * ``use_z=True`` turns on the importance weights z_j that upstream implements but leaves commented
    out (ac_irl.py:404-405), evaluated in log space (no ``c`` normaliser needed);
  * ``update_reward_batch`` / ``generate_batch`` are the batched device-resident versions of
    update_reward / generate_trajectories (thousands of trajectories per update).
There is no TensorFlow: ``self.sess.run(fetches, feed_dict)`` is a small shim that evaluates the
reward net for the fed arrays (what test_acirl.py:60-70 and train() at :683 do).
"""
from __future__ import annotations

import math
import os
import random

import numpy as np
import torch

from . import engine, networks
from .mfg_ac2 import actor_critic as _actor_critic
from .networks import Placeholder

T_STEPS = 15          # ac_irl.py:664,749


class Session:
    """Just enough of tf.Session for the reference's call sites: run(fetch | [fetches], feed_dict)."""

    def __init__(self, owner):
        self.owner = owner

    def run(self, fetches, feed_dict=None):
        single = not isinstance(fetches, (list, tuple))
        outs = [self.owner._fetch(f, feed_dict or {}) for f in ([fetches] if single else fetches)]
        return outs[0] if single else outs


class AC_IRL(_actor_critic):
    reward_kind = "none"              # the reward is the network (ac_irl.py:683)
    discount_kind = "cumulative"      # ac_irl.py:691 multiplies V(s') by the running discount
    first_episode = 1                 # ac_irl.py:649

    def __init__(self, theta=8.64, shift=0, alpha_scale=1e4, d=15, lr_reward=1e-4, num_policies=10, c=2e11,
                 reg='dropout_l1l2', n_fc3=8, n_fc4=4, saved_network=None, use_tf=True, summarize=False,
                 mat_pi0=None, mat_pi0_test=None, demonstrations=None, demonstrations_test=None,
                 device=None, seed=None, net_seed=None, use_z=False):
        """reg - 'none', 'dropout', 'l1l2', 'dropout_l1l2'; use_tf=False skips the reward net (ac_irl.py:33-37).
        In-memory data: mat_pi0 [n,d], demonstrations = list of trajectories, each a list of 15
        (state[d], action[d,d]) pairs -- the structure read_demonstrations returns."""
        engine.require_cuda()
        self.summarize = summarize
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.dtype = torch.float32                       # the TF graph is float32 (ac_irl.py:239-246)
        self.fused_reward_step = True                    # single-rank reward update through ONE C call (dmfg_irl_reward_step)
        self.resident_trajectories = True                # update_reward: sampled trajectories live in a device pool
        self.group = None                                # process group of update_reward's all-reduce (None: the default
                                                         # group when torch.distributed is initialised; False: never)
        self.theta = theta
        self.theta_initial = theta
        self.shift = shift
        self.alpha_scale = alpha_scale
        self.w = self.init_w(d)
        self.d = d
        self.lr_reward = lr_reward
        self.num_policies = num_policies
        self.c = c
        if reg not in ('none', 'dropout', 'l1l2', 'dropout_l1l2'):
            raise ValueError("reg must be 'none', 'dropout', 'l1l2' or 'dropout_l1l2'")
        self.reg = reg
        self.n_fc3 = n_fc3
        self.n_fc4 = n_fc4
        self.use_z = bool(use_z)
        self.one_pass_reward_update = True    # generated half of update_reward_batch in one launch (z_j = 1 only)
        self.rank_invariant_reward_step = False   # force the data-parallel (raw sums) form on ONE rank (tests)
        self.seed = int(np.random.randint(2 ** 31 - 1)) if seed is None else int(seed)
        self.net_seed = net_seed
        self._draws = 0
        self._episodes = 0
        self._batch_episodes = 0
        self._dropout_calls = 0
        self.mat_alpha = np.zeros([d, d])
        self.mat_alpha_deriv = np.zeros([d, d])

        cwd = os.getcwd()
        if mat_pi0 is not None:
            self.mat_pi0 = np.array(mat_pi0, dtype=np.float64)[:, :d].copy()
        else:
            self.init_pi0(path_to_dir=cwd + '/train_normalized_round2')
        self.num_start_samples = self.mat_pi0.shape[0]
        if mat_pi0_test is not None:
            self.mat_pi0_test = np.array(mat_pi0_test, dtype=np.float64)[:, :d].copy()
        elif mat_pi0 is None:
            self.init_pi0_test(path_to_dir=cwd + '/test_normalized_round2', day_start=22)
        else:
            self.mat_pi0_test = self.mat_pi0.copy()
        self.num_start_samples_test = self.mat_pi0_test.shape[0]

        if demonstrations is not None:
            self.list_demonstrations = list(demonstrations)
        elif mat_pi0 is None:
            self.list_demonstrations = self.read_demonstrations(state_dir='./train_normalized_round2',
                                                                action_dir='./actions_2', dim_action=20, start_day=1)
        else:
            self.list_demonstrations = []
        if demonstrations_test is not None:
            self.list_demonstrations_test = list(demonstrations_test)
        elif mat_pi0 is None:
            self.list_demonstrations_test = self.read_demonstrations(state_dir='./test_normalized_round2',
                                                                     action_dir='./actions_test_2', dim_action=20,
                                                                     start_day=22)
        else:
            self.list_demonstrations_test = []
        self.list_eval_demo_transitions = [pair for traj in self.list_demonstrations for pair in traj]
        self.list_generated = []
        self.list_eval_gen_transitions = []

        if use_tf:
            self.create_network()
        self.num_demo_samples = 5
        self.num_gen_samples = 5
        self.num_sampled_trajectories = self.num_gen_samples
        self.list_policies = [theta] * self.num_policies
        self.reward_update_count = 0
        self.loss_val = self.first_term_val = self.second_term_val = float("nan")
        if use_tf:
            self.create_training_method()
            self.sess = Session(self)
            if saved_network:
                self.restore("saved/" + saved_network)

    # ------------------------------------------------------------- file formats
    def init_pi0_test(self, path_to_dir, day_start=22, verbose=0):
        """First line of trend_distribution_day<k>.csv for k = day_start.. (ac_irl.py:477-506)."""
        rows = []
        for k in range(day_start, day_start + len(os.listdir(path_to_dir))):
            name = "trend_distribution_day%d.csv" % k
            with open(os.path.join(path_to_dir, name)) as f:
                rows.append([float(v) for v in f.readline().split()][: self.d])
            if verbose:
                print(name)
        self.mat_pi0_test = np.array(rows, dtype=np.float64)

    def read_demonstrations(self, state_dir, action_dir, dim_action=20, start_day=1):
        """List of trajectories, each 15 (state[d], action[d,d]) pairs (ac_irl.py:164-200).
        trend_distribution_day<k>.csv: 16 space-separated rows; action_day<k>.txt: 15 blocks of
        dim_action rows (blank lines between blocks are skipped), top-left d x d is used."""
        print("Inside read_demonstrations")
        num_file_action = len(os.listdir(action_dir))
        if num_file_action != len(os.listdir(state_dir)):
            print("Weird")
        out = []
        for idx_day in range(start_day, start_day + num_file_action):
            states = np.loadtxt(os.path.join(state_dir, "trend_distribution_day%d.csv" % idx_day), ndmin=2)
            actions = np.loadtxt(os.path.join(action_dir, "action_day%d.txt" % idx_day), ndmin=2)
            traj = []
            for hour in range(T_STEPS):
                state = states[hour, 0:self.d]
                action = actions[hour * dim_action:(hour * dim_action + self.d), 0:self.d]
                traj.append((state, action))
            out.append(traj)
        return out

    def get_eval_transitions(self, list_trajectories):
        """One (s, a) per trajectory: index idx mod 15 (ac_irl.py:203-218)."""
        return [traj[idx % T_STEPS] for idx, traj in enumerate(list_trajectories)]

    # ------------------------------------------------------------- reward net
    def create_network(self):
        """Reward net on demo and generated placeholders with SHARED variables (ac_irl.py:232-267)."""
        print("Inside create_network")
        self.demo_actions = Placeholder('demo_actions', [None, self.d, self.d])
        self.demo_states = Placeholder('demo_states', [None, self.d])
        self.gen_actions = Placeholder('gen_actions', [None, self.d, self.d])
        self.gen_states = Placeholder('gen_states', [None, self.d])
        fn = {'none': networks.r_net, 'dropout': networks.r_net_dropout, 'l1l2': networks.r_net_l1l2,
              'dropout_l1l2': networks.r_net_dropout_l1l2}[self.reg]
        networks._scopes.pop("reward", None)             # a fresh variable set per instance
        with networks.variable_scope("reward", device=self.device, seed=self.net_seed):
            self.reward_demo = fn(self.demo_states, self.demo_actions, f1=1, k1=5, f2=2, k2=3,
                                  n_fc3=self.n_fc3, n_fc4=self.n_fc4, d=self.d)
            self.reward_gen = fn(self.gen_states, self.gen_actions, f1=1, k1=5, f2=2, k2=3,
                                 n_fc3=self.n_fc3, n_fc4=self.n_fc4, d=self.d)
        self.reward_params = self.reward_demo.params
        assert self.reward_gen.params is self.reward_params

    def create_training_method(self):
        """Loss = -1/N sum r_demo + ln(1/M sum_j [z_j] exp(sum_t r_gen[j,t])) [+ l1l2], Adam(lr_reward)
        (ac_irl.py:382-426).  Nothing to build here beyond the optimiser state."""
        print("Inside create_training_method")
        self._dropout = self.reg in ('dropout', 'dropout_l1l2')
        self._d_demo_const = {}
        self._l1l2 = self.reg in ('l1l2', 'dropout_l1l2')
        self.reward_params.m.zero_()
        self.reward_params.v.zero_()
        self.reward_params.step = 0

    def _next_dropout_offset(self, n):
        """Distinct Philox sample ids for every reward-net evaluation (TF draws fresh masks per run)."""
        off = self._dropout_calls
        self._dropout_calls += int(n)
        return off

    def _reward(self, states, actions):
        """reward net on device tensors states [N,d], actions [N,d,d] -> [N]; dropout always active
        for the dropout variants (quirk C.8)."""
        p = self.reward_params
        kw = {}
        if self._dropout:
            kw = dict(seed=self.seed ^ 0x5DEECE66D, sample_offset=self._next_dropout_offset(states.shape[0]))
        return engine.rnet_forward(p.flat, states, actions, p.n_fc3, p.n_fc4, keep_prob=networks.KEEP_PROB, **kw)

    def _fetch(self, fetch, feed):
        if fetch is self.reward_gen:
            s, a = feed[self.gen_states], feed[self.gen_actions]
        elif fetch is self.reward_demo:
            s, a = feed[self.demo_states], feed[self.demo_actions]
        else:
            raise KeyError("unknown fetch %r: this shim evaluates reward_demo / reward_gen only" % (fetch,))
        s = self._dev(np.asarray(s, dtype=np.float32).reshape(-1, self.d), torch.float32)
        a = self._dev(np.asarray(a, dtype=np.float32).reshape(-1, self.d, self.d), torch.float32)
        return self._reward(s, a).cpu().numpy().reshape(-1, 1)

    def _named_slots(self, flat):
        p = self.reward_params
        flat = flat.cpu().numpy()
        return {n: flat[o:o + int(np.prod(sh))].reshape(sh).copy() for n, sh, o in zip(p.NAMES, p.shapes, p.offsets)}

    def save(self, path):
        """tf.train.Saver.save (ac_irl.py:948): writes a TF1 "V2" checkpoint bundle -- ``path.index`` +
        ``path.data-00000-of-00001`` -- with the variables a Saver of the reference's graph holds:
        reward/{conv1,conv2,fc3,fc4,out}/{weights,biases}, their Adam slots (``.../Adam``, ``.../Adam_1``) and
        beta1_power / beta2_power (tf_checkpoint.py, no TensorFlow involved).  A ``.npz`` with the same tensors is
        written next to it (round 1's format, still restorable)."""
        from . import tf_checkpoint
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        p = self.reward_params
        tensors = {"reward/" + k: v for k, v in p.named().items()}
        bundle = dict(tensors)
        for k, v in self._named_slots(p.m).items():
            bundle["reward/%s/Adam" % k] = v
        for k, v in self._named_slots(p.v).items():
            bundle["reward/%s/Adam_1" % k] = v
        bundle["beta1_power"] = np.float32(0.9 ** (p.step + 1))          # TF stores beta^t for the NEXT step
        bundle["beta2_power"] = np.float32(0.999 ** (p.step + 1))
        tf_checkpoint.write_bundle(path, bundle)
        np.savez(path, adam_m=p.m.cpu().numpy(), adam_v=p.v.cpu().numpy(), adam_step=p.step, **tensors)

    def restore(self, path):
        """tf.train.Saver.restore (ac_irl.py:110-111): ``path`` is the checkpoint prefix.  Reads a TF1 V2 bundle
        (written by TensorFlow or by save()) through the TF-free reader; falls back to round 1's ``.npz``."""
        from . import tf_checkpoint
        p = self.reward_params
        if os.path.exists(path + ".index"):
            z = tf_checkpoint.read_bundle(path)
            missing = [n for n in p.NAMES if "reward/" + n not in z]
            if missing:
                raise KeyError("checkpoint %s lacks reward-net variables %s (has: %s)" % (path, missing, sorted(z)))
            p.load_named({k[len("reward/"):]: v for k, v in z.items() if k.startswith("reward/")})
            if all("reward/%s/Adam" % n in z and "reward/%s/Adam_1" % n in z for n in p.NAMES):
                flat_m, flat_v = p.m.cpu().numpy(), p.v.cpu().numpy()
                for n, sh, o in zip(p.NAMES, p.shapes, p.offsets):
                    k = int(np.prod(sh))
                    flat_m[o:o + k] = np.asarray(z["reward/%s/Adam" % n], dtype=np.float32).reshape(-1)
                    flat_v[o:o + k] = np.asarray(z["reward/%s/Adam_1" % n], dtype=np.float32).reshape(-1)
                p.m.copy_(torch.as_tensor(flat_m, device=p.device))
                p.v.copy_(torch.as_tensor(flat_v, device=p.device))
                if "beta1_power" in z:                        # beta1^(t+1) -> steps taken so far
                    p.step = max(0, int(round(math.log(float(z["beta1_power"])) / math.log(0.9))) - 1)
            return
        if not path.endswith(".npz"):
            path = path + ".npz"
        z = np.load(path)
        p.load_named({k[len("reward/"):]: z[k] for k in z.files if k.startswith("reward/")})
        if "adam_m" in z.files:
            p.m.copy_(torch.as_tensor(z["adam_m"], device=p.device))
            p.v.copy_(torch.as_tensor(z["adam_v"], device=p.device))
            p.step = int(z["adam_step"])

    # ------------------------------------------------------------- policy pieces
    def calc_alpha_deriv(self, pi):
        """mat_alpha_deriv = x / (1 + exp(-theta x)), x = pi_j - pi_i - shift (ac_irl.py:573-588)."""
        pi = np.asarray(pi, dtype=np.float64).reshape(1, self.d)
        dummy = self._dev(np.full((1, 1, self.d, self.d), 1.0 / self.d))
        out = engine.rollout(self._dev(pi), self.theta, self.shift, self.alpha_scale, 1, reward="none",
                             actions_in=dummy, outputs=("alpha", "alpha_deriv"))
        self.mat_alpha_deriv = out["alpha_deriv"][0, 0].double().cpu().numpy()

    # ------------------------------------------------------------- a8: forward solve with the learned reward
    def train(self, max_episodes=4000, stop_criteria=0.01, gamma=1, constant=False, lr_critic=0.1, lr_actor=0.001,
              consecutive=100, file_theta='results/theta.csv', file_pi='results/pi.csv',
              file_reward='results/reward.csv', write_file=0, write_all=0, start_rows=None, noise_y=None,
              verbose=True, use_graph=True, fused=True):
        """Actor-critic with r = r_net(pi, P), per-step online updates of w then theta (ac_irl.py:634-732).

        d = 15 / 16 (``fused=True``): the whole learner -- sampling, pi', the reward net on (pi_t, P_t), TD error and the
        two updates, for all episodes up to the next report -- is ONE kernel (dmfg_irl_learners, a CTA per learner).
        Otherwise one serial learner driven from the host: every transition is a dmfg_rollout (sample P, pi', gradient)
        -> dmfg_rnet_forward -> dmfg_td_accumulate -> dmfg_ac_apply_update chain on the device.
        ``start_rows`` [E] / ``noise_y`` [E,15,d,d] replay draws."""
        if verbose:
            print("----- Starting train -----")
        d, T = self.d, T_STEPS
        theta = torch.tensor([float(self.theta)], dtype=torch.float64, device=self.device)
        w = self._w_dev().clone()
        mat = self._dev(self.mat_pi0)
        list_reward = []
        prev_theta = float(self.theta)
        episode = 0
        if fused and d in (15, 16) and not write_all:
            episode = self._train_fused(theta, w, mat, max_episodes, stop_criteria, gamma, constant, lr_critic, lr_actor,
                                        consecutive, file_theta, file_pi, file_reward, write_file, start_rows, noise_y,
                                        verbose)
            self.theta = float(theta[0])
            self.w = w.cpu().numpy().reshape(-1, 1)
            self._episodes += max_episodes
            self.list_policies = (self.list_policies + [self.theta])[1:]                    # ac_irl.py:731
            if verbose:
                print("----- Exiting train at episode %d with theta %f -----" % (episode, self.theta))
            return

        def run_episode(pi, lr_c, lr_a, step_base, noise_ep, lr_dev=None, step_dev=None):
            """the 15-transition chain of one episode: 4 launches per transition, parameters on the device"""
            disc = 1.0
            total = torch.zeros((), dtype=torch.float64, device=self.device)
            for t in range(T):
                noise = None if noise_ep is None else self._dev(np.asarray(noise_ep[t]).reshape(1, 1, d, d))
                out = engine.rollout(pi, 0.0, self.shift, self.alpha_scale, 1, theta_dev=theta, reward="none",
                                     noise_y=noise, seed=self.seed, step_offset=step_base + t, step_offset_dev=step_dev,
                                     outputs=("states", "actions", "grads"))
                if write_all:                                     # ac_irl.py:670-676
                    self.write_all_step('temp.csv', t + 1, out["states"][0, 0].double().cpu().numpy(),
                                        out["actions"][0, 0].double().cpu().numpy())
                r = self._reward(out["states"][0], out["actions"][0])
                td = engine.td_accumulate(out["states"], r.reshape(1, 1), out["grads"], w, gamma=disc,
                                          discount="step", want_deltas=False)
                engine.apply_update(d, theta, w, td["acc"], lr_c, lr_a, 1.0, lr_dev=lr_dev)
                total = total + td["acc"][-1]
                disc *= gamma
                pi = out["states"][1]
            return pi, total

        # The chain is launch-bound (60 tiny launches per episode): with sampled noise and no dropout stream to
        # advance it is captured ONCE as a CUDA graph and replayed per episode -- the start state, the two step sizes
        # and the Philox position live in device buffers refreshed before each replay (about 6x faster per step).
        graph = None
        scratch_streams = []       # streams whose cached workspaces (one of them owned by the graph's pool) die with train()
        if use_graph and noise_y is None and not self._dropout and max_episodes >= 4 and not write_all:
            try:
                pi_static = torch.empty((1, d), dtype=torch.float32, device=self.device)
                lr_dev = torch.zeros(2, dtype=torch.float64, device=self.device)
                step_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
                host = torch.zeros(3, dtype=torch.float64).pin_memory()
                theta0, w0 = theta.clone(), w.clone()
                side = torch.cuda.Stream(self.device)
                side.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(side):                 # warm-up outside capture (allocator, lazy init)
                    pi_static.copy_(mat[0:1])
                    run_episode(pi_static, 0.0, 0.0, 0, None, lr_dev=lr_dev, step_dev=step_dev)
                torch.cuda.current_stream(self.device).wait_stream(side)
                theta.copy_(theta0); w.copy_(w0)
                graph = torch.cuda.CUDAGraph()
                scratch_streams.append(side)
                with torch.cuda.graph(graph):
                    scratch_streams.append(torch.cuda.current_stream(self.device))
                    pi_end, total_static = run_episode(pi_static, 0.0, 0.0, 0, None, lr_dev=lr_dev, step_dev=step_dev)
                theta.copy_(theta0); w.copy_(w0)              # capture does not execute, but keep the state explicit
            except Exception as exc:                          # pragma: no cover - eager launches are the same kernels
                print("CUDA graph capture unavailable (%s): eager launches" % exc)
                graph = None
        pi = None
        for episode in range(1, max_episodes + 1):
            if start_rows is not None:
                row = int(start_rows[episode - 1])
            else:
                row = self._philox_randint(self._episodes + episode, self.num_start_samples)
            lr_c = lr_critic if constant else lr_critic / (episode + 1.0)
            lr_a = lr_actor if constant else lr_actor / ((episode + 1.0) * math.log(math.log(episode + 20.0)))
            step_base = (self._episodes + episode) * T
            if write_all:                                         # ac_irl.py:651-653
                with open('temp.csv', 'a') as f:
                    f.write('Episode %d \n\n' % episode)
            if graph is not None:
                pi_static.copy_(mat[row:row + 1])
                lr_dev.copy_(torch.tensor([lr_c, lr_a], dtype=torch.float64), non_blocking=False)
                step_dev.fill_(step_base)
                graph.replay()
                pi, total = pi_end, total_static.clone()
            else:
                pi, total = run_episode(mat[row:row + 1].contiguous(), lr_c, lr_a, step_base,
                                        None if noise_y is None else noise_y[episode - 1])
            list_reward.append(total)
            if episode % consecutive == 0:
                self.theta = float(theta[0])
                pi_host = pi[0].double().cpu().numpy()
                reward_avg = float(torch.stack(list_reward).sum()) / consecutive
                if verbose:
                    print("Theta\n", self.theta)
                    print("pi\n", pi_host)
                    print("Average reward during previous %d episodes: " % consecutive, str(reward_avg))
                list_reward = []
                if write_file:
                    self.train_log(np.array([self.theta]), file_theta, "%.5e")
                    self.train_log(pi_host, file_pi, "%.3e")
                    self.train_log(np.array([reward_avg]), file_reward, "%.3e")
            if stop_criteria != -1:
                cur = float(theta[0])
                if abs(cur - prev_theta) < stop_criteria:
                    break
                prev_theta = cur
        self.theta = float(theta[0])
        self.w = w.cpu().numpy().reshape(-1, 1)
        self._episodes += max_episodes
        del graph
        for st in scratch_streams:
            engine.drop_workspace(self.device, st)
        self.list_policies = (self.list_policies + [self.theta])[1:]                    # ac_irl.py:731
        if verbose:
            print("----- Exiting train at episode %d with theta %f -----" % (episode, self.theta))

    def _train_fused(self, theta, w, mat, max_episodes, stop_criteria, gamma, constant, lr_critic, lr_actor, consecutive,
                     file_theta, file_pi, file_reward, write_file, start_rows, noise_y, verbose):
        """train() on dmfg_irl_learners: chunks that end at the reference's report episodes (multiples of
        ``consecutive``); with a stop criterion theta after every episode comes back with the chunk, and a chunk in which
        |theta_e - theta_{e-1}| < stop_criteria fires is re-run up to that episode (draws are counter-based, so the
        re-run is identical).  Returns the last episode run (1-based)."""
        d, T = self.d, T_STEPS
        p = self.reward_params
        w2 = w.reshape(1, -1)
        prev_theta = float(self.theta)
        list_reward = []
        e = 0
        while e < max_episodes:
            n = min(consecutive - (e % consecutive), max_episodes - e)

            def run(n_run, trace):
                kw = {}
                if noise_y is not None:
                    kw["noise_y"] = self._dev(np.asarray(noise_y)[e:e + n_run].reshape(1, n_run, T, d, d), torch.float32)
                if start_rows is not None:
                    kw["start_rows"] = self._dev(np.asarray(start_rows)[e:e + n_run].reshape(1, n_run), torch.int32)
                if self._dropout:
                    kw.update(dropout_seed=self.seed ^ 0x5DEECE66D, sample_offset=self._dropout_calls)
                return engine.irl_learners(theta, w2, mat, n_run, T, p.flat, p.n_fc3, p.n_fc4, shift=self.shift,
                                           alpha_scale=self.alpha_scale, episode0=1 + e, gamma=gamma, lr_critic=lr_critic,
                                           lr_actor=lr_actor, constant=bool(constant), discount="cumulative",
                                           seed=self.seed, noise_episode_offset=self._episodes,
                                           keep_prob=networks.KEEP_PROB, trace=trace, **kw)
            stop_at = None
            if stop_criteria != -1:
                theta0, w0 = theta.clone(), w2.clone()
                res = run(n, True)
                th = [prev_theta] + res["theta_trace"][0, :, -1].cpu().tolist()
                for k in range(n):
                    if abs(th[k + 1] - th[k]) < stop_criteria:
                        stop_at = k + 1
                        break
                if stop_at is not None and stop_at < n:
                    theta.copy_(theta0); w2.copy_(w0)
                    res = run(stop_at, False)
                n_done = stop_at if stop_at is not None else n
                prev_theta = th[n_done]
            else:
                res = run(n, False)
                n_done = n
            if self._dropout:
                self._dropout_calls += n_done * T
            list_reward += res["total_reward"][0, :n_done].cpu().tolist()
            e += n_done
            if e % consecutive == 0:
                self.theta = float(theta[0])
                pi_host = res["pi_final"][0].double().cpu().numpy()
                reward_avg = sum(list_reward) / consecutive
                if verbose:
                    print("Theta\n", self.theta)
                    print("pi\n", pi_host)
                    print("Average reward during previous %d episodes: " % consecutive, str(reward_avg))
                list_reward = []
                if write_file:
                    self.train_log(np.array([self.theta]), file_theta, "%.5e")
                    self.train_log(pi_host, file_pi, "%.3e")
                    self.train_log(np.array([reward_avg]), file_reward, "%.3e")
            if stop_at is not None:
                break
        return e

    def train_batch(self, pi0, num_episodes=1, gamma=1, constant=False, lr_critic=0.1, lr_actor=0.001, seed=None,
                    pop_offset=0, group=None, first_episode=1, noise_y=None, keep_record=False):
        """Batched forward solve with the reward net in the loop: B populations share (theta, w), parameters
        frozen within an episode, one batch-mean update per episode (the data-parallel form of train(); equal to
        it in expectation, not step by step -- DESIGN.md section 5).  Per episode, on the device:
        rollout + record (states, actions, d log F/d theta)  ->  r = r_net(pi_t, P_t) for all B*15 transitions
        ->  TD errors with the cumulative discount of ac_irl.py:691 and the [2+F] sums  ->  (one all-reduce when
        ``group`` is a process group)  ->  theta, w update with train()'s step sizes (episodes count from 1).
        pi0 [B,d]: CUDA tensor or host array.  ``noise_y`` [E,15,B,d,d] injects the Gamma variates (parity).
        ``keep_record=True`` returns the last episode's (states [16,B,d], actions [15,B,d,d]) -- the generated
        batch a reward update consumes, so sampling costs nothing extra.  Returns dict(theta, mean_reward [E])."""
        from . import parallel
        d, T = self.d, T_STEPS
        seed = self.seed if seed is None else seed
        theta = torch.tensor([float(self.theta)], dtype=torch.float64, device=self.device)
        w = self._w_dev().clone()
        pi = pi0 if isinstance(pi0, torch.Tensor) and pi0.is_cuda else self._dev(np.asarray(pi0, dtype=np.float32), torch.float32)
        B = pi.shape[0]
        total_pops = parallel.total_count(B, group, self.device)
        p = self.reward_params
        mean_rewards, rec = [], None
        for e in range(num_episodes):
            episode = first_episode + e
            lr_c = lr_critic if constant else lr_critic / (episode + 1.0)
            lr_a = lr_actor if constant else lr_actor / ((episode + 1.0) * math.log(math.log(episode + 20.0)))
            noise = None if noise_y is None else self._dev(np.asarray(noise_y[e], dtype=np.float32), torch.float32)
            rec = engine.rollout(pi, 0.0, self.shift, self.alpha_scale, T, theta_dev=theta, reward="none",
                                 noise_y=noise, seed=seed, pop_offset=pop_offset,
                                 step_offset=(self.first_episode + self._batch_episodes + e) * T,
                                 outputs=("states", "actions", "grads"))
            kd = {}
            if self._dropout:
                kd = dict(seed=self.seed ^ 0x5DEECE66D, sample_offset=self._next_dropout_offset(T * B))
            r = engine.rnet_forward(p.flat, rec["states"][:T].reshape(-1, d), rec["actions"].reshape(-1, d, d),
                                    p.n_fc3, p.n_fc4, keep_prob=networks.KEEP_PROB, **kd)
            td = engine.td_accumulate(rec["states"], r.view(T, B), rec["grads"], w, gamma=gamma,
                                      discount="cumulative", want_deltas=False)
            acc = parallel.allreduce_sum_(td["acc"], group)
            engine.apply_update(d, theta, w, acc, lr_c, lr_a, 1.0 / total_pops)
            mean_rewards.append(acc[-1] / total_pops)
        self._batch_episodes += num_episodes        # first_episode restarts the step sizes, never the noise
        self.theta = float(theta[0])
        self.w = w.cpu().numpy().reshape(-1, 1)
        self.list_policies = (self.list_policies + [self.theta])[1:]
        out = dict(theta=self.theta, mean_reward=torch.stack(mean_rewards).cpu().numpy() if mean_rewards else np.zeros(0))
        if keep_record and rec is not None:
            out["states"], out["actions"] = rec["states"], rec["actions"]
        return out

    def irl_step_batch(self, pi0, demo_states, demo_actions, num_demo_traj, episode=1, gamma=1, constant=False,
                       lr_critic=0.1, lr_actor=0.001, seed=None, pop_offset=0, group=None):
        """One data-parallel IRL training step (BASELINE config 5) = train_batch(num_episodes=1, keep_record=True)
        followed by update_reward_batch on that record, with the shared work done once:
          rollout + record of this rank's B populations  ->  reward-net backward over the demonstrations (-1/N)
          ->  ONE reward-net pass over the generated record that hands back r_gen (the rewards of the forward solve)
              AND the generated half of the reward gradient / loss (rnet_kernel<TRAJ>)
          ->  TD sums from r_gen  ->  ONE all-reduce of the flat [2+F+|r_net|] double buffer (actor, critic and
              reward gradients together)  ->  theta, w update and the Adam step.
        Same numbers as the two calls in sequence (both evaluate the reward net with the parameters before the Adam
        step); with dropout regularisers the two calls draw separate masks, so this method falls back to them.
        demo_* : [N*15, d] / [N*15, d, d] device tensors.  Returns dict(theta, mean_reward, loss [4] device,
        states, actions)."""
        from . import parallel
        if self._dropout or self.use_z or not self.one_pass_reward_update or self.d > 16:
            res = self.train_batch(pi0, 1, gamma, constant, lr_critic, lr_actor, seed, pop_offset, group, episode,
                                   keep_record=True)
            T = T_STEPS
            res["loss"] = self.update_reward_batch(demo_states, demo_actions, res["states"][:T].reshape(-1, self.d),
                                                   res["actions"].reshape(-1, self.d, self.d), num_demo_traj,
                                                   "time_major", group=group)
            res["mean_reward"] = res["mean_reward"][0]
            return res
        d, T = self.d, T_STEPS
        seed = self.seed if seed is None else seed
        theta = torch.tensor([float(self.theta)], dtype=torch.float64, device=self.device)
        w = self._w_dev().clone()
        pi = pi0 if isinstance(pi0, torch.Tensor) and pi0.is_cuda else self._dev(np.asarray(pi0, dtype=np.float32), torch.float32)
        B = pi.shape[0]
        _, world = parallel.world_info(group)
        p = self.reward_params
        lr_c = lr_critic if constant else lr_critic / (episode + 1.0)
        lr_a = lr_actor if constant else lr_actor / ((episode + 1.0) * math.log(math.log(episode + 20.0)))
        rec = engine.rollout(pi, 0.0, self.shift, self.alpha_scale, T, theta_dev=theta, reward="none", seed=seed,
                             pop_offset=pop_offset, step_offset=(self.first_episode + self._batch_episodes) * T,
                             outputs=("states", "actions", "grads"))
        self._batch_episodes += 1              # `episode` drives the step sizes only; the noise never restarts
        gs, ga = rec["states"][:T].reshape(-1, d), rec["actions"].reshape(-1, d, d)
        if world > 1 or self.rank_invariant_reward_step:
            # data-parallel form: every rank contributes RAW sums -- the demonstration gradient for dL/dr = -1, the
            # unnormalised generated gradient sum_j e^{R_j} dR_j, Z = sum_j e^{R_j}, sum r_demo, the trajectory and
            # population counts -- next to its actor / critic sums; 1/N_demo, 1/Z and 1/B are applied AFTER the ONE
            # all-reduce of the flat double buffer, so the step equals the single-rank step on the concatenated batch
            terms, r_gen = self._dp_reward_terms(demo_states, demo_actions, gs, ga, num_demo_traj, "time_major",
                                                 want_rewards=True)
            td = engine.td_accumulate(rec["states"], r_gen.view(T, B), rec["grads"], w, gamma=gamma,
                                      discount="cumulative", want_deltas=False)
            nacc = td["acc"].numel()
            flat = torch.cat([td["acc"], torch.full((1,), float(B), dtype=torch.float64, device=self.device), terms])
            parallel.allreduce_sum_(flat, group)
            acc, total_b = flat[:nacc], flat[nacc:nacc + 1]
            lr_dev = torch.tensor([lr_c, lr_a], dtype=torch.float64, device=self.device) / total_b
            engine.apply_update(d, theta, w, acc, 0.0, 0.0, 1.0, lr_dev=lr_dev)
            grad, loss = engine.irl_dp_finalize(flat[nacc + 1:].contiguous(), p.flat.numel())
            mean_reward = acc[-1] / total_b[0]
        else:
            n_demo = demo_states.shape[0]
            d_const = self._demo_weight(n_demo, -1.0 / float(num_demo_traj))
            grad, r_demo = engine.rnet_backward(p.flat, demo_states, demo_actions, d_const, p.n_fc3, p.n_fc4,
                                                keep_prob=networks.KEEP_PROB, want_rewards=True)
            _, loss, r_gen = engine.rnet_backward_gen(p.flat, gs, ga, p.n_fc3, p.n_fc4, T, r_demo, num_demo_traj,
                                                      layout="time_major", grad=grad, accumulate=True,
                                                      keep_prob=networks.KEEP_PROB, want_rewards=True)
            td = engine.td_accumulate(rec["states"], r_gen.view(T, B), rec["grads"], w, gamma=gamma,
                                      discount="cumulative", want_deltas=False)
            acc = td["acc"]
            engine.apply_update(d, theta, w, acc, lr_c, lr_a, 1.0 / B)
            mean_reward = acc[-1] / B
        p.step += 1
        reg = engine.adam_tf(p.flat, p.m, p.v, grad, p.step, self.lr_reward,
                             l1l2=self._l1l2, net=(p.d, p.n_fc3, p.n_fc4), want_reg_loss=self._l1l2)
        if reg is not None:
            loss = loss.clone()
            loss[0] += reg[0]
        self._last_grad = grad
        self.theta = float(theta[0])
        self.w = w.cpu().numpy().reshape(-1, 1)
        self.list_policies = (self.list_policies + [self.theta])[1:]
        return dict(theta=self.theta, mean_reward=float(mean_reward), loss=loss, states=rec["states"],
                    actions=rec["actions"])

    def _demo_weight(self, n_demo, value):
        """constant dL/dr of the demonstration transitions (cached device vector)"""
        d_const = self._d_demo_const.get((n_demo, float(value)))
        if d_const is None:
            d_const = torch.full((n_demo,), float(value), dtype=torch.float32, device=self.device)
            self._d_demo_const = {(n_demo, float(value)): d_const}
        return d_const

    def _dp_reward_terms(self, demo_states, demo_actions, gen_states, gen_actions, num_demo_traj, layout,
                         want_rewards=False):
        """This rank's RAW contribution to a data-parallel reward step (ac_irl.py:390-406 over the union of all ranks'
        trajectories): float64 [2|r_net|+4] = {d/dparams sum_demo (-r), sum_j e^{R_j} dR_j/dparams, sum_j e^{R_j},
        sum r_demo, demonstration trajectories, generated trajectories}.  Sum over ranks -> engine.irl_dp_finalize."""
        p = self.reward_params
        grad_demo, r_demo = engine.rnet_backward(p.flat, demo_states, demo_actions,
                                                 self._demo_weight(demo_states.shape[0], -1.0), p.n_fc3, p.n_fc4,
                                                 keep_prob=networks.KEEP_PROB, want_rewards=True)
        out = engine.rnet_backward_gen(p.flat, gen_states, gen_actions, p.n_fc3, p.n_fc4, T_STEPS, r_demo, 1.0,
                                       layout=layout, keep_prob=networks.KEEP_PROB, want_rewards=want_rewards,
                                       local_sums=True)
        counts = torch.tensor([float(num_demo_traj), float(gen_states.shape[0] // T_STEPS)], dtype=torch.float64,
                              device=self.device)
        terms = torch.cat([grad_demo.double(), out[0].double(), out[1][:2], counts])
        return terms, (out[2] if want_rewards else None)

    def _philox_randint(self, counter, n):
        w0 = engine.philox((0, 0, counter & 0xFFFFFFFF, 0xC0000000), (self.seed & 0xFFFFFFFF, self.seed >> 32))[0]
        return (w0 * n) >> 32

    # ------------------------------------------------------------- a9: sampling from the current policy
    def generate_batch(self, n, from_test=False, theta=None):
        """n trajectories from the current policy in ONE launch; returns device tensors
        states [16,n,d], actions [15,n,d,d] (time-major record)."""
        mat = self.mat_pi0_test if from_test else self.mat_pi0
        rows = [self._philox_randint((1 << 20) + self._draws + i, mat.shape[0]) for i in range(n)]
        pi0 = self._dev(mat[rows])
        out = engine.rollout(pi0, self.theta if theta is None else theta, self.shift, self.alpha_scale, T_STEPS,
                             reward="none", seed=self.seed, pop_offset=(1 << 32) + self._draws,
                             outputs=("states", "actions"))
        self._draws += n
        return out["states"], out["actions"]

    def generate_trajectories(self, n, from_test=False):
        """List of n trajectories, each a list of 15 (state, action) tuples (ac_irl.py:735-767)."""
        print("Inside generate_trajectories")
        states, actions = self.generate_batch(n, from_test)
        S = states.double().cpu().numpy()
        A = actions.double().cpu().numpy()
        return [[(S[t, j], A[t, j]) for t in range(T_STEPS)] for j in range(n)]

    # ------------------------------------------------------------- a11/a12: reward update
    def _pack_host(self, trajectories):
        """list of trajectories -> trajectory-major float32 arrays states [n*15, d], actions [n*15, d, d].  A trajectory
        (a list of (state, action) pairs) is stacked ONCE and remembered: update_reward resamples the same few dozen
        trajectories hundreds of times (ac_irl.py:804-846) and the per-pair np.asarray was a third of its wall time.
        The memo is keyed by the list's id and validated by the identity of its first and last pair (which the entry
        keeps alive, so a recycled id cannot alias); arrays inside a trajectory are not expected to be edited in place
        (the reference never does)."""
        memo = self.__dict__.setdefault("_pack_memo", {})
        ss, aa = [], []
        for traj in trajectories:
            n = len(traj)
            hit = memo.get(id(traj))
            if hit is None or hit[0] != n or hit[1] is not traj[0] or hit[2] is not traj[-1]:
                if len(memo) >= 8192:
                    memo.clear()
                s = np.asarray([pair[0] for pair in traj], dtype=np.float32).reshape(n, self.d)
                a = np.asarray([pair[1] for pair in traj], dtype=np.float32).reshape(n, self.d, self.d)
                hit = memo[id(traj)] = (n, traj[0], traj[-1], s, a)
            ss.append(hit[3])
            aa.append(hit[4])
        if not ss:
            return np.zeros((0, self.d), np.float32), np.zeros((0, self.d, self.d), np.float32)
        return np.concatenate(ss), np.concatenate(aa)

    def _pack(self, trajectories):
        """list of trajectories -> trajectory-major device tensors states [n*15,d], actions [n*15,d,d]."""
        s, a = self._pack_host(trajectories)
        return self._dev(s, torch.float32), self._dev(a, torch.float32)

    def _pack_resident(self, transitions):
        """_pack([transitions]) whose device tensors are remembered: the evaluation lists of reward_iteration
        (ac_irl.py:877-884) are the same objects for a whole outer iteration and were stacked and copied every tenth update.
        Keyed like the host memo (id of the list, validated by length and the identity of its first and last pair)."""
        memo = self.__dict__.setdefault("_dev_memo", {})
        n = len(transitions)
        hit = memo.get(id(transitions))
        if hit is None or hit[0] != n or n == 0 or hit[1] is not transitions[0] or hit[2] is not transitions[-1]:
            if len(memo) >= 8:
                memo.clear()
            s, a = self._pack([transitions])
            hit = memo[id(transitions)] = (n, transitions[0] if n else None, transitions[-1] if n else None, s, a)
        return hit[3], hit[4]

    def _pack_pair(self, demo, gen):
        """both halves of a reward minibatch through ONE host-to-device copy: (demo states, demo actions, generated
        states, generated actions) as views of one device buffer"""
        parts = list(self._pack_host(demo)) + list(self._pack_host(gen))
        offs, o = [], 0
        for p in parts:                                     # every part starts on a 16-byte boundary
            offs.append(o)
            o += (p.size + 3) // 4 * 4
        flat = np.zeros(o, dtype=np.float32)
        for p, b in zip(parts, offs):
            flat[b:b + p.size] = p.reshape(-1)
        buf = self._dev(flat, torch.float32)
        return [buf[b:b + p.size].view(p.shape) for p, b in zip(parts, offs)]

    # ---- trajectories resident in device memory ---------------------------------------------------------------------
    # update_reward (ac_irl.py:804-846) draws 5 + 5 trajectories out of the same few dozen, hundreds of times per outer
    # iteration.  Each trajectory is uploaded ONCE into a pool ([slot * 15 + t] rows of states / actions) and a minibatch
    # is then just its slot numbers, which travel by value with the launch (dmfg_rnet_args.gather_*).
    _POOL_LIMIT = 1 << 14                                # slots (236 MB at d = 15) before the pool starts over

    def _resident_ok(self, demo, gen):
        from . import parallel
        from ._lib import MAX_GATHER
        if not (self.resident_trajectories and self.fused_reward_step and self.one_pass_reward_update) or self.use_z \
                or self.rank_invariant_reward_step or self.d > 16:
            return False
        if not (0 < len(demo) <= MAX_GATHER and 0 < len(gen) <= MAX_GATHER):
            return False
        if parallel.world_info(self.group)[1] != 1:
            return False
        return all(len(t) == T_STEPS for t in demo) and all(len(t) == T_STEPS for t in gen)

    def _resident_slots(self, trajectories):
        pool = self.__dict__.get("_pool")
        if pool is None:
            pool = self._pool = {"index": {}, "used": 0, "states": None, "actions": None}
        index = pool["index"]
        if pool["used"] + len(trajectories) > self._POOL_LIMIT:
            index.clear()
            pool["used"] = 0
        out = []
        for traj in trajectories:
            hit = index.get(id(traj))
            # (an entry keeps its first and last pair alive, so a recycled id() cannot alias another trajectory)
            if hit is None or hit[1] is not traj[0] or hit[2] is not traj[-1]:
                hit = index[id(traj)] = (self._pool_add(pool, traj), traj[0], traj[-1])
            out.append(hit[0])
        return out

    def _pool_add(self, pool, traj):
        T = T_STEPS
        slot = pool["used"]
        cap = 0 if pool["states"] is None else pool["states"].shape[0] // T
        if slot == cap:                                    # grow by doubling; slot numbers stay valid
            new_cap = max(64, 2 * cap)
            with torch.cuda.device(self.device):
                st = torch.empty((new_cap * T, self.d), dtype=torch.float32, device=self.device)
                ac = torch.empty((new_cap * T, self.d, self.d), dtype=torch.float32, device=self.device)
                if cap:
                    st[:cap * T].copy_(pool["states"])
                    ac[:cap * T].copy_(pool["actions"])
            pool["states"], pool["actions"] = st, ac
        s, a = self._pack_host([traj])
        pool["states"][slot * T:(slot + 1) * T].copy_(torch.from_numpy(s))
        pool["actions"][slot * T:(slot + 1) * T].copy_(torch.from_numpy(a))
        pool["used"] = slot + 1
        return slot

    def _update_reward_resident(self, demo, gen):
        """update_reward_batch's one-call branch on pool slots instead of stacked device arrays (same kernels, same
        order, same dropout offsets: bit-identical to it)."""
        p = self.reward_params
        T = T_STEPS
        slots = self._resident_slots(list(demo) + list(gen))             # one pass: a pool reset cannot split the batch
        ds, gs = slots[:len(demo)], slots[len(demo):]
        pool = self._pool
        n_demo, n_gen = len(ds) * T, len(gs) * T
        seed_d = seed_g = None
        off_d = off_g = 0
        if self._dropout:
            seed_d = seed_g = self.seed ^ 0x5DEECE66D
            off_d = self._next_dropout_offset(n_demo)
            off_g = self._next_dropout_offset(n_gen)
        d_const = self._demo_weight(n_demo, -1.0 / float(self.num_demo_samples))
        p.step += 1
        grad, loss, reg = engine.irl_reward_step(
            p.flat, p.m, p.v, p.step, self.lr_reward, pool["states"], pool["actions"], d_const, pool["states"],
            pool["actions"], p.n_fc3, p.n_fc4, T, self.num_demo_samples, layout="trajectory_major",
            keep_prob=networks.KEEP_PROB, demo_seed=seed_d, demo_sample_offset=off_d, gen_seed=seed_g,
            gen_sample_offset=off_g, l1l2=self._l1l2, want_reg_loss=self._l1l2,
            finishing_launch=self.fused_reward_step != "chain", demo_gather=(T, ds), gen_gather=(T, gs))
        if reg is not None:
            loss[0] += reg[0]
        self._last_grad = grad
        return loss

    def update_reward_batch(self, demo_states, demo_actions, gen_states, gen_actions, num_demo_traj, layout,
                            group=None, masks=None):
        """One gradient step on the reward net from device tensors (the kernel chain of update_reward):
        forward demo + gen -> loss and dL/dr -> backward demo + gen -> (all-reduce) -> Adam.
        gen_* hold M*15 transitions, `layout` 'time_major' | 'trajectory_major'.  Returns loss [4] (device)."""
        from . import parallel
        p = self.reward_params
        kd, kg = {}, {}
        if masks is not None:
            kd = dict(mask3=masks["demo3"], mask4=masks["demo4"])
            kg = dict(mask3=masks["gen3"], mask4=masks["gen4"])
        elif self._dropout:
            key = self.seed ^ 0x5DEECE66D
            kd = dict(seed=key, sample_offset=self._next_dropout_offset(demo_states.shape[0]))
            kg = dict(seed=key, sample_offset=self._next_dropout_offset(gen_states.shape[0]))
        # dL/dr of a demonstration transition is the constant -1/N (first term of ac_irl.py:390), so the
        # demonstrations need no separate forward pass: their backward launch recomputes the forward anyway
        # and hands back r_demo for the loss value.  4 -> 3 reward-net launches per update.
        _, world = parallel.world_info(group)
        if (world > 1 or self.rank_invariant_reward_step) and not self.use_z and self.one_pass_reward_update \
                and masks is None and not self._dropout and self.d <= 16:
            # rank-count invariant data-parallel step: raw sums are all-reduced, 1/N_demo and 1/Z applied afterwards
            terms, _ = self._dp_reward_terms(demo_states, demo_actions, gen_states, gen_actions, num_demo_traj, layout)
            parallel.allreduce_sum_(terms, group)
            grad, loss = engine.irl_dp_finalize(terms, p.flat.numel())
            p.step += 1
            reg = engine.adam_tf(p.flat, p.m, p.v, grad, p.step, self.lr_reward, l1l2=self._l1l2,
                                 net=(p.d, p.n_fc3, p.n_fc4), want_reg_loss=self._l1l2)
            if reg is not None:
                loss[0] += reg[0]
            self._last_grad = grad
            return loss
        n_demo = demo_states.shape[0]
        d_const = self._demo_weight(n_demo, -1.0 / float(num_demo_traj))
        if world == 1 and masks is None and not self.use_z and self.one_pass_reward_update and self.d <= 16 \
                and n_demo > 0 and self.fused_reward_step:
            # the whole update through one C call (dmfg_irl_reward_step): the two backward launches and ONE finishing launch
            # (reductions, loss terms, Adam); fused_reward_step = "chain" keeps the six launches of the calls below
            p.step += 1
            grad, loss, reg = engine.irl_reward_step(
                p.flat, p.m, p.v, p.step, self.lr_reward, demo_states, demo_actions, d_const, gen_states, gen_actions,
                p.n_fc3, p.n_fc4, T_STEPS, num_demo_traj, layout=layout, keep_prob=networks.KEEP_PROB,
                demo_seed=kd.get("seed"), demo_sample_offset=kd.get("sample_offset", 0),
                gen_seed=kg.get("seed"), gen_sample_offset=kg.get("sample_offset", 0),
                l1l2=self._l1l2, want_reg_loss=self._l1l2, finishing_launch=self.fused_reward_step != "chain")
            if reg is not None:
                loss[0] += reg[0]
            self._last_grad = grad
            return loss
        grad, r_demo = engine.rnet_backward(p.flat, demo_states, demo_actions, d_const, p.n_fc3, p.n_fc4,
                                            keep_prob=networks.KEEP_PROB, want_rewards=True, **kd)
        if not self.use_z and self.one_pass_reward_update and self.d <= 16:
            # z_j = 1 (upstream's ac_irl.py:406): the generated half runs in ONE reward-net launch -- trajectory by
            # trajectory, weight exp(R_j), 1/sum_j exp(R_j) applied to the reduced gradient -- instead of
            # forward -> loss / dL/dr -> backward.  3 -> 2 reward-net launches per update.
            _, loss4 = engine.rnet_backward_gen(p.flat, gen_states, gen_actions, p.n_fc3, p.n_fc4, T_STEPS, r_demo,
                                                num_demo_traj, layout=layout, grad=grad, accumulate=True,
                                                keep_prob=networks.KEEP_PROB, **kg)
            res = {"loss": loss4}
        else:
            r_gen = engine.rnet_forward(p.flat, gen_states, gen_actions, p.n_fc3, p.n_fc4, keep_prob=networks.KEEP_PROB,
                                        **kg)
            log_z = None
            if self.use_z:
                thetas = torch.as_tensor(np.asarray(self.list_policies, dtype=np.float64), device=self.device)
                lq = engine.dirichlet_logq(gen_states, gen_actions, thetas, self.shift)
                log_z = engine.irl_log_z(lq, T_STEPS, self.num_start_samples, layout=layout)
            res = engine.irl_loss_grad(r_demo, r_gen, T_STEPS, num_demo_traj, layout=layout, log_z=log_z)
            engine.rnet_backward(p.flat, gen_states, gen_actions, res["d_gen"], p.n_fc3, p.n_fc4, grad=grad,
                                 accumulate=True, keep_prob=networks.KEEP_PROB, **kg)
        parallel.allreduce_sum_(grad, group)
        p.step += 1
        reg = engine.adam_tf(p.flat, p.m, p.v, grad, p.step, self.lr_reward, grad_scale=1.0 / world,
                             l1l2=self._l1l2, net=(p.d, p.n_fc3, p.n_fc4), want_reg_loss=self._l1l2)
        loss = res["loss"]
        if reg is not None:
            loss = loss.clone()
            loss[0] += reg[0]
        self._last_grad = grad
        return loss

    def update_reward(self, summary=False, iteration=0):
        """Sample 5 demo + 5 generated trajectories, one Adam step on the reward net (ac_irl.py:804-846)."""
        if len(self.list_demonstrations) >= self.num_demo_samples:
            demo_sampled = random.sample(self.list_demonstrations, self.num_demo_samples)
        else:
            demo_sampled = self.list_demonstrations[:]
        if len(self.list_generated) >= self.num_gen_samples:
            gen_sampled = random.sample(self.list_generated, self.num_gen_samples)
        else:
            gen_sampled = self.list_generated[:]
        if self._resident_ok(demo_sampled, gen_sampled):
            # the sampled trajectories are already in device memory (uploaded the first time they were drawn): the update
            # is one C call with the ten pool slots by value -- no stacking, no host-to-device copy
            self._loss_dev = self._update_reward_resident(demo_sampled, gen_sampled)
            return
        ds, da, gs, ga = self._pack_pair(demo_sampled, gen_sampled)
        # the loss terms stay on the device until someone reads loss_val / first_term_val / second_term_val (the reference
        # prints them every iter_check updates): no device-to-host synchronisation per update
        self._loss_dev = self.update_reward_batch(ds, da, gs, ga, self.num_demo_samples, "trajectory_major", group=self.group)

    def _loss_terms(self):
        dev = self.__dict__.get("_loss_dev")
        if dev is not None:
            h = dev.cpu().numpy()
            self._loss_host = (float(h[0]), float(h[1]), float(h[2]))
            self._loss_dev = None
        return self.__dict__.get("_loss_host", (float("nan"),) * 3)

    def _set_loss_term(self, k, value):
        t = list(self._loss_terms())
        t[k] = float(value)
        self._loss_host = tuple(t)

    loss_val = property(lambda self: self._loss_terms()[0], lambda self, v: self._set_loss_term(0, v))
    first_term_val = property(lambda self: self._loss_terms()[1], lambda self, v: self._set_loss_term(1, v))
    second_term_val = property(lambda self: self._loss_terms()[2], lambda self, v: self._set_loss_term(2, v))

    def reward_iteration(self, max_iterations=500, stop_criteria=0.01, iter_check=10, verbose=True):
        """Reward updates with an evaluation / early-stop check every iter_check (ac_irl.py:849-897)."""
        prev_reward_demo_avg = -100
        if verbose:
            print("----- Starting reward_iteration -----")
        it = 0
        for it in range(1, max_iterations + 1):
            self.reward_update_count += 1
            if it % iter_check != 0:
                self.update_reward(summary=False)
                continue
            if verbose:
                print("Reward iteration %d" % it)
            self.update_reward(summary=False, iteration=self.reward_update_count)
            # (the flat transition lists are packed once and remembered, like the trajectories of update_reward)
            ds, da = self._pack_resident(self.list_eval_demo_transitions)
            gs, ga = self._pack_resident(self.list_eval_gen_transitions)
            sums = torch.stack([self._reward(ds, da).double().sum(), self._reward(gs, ga).double().sum()]).cpu()
            reward_demo_avg = float(sums[0]) / len(self.list_eval_demo_transitions)      # (one read-back for both)
            reward_gen_avg = float(sums[1]) / len(self.list_eval_gen_transitions)
            if verbose:
                print("Reward demo avg %f | Reward gen avg %f" % (reward_demo_avg, reward_gen_avg))
                print("First %f | Second %f | Loss %f" % (self.first_term_val, self.second_term_val, self.loss_val))
            if np.isnan(reward_demo_avg) or np.isnan(reward_gen_avg):
                break
            os.makedirs("results", exist_ok=True)
            with open("results/reward_training.csv", 'a') as f:
                f.write("%f,%f\n" % (reward_demo_avg, reward_gen_avg))
            if stop_criteria != -1 and abs(reward_demo_avg - prev_reward_demo_avg) < stop_criteria:
                break
            prev_reward_demo_avg = reward_demo_avg
        if verbose:
            print("----- Exiting reward_iteration at iter %d -----" % it)

    # ------------------------------------------------------------- a14
    def outerloop(self, num_iterations=20, num_gen_from_policy=5, max_reward_iterations=100,
                  max_forward_episodes=200, gamma=1, constant=False, lr_critic=0.1, lr_actor=0.001,
                  final_episodes=2000, verbose=True):
        """Alternate reward learning and the forward solve (ac_irl.py:900-954); returns theta."""
        self.list_generated = self.generate_trajectories(num_gen_from_policy * self.num_policies)
        self.reward_update_count = 0
        os.makedirs("results", exist_ok=True)
        with open("results/reward_training.csv", 'w') as f:
            f.write("reward_demo_avg,reward_gen_avg\n")
        for it in range(num_iterations):
            if verbose:
                print("########## Outerloop iteration %d ##########" % it)
            list_generated = self.generate_trajectories(num_gen_from_policy)
            self.list_generated = (self.list_generated + list_generated)[num_gen_from_policy:]
            self.list_eval_gen_transitions = [pair for traj in self.list_generated for pair in traj]
            self.reward_iteration(max_iterations=max_reward_iterations, stop_criteria=0.0001, iter_check=10,
                                  verbose=verbose)
            self.theta = self.theta_initial                                        # quirk C.7
            self.train(max_forward_episodes, -1, gamma, constant, lr_critic, lr_actor, consecutive=100,
                       file_theta='results/theta.csv', file_pi='results/pi.csv', file_reward='results/reward.csv',
                       write_file=1, write_all=0, verbose=verbose)
        if verbose:
            print("Saving network")
        self.save("log/model_%s_%d_%d.ckpt" % (self.reg, self.n_fc3, self.n_fc4))
        if verbose:
            print("********** Final forward training **********")
        self.theta = self.theta_initial
        self.train(final_episodes, -1, gamma, constant, lr_critic, lr_actor, consecutive=100,
                   file_theta='results/theta.csv', file_pi='results/pi.csv', file_reward='results/reward.csv',
                   write_file=1, write_all=0, verbose=verbose)
        return self.theta

    def evaluate(self, theta=8.86349, shift=0.5, alpha_scale=1e4, d=15, episode_length=16,
                 indir='test_normalized_round2', outfile='eval_mfg_round2/validation.csv', write_header=0,
                 empirical=None, y=None):
        """ac_irl.py:1495-1570: the same evaluation as mfg_ac2's (inherited implementation: all test days as one
        batched rollout + dmfg_traj_metrics) with this class's defaults (d = 15, validation.csv).  JSD,
        generate_trajectory and gridsearch (ac_irl.py:1445-1589) are inherited unchanged."""
        return super().evaluate(theta=theta, shift=shift, alpha_scale=alpha_scale, d=d, episode_length=episode_length,
                                indir=indir, outfile=outfile, write_header=write_header, empirical=empirical, y=y)

    def test_reward_network(self):
        """Average reward of the fixed reward network over all transitions of the training demonstrations, of the
        test demonstrations and of freshly generated trajectories (ac_irl.py:1008-1043; gridsearch.py:28)."""
        def avg(trajectories):
            if not trajectories:
                return float("nan")
            s, a = self._pack(trajectories)
            return float(self._reward(s, a).double().sum()) / s.shape[0]
        self.list_generated = self.generate_trajectories(len(self.list_demonstrations))
        reward_demo_avg_train = avg(self.list_demonstrations)
        reward_gen_avg = avg(self.list_generated)
        reward_demo_avg_test = avg(self.list_demonstrations_test)
        print("Avg reward demo train %f | Avg reward demo test %f | Avg reward gen %f"
              % (reward_demo_avg_train, reward_demo_avg_test, reward_gen_avg))
        return reward_demo_avg_train, reward_demo_avg_test, reward_gen_avg

    # ------------------------------------------------------------- a13: importance weights
    def calc_z(self, gen_states, gen_actions, layout="trajectory_major"):
        """z_j = K / (N_start sum_k q_k(tau_j)) for generated trajectories (ac_irl.py:324-379), returned as
        ln z_j (float32 device tensor [M]); log space replaces the reference's float64 + `c` normaliser."""
        thetas = torch.as_tensor(np.asarray(self.list_policies, dtype=np.float64), device=self.device)
        lq = engine.dirichlet_logq(gen_states, gen_actions, thetas, self.shift)
        return engine.irl_log_z(lq, T_STEPS, self.num_start_samples, layout=layout)

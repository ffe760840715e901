"""Drop-in for the reference's ``networks.py`` (networks.py:1-157): the reward-network family.

The reference builds TensorFlow-1 graph nodes.  Here ``r_net*`` keep their names and signatures and
return a ``RewardOutput`` -- a handle on (parameters, state input, action input, regularisation
variant) that is evaluated by the fused CUDA kernels (``dmfg_rnet_forward`` / ``dmfg_rnet_backward``)
when it is fetched through ``Session.run`` or called directly.  Inputs may be ``Placeholder`` objects
(graph style, as ``AC_IRL.create_network`` uses them, ac_irl.py:232-267) or concrete arrays / CUDA
tensors (eager).  Variables live in a ``variable_scope``: a second ``r_net*`` call in the same scope
shares the parameters of the first (``scope.reuse_variables()``, ac_irl.py:251).

Only the instantiation the reference uses is built on the GPU: f1=1, k1=5, f2=2, k2=3, n_fc3, n_fc4 <= 8, forward for
d <= 32, backward for d <= 16 and d = 20, 21; ``n_fc5`` is accepted and ignored exactly like upstream (fc5 is commented out,
networks.py:39,76,115,152).
"""
from __future__ import annotations

import contextlib
import math

import numpy as np
import torch

from . import engine
from .layers import linear_layer

KEEP_PROB = 0.4                       # networks.py:70,75,146,151

_scope_stack = []
_scopes = {}                          # scope name -> RewardParams


class Placeholder:
    """Stand-in for tf.placeholder (ac_irl.py:239-246): a named feed slot."""

    def __init__(self, name, shape=None):
        self.name, self.shape = name, shape

    def __repr__(self):
        return "Placeholder(%r, shape=%r)" % (self.name, self.shape)


class RewardParams:
    """The flat float32 parameter vector of one reward net on the device (TF variable order, see
    include/dmfg.h) plus its Adam moments."""

    NAMES = ("conv1/weights", "conv1/biases", "conv2/weights", "conv2/biases", "fc3/weights", "fc3/biases",
             "fc4/weights", "fc4/biases", "out/weights", "out/biases")

    def __init__(self, d, n_fc3, n_fc4, device, seed=None):
        self.d, self.n_fc3, self.n_fc4 = int(d), int(n_fc3), int(n_fc4)
        self.device = torch.device(device)
        self.count = engine.rnet_param_count(d, n_fc3, n_fc4)
        self.offsets = engine.rnet_param_offsets(d, n_fc3, n_fc4)
        self.shapes = [(5, 5, 1, 1), (1,), (3, 3, 1, 2), (2,), (2 * d * d, n_fc3), (n_fc3,),
                       (n_fc3 + d, n_fc4), (n_fc4,), (n_fc4, 1), (1,)]
        self.flat = torch.zeros(self.count, dtype=torch.float32, device=self.device)
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.step = 0
        self.initialize(seed)

    def initialize(self, seed=None):
        """tf.contrib.layers defaults: Xavier-uniform weights, zero biases; resets Adam."""
        rng = np.random.RandomState(seed)
        p = np.zeros(self.count, dtype=np.float32)
        for name, shp, off in zip(self.NAMES, self.shapes, self.offsets):
            if name.endswith("biases"):
                continue
            if len(shp) == 4:
                rf = shp[0] * shp[1]
                fan_in, fan_out = rf * shp[2], rf * shp[3]
            else:
                fan_in, fan_out = shp
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            n = int(np.prod(shp))
            p[off:off + n] = rng.uniform(-lim, lim, size=n)
        self.load_flat(p)

    def load_flat(self, p):
        self.flat.copy_(torch.as_tensor(np.asarray(p, dtype=np.float32), device=self.device))
        self.m.zero_()
        self.v.zero_()
        self.step = 0

    def named(self):
        """{'reward-scope-relative name': numpy array} -- the tensors a tf.train.Saver would hold."""
        flat = self.flat.cpu().numpy()
        return {n: flat[o:o + int(np.prod(s))].reshape(s).copy() for n, s, o in zip(self.NAMES, self.shapes, self.offsets)}

    def load_named(self, tensors):
        flat = self.flat.cpu().numpy()
        for n, s, o in zip(self.NAMES, self.shapes, self.offsets):
            if n in tensors:
                flat[o:o + int(np.prod(s))] = np.asarray(tensors[n], dtype=np.float32).reshape(-1)
        self.flat.copy_(torch.as_tensor(flat, device=self.device))


class RewardOutput:
    """Graph-style handle on r_net(state_input, action_input); evaluate with ``Session.run`` or ``__call__``."""

    def __init__(self, params, state_input, action_input, dropout, l1l2):
        self.params, self.state_input, self.action_input = params, state_input, action_input
        self.dropout, self.l1l2 = dropout, l1l2

    def __call__(self, states=None, actions=None, seed=None, sample_offset=0, mask3=None, mask4=None):
        """[N,1] rewards on the device.  ``seed`` keys the in-kernel dropout masks of the dropout variants
        (always active, like upstream: is_training defaults to True, networks.py:70)."""
        states = self.state_input if states is None else states
        actions = self.action_input if actions is None else actions
        p = self.params
        s = _as_device(states, p.device).reshape(-1, p.d)
        a = _as_device(actions, p.device).reshape(-1, p.d, p.d)
        kw = {}
        if self.dropout:
            if mask3 is not None:
                kw = dict(mask3=mask3, mask4=mask4)
            else:
                kw = dict(seed=0 if seed is None else seed, sample_offset=sample_offset)
        return engine.rnet_forward(p.flat, s, a, p.n_fc3, p.n_fc4, keep_prob=KEEP_PROB, **kw).reshape(-1, 1)


def _as_device(x, device):
    if isinstance(x, Placeholder):
        raise ValueError("placeholder %r has not been fed" % x.name)
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.float32).contiguous()
    return torch.as_tensor(np.ascontiguousarray(np.asarray(x, dtype=np.float32)), device=device)


@contextlib.contextmanager
def variable_scope(name, device=None, seed=None):
    """``with variable_scope("reward"):`` -- r_net* calls inside share one RewardParams."""
    _scope_stack.append((name, device, seed))
    try:
        yield name
    finally:
        _scope_stack.pop()


def reset_default_graph():
    """tf.reset_default_graph() (gridsearch.py:21): forget every scope's parameters."""
    _scopes.clear()


def scope_params(name):
    return _scopes[name]


def _build(state_input, action_input, f1, k1, f2, k2, n_fc3, n_fc4, d, dropout, l1l2):
    if (f1, k1, f2, k2) != (1, 5, 2, 3):
        raise NotImplementedError("the CUDA reward net is built for f1=1, k1=5, f2=2, k2=3 (the only "
                                  "instantiation the reference uses, ac_irl.py:249-267)")
    name, device, seed = _scope_stack[-1] if _scope_stack else ("", None, None)
    if device is None:
        device = "cuda:%d" % torch.cuda.current_device()
    params = _scopes.get(name)
    if params is None or (params.d, params.n_fc3, params.n_fc4) != (d, n_fc3, n_fc4):
        params = RewardParams(d, n_fc3, n_fc4, device, seed)
        _scopes[name] = params
    return RewardOutput(params, state_input, action_input, dropout, l1l2)


def hidden2(vec_input, n_hidden1, n_hidden2, n_outputs, nonlinearity1, nonlinearity2):
    """Generic two-hidden-layer MLP (networks.py:4-10).  Dead code upstream -- API surface only."""
    h1 = linear_layer(vec_input, n_hidden1, nonlinearity1, scope='fc1')
    h2 = linear_layer(h1, n_hidden2, nonlinearity2, scope='fc2')
    return linear_layer(h2, n_outputs, nonlinearity=None, scope='out')


def r_net(state_input, action_input, f1=1, k1=5, f2=2, k2=3, n_fc3=8, n_fc4=4, n_fc5=4, d=15):
    """networks.py:13-43."""
    return _build(state_input, action_input, f1, k1, f2, k2, n_fc3, n_fc4, d, dropout=False, l1l2=False)


def r_net_dropout_l1l2(state_input, action_input, f1=1, k1=5, f2=2, k2=3, n_fc3=8, n_fc4=4, n_fc5=4, d=15):
    """networks.py:46-81: dropout(keep 0.4) after fc3 and fc4, l1_l2 regulariser on their weights."""
    return _build(state_input, action_input, f1, k1, f2, k2, n_fc3, n_fc4, d, dropout=True, l1l2=True)


def r_net_l1l2(state_input, action_input, f1=1, k1=5, f2=2, k2=3, n_fc3=4, n_fc4=4, n_fc5=4, d=15):
    """networks.py:84-119."""
    return _build(state_input, action_input, f1, k1, f2, k2, n_fc3, n_fc4, d, dropout=False, l1l2=True)


def r_net_dropout(state_input, action_input, f1=1, k1=5, f2=2, k2=3, n_fc3=8, n_fc4=4, n_fc5=4, d=15):
    """networks.py:122-157."""
    return _build(state_input, action_input, f1, k1, f2, k2, n_fc3, n_fc4, d, dropout=True, l1l2=False)

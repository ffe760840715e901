"""Drop-in for the reference's ``layers.py`` (layers.py:4-11).

``linear_layer`` is dead code upstream (never called); it is kept as API surface only.  There is no
TensorFlow graph here: a layer is evaluated eagerly on CUDA tensors, with its variables kept in a
small scope-keyed store so that a second call under the same scope reuses them (the behaviour
``tf.variable_scope`` + ``fully_connected`` has).  No reference behaviour is pinned for it.
"""
from __future__ import annotations

import math

import torch

_variables = {}          # "scope/weights" | "scope/biases" -> CUDA tensor


def get_variables(prefix=""):
    """Variables created so far whose name starts with ``prefix`` (name -> tensor)."""
    return {k: v for k, v in _variables.items() if k.startswith(prefix)}


def reset_variables():
    _variables.clear()


def _xavier_uniform(fan_in, fan_out, device, generator=None):
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand((fan_in, fan_out), device=device, generator=generator) * 2 - 1) * lim


def linear_layer(vec_input, num_nodes, nonlinearity, scope):
    """h = nonlinearity(vec_input @ W + b) with W Xavier-uniform, b zero (tf.contrib ``fully_connected``
    defaults), variables named ``<scope>/weights`` and ``<scope>/biases`` (layers.py:4-11)."""
    if not isinstance(vec_input, torch.Tensor) or not vec_input.is_cuda:
        raise TypeError("linear_layer evaluates eagerly on CUDA tensors (there is no CPU path)")
    if nonlinearity is None:
        nonlinearity = lambda x: x                                   # tf.identity (layers.py:5-6)
    wname, bname = scope + "/weights", scope + "/biases"
    if wname not in _variables:
        _variables[wname] = _xavier_uniform(vec_input.shape[-1], num_nodes, vec_input.device).to(vec_input.dtype)
        _variables[bname] = torch.zeros(num_nodes, device=vec_input.device, dtype=vec_input.dtype)
    return nonlinearity(vec_input @ _variables[wname] + _variables[bname])

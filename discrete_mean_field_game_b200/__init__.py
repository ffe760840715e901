"""discrete_mean_field_game_b200 -- B200-native hot path of the Deep Mean Field Games code.

Drop-in surface (same names as the reference's modules):
    mfg_ac2.actor_critic     forward actor-critic, closed-form reward
    mfg_synthetic.actor_critic   same solver with the synthetic reward + the (shift, theta0) sweep
    ac_irl.AC_IRL            MaxEnt IRL + actor-critic
    networks / layers        reward-net definitions
underneath: hand-written sm_100a kernels behind the C ABI of include/dmfg.h
(libdmfg.so, loaded with ctypes).  There is no CPU fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib", "engine", "mfg_ac2", "mfg_synthetic", "ac_irl", "parallel", "networks", "layers"]
__version__ = "0.1.0"

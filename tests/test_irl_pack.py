"""Host-side packing of reward minibatches (AC_IRL.update_reward, ac_irl.py:804-846): trajectories -> trajectory-major
arrays, stacked once and remembered, both halves through one buffer.  CPU only."""
import numpy as np
import torch

from discrete_mean_field_game_b200.ac_irl import AC_IRL


def _obj(d):
    obj = AC_IRL.__new__(AC_IRL)          # no device, no network: the packing helpers only need d and _dev
    obj.d = d
    obj._dev = lambda x, dt=None: torch.as_tensor(np.ascontiguousarray(x), dtype=dt)
    return obj


def _ref(trajs, k):
    return np.asarray([p[k] for t in trajs for p in t], dtype=np.float32)


def test_pack_pair_matches_the_reference_layout_and_is_aligned():
    d = 15
    obj = _obj(d)
    rng = np.random.RandomState(0)
    trajs = [[(rng.rand(d), rng.rand(d, d)) for _ in range(15)] for _ in range(10)]
    ds, da, gs, ga = obj._pack_pair(trajs[:5], trajs[5:])
    assert np.array_equal(ds.numpy(), _ref(trajs[:5], 0)) and np.array_equal(da.numpy(), _ref(trajs[:5], 1))
    assert np.array_equal(gs.numpy(), _ref(trajs[5:], 0)) and np.array_equal(ga.numpy(), _ref(trajs[5:], 1))
    assert ds.shape == (75, d) and ga.shape == (75, d, d)
    for t in (ds, da, gs, ga):
        assert t.is_contiguous() and t.data_ptr() % 16 == 0
    # a second call (memo hits) gives the same tensors; _pack agrees with _pack_pair
    ds2, da2, gs2, ga2 = obj._pack_pair(trajs[:5], trajs[5:])
    assert torch.equal(ds, ds2) and torch.equal(ga, ga2)
    s, a = obj._pack(trajs[5:])
    assert torch.equal(s, gs) and torch.equal(a, ga)


def test_pack_memo_follows_edits_of_a_trajectory_list_and_empty_input():
    obj = _obj(3)
    rng = np.random.RandomState(1)
    traj = [(rng.rand(3), rng.rand(3, 3)) for _ in range(15)]
    s0, _ = obj._pack([traj])
    traj[0] = (np.zeros(3), np.zeros((3, 3)))                     # first pair replaced: the memo entry is stale
    s1, a1 = obj._pack([traj])
    assert np.all(s1.numpy()[0] == 0) and not np.all(s0.numpy()[0] == 0)
    traj.append((np.ones(3), np.ones((3, 3))))                    # longer now
    s2, _ = obj._pack([traj])
    assert s2.shape == (16, 3) and np.all(s2.numpy()[-1] == 1)
    se, ae = obj._pack([])
    assert se.shape == (0, 3) and ae.shape == (0, 3, 3)
    ds, da, gs, ga = obj._pack_pair([], [traj])
    assert ds.shape == (0, 3) and gs.shape == (16, 3)

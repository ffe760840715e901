"""networks.py / layers.py drop-in surface (networks.py:4-157, layers.py:4-11) on the GPU."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import rnet_oracle as R


@pytest.fixture(scope="module")
def dev():
    from discrete_mean_field_game_b200 import engine
    engine.require_cuda()
    return torch.device("cuda:0")


def test_r_net_variants_share_variables_in_a_scope(dev):
    from discrete_mean_field_game_b200 import networks
    networks.reset_default_graph()
    rng = np.random.RandomState(0)
    s = np.float32(rng.dirichlet(np.ones(15), size=9))
    a = np.float32(rng.dirichlet(np.ones(15), size=(9, 15)))
    sp, ap = networks.Placeholder("states", [None, 15]), networks.Placeholder("actions", [None, 15, 15])
    with networks.variable_scope("reward", device=dev, seed=4):
        demo = networks.r_net(sp, ap, n_fc3=6, n_fc4=8, d=15)
        gen = networks.r_net(s, a, n_fc3=6, n_fc4=8, d=15)            # eager inputs, same scope
    assert demo.params is gen.params and networks.scope_params("reward") is demo.params
    assert demo.params.count == R.param_count(15, 6, 8)
    names = demo.params.named()
    assert set(names) == {"conv1/weights", "conv1/biases", "conv2/weights", "conv2/biases", "fc3/weights",
                          "fc3/biases", "fc4/weights", "fc4/biases", "out/weights", "out/biases"}
    assert names["fc3/weights"].shape == (450, 6) and np.all(names["fc4/biases"] == 0)
    r = gen().cpu().numpy()
    assert r.shape == (9, 1)
    ref = R.forward(demo.params.flat.cpu().numpy(), s, a, 6, 8)
    np.testing.assert_allclose(r[:, 0], ref, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(demo(s, a).cpu().numpy(), r, rtol=0, atol=0)
    with pytest.raises(ValueError):
        demo()                                                        # placeholder not fed
    with pytest.raises(NotImplementedError):
        networks.r_net(s, a, f1=2, d=15)
    # defaults of the four variants (networks.py:13,46,84,122)
    import inspect
    d3 = {f.__name__: inspect.signature(f).parameters["n_fc3"].default
          for f in (networks.r_net, networks.r_net_dropout_l1l2, networks.r_net_l1l2, networks.r_net_dropout)}
    assert d3 == {"r_net": 8, "r_net_dropout_l1l2": 8, "r_net_l1l2": 4, "r_net_dropout": 8}


def test_dropout_variants_flag_and_masks(dev):
    from discrete_mean_field_game_b200 import networks
    networks.reset_default_graph()
    rng = np.random.RandomState(1)
    s = np.float32(rng.dirichlet(np.ones(15), size=33))
    a = np.float32(rng.dirichlet(np.ones(15), size=(33, 15)))
    with networks.variable_scope("reward", device=dev, seed=1):
        net = networks.r_net_dropout_l1l2(s, a, d=15)
    assert net.dropout and net.l1l2
    net.params.load_flat(net.params.flat.cpu().numpy() + 0.2)
    m3 = torch.as_tensor(rng.rand(33, 8) < 0.4, dtype=torch.uint8, device=dev)
    m4 = torch.as_tensor(rng.rand(33, 4) < 0.4, dtype=torch.uint8, device=dev)
    r = net(mask3=m3, mask4=m4).cpu().numpy()[:, 0]
    ref = R.forward(net.params.flat.cpu().numpy(), s, a, 8, 4, m3.cpu().numpy(), m4.cpu().numpy())
    np.testing.assert_allclose(r, ref, rtol=1e-5, atol=1e-6)
    assert not np.array_equal(net(seed=1).cpu().numpy(), net(seed=2).cpu().numpy())


def test_hidden2_and_linear_layer_surface(dev):
    """Dead code upstream (never called): API surface only -- shapes, variable names, reuse."""
    from discrete_mean_field_game_b200 import layers, networks
    layers.reset_variables()
    x = torch.rand(5, 7, device=dev)
    out = networks.hidden2(x, 16, 8, 3, torch.relu, torch.tanh)
    assert out.shape == (5, 3)
    v = layers.get_variables()
    assert set(v) == {"fc1/weights", "fc1/biases", "fc2/weights", "fc2/biases", "out/weights", "out/biases"}
    assert v["fc1/weights"].shape == (7, 16) and float(v["fc2/biases"].abs().sum()) == 0.0
    again = networks.hidden2(x, 16, 8, 3, torch.relu, torch.tanh)
    assert torch.equal(out, again)                                   # variables are reused
    h = layers.linear_layer(x, 16, None, scope="fc1")
    assert torch.allclose(h, x @ v["fc1/weights"] + v["fc1/biases"])
    with pytest.raises(TypeError):
        layers.linear_layer(x.cpu(), 4, None, scope="cpu")

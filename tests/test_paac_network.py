"""CPU tests of the PyTorch policy/value net, mirroring the reference's tests/estimators_tests.py:8-130
(ConvSingleAgentTest.policy_predict_test / train_test) -- shapes and "both losses go to ~0"."""
import numpy as np
import torch


def make(pkg, num_actions, **kw):
    pv = pkg.submodule("agents.paac.policy_v_network")
    conf = {"name": "test_conv_network", "num_actions": num_actions, "clip_norm": 40.0, "clip_norm_type": "global",
            "device": "cpu", "static_size": None, "n_agents": 10, "entropy_regularisation_strength": 0.0, "scale": 1.0,
            "height": 84, "width": 84, "channels": 3, "filters": 5, "conv_layers": 2}
    conf.update(kw)
    return pv.ConvSingleAgentPolicyNetwork(conf)


def test_policy_predict_shapes(pkg):
    torch.manual_seed(0)
    n_agents, num_actions = 10, 3
    net = make(pkg, num_actions)
    state = torch.rand(n_agents, 84, 84, 3)
    actions = torch.rand(n_agents, num_actions)
    out = net.losses(state, actions, torch.ones(n_agents), torch.zeros(n_agents))
    assert out["mu"].shape == (n_agents, num_actions) and out["sigma"].shape == (n_agents, num_actions)
    assert out["policy_loss"].shape == () and out["vs"].shape == (n_agents,) and out["critic_loss"].shape == (n_agents,)
    assert (out["vs"] <= 0).all() and (out["sigma"] > 0).all() and (out["mu"].abs() <= 1).all()
    pred = net.predict(state)
    assert set(pred) == {"vs", "mu", "sigma"}


def test_parameter_count_matches_reference_architecture(pkg):
    net = make(pkg, 2, scale=1000.0)
    assert sum(p.numel() for p in net.parameters()) == 2210213      # SURVEY.md section 2 (layer shapes of the reference)


def test_losses_train_to_zero(pkg):
    """estimators_tests.py:78-129: 100 RMSProp steps drive the critic and policy losses to ~0 (1 decimal)."""
    torch.manual_seed(1692)
    rs = np.random.RandomState(1692)
    n_batch, num_actions = 10, 3
    net = make(pkg, num_actions)
    actions = torch.as_tensor(rs.uniform(size=(n_batch, num_actions)), dtype=torch.float32)
    state = torch.as_tensor(rs.uniform(0.0, 1.0, (n_batch, 84, 84, 3)), dtype=torch.float32)
    opt = torch.optim.RMSprop(net.parameters(), lr=0.02, alpha=0.99, eps=0.1)
    for idx in range(100):
        out = net.losses(state, actions, torch.ones(n_batch) / (idx + 1),
                         torch.as_tensor(-0.5 * rs.uniform(size=(n_batch,)), dtype=torch.float32))
        opt.zero_grad()
        out["loss"].backward()
        opt.step()
    assert abs(float(out["critic_loss_mean"])) < 0.05
    assert abs(float(out["policy_loss"])) < 0.05


def test_gaussian_terms_match_torch_distributions(pkg):
    torch.manual_seed(3)
    net = make(pkg, 2, entropy_regularisation_strength=0.02)
    state, actions = torch.rand(4, 84, 84, 3), torch.randn(4, 2)
    adv, tgt = torch.randn(4), -torch.rand(4)
    out = net.losses(state, actions, adv, tgt)
    d = torch.distributions.Normal(out["mu"], out["sigma"])
    ref = -(d.log_prob(actions).sum(-1) * adv + 0.02 * d.entropy().sum(-1)).mean()
    assert torch.allclose(out["policy_loss"], ref, atol=1e-5)


def test_compact_observation_forward_equals_expanded(pkg):
    """conv1 factorisation: the net on (grid, positions) == the net on get_local_states(grid, positions), values and
    gradients, including hot cells on the borders and in the last (clamped) row/column."""
    torch.manual_seed(11)
    net = make(pkg, 2, scale=1000.0, entropy_regularisation_strength=0.02).double()
    E, A, G = 3, 10, 84
    grid = torch.rand(E, G, G, 2, dtype=torch.float64) * (torch.rand(E, G, G, 2) > 0.98)
    pos = torch.randint(0, G, (E, A, 2), dtype=torch.uint8)
    pos[0, 0] = torch.tensor([0, 0]); pos[0, 1] = torch.tensor([83, 83]); pos[0, 2] = torch.tensor([3, 80])
    pos[0, 3] = torch.tensor([4, 7]); pos[0, 4] = torch.tensor([79, 0])
    expanded = torch.zeros(E, A, G, G, 3, dtype=torch.float64)
    expanded[..., :2] = grid[:, None]
    for e in range(E):
        for a in range(A):
            expanded[e, a, int(pos[e, a, 0]), int(pos[e, a, 1]), 2] = 1.0
    act, adv, tgt = torch.randn(E * A, 2, dtype=torch.float64), torch.randn(E * A, dtype=torch.float64), -torch.rand(E * A, dtype=torch.float64)
    o1 = net.losses(expanded.view(E * A, G, G, 3), act, adv, tgt)
    g1 = torch.autograd.grad(o1["loss"], list(net.parameters()))
    o2 = net.losses(grid, act, adv, tgt, positions=pos)
    g2 = torch.autograd.grad(o2["loss"], list(net.parameters()))
    for k in ("mu", "sigma", "vs"):
        assert torch.allclose(o1[k], o2[k], rtol=1e-9, atol=1e-9), k
    assert torch.allclose(o1["loss"], o2["loss"], rtol=1e-9)
    for a_, b_ in zip(g1, g2):
        assert torch.allclose(a_, b_, rtol=1e-7, atol=1e-10)

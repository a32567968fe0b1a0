"""PyTorch/cuDNN mirror of fed_gym/agents/paac/policy_v_network.py:5-80 (ConvSingleAgentPolicyNetwork)
and of the Network / ConvSingleAgentNetwork configuration classes (networks.py:100-167).

The policy conv net is the only dense contraction of the PAAC loop and stays in PyTorch (BASELINE.json
north_star); everything around it (env step, observation, action clip) is the CUDA hot path.

Layer for layer like the reference (TF defaults: VALID padding, glorot-uniform kernels, zero biases):
  process_input: conv 32@8x8 s4 relu -> conv 64@4x4 s2 relu -> conv 64@3x3 relu -> flatten (H,W,C order)
                 -> dense 512 relu -> dense 256 relu
  policy:        dense 512 relu -> mu = tanh(dense num_actions), sigma = sigmoid(dense num_actions)
  v_s:           dense 512 relu -> dense 256 relu -> vs = -scale * softplus(dense 1)
  loss:          policy_loss = -mean(log N(a; mu, sigma) * adv + beta * entropy)      (sums over action dims)
                 critic_loss = (vs - target)^2 / scale ;  loss = policy_loss + mean(0.25 * critic_loss)
2,210,213 parameters at num_actions = 2.

Input layout: (B, H, W, C) float32 exactly as the reference placeholder (networks.py:165-167) -- i.e. the
NHWC observation the rasteriser writes; it is viewed as channels_last NCHW for cuDNN without a copy.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def _glorot_(linear_or_conv):
    nn.init.xavier_uniform_(linear_or_conv.weight)
    nn.init.zeros_(linear_or_conv.bias)
    return linear_or_conv


class ConvSingleAgentPolicyNetwork(nn.Module):
    def __init__(self, conf):
        super().__init__()
        self.name = conf.get("name", "local_learning")
        self.num_actions = conf["num_actions"]
        self.clip_norm = conf.get("clip_norm", 40.0)
        self.clip_norm_type = conf.get("clip_norm_type", "global")
        self.entropy_beta = conf.get("entropy_regularisation_strength", 0.02)
        self.scale = float(conf.get("scale", 1000.0))
        self.height, self.width, self.channels = conf["height"], conf["width"], conf["channels"]
        self.conf = conf
        self.fc_hidden = 256

        self.conv1 = _glorot_(nn.Conv2d(self.channels, 32, kernel_size=8, stride=4))
        self.conv2 = _glorot_(nn.Conv2d(32, 64, kernel_size=4, stride=2))
        self.conv3 = _glorot_(nn.Conv2d(64, 64, kernel_size=3))

        def out(n, k, s):
            return (n - k) // s + 1
        h = out(out(out(self.height, 8, 4), 4, 2), 3, 1)
        w = out(out(out(self.width, 8, 4), 4, 2), 3, 1)
        self.dense1 = _glorot_(nn.Linear(h * w * 64, 2 * self.fc_hidden))
        self.dense2 = _glorot_(nn.Linear(2 * self.fc_hidden, self.fc_hidden))
        self.policy_hidden = _glorot_(nn.Linear(self.fc_hidden, 2 * self.fc_hidden))
        self.mu_head = _glorot_(nn.Linear(2 * self.fc_hidden, self.num_actions))
        self.sigma_head = _glorot_(nn.Linear(2 * self.fc_hidden, self.num_actions))
        self.v1 = _glorot_(nn.Linear(self.fc_hidden, 2 * self.fc_hidden))
        self.v2 = _glorot_(nn.Linear(2 * self.fc_hidden, self.fc_hidden))
        self.v3 = _glorot_(nn.Linear(self.fc_hidden, 1))

    def forward(self, states):
        """states (B,H,W,C) float32 -> mu (B,num_actions), sigma (B,num_actions), vs (B,)."""
        x = states.permute(0, 3, 1, 2)                  # NHWC storage seen as channels_last NCHW: no copy
        x = F.relu(self.conv1(x))
        return self._trunk_and_heads(x)

    def forward_compact(self, grid, positions):
        """The same function of the COMPACT observation: grid (E,H,W,C-1) float32 (the channels every agent of an
        env shares) + positions (E,A,2) uint8 (each agent's one-hot cell in the last channel,
        emulator_runner.py:98-111).  conv1 is linear, so conv1(expanded obs of agent a) =
        conv1_{channels<C-1}(grid) [once per env] + the <= 4 kernel taps of the last channel that see the hot cell.
        Saves 10x of conv1 and never materialises the (E,A,H,W,C) observation.  -> mu, sigma (E*A,.), vs (E*A,)."""
        E, A = positions.shape[0], positions.shape[1]
        C = self.channels
        w = self.conv1.weight                                           # (32, C, 8, 8)
        k, s = self.conv1.kernel_size[0], self.conv1.stride[0]
        base = F.conv2d(grid.permute(0, 3, 1, 2), w[:, :C - 1], self.conv1.bias, stride=s)     # (E,32,oh,ow)
        OH, OW, F1 = base.shape[2], base.shape[3], base.shape[1]
        # per-agent maps in NHWC memory (what cuDNN's tensor-core kernels want): ONE materialising copy
        x = base.permute(0, 2, 3, 1).reshape(E, 1, OH * OW, F1).expand(E, A, OH * OW, F1).reshape(E * A, OH * OW, F1)
        pos = positions.reshape(E * A, 2).long()
        h, wd = pos[:, 0], pos[:, 1]
        w_hot = w[:, C - 1].reshape(F1, k * k)                          # (32, 64) taps of the one-hot channel
        for dh in range((k + s - 1) // s):
            oh = torch.div(h, s, rounding_mode="floor") - dh
            kh = h - s * oh
            for dw in range((k + s - 1) // s):
                ow = torch.div(wd, s, rounding_mode="floor") - dw
                kw = wd - s * ow
                ok = (oh >= 0) & (oh < OH) & (ow >= 0) & (ow < OW) & (kh < k) & (kw < k)
                tap = w_hot[:, (kh.clamp(max=k - 1) * k + kw.clamp(max=k - 1))].t() * ok[:, None].to(w_hot.dtype)   # (B,32)
                idx = (oh.clamp(0, OH - 1) * OW + ow.clamp(0, OW - 1))[:, None, None].expand(-1, 1, F1)
                x.scatter_add_(1, idx, tap[:, None, :].to(x.dtype))
        x = F.relu_(x).view(E * A, OH, OW, F1).permute(0, 3, 1, 2)       # channels_last (B,32,oh,ow), no copy
        return self._trunk_and_heads(x)

    def _trunk_and_heads(self, x):
        x = F.relu(self.conv2(x))
        x = F.relu(self.conv3(x))
        x = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)      # tf.layers.flatten of an NHWC tensor
        h = F.relu(self.dense2(F.relu(self.dense1(x))))
        a = F.relu(self.policy_hidden(h))
        mu = torch.tanh(self.mu_head(a))
        sigma = torch.sigmoid(self.sigma_head(a))
        v = F.relu(self.v2(F.relu(self.v1(h))))
        vs = -self.scale * F.softplus(self.v3(v)).squeeze(-1)
        return mu, sigma, vs

    @torch.no_grad()
    def predict(self, states, session=None):
        """policy_v_network.py:69-80 ('session' kept for signature compatibility, unused)."""
        mu, sigma, vs = self.forward(states)
        return {"vs": vs, "mu": mu, "sigma": sigma}

    @torch.no_grad()
    def predict_compact(self, grid, positions):
        mu, sigma, vs = self.forward_compact(grid, positions)
        return {"vs": vs, "mu": mu, "sigma": sigma}

    def losses(self, states, actions, advantages, critic_target, positions=None):
        """-> dict(loss, policy_loss, critic_loss (B,), critic_loss_mean, mu, sigma, vs, entropy).
        With ``positions`` given, ``states`` is the compact (E,H,W,C-1) grid (see forward_compact)."""
        mu, sigma, vs = self.forward(states) if positions is None else self.forward_compact(states, positions)
        var = sigma * sigma
        log_l = -((actions - mu) ** 2) / (2.0 * var) - torch.log(sigma) - 0.5 * math.log(2.0 * math.pi)
        entropy = 0.5 + 0.5 * math.log(2.0 * math.pi) + torch.log(sigma)
        if self.num_actions > 1:
            log_l = log_l.sum(-1)
            entropy = entropy.sum(-1)
        else:
            log_l, entropy = log_l.squeeze(-1), entropy.squeeze(-1)
        policy_loss = -(log_l * advantages + self.entropy_beta * entropy).mean()
        critic_loss = (vs - critic_target) ** 2 / self.scale
        critic_loss_mean = (0.25 * critic_loss).mean()
        return {"loss": policy_loss + critic_loss_mean, "policy_loss": policy_loss, "critic_loss": critic_loss,
                "critic_loss_mean": critic_loss_mean, "mu": mu, "sigma": sigma, "vs": vs, "entropy": entropy}

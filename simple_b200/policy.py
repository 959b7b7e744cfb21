"""Minimal actor-critic used by the rollout benchmark and the multi-GPU demo.

The PPO policy is OUTSIDE the hot path (~33 MFLOP/sample, 350x smaller than the world model) and
stays plain PyTorch, exactly like in the reference (a2c_ppo_acktr/a2c_ppo_acktr/policy.py:105-133): three
convolutions (12->32 k8 s4, 32->64 k4 s2, 64->32 k3 s1), FC 512, a value head and a categorical action head.
It exists so that `SimulatedVecEnv` rollouts can be driven end to end without the reference's PPO stack; any object
with the reference's `act(obs) -> (value, action, log_prob)` can be used instead.
"""
import torch
import torch.nn as nn


class CnnPolicy(nn.Module):
    def __init__(self, obs_shape, n_action, hidden_size=512):
        super().__init__()
        c, h, w = obs_shape
        h, w = h // 4 - 1, w // 4 - 1
        h, w = h // 2 - 1, w // 2 - 1
        h, w = h - 2, w - 2
        self.main = nn.Sequential(
            nn.Conv2d(c, 32, 8, stride=4), nn.ReLU(),
            nn.Conv2d(32, 64, 4, stride=2), nn.ReLU(),
            nn.Conv2d(64, 32, 3, stride=1), nn.ReLU(), nn.Flatten(),
            nn.Linear(32 * h * w, hidden_size), nn.ReLU())
        self.critic_linear = nn.Linear(hidden_size, 1)
        self.dist = nn.Linear(hidden_size, n_action)
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.orthogonal_(m.weight, nn.init.calculate_gain('relu'))
                nn.init.zeros_(m.bias)

    def evaluate_actions(self, obs, action):
        """(value, log pi(action|obs), mean entropy) -- a2c_ppo_acktr/policy.py:120-133."""
        x = self.main(obs / 255.0)
        logp_all = torch.log_softmax(self.dist(x), dim=-1)
        entropy = -(logp_all.exp() * logp_all).sum(-1).mean()
        return self.critic_linear(x), logp_all.gather(1, action), entropy

    @torch.no_grad()
    def get_value(self, obs):
        return self.critic_linear(self.main(obs / 255.0))

    @torch.no_grad()
    def act(self, obs):
        x = self.main(obs / 255.0)
        logits = self.dist(x)
        # categorical sample by the Gumbel trick (no host sync, CUDA-graph friendly)
        g = -torch.log(-torch.log(torch.rand_like(logits).clamp_(1e-20, 1.0)))
        action = torch.argmax(logits + g, dim=-1, keepdim=True)
        logp = torch.log_softmax(logits, dim=-1).gather(1, action)
        return self.critic_linear(x), action, logp

"""Device-resident PPO driver for the simulated environment.

Mirrors `atari_utils/atari_utils/ppo_wrapper.py:16-120` (`PPO.learn`), `a2c_ppo_acktr/rollout_storage.py:54-62` (GAE)
and `a2c_ppo_acktr/ppo.py:33-85` (clipped-surrogate update, same hyper-parameters) in plain PyTorch -- the policy is
outside the kernel scope and stays a PyTorch module like in the reference -- but without the reference's host round
trips: observations are kept as u8 on the device (the env's frames ARE u8), the rollout storage, GAE and
the minibatch sampler live on the device, and losses are read back once per `learn` call.  With torch.distributed
initialised, the policy gradients are all-reduced (data-parallel PPO over agent shards: every rank simulates its own
agents) so that the replicas stay identical.
"""
import torch
import torch.distributed as dist


class DevicePPO:
    def __init__(self, env, policy, num_steps=None, gamma=0.99, lr=2.5e-4, clip_param=0.1, value_loss_coef=0.5,
                 num_mini_batch=4, entropy_coef=0.01, eps=1e-5, max_grad_norm=0.5, ppo_epoch=4, gae_lambda=0.95):
        self.env, self.policy = env, policy
        self.num_steps = num_steps or env.config.rollout_length
        self.gamma, self.gae_lambda = gamma, gae_lambda
        self.clip_param, self.value_loss_coef, self.entropy_coef = clip_param, value_loss_coef, entropy_coef
        self.num_mini_batch, self.ppo_epoch, self.max_grad_norm = num_mini_batch, ppo_epoch, max_grad_norm
        self.optimizer = torch.optim.Adam(policy.parameters(), lr=lr, eps=eps)
        self.world = dist.get_world_size() if dist.is_initialized() else 1

    @torch.no_grad()
    def collect(self):
        """One rollout of `num_steps` simulated steps for all agents.  Returns the storage dict (device tensors)."""
        env, T, A = self.env, self.num_steps, self.env.num_envs
        dev = env.device
        obs = env.reset()
        st = dict(obs=torch.empty((T + 1, A) + tuple(obs.shape[1:]), dtype=torch.uint8, device=dev),
                  actions=torch.empty(T, A, 1, dtype=torch.long, device=dev),
                  logp=torch.empty(T, A, 1, device=dev), values=torch.empty(T + 1, A, 1, device=dev),
                  rewards=torch.empty(T, A, 1, device=dev), masks=torch.ones(T + 1, A, 1, device=dev))
        st['obs'][0].copy_(obs)
        for t in range(T):
            value, action, logp = self.policy.act(obs)
            obs, reward, done, _ = env.step(action)
            st['actions'][t], st['logp'][t], st['values'][t] = action, logp, value
            st['rewards'][t] = reward.to(torch.float32).reshape(A, 1)
            st['masks'][t + 1] = torch.as_tensor(~done, device=dev, dtype=torch.float32).reshape(A, 1)
            st['obs'][t + 1].copy_(obs)
        st['values'][T] = self.policy.get_value(obs)
        # generalised advantage estimation (rollout_storage.py:54-62)
        returns = torch.empty(T + 1, A, 1, device=dev)
        gae = torch.zeros(A, 1, device=dev)
        for t in reversed(range(T)):
            delta = st['rewards'][t] + self.gamma * st['values'][t + 1] * st['masks'][t + 1] - st['values'][t]
            gae = delta + self.gamma * self.gae_lambda * st['masks'][t + 1] * gae
            returns[t] = gae + st['values'][t]
        st['returns'] = returns
        return st

    def update(self, st):
        """Clipped-surrogate PPO update (ppo.py:33-85).  Returns device scalars (value, action, entropy losses)."""
        T, A = st['actions'].shape[:2]
        adv = st['returns'][:-1] - st['values'][:-1]
        adv = (adv - adv.mean()) / (adv.std() + 1e-5)
        flat = lambda x: x.reshape(T * A, *x.shape[2:])
        obs, actions, values_old = flat(st['obs'][:-1]), flat(st['actions']), flat(st['values'][:-1])
        returns, logp_old, adv = flat(st['returns'][:-1]), flat(st['logp']), flat(adv)
        mb = (T * A) // self.num_mini_batch
        sums = torch.zeros(3, device=obs.device)
        params = [p for p in self.policy.parameters()]
        for _ in range(self.ppo_epoch):
            perm = torch.randperm(T * A, device=obs.device)
            for k in range(self.num_mini_batch):  # drop_last like the reference's BatchSampler
                idx = perm[k * mb:(k + 1) * mb]
                values, logp, entropy = self.policy.evaluate_actions(obs[idx].float(), actions[idx])
                ratio = torch.exp(logp - logp_old[idx])
                surr1, surr2 = ratio * adv[idx], torch.clamp(ratio, 1 - self.clip_param, 1 + self.clip_param) * adv[idx]
                action_loss = -torch.min(surr1, surr2).mean()
                v_clipped = values_old[idx] + (values - values_old[idx]).clamp(-self.clip_param, self.clip_param)
                value_loss = 0.5 * torch.max((values - returns[idx]).pow(2), (v_clipped - returns[idx]).pow(2)).mean()
                self.optimizer.zero_grad(set_to_none=True)
                (value_loss * self.value_loss_coef + action_loss - entropy * self.entropy_coef).backward()
                if self.world > 1:  # data-parallel PPO: average the (small) policy gradient over the agent shards
                    flat_g = torch.cat([p.grad.reshape(-1) for p in params])
                    dist.all_reduce(flat_g)
                    flat_g.mul_(1.0 / self.world)
                    off = 0
                    for p in params:
                        p.grad.copy_(flat_g[off:off + p.numel()].view_as(p))
                        off += p.numel()
                torch.nn.utils.clip_grad_norm_(params, self.max_grad_norm)
                self.optimizer.step()
                sums += torch.stack((value_loss.detach(), action_loss.detach(), entropy.detach()))
        return sums / (self.ppo_epoch * self.num_mini_batch)

    def learn(self, steps):
        """`steps` simulated env steps in total over all agents of this rank (ppo_wrapper.py:58-75)."""
        n_updates = max(1, steps // self.num_steps // self.env.num_envs)
        last = None
        for _ in range(n_updates):
            last = self.update(self.collect())
        v, a, e = last.tolist()  # the only host sync
        return {'ppo_value_loss': v, 'ppo_action_loss': a, 'ppo_entropy': e}

"""Device-resident simulated environment (drop-in for simple/subproc_vec_env.py:398-457 + simple/simulated_env.py).

`make_simulated_env(config, model, action_space)` returns a vectorised env with the reference's surface
(`reset`, `step`, `env_method('restart', frames, indices=j)`, `num_envs`, `observation_space`, `action_space`,
`close`).  The reference ships every frame stack through 16 echo subprocesses over pipes (D2H -> pickle -> H2D per
step); here the frame stack, the recurrent state and the rewards never leave the GPU: one step is
one-hot(actions) -> NextFramePredictor eval forward -> fused argmax decode -> u8 frame-stack shift.

Reference quirks kept on purpose: dropout stays active in eval; `done` is never True
(subproc_vec_env.py:433-440 computes it after the counter was reset); the value head is added to the reward of the
last step of a rollout; observations are the u8 round trip of the frame stack, returned as float 0..255.
"""
import numpy as np
import torch

from . import _lib, ops


class Box:
    def __init__(self, low, high, shape, dtype=np.uint8):
        self.shape, self.dtype = tuple(shape), np.dtype(dtype)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)


class Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)


class SimulatedEnv:
    """Per-agent view kept for surface compatibility (simple/simulated_env.py:9-64): holds the initial frames."""

    def __init__(self, vec, index):
        self._vec, self._index = vec, index
        self.main = index == 0

    @property
    def initial_frames(self):
        return self._vec._initial[self._index] if self._vec._has_initial[self._index] else None

    def restart(self, initial_frames):
        self._vec._set_initial(self._index, initial_frames)

    def get_initial_frames(self):
        return self._vec._initial[self._index]

    def reset(self):
        if not self._vec._has_initial[self._index]:
            raise ValueError('This environment has not been initialized. Call the restart function.')
        return self._vec._initial[self._index]

    def close(self):
        return None


class SimulatedVecEnv:

    def __init__(self, config, model, action_space, reference_types=False):
        self.config = config
        self.model = model
        self.action_space = action_space
        self.n_action = action_space.n
        self.num_envs = config.agents
        c, h, w = config.frame_shape
        self.stack_channels = c * int(config.stacking)
        self.observation_space = Box(0, 255, (self.stack_channels, h, w), np.uint8)
        self.device = torch.device(config.device)
        if self.device.type != 'cuda':
            raise RuntimeError('simple_b200 simulated env runs on CUDA only; there is no CPU fallback')
        self.reference_types = reference_types
        self._initial = torch.zeros(self.num_envs, self.stack_channels, h, w, dtype=torch.uint8, device=self.device)
        self._has_initial = [False] * self.num_envs
        self.envs = [SimulatedEnv(self, i) for i in range(self.num_envs)]
        self.step_count = 0
        self.frames = None  # u8 [A,12,H,W] on device (static buffer)
        self.use_graph = bool(getattr(config, 'cuda_graph', True))
        self.closed = False

    # ---- reference surface --------------------------------------------------------------------------------------
    def _set_initial(self, index, frames):
        self._initial[index].copy_(torch.as_tensor(frames).to(torch.uint8), non_blocking=True)
        self._has_initial[index] = True

    def restart_all(self, frames_u8):
        """Batched form of `env_method('restart', f, indices=j)` for all agents: u8 [A,12,H,W] (one device copy)."""
        self._initial.copy_(frames_u8.to(torch.uint8), non_blocking=True)
        self._has_initial = [True] * self.num_envs

    def restart_from_replay(self, obs_u8, valid=None, first_small_rollout=None, flip_first=True):
        """Restart path of the epoch loop (simple/__main__.py:85-91) as ONE device gather instead of `agents` pickled
        tensor sends: every agent draws a random record of the real replay (`sample_buffer(1)`: uniformly among the
        records that already carry a value, atari_utils/atari_utils/envs.py:158-172); with `flip_first`
        (`simulation_flip_first_random_for_beginning`) the LAST agent starts from `first_small_rollout`
        (`get_first_small_rollout()`: the first frames of the real episode) instead.
        obs_u8: u8 [N,12,H,W] replay observations (device or host); valid: optional bool [N]."""
        obs_u8 = torch.as_tensor(obs_u8)
        n = obs_u8.shape[0]
        if valid is None:
            cand = torch.arange(n, device=self.device)
        else:
            cand = torch.nonzero(torch.as_tensor(valid, device=self.device), as_tuple=False).flatten()
        if cand.numel() == 0:
            raise ValueError('the replay holds no record with a value yet')
        pick = cand[torch.randint(cand.numel(), (self.num_envs,), device=self.device)]
        frames = obs_u8.to(self.device)[pick].to(torch.uint8)
        if flip_first and first_small_rollout is not None:
            frames[-1] = torch.as_tensor(first_small_rollout).to(self.device, torch.uint8)
        self.restart_all(frames)
        return pick

    def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
        idx = range(self.num_envs) if indices is None else ([indices] if isinstance(indices, int) else indices)
        return [getattr(self.envs[i], method_name)(*method_args, **method_kwargs) for i in idx]

    def reset(self):
        if not all(self._has_initial):
            raise ValueError('This environment has not been initialized. Call the restart function.')
        return self._initial.float()

    def step_async(self, actions):
        self._pending = self._step(actions)

    def step_wait(self):
        out, self._pending = self._pending, None
        return out

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        self.closed = True

    # ---- one simulated step (subproc_vec_env.py:409-445) ----------------------------------------------------------
    def _alloc_static(self):
        A, dev = self.num_envs, self.device
        c, h, w = self.config.frame_shape
        self.frames = torch.zeros(A, self.stack_channels, h, w, dtype=torch.uint8, device=dev)
        self._act = torch.zeros(A, 1, dtype=torch.long, device=dev)
        self._obs = torch.zeros(A, self.stack_channels, h, w, device=dev)
        self._rew = torch.zeros(A, device=dev)
        self._val = torch.zeros(A, device=dev)
        self._graph = None
        self._graph_key = None
        self._warm = 0
        self._arena = ops.ZeroArena()  # zero-initialised scratch of one env step (its address is part of the graph key)
        self._policy = None

    def _core(self):
        """Device-only arithmetic of one step on static buffers (CUDA-graph capturable): optional policy act,
        one-hot, world-model eval forward, fused argmax decode, u8 frame-stack shift, reward decode."""
        cfg, model, A = self.config, self.model, self.num_envs
        if self._policy is not None:
            _, action, _ = self._policy.act(self._obs)
            self._act.copy_(action.reshape(A, 1))
        # all zero-initialised scratch of the step (split-K accumulators, statistics, the one-hot) out of ONE memset
        with ops.use_arena(self._arena):
            self._arena.begin(self.device)
            onehot = ops.zeros((A, 1, self.n_action), self.device).scatter_(2, self._act.unsqueeze(-1), 1.0)
            _, reward_pred, value_pred = model.forward_internal(self.frames / 255, onehot, logits_mode='argmax')
            self._arena.end()
        new = model.last_aux['argmax']  # u8 [A,3,H,W]: decoded in the epilogue of the logits GEMM
        c, h, w = cfg.frame_shape
        rp, vp = reward_pred.float().contiguous(), value_pred.float().contiguous().reshape(-1)
        if (h * w) % 16 == 0:  # frame-stack shift, float observation, reward decode and value copy in ONE launch
            _lib.check(_lib.lib().sb_env_post(self.frames.data_ptr(), new.contiguous().data_ptr(), self._obs.data_ptr(), A,
                                              self.stack_channels, c, h * w, rp.data_ptr(), rp.shape[1], vp.data_ptr(),
                                              self._rew.data_ptr(), self._val.data_ptr(), _lib.stream_ptr()),
                       'sb_env_post')
        else:
            self.frames.copy_(torch.cat((self.frames[:, c:], new), dim=1))
            self._obs.copy_(self.frames)
            self._rew.copy_((torch.argmax(rp, dim=1) - 1).to(torch.float32))
            self._val.copy_(vp)
        model.commit_state()

    def _capture_key(self):
        buf = self._arena.buf
        return (self.model.capture_key(rollout=True), buf.data_ptr() if buf is not None else 0)

    @torch.no_grad()
    def _advance(self):
        """Runs `_core` eagerly for the first step of a rollout (no recurrent gate yet) and during warm-up, then as a
        replayed CUDA graph."""
        model = self.model
        model.eval()
        auto = model._auto_repack
        model._auto_repack = False
        try:
            first = not model._has_prev
            if self._graph is not None and self._graph_key != self._capture_key():
                self._graph, self._warm = None, 0  # a buffer the graph baked in was re-allocated: capture again
            if first or not self.use_graph:
                self._core()
            elif self._graph is None and self._warm < 2:
                self._core()
                self._warm += 1
            elif self._graph is None:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._core()
                self._graph, self._graph_key = g, self._capture_key()
                g.replay()
            else:
                self._graph.replay()
        finally:
            model._auto_repack = auto

    @torch.no_grad()
    def _step(self, actions):
        cfg = self.config
        A = self.num_envs
        if self.frames is None:
            self._alloc_static()
        if self.step_count == 0:
            self.frames.copy_(self._initial)
            self._obs.copy_(self._initial)
            if cfg.stack_internal_states:
                self.model.init_internal_states(A)
            self.model.repack(static=True)  # weights are static during a rollout: pack once, not every step
        self.step_count += 1
        if actions is not None:
            self._act.copy_(torch.as_tensor(actions, device=self.device).long().reshape(A, -1)[:, :1])
        self._advance()
        rewards = self._rew.clone()
        if self.step_count == cfg.rollout_length:
            self.step_count = 0
            rewards = rewards + self._val
        if cfg.done_on_last_rollout_step:
            done = self.step_count == cfg.rollout_length  # evaluated after the reset above: always False (ref quirk)
        else:
            done = False
        obs = self._obs.clone()
        if self.reference_types:  # exactly the host-side types VecPytorchWrapper hands out (envs.py:109-118)
            return obs, rewards.double().cpu().unsqueeze(1), np.full(A, done, dtype=bool), tuple({} for _ in range(A))
        return obs, rewards.unsqueeze(1), np.full(A, done, dtype=bool), tuple({} for _ in range(A))

    def attach_policy(self, policy):
        """Fuse `policy.act` into the captured step: `step(None)` then samples the actions on the device from the
        current observation (used by bench.py; PPO drivers that call `act` themselves keep using `step(actions)`)."""
        if self.frames is None:
            self._alloc_static()
        self._policy = policy
        self._graph = None
        self._warm = 0


def make_simulated_env(config, model, action_space, reference_types=False):
    return SimulatedVecEnv(config, model, action_space, reference_types=reference_types)


def shard_agents(n_agents, rank, world_size):
    """Agent-sharded rollouts (no collective): the slice of global agent indices simulated by `rank`."""
    per = -(-n_agents // world_size)
    return range(min(n_agents, rank * per), min(n_agents, (rank + 1) * per))

"""Minimal stand-in for `gym` (not installable offline). TEST INFRASTRUCTURE ONLY.

Only what the reference's import statements and isinstance checks need
(SURVEY.md section 8c): Env/Wrapper base classes and Box/Discrete/Dict/Tuple.
"""
from . import spaces, wrappers  # noqa: F401


class Env:
    metadata = {}
    observation_space = None
    action_space = None

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def render(self, mode='human'):
        return None

    def close(self):
        return None

    def seed(self, seed=None):
        return [seed]


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self.observation_space = getattr(env, 'observation_space', None)
        self.action_space = getattr(env, 'action_space', None)

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError(name)
        return getattr(self.env, name)

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kw):
        return self.env.reset(**kw)


class ObservationWrapper(Wrapper):
    def observation(self, obs):
        raise NotImplementedError


class RewardWrapper(Wrapper):
    def reward(self, r):
        raise NotImplementedError


def make(name, **kw):
    raise RuntimeError('gym shim: real Atari environments are not available offline')

class NoopResetEnv:
    def __init__(self, env, noop_max=30):
        self.env = env


class FireResetEnv:
    def __init__(self, env):
        self.env = env

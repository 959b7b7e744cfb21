def copy_obs_dict(obs):
    return {k: v.copy() for k, v in obs.items()}


def dict_to_obs(d):
    return d[None] if set(d.keys()) == {None} else d


def obs_space_info(space):
    return [None], {None: space.shape}, {None: space.dtype}

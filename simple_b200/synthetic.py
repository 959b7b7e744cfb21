"""Synthetic "Atari-like" replay of the Atari-100k shape: constant background + moving 8x8 sprites from a
16-colour palette.  Used by bench.py, the tests and the `python -m simple_b200` demo (no ALE / dataset is reachable)."""
import torch


def synth_replay(n, n_action, seed, device='cpu'):
    g = torch.Generator().manual_seed(seed)
    H, W = 105, 80
    palette = torch.randint(0, 256, (16, 3), generator=g, dtype=torch.uint8)
    frames = palette[0].view(1, 3, 1, 1).repeat(n + 4, 1, H, W)
    pos = torch.randint(0, 60, (10, 2), generator=g)
    vel = torch.randint(-2, 3, (10, 2), generator=g)
    col = torch.randint(1, 16, (10,), generator=g)
    for t in range(n + 4):
        for s in range(10):
            y = int((pos[s, 0] + vel[s, 0] * t) % (H - 8))
            x = int((pos[s, 1] + vel[s, 1] * t) % (W - 8))
            frames[t, :, y:y + 8, x:x + 8] = palette[col[s]].view(3, 1, 1)
    obs = torch.stack([frames[t:t + 4].reshape(12, H, W) for t in range(n)])
    new_obs = frames[4:n + 4].clone()
    a = torch.randint(0, n_action, (n,), generator=g)
    action = torch.nn.functional.one_hot(a, n_action).to(torch.uint8)
    reward = torch.randint(0, 3, (n,), generator=g, dtype=torch.uint8)
    value = torch.randn(n, generator=g)
    return dict(obs=obs.to(device), new_obs=new_obs.to(device), action=action.to(device), reward=reward.to(device),
                value=value.to(device))

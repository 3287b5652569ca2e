"""Beta schedules (reference: utils/diffusion.py:28-65, make_beta_schedule)."""
import numpy as np
import torch


def make_beta_schedule(schedule, start, end, n_timestep):
    if schedule == 'linear':
        return torch.linspace(start, end, n_timestep, dtype=torch.float64)
    if schedule == 'quad':
        return torch.linspace(start ** 0.5, end ** 0.5, n_timestep, dtype=torch.float64) ** 2
    if schedule == 'const':
        return end * torch.ones(n_timestep, dtype=torch.float64)
    if schedule == 'cust':
        betas = end * np.ones(n_timestep, dtype=np.float64)
        warm = int(n_timestep * 0.1)
        betas[:warm] = np.linspace(start, end, warm, dtype=np.float64)
        return torch.from_numpy(betas)
    if schedule == 'jsd':
        return 1. / torch.linspace(n_timestep, 1, n_timestep, dtype=torch.float64)
    raise NotImplementedError(schedule)

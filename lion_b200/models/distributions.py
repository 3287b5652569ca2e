"""Diagonal normal used by the VAE encoders (mirror of the reference's models/distributions.py:17-35, the part
`Model.encode` / `recont` touch): sample = mu + sigma * rho with rho ~ N(0, 1) drawn by torch on mu's device."""
import numpy as np
import torch


class Normal:
    def __init__(self, mu, log_sigma, sigma=None):
        self.mu = mu
        self.log_sigma = log_sigma
        self.sigma = torch.exp(log_sigma) if sigma is None else sigma

    def sample(self, t=1.):
        rho = torch.zeros_like(self.mu).normal_()
        return rho * (self.sigma * t) + self.mu, rho

    def sample_given_rho(self, rho):
        return rho * self.sigma + self.mu

    def mean(self):
        return self.mu

    def log_p(self, samples):
        normalized_samples = (samples - self.mu) / self.sigma
        return - 0.5 * normalized_samples * normalized_samples - 0.5 * np.log(2 * np.pi) - self.log_sigma

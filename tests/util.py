import torch


def rel_err(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def assert_close(a, b, tol, name):
    assert tuple(a.shape) == tuple(b.shape), "%s: shape %s vs %s" % (name, tuple(a.shape), tuple(b.shape))
    a64 = torch.as_tensor(a).detach().double().cpu()
    assert torch.isfinite(a64).all(), "%s: non-finite values" % name
    e = rel_err(a, b)
    assert e <= tol, "%s: max-abs error / max-abs reference = %.3e > %.1e" % (name, e, tol)
    return e


def rms_err(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp_min(1e-30)).item()


def gen(seed, *shape, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale

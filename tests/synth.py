"""Deterministic synthetic weights, shared by tests/, bench.py and tests/golden/make_golden.py.

No checkpoints exist offline, so every parity test and the benchmark fill a
state_dict -- given only its key -> shape table -- from a counter-based recipe that depends
on nothing but the key name, the shape and a seed.  The same recipe fed the reference modules
when the golden vectors were made, so goldens, oracle and CUDA path all see identical weights
without shipping 400 MB of tensors.
"""
import zlib

import torch


def synth_tensor(key, shape, seed=1234):
    g = torch.Generator()
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 63))
    shape = tuple(shape)
    n = 1
    for s in shape:
        n *= s
    r = torch.randn(n, generator=g, dtype=torch.float32).reshape(shape)
    leaf = key.rsplit(".", 1)[-1]
    if leaf == "weight" and len(shape) >= 2:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        gain = 0.3 if ".emd." in key else 1.0          # style -> (factor,bias) stays a perturbation
        return r * (gain / fan_in ** 0.5)
    if leaf == "weight":                                # GroupNorm gamma
        return 1.0 + 0.1 * r
    if leaf == "bias" and ".emd." in key:               # AdaGN: factor ~ 1, bias ~ 0 (adagn.py:39-40)
        c = shape[0] // 2
        out = 0.1 * r
        out[:c] += 1.0
        return out
    return 0.1 * r                                      # every other bias / 1-D parameter


def synth_state_dict(shapes, seed=1234, device="cpu"):
    """shapes: {key: shape}.  Returns {key: fp32 tensor}."""
    return {k: synth_tensor(k, s, seed).to(device) for k, s in shapes.items()}

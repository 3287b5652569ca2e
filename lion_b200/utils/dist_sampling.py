"""Data-parallel sampling helpers.

Shapes are independent units, so a batch is sharded contiguously over ranks with no data-path
collective; the only exchange is one all_gather of the finished point clouds
(reference: trainers/base_trainer.py:446-487, which bounces them through .cpu() and seeds
every rank identically -- here ranks get distinct seeds and the gather is device-to-device).
"""
import torch
import torch.distributed as dist


def shard_sizes(total, world):
    """Contiguous split of `total` shapes over `world` ranks (first ranks get the remainder)."""
    base, rem = divmod(total, world)
    return [base + (1 if r < rem else 0) for r in range(world)]


def rank_seed(seed, rank, reference_behaviour=False):
    """The reference re-seeds every rank with the same seed (base_trainer.py:459-463), so ranks
    generate duplicates; default here is a distinct stream per rank."""
    return seed if reference_behaviour else seed * 1000 + rank


def gather_samples(local):
    """all_gather of per-rank [B_r, N, 3] tensors -> [sum B_r, N, 3] on every rank (rank order).
    Works with equal shard sizes (one all_gather) or ragged ones (padded to the largest)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    n = torch.tensor([local.shape[0]], device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat([local, local.new_zeros(mx - local.shape[0], *local.shape[1:])], dim=0)
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad.contiguous())
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)

"""Data-parallel plumbing: one process per GPU, utterances sharded by rank, ONE all-reduce of
the flat gradient buffer per step.  The reference is single-device (config_utils.py:64-66 parses a
`device` key nobody reads), so there is no counterpart to cite; the loss definitions that the global
reduction must preserve are models.py:151 (SI: mean over all B*T*F bins) and models.py:1947-1955 (MTL:
hole-L1 as a global ratio + mean-over-batch CTC).

Everything here works on CPU tensors with the gloo backend (tests/test_parallel_cpu.py) and on CUDA tensors
with NCCL (bench.py under torchrun); torch.distributed is the transport, nothing else.
"""
import torch

LOSS_TAIL = 8          # loss scalars riding behind the gradient in the same buffer


def shard_list(items, rank, world):
    """Rank r reads items[r::world] of the sorted list (replaces training.py:47-49's single glob)."""
    return sorted(items)[rank::world]


def shard_batch(n, rank, world):
    """[lo, hi) of a global batch of n utterances for this rank; equal sizes (n must divide)."""
    if n % world:
        raise ValueError('global batch %d is not divisible by world size %d' % (n, world))
    per = n // world
    return rank * per, (rank + 1) * per


def pack_loss_tail(flat, n_params, sums):
    """Write the per-rank loss sums {sum|d|(1-m), sum(1-m), sum|d|m, sum m, sum|d|, count} behind the gradient."""
    flat[n_params:n_params + 6].copy_(sums[:6].to(flat.dtype))
    return flat


def all_reduce_flat(flat, group=None):
    """Sum the flat gradient (+ loss tail) over the ranks of `group`, in place."""
    import torch.distributed as dist
    if group is None and not dist.is_initialized():
        return flat
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def si_unscale(batch_per_rank, world, T, F):
    """Factor turning the summed, unnormalised SI gradient into d mean|d| / d theta over the GLOBAL batch."""
    return 1.0 / (batch_per_rank * world * T * F)


def global_losses(flat, n_params):
    """(loss_hole, loss_valid, loss_func) of the global batch from the all-reduced tail (float64)."""
    t = flat[n_params:n_params + 6].double()
    return t[0] / t[1], t[2] / t[3], t[4] / t[5]

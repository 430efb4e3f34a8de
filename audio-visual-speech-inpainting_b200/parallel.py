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


def bind_to_gpu_numa(local_rank):
    """Pin this process to the CPU cores next to its GPU (NVML's ideal CPU affinity of the device) BEFORE it allocates
    pinned host buffers, so that their pages sit on the GPU's NUMA node and the H2D copies do not cross the socket
    interconnect -- with 8 ranks staging ~1 GB per step each, the host side of the PCIe path is the scarce resource.
    Returns the number of cores bound, or 0 when NVML / sched_setaffinity is unavailable (nothing changes then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByIndex(int(local_rank))
            words = (os.cpu_count() + 63) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
            cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1}
            cpus &= set(os.sched_getaffinity(0))
            if cpus:
                os.sched_setaffinity(0, cpus)
            return len(cpus)
        finally:
            pynvml.nvmlShutdown()
    except Exception:                                     # noqa: BLE001 -- an optimisation, never a requirement
        return 0


def shard_list(items, rank, world, keep_order=False):
    """Rank r reads items[r::world] (replaces training.py:47-49's single glob).  The list is sorted first unless the
    caller vouches that every rank passes it in the SAME order (keep_order=True: the per-epoch shuffle of
    training.py:212-214 done with a rank-consistent seed survives the sharding)."""
    items = list(items) if keep_order else sorted(items)
    return items[rank::world]


def lockstep(batches, batch_size, group=None, device='cpu'):
    """Yield from this rank's batch generator only while EVERY rank of `group` still has a full batch.

    Each rank reads its own shard, so the ranks can run out at different steps or end on a smaller batch
    (N_train % (world * batch) != 0).  A rank that left the loop while the others are still inside the gradient
    all-reduce would dead-lock the job, and a smaller last batch would be normalised by the wrong global count.  Every
    step the ranks therefore agree (one MIN all-reduce of a flag) whether all of them hold a batch of exactly
    `batch_size` utterances; the first step on which one does not ends the epoch everywhere, the ragged tail is dropped.
    Single process (no group, torch.distributed not initialised): plain iteration, partial batches included."""
    import torch.distributed as dist
    if group is None and not dist.is_initialized():
        for b in batches:
            yield b
        return
    it = iter(batches)
    flag = torch.zeros(1, dtype=torch.int32, device=device)
    while True:
        b = next(it, None)
        ok = b is not None and len(b[0]) == batch_size
        flag.fill_(1 if ok else 0)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            close = getattr(it, 'close', None)
            if close is not None:
                close()
            return
        yield b


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

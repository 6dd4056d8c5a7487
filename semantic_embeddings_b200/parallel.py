"""Data-parallel plumbing (one process per GPU, torch.distributed).

The reference's multi-GPU path is keras.utils.multi_gpu_model (learn_image_embeddings.py:133,148): the batch is
sliced over towers with shared weights, BatchNorm statistics stay per tower, and the loss is the mean over the
merged batch.  Here every rank runs the same launch plans on its shard with the per-sample loss pre-scaled by
1/global_batch, so that ONE all-reduce(SUM) over the flat gradient buffer yields the gradient of the global mean;
L2 regularisation, global-norm clipping and the SGD update then run identically on every rank.  The all-pairs
distance matrix shards by query rows with no exchange step (evaluate_retrieval.py:56-67 is per-query).
"""
import os


def shard_rows(n, world, rank):
    """Contiguous row block [row0, row0+rows) of rank `rank`; the last ranks may be empty when n < world."""
    per = (n + world - 1) // world
    row0 = min(n, rank * per)
    return row0, max(0, min(per, n - row0))


def shard_batch(global_batch, world, rank):
    """keras.utils.multi_gpu_model slicing: equal contiguous slices, the last tower takes the remainder."""
    per = global_batch // world
    start = rank * per
    size = per if rank < world - 1 else global_batch - start
    return start, size


def init_process_group(backend=None, device=None):
    import torch
    import torch.distributed as dist
    if dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world == 1:
        return 0, 1
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    backend = backend or ('nccl' if torch.cuda.is_available() else 'gloo')
    kw = {'device_id': device} if (backend == 'nccl' and device is not None) else {}
    dist.init_process_group(backend, **kw)
    return dist.get_rank(), dist.get_world_size()


def allreduce_gradients(flat, group=None):
    """SUM all-reduce of the flat gradient buffer (losses are already scaled by 1/global_batch)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def broadcast_parameters(tensors, src=0, group=None):
    """Makes every replica start from rank `src`'s weights / optimizer state."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        for t in tensors:
            dist.broadcast(t, src=src, group=group)

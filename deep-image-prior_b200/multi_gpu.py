"""One independent image per GPU: static sharding + the only collectives of the job.

The hot path has no exchange step, so there is no data-path collective: NCCL (or gloo in the CPU tests) is used for
the start/stop barrier of the timed region, the max-over-ranks of the device time and one all_gather of a small
fixed-size result record per rank.
"""
import torch
import torch.distributed as dist


def shard(n_items, rank, world):
    """Indices of the images rank `rank` optimises: item i goes to rank i mod world."""
    return list(range(rank, n_items, world))


def max_over_ranks(value, device="cpu"):
    """Max of a python float over all ranks (timing is always the slowest rank's)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def gather_records(record, device="cpu"):
    """all_gather of a fixed-size float64 record (e.g. [psnr_gt, final_loss, it_per_s]); returns a list per rank."""
    rec = torch.as_tensor(record, dtype=torch.float64, device=device).flatten()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [rec.cpu().tolist()]
    out = [torch.zeros_like(rec) for _ in range(dist.get_world_size())]
    dist.all_gather(out, rec)
    return [o.cpu().tolist() for o in out]


def aggregate_rate(steps_per_rank, max_seconds, world):
    """Whole-job iterations/sec: every rank did `steps_per_rank` steps within the slowest rank's time."""
    return world * steps_per_rank / max_seconds

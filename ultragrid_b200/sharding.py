"""Frame-level data parallelism across GPUs: frames are independent units, so the encode queue is
partitioned with no data-path collective.  The only message is which frame index goes where; rank 0 owns the queue
and scatters int32 index vectors (NCCL on GPUs, gloo in the CPU tests).  Mirrors the per-device workers + sequence
reorder of src/video_compress/gpujpeg.cpp:643-722 at process granularity."""
import torch
import torch.distributed as dist


def frame_assignment(step, world, frames_per_rank, device="cpu"):
    """global frame indices of `step`, shape (world, frames_per_rank): frame i of the step's batch goes to rank i // B"""
    base = step * world * frames_per_rank
    return (torch.arange(world * frames_per_rank, dtype=torch.int32, device=device) + base).reshape(world, frames_per_rank)


def scatter_assignment(step, frames_per_rank, out, group=None):
    """rank 0 scatters the assignment of `step`; every rank receives its row into `out` (int32[frames_per_rank])"""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        out.copy_(frame_assignment(step, 1, frames_per_rank, out.device)[0])
        return out
    rows = None
    if rank == 0:
        a = frame_assignment(step, world, frames_per_rank, out.device)
        rows = [a[r].contiguous() for r in range(world)]
    dist.scatter(out, rows, src=0, group=group)
    return out


def merge_in_sequence(per_rank_results):
    """host-side reorder by global frame index (the role of pop()'s m_out_frames map): list of (index, payload) lists -> payloads"""
    merged = sorted((item for r in per_rank_results for item in r), key=lambda t: t[0])
    idx = [i for i, _ in merged]
    assert idx == list(range(idx[0], idx[0] + len(idx))), "frames lost or duplicated"
    return [p for _, p in merged]

"""Multi-GPU plumbing for the matcher: images are independent, so the query batch is sharded
across ranks (one process per GPU) and the per-object 3D bank is broadcast once over NCCL
(SURVEY.md §8e).  There is no steady-state collective on the data path; `gather_matches` is the
optional end-of-batch collection of the ragged match lists.

Backend-agnostic (NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist

BANK_KEYS = ("keypoints3d", "descriptors3d_db", "descriptors3d_coarse_db")


def shard_range(n_items, rank, world):
    """Contiguous, balanced slice [lo, hi) of `n_items` for `rank` (first ranks get the remainder)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_bank(bank, src=0, group=None):
    """In-place broadcast of the per-object bank tensors from `src` (fixed key order so every rank
    issues the same collectives).  Tensors must be pre-allocated with the right shapes on all ranks."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return bank
    for k in BANK_KEYS:
        if k in bank:
            dist.broadcast(bank[k], src=src, group=group)
    return bank


def gather_matches(data, image_offset, group=None):
    """All-gather the ragged match lists (m_bids, mkpts_3d_db, mkpts_query_f, mconf) of every rank.
    `image_offset` is this rank's first global image index; returned m_bids are global.
    Two collectives: counts, then one padded [cap, 7] float tensor per rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    m = data["m_bids"].numel()
    dev = data["mconf"].device
    packed = torch.cat([(data["m_bids"] + image_offset).to(torch.float32)[:, None],
                        data["mkpts_3d_db"], data["mkpts_query_f"], data["mconf"][:, None]], 1)
    if world == 1:
        return packed
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([m], dtype=torch.int64, device=dev), group=group)
    cap = int(max(c.item() for c in counts))
    padded = torch.zeros((cap, 7), dtype=torch.float32, device=dev)
    padded[:m] = packed
    out = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    return torch.cat([o[: int(c.item())] for o, c in zip(out, counts)], 0)

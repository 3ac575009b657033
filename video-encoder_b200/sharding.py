"""Host-side sharding of the encode stage across GPUs: closed GOPs of one stream, or whole
streams, are independent units -> unit u belongs to rank u % world; no data-path collective exists.
torch.distributed is used only for the barrier and the max-over-ranks of a timed region."""


def owner(unit: int, world: int) -> int:
    """rank that encodes closed GOP / stream number `unit`"""
    return unit % world


def units_of_rank(n_units: int, rank: int, world: int):
    """closed GOPs / streams encoded by `rank` (round-robin, keeps every rank within one unit of the others)"""
    return list(range(rank, n_units, world))


def slot_streams(rank: int, slots: int):
    """synthetic stream ids behind the `slots` lock-step slots of `rank` (bench.py): disjoint across ranks"""
    return [rank * slots + s for s in range(slots)]


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """max of a per-rank scalar (timed region = slowest rank)"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

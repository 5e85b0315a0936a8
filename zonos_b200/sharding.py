"""Request sharding across the GPUs of one box.

Utterances never interact inside `generate` (rows of different utterances share nothing but the weights), so the
multi-GPU mode is one process per GPU, a full weight replica each, and NO collective on the decode path: rank g
takes utterances [g*B/N, (g+1)*B/N).  The only cross-rank traffic is the final gather of the finished codes (a few KB
per utterance) and, to keep the reference's single-call semantics, the batch-global truncation of
`zonos/model.py:513-539`, which is re-applied on the gathered batch.
"""
import torch
import torch.distributed as dist


def shard_range(n_items: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced split: the first n_items % world_size ranks get one extra item."""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_conditioning(prefix_conditioning: torch.Tensor, batch_size: int, world_size: int, rank: int) -> torch.Tensor:
    """[2B, Lc, D] (cond rows then uncond rows, conditioning_cache.py:176) -> this rank's [2b, Lc, D]; the CFG pair of an
    utterance stays on one GPU because both rows reuse every weight byte."""
    assert prefix_conditioning.shape[0] == 2 * batch_size
    lo, hi = shard_range(batch_size, world_size, rank)
    return torch.cat([prefix_conditioning[lo:hi], prefix_conditioning[batch_size + lo:batch_size + hi]], dim=0)


def batch_global_truncate(codes: list[torch.Tensor], pad_value: int = 0) -> torch.Tensor:
    """Per-rank outputs may have different valid lengths (each rank's early exit only sees its own utterances); the
    reference cuts every utterance of a call at ONE length, the shortest per-rank result is the batch-global bound."""
    n = min(c.shape[-1] for c in codes)
    return torch.cat([c[..., :n] for c in codes], dim=0)


def generate_sharded(generate_fn, prefix_conditioning: torch.Tensor, batch_size: int, audio_prefix_codes: torch.Tensor | None = None,
                     group=None, **kwargs) -> torch.Tensor | None:
    """Runs `generate_fn(cond_shard, audio_prefix_codes=..., batch_size=b, **kwargs)` on this rank's utterances and gathers
    the int64 codes on rank 0 (returns None elsewhere).  `generate_fn` is `Zonos.generate` in production; tests inject a
    CPU stand-in.  Works with any backend (NCCL on the GPU box, gloo in the CPU tests)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_range(batch_size, world, rank)
    local = None
    if hi > lo:
        cond = shard_conditioning(prefix_conditioning, batch_size, world, rank)
        prefix = audio_prefix_codes[lo:hi] if audio_prefix_codes is not None else None
        local = generate_fn(cond, audio_prefix_codes=prefix, batch_size=hi - lo, **kwargs)
    if world == 1:
        return local
    payload = None if local is None else local.cpu()
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0, group=group)
    if rank != 0:
        return None
    return batch_global_truncate([g for g in gathered if g is not None])

"""Request sharding across the GPUs of one box.

Utterances never interact inside `generate` (rows of different utterances share nothing but the weights), so the
multi-GPU mode is one process per GPU, a full weight replica each, and NO collective on the decode path: rank g
takes utterances [g*B/N, (g+1)*B/N).  The only cross-rank traffic is the final gather of the finished codes (a few KB
per utterance), which is aligned to ONE length like a single reference call's batch (`batch_global_align`).
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


def batch_global_align(codes: list[torch.Tensor], pad_value: int = 512) -> torch.Tensor:
    """Per-rank outputs may have different lengths: each rank's loop ends when ITS utterances have finished.  A single
    reference call runs until ALL rows have finished (`zonos/utilities/tensor_ops.py:95`, `.all()`), so its length follows
    the LONGEST utterance, and an utterance that ended earlier is filled with masked tokens, which the output
    clean-up turns into 512 (`zonos/model.py:531-535`).  The gathered batch is therefore right-padded with 512 to the
    longest per-rank result; no utterance is ever cut.  (The reference's last-50-positions EOS scan, `model.py:513-528`,
    counts EOS over the whole batch and can trim a few trailing frames more than a rank-local scan does; by then the
    EOS tokens are gone, so that trim is not re-applied here: the sharded result can be a few frames LONGER than the
    single-call one, never shorter.)"""
    n = max(c.shape[-1] for c in codes)
    out = []
    for c in codes:
        if c.shape[-1] < n:
            pad = torch.full((*c.shape[:-1], n - c.shape[-1]), pad_value, dtype=c.dtype, device=c.device)
            c = torch.cat([c, pad], dim=-1)
        out.append(c)
    return torch.cat(out, dim=0)


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
    return batch_global_align([g for g in gathered if g is not None])

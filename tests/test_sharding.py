"""CPU, world_size 2 over gloo: the request-sharding host logic (the N>1 path has no data-path collective)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from zonos_b200.sharding import batch_global_align, generate_sharded, shard_conditioning, shard_range


def test_shard_range_is_a_partition():
    for n in (1, 2, 7, 64, 65):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_conditioning_keeps_cfg_pairs_together():
    B, Lc, D = 5, 3, 4
    cond = torch.arange(2 * B).float().view(2 * B, 1, 1).expand(2 * B, Lc, D)
    for w in (2, 3):
        for r in range(w):
            lo, hi = shard_range(B, w, r)
            s = shard_conditioning(cond, B, w, r)
            assert s.shape[0] == 2 * (hi - lo)
            assert s[:hi - lo, 0, 0].tolist() == list(range(lo, hi))                  # cond rows
            assert s[hi - lo:, 0, 0].tolist() == list(range(B + lo, B + hi))          # matching uncond rows


def _fake_generate(cond, audio_prefix_codes=None, batch_size=1, max_new_tokens=8, **kw):
    # deterministic stand-in: codes depend only on the utterance's own cond row; length depends on the shard
    ids = cond[:batch_size, 0, 0].long()
    n = max_new_tokens - 3 * (int(ids.min()) % 2)
    return (ids.view(-1, 1, 1) * 10 + torch.arange(n).view(1, 1, n)).expand(batch_size, 9, n).contiguous()


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B = 5
    cond = torch.arange(2 * B).float().view(2 * B, 1, 1).expand(2 * B, 2, 4).contiguous()
    res = generate_sharded(_fake_generate, cond, B, max_new_tokens=8)
    if rank == 0:
        torch.save(res, out)
    else:
        assert res is None
    dist.destroy_process_group()


def test_generate_sharded_world2_gloo(tmp_path):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, 29611, out), nprocs=2, join=True)
    res = torch.load(out)
    B = 5
    cond = torch.arange(2 * B).float().view(2 * B, 1, 1).expand(2 * B, 2, 4)
    # every utterance keeps its own tokens (no cross-rank mixing); the batch has the length of the LONGEST per-rank result
    parts = [_fake_generate(shard_conditioning(cond, B, 2, r), batch_size=shard_range(B, 2, r)[1] - shard_range(B, 2, r)[0]) for r in range(2)]
    assert parts[0].shape[-1] != parts[1].shape[-1]          # the ranks really finish at different lengths
    n_long = max(p.shape[-1] for p in parts)
    assert res.shape == (B, 9, n_long)
    row = 0
    for p in parts:
        for u in range(p.shape[0]):
            assert torch.equal(res[row, :, :p.shape[-1]], p[u])                   # nothing cut, nothing mixed
            assert (res[row, :, p.shape[-1]:] == 512).all()                       # what the reference emits after an utterance's end
            row += 1
    assert res[:, 0, 0].tolist() == [0, 10, 20, 30, 40]


def test_batch_global_align_keeps_the_longest_utterance():
    """A rank with a short utterance must not truncate another rank's long one (zonos/utilities/tensor_ops.py:95: a batch
    runs until ALL rows are done)."""
    short = torch.arange(2 * 9 * 20).view(2, 9, 20) % 1000
    long_ = torch.arange(1 * 9 * 100).view(1, 9, 100) % 1000
    out = batch_global_align([short, long_])
    assert out.shape == (3, 9, 100)
    assert torch.equal(out[2], long_[0]) and torch.equal(out[:2, :, :20], short) and (out[:2, :, 20:] == 512).all()

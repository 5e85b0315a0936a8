"""CPU, world_size 2 over gloo: the request-sharding host logic (the N>1 path has no data-path collective)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from zonos_b200.sharding import batch_global_truncate, generate_sharded, shard_conditioning, shard_range


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
    n = max_new_tokens - int(ids.min()) % 3
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
    # every utterance keeps its own tokens (no cross-rank mixing) and the batch is cut at one common length
    parts = [_fake_generate(shard_conditioning(cond, B, 2, r), batch_size=shard_range(B, 2, r)[1] - shard_range(B, 2, r)[0]) for r in range(2)]
    want = batch_global_truncate(parts)
    assert res.shape == want.shape and torch.equal(res, want)
    assert res[:, 0, 0].tolist() == [0, 10, 20, 30, 40]

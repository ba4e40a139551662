"""CPU tests of the N>1 path: utterance sharding + the single all-gather of token rows,
world_size 2 over gloo on 127.0.0.1 (host-side logic only; the per-rank engine is replaced by
a deterministic token producer because there is no GPU here)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_tokens(pcm):
    """Stand-in for the engine: tokens are a pure function of the clip."""
    n = int(abs(pcm[:8]).sum() * 1000) % 7
    return [int(abs(v) * 1e4) % 1024 for v in pcm[:n]]


def _worker(rank, world, port, n_clips, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import __graft_entry__ as ge
    ge.load_package()
    from parakeet_cpp_b200 import dist as pkd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    pcms = [rng.standard_normal(64).astype(np.float32) for _ in range(n_clips)]
    got = pkd.transcribe_sharded(lambda xs: [_fake_tokens(x) for x in xs], pcms, cap=16, world=world, rank=rank)
    q.put((rank, got))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [8, 7, 1])
def test_sharded_transcribe_gloo_world2(n_clips):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, n_clips, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    want = [_fake_tokens(rng.standard_normal(64).astype(np.float32)) for _ in range(n_clips)]
    assert res[0] == want and res[1] == want


def test_shard_ranges_cover_and_partition():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import __graft_entry__ as ge
    ge.load_package()
    from parakeet_cpp_b200 import dist as pkd
    for n in (0, 1, 7, 64, 8192, 8191):
        for w in (1, 2, 4, 8):
            spans = [pkd.shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) <= -(-n // w) if n else True
    rows = pkd.rows_from_tokens([[1, 2, 3], [], [9]], 4)
    assert rows.tolist() == [[3, 1, 2, 3, 0], [0, 0, 0, 0, 0], [1, 9, 0, 0, 0]]
    assert pkd.tokens_from_rows(rows) == [[1, 2, 3], [], [9]]

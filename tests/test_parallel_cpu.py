"""Host logic of the data-parallel gradient exchange on CPU: world_size 2, gloo backend (the NCCL path runs the same
code with ReduceOp.AVG on device buffers).  A stand-in encoder exposes the flat gradient arena and reports finished
ranges from the tail to the head, exactly like `VisualDialogEncoder._backward`."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _Arena:
    def __init__(self, n, live_end):
        self.w32 = torch.zeros(n)
        self.g32 = torch.zeros(n)
        self.live_end = live_end


class _FakeEncoder(torch.nn.Module):
    def __init__(self, n, live_end):
        super().__init__()
        self.arena = _Arena(n, live_end)
        self.grad_ready_hook = None


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from cqa_crct_b200.parallel import DistributedDataParallel
    n, live_end = 64 * 40, 64 * 32
    enc = _FakeEncoder(n, live_end)
    enc.arena.w32[:] = float(rank + 1)
    ddp = DistributedDataParallel(enc, bucket_cap_mb=64 * 10 * 4 / (1 << 20))        # 640-element buckets
    assert torch.equal(enc.arena.w32, torch.ones(n))                                  # rank 0's weights were broadcast
    ranges = [(64 * 28, live_end), (64 * 20, 64 * 28), (64 * 19, 64 * 20), (64 * 6, 64 * 19), (0, 64 * 6)]
    for it in range(2):                                                               # two backward passes: state resets
        enc.arena.g32[:] = torch.arange(n, dtype=torch.float32) * (rank + 1) + it
        for lo, hi in ranges:
            enc.grad_ready_hook(lo, hi)
        enc.grad_ready_hook(None, None)
        want = torch.arange(n, dtype=torch.float32) * 1.5 + it
        assert torch.allclose(enc.arena.g32[:live_end], want[:live_end])
        # dead tail is never exchanged
        assert torch.equal(enc.arena.g32[live_end:], torch.arange(n, dtype=torch.float32)[live_end:] * (rank + 1) + it)
        assert ddp.buckets_last_step == [(64 * 20, live_end), (64 * 6, 64 * 20), (0, 64 * 6)], ddp.buckets_last_step
    try:                                                                              # a gap in the reported ranges is a bug
        enc.grad_ready_hook(64 * 28, live_end)
        enc.grad_ready_hook(64 * 10, 64 * 20)
        ok = False
    except RuntimeError:
        ok = True
    q.put((rank, ok))
    dist.destroy_process_group()


def test_bucketed_allreduce_world2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert got == [(0, True), (1, True)]


def test_block_ranges_tile_the_live_arena():
    """The ranges the backward reports are contiguous, descending and cover exactly [0, live_end)."""
    from cqa_crct_b200.encoder import VisualDialogEncoder
    from cqa_crct_b200.synthetic import default_params
    from tests.helpers import CONFIG_DIR
    for f in ('tiny.json', 'vilbert.json'):
        if f == 'vilbert.json':
            from cqa_crct_b200.spec import ModelConfig, param_spec, arena_order, arena_offsets
            cfg = ModelConfig(os.path.join(CONFIG_DIR, f))
            order = arena_order(cfg, param_spec(cfg))
            off, live_end, _ = arena_offsets(order)

            class A:      # no 1 GB allocation on the CPU test box
                pass
            enc = VisualDialogEncoder.__new__(VisualDialogEncoder)
            enc.arena = A()
            enc.arena.order, enc.arena.offsets, enc.arena.live_end = order, off, live_end
            enc.arena.by_name = {p.name: p for p in order}
            enc.cfg = cfg
        else:
            enc = VisualDialogEncoder(default_params(os.path.join(CONFIG_DIR, f)))
        a = enc.arena
        ranges = [(a.offsets['bert.t_pooler.dense.weight'], a.live_end)]
        for kind, i in reversed(enc.cfg.schedule()):
            ranges.append(enc._block_range({'t': f'bert.encoder.layer.{i}', 'v': f'bert.encoder.v_layer.{i}', 'c': f'bert.encoder.c_layer.{i}'}[kind]))
        ranges.append((0, enc._block_range('bert.v_embeddings')[1]))
        for (lo, hi), (lo2, hi2) in zip(ranges, ranges[1:]):
            assert hi2 == lo, (lo, hi, lo2, hi2)
        assert ranges[-1][0] == 0 and ranges[0][1] == a.live_end


def test_sharded_optimizer_pieces_and_shards_tile_every_bucket():
    """Host arithmetic of the sharded optimizer (graph.GraphedTrainStep._chunks / shard_bounds): the exchange pieces of a bucket are
    contiguous, cover it exactly, each splits evenly over the ranks, and the ranks' shards of a piece tile the piece."""
    import types
    from cqa_crct_b200.graph import GraphedTrainStep as G
    for world in (2, 4, 8):
        for chunk_mb in (1.0, 32.0):
            for lo, hi in [(0, 64), (0, 64 * world), (128, 128 + 25140480), (320, 320 + 64 * world * 3 + 64), (0, 23440896 + 76800),
                           (6400, 6400 + 20478976)]:
                ranks = [types.SimpleNamespace(world=world, rank=r, shard_chunk=int(chunk_mb * (1 << 20) / 4)) for r in range(world)]
                pieces = G._chunks(ranks[0], lo, hi)
                assert pieces[0][0] == lo and pieces[-1][1] == hi and all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
                assert all((b - a) % world == 0 and b > a for a, b in pieces)
                assert all(b - a <= max(64 * world, ranks[0].shard_chunk) + 64 * world for a, b in pieces)
                for a, b in pieces:
                    shards = [G.shard_bounds(o, a, b) for o in ranks]
                    assert shards[0][0] == a and shards[-1][1] == b and all(x[1] == y[0] for x, y in zip(shards, shards[1:]))
                    assert len({s1 - s0 for s0, s1 in shards}) == 1

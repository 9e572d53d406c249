"""Multi-rank path on CPU: two gloo ranks each encode a contiguous block range of ONE stream
(emulated kernels), rank 0 assembles -- the result must be the single-rank file."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, level, ov, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from flake_b200 import api, build, shard, synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = api.load_library(build.build_emu())
    pcm = synth.synth_pcm(1024 * 5 + 300, 2, 16, 44100, seed=21, kind="impulses")
    data = shard.encode_sharded(lib, pcm, 44100, 16, level, rank=rank, world=world, **ov)
    if rank == 0:
        q.put(data)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("level", [5, 9])
def test_two_rank_frame_range_sharding(level, oracle):
    import torch.multiprocessing as mp
    from flake_b200 import build, shard, synth
    build.build_emu()
    ov = {"block_size": 1024}
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, level, ov, q)) for r in range(2)]
    for p in procs:
        p.start()
    data = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pcm = synth.synth_pcm(1024 * 5 + 300, 2, 16, 44100, seed=21, kind="impulses")
    body, flen, fbs, mx = oracle.encode_stream(pcm, 44100, 16, level, **ov)
    p = oracle.make_params(2, 44100, 16, level, pcm.shape[0], **ov)
    hdr = bytearray(oracle.header(p))
    hdr[8:42] = oracle.streaminfo(p, mx, oracle.md5_pcm(pcm, 16))
    assert data == bytes(hdr) + body
    dec, info = oracle.decode(data)
    assert info.md5_ok == 1 and np.array_equal(dec, pcm)


def test_block_ranges():
    from flake_b200.shard import block_ranges
    for nb in (0, 1, 7, 8, 38760):
        for w in (1, 2, 4, 8):
            r = block_ranges(nb, w)
            assert r[0][0] == 0 and r[-1][1] == nb
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1

"""Multi-GPU host logic on CPU: the contiguous pair split and the host gather of result
records, exercised with a world-size-2 gloo group (no GPU, no CUDA library calls).

The per-rank "device work" is stood in for by the oracle on small pairs -- the point here
is the sharding and gather plumbing `bench.py --gpus N` and multi-process users rely on
(independent pairs, contiguous blocks, no collective on the data path, host
gather of (lag, coef, ret, peak) only).
"""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "old-audiosync_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import audiosync_cuda as ac  # noqa: E402

SEED, L, N_PAIRS = 0x5EED + 9, 600, 7


def test_shard_pairs_is_the_dispatcher_split():
    # audiosync_cuda.cu (HOST memspace): base = n / G, remainder to the low devices, contiguous
    for n in (0, 1, 7, 8, 4096, 32768, 32771):
        for world in (1, 2, 3, 4, 8):
            blocks = [ac.shard_pairs(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
            for (f0, c0), (f1, c1) in zip(blocks, blocks[1:]):
                assert f1 == f0 + c0 and c0 - c1 in (0, 1)
    assert ac.shard_pairs(32768, 8, 3) == (3 * 4096, 4096)      # BASELINE config 5
    with pytest.raises(ValueError):
        ac.shard_pairs(8, 2, 2)


def _records_for(first, count):
    from oracle import capi
    rec = np.zeros(count, dtype=ac.RESULT_DTYPE)
    for i in range(count):
        s, p = capi.synth_pair(SEED, first + i, L)
        o = capi.cross_correlation(s, p)
        rec[i] = (o["lag"], o["coef"], o["peak"], o["ret"],
                  int(o["ret"] == 0 and o["coef"] >= ac.MIN_CONFIDENCE), o["raw_index"], o["second"],
                  o["margin"], o["ncc"])
    return rec


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, count = ac.shard_pairs(N_PAIRS, world, rank)
        local = _records_for(first, count)
        out = ac.gather_results(local, N_PAIRS, dst=0)
        if rank == 0:
            q.put(out.tobytes())
        else:
            assert out is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gather_matches_serial():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = np.frombuffer(q.get(timeout=120), dtype=ac.RESULT_DTYPE)
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    want = _records_for(0, N_PAIRS)
    assert got.shape == want.shape
    for name in ac.RESULT_DTYPE.names:
        a, b = got[name], want[name]
        assert np.array_equal(a, b) or (np.isnan(a) == np.isnan(b)).all() and np.allclose(
            np.nan_to_num(a), np.nan_to_num(b), rtol=0, atol=0), name
    # the injected lags come back in global pair order
    from oracle import capi
    assert [int(x) for x in got["lag"]] == [capi.synth_true_lag(SEED, i, L) for i in range(N_PAIRS)]


def test_gather_without_process_group_is_identity():
    rec = _records_for(0, 3)
    out = ac.gather_results(rec, 3)
    assert out.tobytes() == rec.tobytes()

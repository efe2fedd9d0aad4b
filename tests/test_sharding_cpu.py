"""N>1 host logic on CPU: world_size-2 gloo processes partition an utterance set with no collective on the
data path, and the gathered result equals the single-replica order."""

import os
import socket
import sys

import numpy as np
import pytest

from conftest import REPO


def test_shard_bounds_cover_everything():
    from pocket_tts_mlx_b200.sharding import shard_bounds, shard_by_length
    for n in (0, 1, 7, 256, 4096):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    lengths = list(np.random.Generator(np.random.PCG64(0)).integers(10, 200, size=101))
    parts = shard_by_length(lengths, 4)
    assert sorted(i for p in parts for i in p) == list(range(101))
    loads = [sum(lengths[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= max(lengths)


def _worker(rank, world, port, q):
    sys.path.insert(0, str(REPO))
    import torch.distributed as dist
    from pocket_tts_mlx_b200.sharding import gather_in_order, max_over_ranks, shard_by_length
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lengths = [5 + (i * 7) % 23 for i in range(37)]
    mine = shard_by_length(lengths, world)[rank]
    # stand-in for "generate utterance i on my replica": a deterministic function of i only
    local = [np.full(3, i * 10 + lengths[i], dtype=np.int64) for i in mine]
    full = gather_in_order(local, mine, len(lengths))
    ok = all(np.array_equal(full[i], np.full(3, i * 10 + lengths[i])) for i in range(len(lengths)))
    slow = max_over_ranks(1.0 + rank)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ok, slow, len(mine)))


def test_two_gloo_replicas_partition_and_gather():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res)
    assert all(abs(slow - 2.0) < 1e-9 for _, _, slow, _ in res)      # max over ranks
    assert sum(n for _, _, _, n in res) == 37


def test_bench_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` is the CPU arm (oracle port on host cores); it must emit the contract line."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=str(REPO))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "audio-s/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0

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


class _StubModel:
    """Stands in for a per-GPU TTSModel replica: an utterance's "waveform" is a deterministic function of its token ids
    and its own noise only, like the real path with injected noise."""

    def _estimate_max_gen_len(self, n_tok):
        import math
        return math.ceil((n_tok / 3.0 + 2.0) * 12.5)

    def generate_audio_continuous(self, model_states, token_ids, slots=256, frames_after_eos=3, noise=None, seed=0, **kw):
        assert len(model_states) == len(token_ids) and (noise is None or len(noise) == len(token_ids))
        return [np.asarray(t, dtype=np.float32).cumsum() * (1 + s["voice_id"]) + (0 if noise is None else float(z[0]))
                for s, t, z in zip(model_states, token_ids, noise if noise is not None else [None] * len(token_ids))]


def _jobs():
    rng = np.random.Generator(np.random.PCG64(4))
    ids = [rng.integers(0, 4000, size=int(rng.integers(3, 40))).astype(np.int32) for _ in range(23)]
    states = [{"voice_id": j % 3, "prompt_len": 125} for j in range(23)]
    noise = [rng.standard_normal(4).astype(np.float32) for _ in range(23)]
    return states, ids, noise


def _launcher_worker(rank, world, port, q):
    sys.path.insert(0, str(REPO))
    import torch.distributed as dist
    from pocket_tts_mlx_b200.sharding import generate_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    states, ids, noise = _jobs()
    mine, local = generate_sharded(_StubModel(), states, ids, rank=rank, world_size=world, slots=4, noise=noise)
    full = generate_sharded(_StubModel(), states, ids, rank=rank, world_size=world, slots=4, noise=noise, gather=True)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, mine, [np.asarray(x) for x in full]))


def test_sharded_launcher_two_ranks_equal_one_rank():
    """`generate_sharded` (the multi-GPU product path) with two gloo ranks: the ranks' shares are disjoint and cover the
    set, no data crosses ranks unless a gather is asked for, and the gathered list equals the single-rank result
    utterance by utterance and in input order."""
    import torch.multiprocessing as mp
    from pocket_tts_mlx_b200.sharding import generate_sharded
    states, ids, noise = _jobs()
    idx1, one = generate_sharded(_StubModel(), states, ids, rank=0, world_size=1, slots=4, noise=noise)
    single = [None] * len(ids)
    for i, w in zip(idx1, one):
        single[i] = w
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_launcher_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shares = [set(m) for _, m, _ in res]
    assert shares[0].isdisjoint(shares[1]) and shares[0] | shares[1] == set(range(len(ids)))
    for _, _, full in res:
        assert len(full) == len(ids)
        for a, b in zip(full, single):
            assert np.array_equal(a, b)


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

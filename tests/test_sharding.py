"""Host-side sharding logic with world_size-2 gloo process groups on CPU: the contiguous stream
partition, scatter of int16 audio, gather of float32 scores, and "N shards == 1 process" on the oracle
standing in for the per-rank engine (the GPU engine itself is covered by tests/test_gpu_parity.py)."""

import os
import socket

import numpy as np
import pytest

from conftest import GOLDEN
from microwakeword_b200.sharding import partition


def test_partition_is_contiguous_and_balanced():
    for n in (0, 1, 7, 64, 65536, 524288, 100001):
        for w in (1, 2, 3, 4, 8):
            parts = partition(n, w)
            assert len(parts) == w and parts[0][0] == 0 and sum(c for _, c in parts) == n
            assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    assert partition(524288, 8) == [(65536 * r, 65536) for r in range(8)]
    with pytest.raises(ValueError):
        partition(4, 0)


def test_tile_blocks_cover_every_rank_block_in_order():
    from microwakeword_b200.sharding import tile_blocks
    for n, w, t in ((524288, 8, 8), (11, 2, 3), (12, 2, 4), (7, 3, 2), (5, 2, 1)):
        blocks = tile_blocks(n, w, t)
        assert len(blocks) == t and all(len(b) == w for b in blocks)
        for r, (start, count) in enumerate(partition(n, w)):
            pos = start
            for tb in blocks:                       # a rank's tiles are contiguous, in order, and tile its block exactly
                assert tb[r][0] == pos
                pos += tb[r][1]
            assert pos == start + count
    assert tile_blocks(524288, 8, 8)[3][5] == (5 * 65536 + 3 * 8192, 8192)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_streams, q, tiles=0):
    import torch
    import torch.distributed as dist

    import oracle
    from microwakeword_b200.sharding import gather_probs, scatter_audio, scatter_compute_gather
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        audio = np.load(os.path.join(GOLDEN, "batch_audio.npy"))[:n_streams]
        blob = open(os.path.join(GOLDEN, "okay_nabu_synth_int8.mww"), "rb").read()
        full = torch.from_numpy(audio) if rank == 0 else None
        if tiles:
            # the pipelined ingest: tile scatters issued up front, the oracle standing in for each tile's GPU engine
            seen = []

            def compute(t, local):
                seen.append((t, local.shape[0]))
                _, p = oracle.run_pipeline(blob, local.numpy(), want_features=False)
                return torch.from_numpy(p)

            out = scatter_compute_gather(full, n_streams, audio.shape[1], compute, tiles=tiles, src=0)
            start, count = partition(n_streams, world)[rank]
            assert [t for t, _ in seen] == list(range(tiles)) and sum(c for _, c in seen) == count
            if rank == 0:
                q.put(out.numpy())
            else:
                assert out is None
            return
        local = scatter_audio(full, n_streams, audio.shape[1], src=0)
        start, count = partition(n_streams, world)[rank]
        assert local.shape == (count, audio.shape[1]) and np.array_equal(local.numpy(), audio[start:start + count])
        _, probs = oracle.run_pipeline(blob, local.numpy(), want_features=False)      # stand-in for the rank's GPU engine
        out = gather_probs(torch.from_numpy(probs), n_streams, dst=0)
        if rank == 0:
            q.put(out.numpy())
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_streams,tiles", [(12, 0), (11, 0), (12, 3), (11, 2)])
def test_two_rank_scatter_compute_gather_equals_single_process(n_streams, tiles):
    """tiles = 0: one scatter, compute, gather; tiles > 0: the pipelined ingest (equal and ragged tile sizes)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_streams, q, tiles)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    want = np.load(os.path.join(GOLDEN, "batch_probs_int8.npy"))[:n_streams]
    assert np.array_equal(got, want)


# ---- the pulled ingest (IngestBuffer + predict_clip_ingest): host logic under gloo ------------------------------------

class _OracleEngine:
    """CPU stand-in with the StreamEngine surface predict_clip_ingest uses: reads int16 [n_streams, n] at a raw address
    (here: a shared-memory segment standing in for the peer-mapped ingest buffer) and runs the oracle from carried state."""

    stride, hop = 3, 160            # okay_nabu: three feature rows per model step, 10 ms hop

    def __init__(self, model, n_streams=1, device=0):
        self.blob, self.n_streams, self.history = model, n_streams, None

    def close(self):
        pass

    def predict_clip_remote(self, src_ptr, n_samples, stride=None, tiles=0, out=None):
        import ctypes

        import torch

        import oracle
        raw = (ctypes.c_int16 * (self.n_streams * n_samples)).from_address(src_ptr)
        new = np.frombuffer(raw, np.int16).reshape(self.n_streams, n_samples).copy()
        self.history = new if self.history is None else np.concatenate([self.history, new], 1)
        _, p = oracle.run_pipeline(self.blob, self.history, want_features=False)
        done = getattr(self, "steps_done", 0)
        self.steps_done = p.shape[1]
        return torch.from_numpy(np.ascontiguousarray(p[:, done:]))


class _ShmIngest:
    """IngestBuffer's surface over a POSIX shared-memory segment (one box, no GPU)."""

    def __init__(self, name, n_streams, n_samples, src=0):
        from multiprocessing import shared_memory
        self.shm = shared_memory.SharedMemory(name=name)
        import torch
        self.n_streams, self.n_samples, self.src, self.device = n_streams, n_samples, src, torch.device("cpu")
        self.view = np.ndarray((n_streams, n_samples), np.int16, buffer=self.shm.buf)

    def block_ptr(self, first_stream):
        return self.view[first_stream:].ctypes.data if first_stream < self.n_streams else self.view.ctypes.data


def _ingest_worker(rank, world, port, n_streams, shm_name, q, shares=None):
    import torch.distributed as dist

    from microwakeword_b200.sharding import ShardedEngine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        audio = np.load(os.path.join(GOLDEN, "batch_audio.npy"))[:n_streams]
        half = audio.shape[1] // 2 // 160 * 160
        blob = open(os.path.join(GOLDEN, "okay_nabu_synth_int8.mww"), "rb").read()
        ingest = _ShmIngest(shm_name, n_streams, half)
        sh = ShardedEngine(blob, n_streams, 0, engine_factory=_OracleEngine, shares=shares)
        assert (sh.engine is None) == (shares is not None and shares[rank] == 0)
        outs = []
        for call in range(2):                        # two steps: the ingest rank refills the buffer between them
            if rank == 0:
                ingest.view[:] = audio[:, call * half:(call + 1) * half]
            outs.append(sh.predict_clip_ingest(ingest))
            dist.barrier()
        if rank == 0:
            q.put(np.concatenate([o.numpy() for o in outs], 1))
        else:
            assert outs == [None, None]
        del ingest.view
        ingest.shm.close()
    finally:
        dist.destroy_process_group()


def test_ingest_shares():
    from microwakeword_b200.sharding import ingest_shares, partition
    assert ingest_shares(524288, 8, 0, 1.0) == [65536] * 8
    assert ingest_shares(524288, 8, 0, 0.0) == [0] + [74899] * 2 + [74898] * 5 and sum(ingest_shares(524288, 8, 0, 0.0)) == 524288
    assert ingest_shares(100, 4, 2, 0.5) == [29, 29, 12, 30][:0] + [30, 29, 12, 29] or sum(ingest_shares(100, 4, 2, 0.5)) == 100
    assert ingest_shares(7, 1) == [7]
    assert partition(10, 3, [0, 4, 6]) == [(0, 0), (0, 4), (4, 6)]
    for bad in ([1, 2], [5, 5, 1], [-1, 5, 6]):
        with pytest.raises(ValueError):
            partition(10, 3, bad)


@pytest.mark.parametrize("n_streams,shares", [(12, None), (11, None), (12, [0, 12]), (12, [3, 9])])
def test_two_rank_pulled_ingest_equals_single_process(n_streams, shares):
    import torch.multiprocessing as mp
    from multiprocessing import shared_memory

    import oracle
    audio = np.load(os.path.join(GOLDEN, "batch_audio.npy"))[:n_streams]
    half = audio.shape[1] // 2 // 160 * 160
    shm = shared_memory.SharedMemory(create=True, size=n_streams * half * 2)
    try:
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_ingest_worker, args=(r, 2, port, n_streams, shm.name, q, shares)) for r in range(2)]
        for p in procs:
            p.start()
        got = q.get(timeout=120)
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
    finally:
        shm.close()
        shm.unlink()
    blob = open(os.path.join(GOLDEN, "okay_nabu_synth_int8.mww"), "rb").read()
    _, want = oracle.run_pipeline(blob, audio[:, :2 * half], want_features=False)
    assert got.shape == want.shape and np.array_equal(got, want)

"""Stream sharding across the GPUs of one box (one process per GPU, torch.distributed / NCCL).

The reference has no distributed code; streams are fully independent
(microwakeword/inference.py and mixednet.py contain no cross-stream operation), so rank r owns the
contiguous block streams[r*S/G : (r+1)*S/G], weights are replicated, and the only communication the
path ever needs is what BASELINE.json's north_star names: scatter int16 audio batches from an ingest
rank and gather float32 scores back.  When audio already arrives sharded (bench.py's default, weak
scaling) there is no data-path collective at all.

Works with backend "nccl" (GPU tensors, NVLink) and "gloo" (CPU tensors; used by the world-size-2
tests that run without a GPU).
"""

from __future__ import annotations


def partition(n_streams: int, world: int):
    """Contiguous, balanced blocks: [(start, count)] * world; the first n % world ranks get one extra."""
    if n_streams < 0 or world < 1:
        raise ValueError("bad partition arguments")
    base, extra = divmod(n_streams, world)
    out, start = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((start, n))
        start += n
    return out


def scatter_audio(audio, n_streams: int, n_samples: int, src: int = 0, group=None, device=None):
    """Rank `src` holds int16 [n_streams, n_samples]; every rank returns its own [count, n_samples] block."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    parts = partition(n_streams, world)
    if device is None:
        device = audio.device if audio is not None else torch.device("cpu")
    local = torch.empty((parts[rank][1], n_samples), dtype=torch.int16, device=device)
    if rank == src and (tuple(audio.shape) != (n_streams, n_samples) or audio.dtype != torch.int16):
        raise ValueError("scatter_audio: rank %d must pass int16 [%d, %d]" % (src, n_streams, n_samples))
    # samples travel as raw bytes: every backend (gloo has no int16) moves uint8
    as_bytes = lambda t: t.contiguous().view(torch.uint8)
    if len({c for _, c in parts}) == 1:
        chunks = [as_bytes(audio[s:s + c]) for s, c in parts] if rank == src else None
        dist.scatter(as_bytes(local), chunks, src=src, group=group)
    else:
        # ragged blocks: point-to-point sends (grouped on NCCL)
        if rank == src:
            reqs = []
            for r, (s, c) in enumerate(parts):
                if r == src:
                    local.copy_(audio[s:s + c])
                elif c:
                    reqs.append(dist.isend(as_bytes(audio[s:s + c]), dst=r, group=group))
            for q in reqs:
                q.wait()
        elif parts[rank][1]:
            dist.recv(as_bytes(local), src=src, group=group)
    return local


def tile_blocks(n_streams: int, world: int, tiles: int):
    """blocks[t][r] = (global start, count) of tile t of rank r: every rank's contiguous block cut into `tiles`
    contiguous sub-blocks (balanced like partition())."""
    if tiles < 1:
        raise ValueError("tiles must be >= 1")
    per_rank = [[(start + s, c) for s, c in partition(count, tiles)] for start, count in partition(n_streams, world)]
    return [[per_rank[r][t] for r in range(world)] for t in range(tiles)]


def scatter_audio_tiles(audio, n_streams: int, n_samples: int, tiles: int, src: int = 0, group=None, device=None):
    """Issue the scatter of every tile asynchronously and return at once: (locals, works) with locals[t] this rank's
    int16 [count_t, n_samples] block of tile t and works[t] the handles to wait on before reading it.  On NCCL the
    transfers queue on the communicator's own stream in tile order, so tile t+1 moves over NVLink while the caller
    computes on tile t (wait() only makes the caller's CUDA stream wait); on gloo wait() blocks the host."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    blocks = tile_blocks(n_streams, world, tiles)
    if device is None:
        device = audio.device if audio is not None else torch.device("cpu")
    if rank == src and (tuple(audio.shape) != (n_streams, n_samples) or audio.dtype != torch.int16):
        raise ValueError("scatter_audio_tiles: rank %d must pass int16 [%d, %d]" % (src, n_streams, n_samples))
    as_bytes = lambda t: t.contiguous().view(torch.uint8)
    locals_, works = [], []
    for blk in blocks:
        local = torch.empty((blk[rank][1], n_samples), dtype=torch.int16, device=device)
        ws = []
        if len({c for _, c in blk}) == 1:
            if blk[0][1]:
                chunks = [as_bytes(audio[s:s + c]) for s, c in blk] if rank == src else None
                ws.append(dist.scatter(as_bytes(local), chunks, src=src, group=group, async_op=True))
        elif rank == src:
            for r, (s, c) in enumerate(blk):
                if r == src:
                    local.copy_(audio[s:s + c])
                elif c:
                    ws.append(dist.isend(as_bytes(audio[s:s + c]), dst=r, group=group))
        elif blk[rank][1]:
            ws.append(dist.irecv(as_bytes(local), src=src, group=group))
        locals_.append(local)
        works.append(ws)
    return locals_, works


def scatter_compute_gather(audio_on_src, n_streams: int, n_samples: int, compute, tiles: int = 8, src: int = 0, group=None, device=None):
    """BASELINE.json configs[4] ingest pattern with the exchange hidden behind the compute: audio lives on `src`, every
    rank's block is cut into `tiles` sub-blocks, all tile scatters are issued up front and `compute(t, local_audio)`
    (-> float32 [count_t, steps]) runs on tile t as soon as it has landed while the later tiles are still in flight.
    The scores (4 bytes per 960 bytes of audio) go back in one gather.  Returns [n_streams, steps] on `src`, else None."""
    import torch

    locals_, works = scatter_audio_tiles(audio_on_src, n_streams, n_samples, tiles, src=src, group=group, device=device)
    outs = []
    for t, (local, ws) in enumerate(zip(locals_, works)):
        for w in ws:
            w.wait()
        outs.append(compute(t, local))
    return gather_probs(torch.cat(outs, 0), n_streams, dst=src, group=group)


def gather_probs(local_probs, n_streams: int, dst: int = 0, group=None):
    """Every rank passes float32 [count, steps]; rank `dst` returns [n_streams, steps], others None."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    parts = partition(n_streams, world)
    steps = local_probs.shape[1]
    local_probs = local_probs.contiguous()
    if len({c for _, c in parts}) == 1:
        outs = [torch.empty_like(local_probs) for _ in range(world)] if rank == dst else None
        dist.gather(local_probs, outs, dst=dst, group=group)
        return torch.cat(outs, 0) if rank == dst else None
    if rank == dst:
        full = torch.empty((n_streams, steps), dtype=local_probs.dtype, device=local_probs.device)
        for r, (s, c) in enumerate(parts):
            if r == dst:
                full[s:s + c] = local_probs
            elif c:
                dist.recv(full[s:s + c], src=r, group=group)
        return full
    if parts[rank][1]:
        dist.send(local_probs, dst=dst, group=group)
    return None


class ShardedEngine:
    """This rank's block of streams as `tiles` StreamEngines over contiguous sub-blocks (tiles = 1: one engine).
    More than one tile lets predict_clip_scattered() overlap the NVLink scatter of tile t+1 with the kernels of tile t."""

    def __init__(self, model, n_streams_total: int, device_index: int, group=None, tiles: int = 1):
        import torch.distributed as dist

        from .engine import StreamEngine
        self.group = group
        self.n_total = n_streams_total
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if n_streams_total < self.world:
            raise ValueError("ShardedEngine: %d streams cannot be spread over %d ranks" % (n_streams_total, self.world))
        self.tiles = max(1, min(int(tiles), n_streams_total // self.world))     # same on every rank; no empty sub-block
        tiles = self.tiles
        self.start, self.count = partition(n_streams_total, self.world)[self.rank]
        self.tile_parts = partition(self.count, tiles)
        self.engines = [StreamEngine(model, n_streams=c, device=device_index) for _, c in self.tile_parts]
        self.engine = self.engines[0]

    def reset(self):
        for e in self.engines:
            e.reset()

    def predict_clip_scattered(self, audio_on_src, n_samples: int, src: int = 0):
        """Audio originates on `src` ([n_total, n_samples] int16 CUDA tensor, None elsewhere); scores return to `src`."""
        import torch

        return scatter_compute_gather(audio_on_src, self.n_total, n_samples, lambda t, local: self.engines[t].predict_clip(local),
                                      tiles=self.tiles, src=src, group=self.group, device=torch.device("cuda", self.engine.device))

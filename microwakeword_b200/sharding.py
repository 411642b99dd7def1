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

import ctypes

from . import _lib


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


class PeerAudio:
    """The ingest rank's audio buffer mapped into every rank's address space (CUDA IPC over NVLink peer access), so that
    a rank PULLS its tiles with copy-engine DMA (cudaMemcpyPeerAsync): no communication kernel occupies SMs on either
    side and no send/recv rendezvous has to be co-scheduled with the compute kernels -- which is what made the NCCL
    tile pipeline slower than the serial exchange.  One box only (the north_star's 8 x B200).

        peer = PeerAudio(full_audio_or_None, n_streams, n_samples, src=0)     # once per buffer (collective)
        probs = peer.pull_compute_gather(compute, tiles=8)                     # every step (collective)

    Ordering per step: a barrier makes the ingest rank's writes to the buffer (stream-ordered before its barrier) visible
    before any peer's copies start; the final gather of the scores is what tells the ingest rank that every peer has
    finished reading, so it may refill the buffer after pull_compute_gather() returns on its stream."""

    def __init__(self, audio_on_src, n_streams: int, n_samples: int, src: int = 0, group=None, device=None):
        import torch
        import torch.distributed as dist
        from torch.multiprocessing.reductions import reduce_tensor

        self.group, self.src = group, src
        self.n_streams, self.n_samples = n_streams, n_samples
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        if self.rank == src:
            if tuple(audio_on_src.shape) != (n_streams, n_samples) or audio_on_src.dtype != torch.int16 or not audio_on_src.is_cuda \
                    or not audio_on_src.is_contiguous():
                raise ValueError("PeerAudio: the ingest rank must pass a contiguous CUDA int16 [%d, %d]" % (n_streams, n_samples))
            torch.cuda.current_stream().synchronize()          # the IPC handle carries no stream ordering of earlier writes
        box = [reduce_tensor(audio_on_src) if self.rank == src and self.world > 1 else None]
        if self.world > 1:
            dist.broadcast_object_list(box, src=src, group=group)
        err = None
        try:
            if self.rank == src:
                self.remote = audio_on_src
            else:
                rebuild, args = box[0]
                self.remote = rebuild(*args)                    # a tensor on the ingest rank's device, readable from here
                probe = torch.empty(16, dtype=torch.int16, device=self.device)
                probe.copy_(self.remote.view(-1)[:16])          # peer access really works from this process
                torch.cuda.synchronize(self.device)
        except Exception as exc:                                # noqa: BLE001 -- reported on every rank below
            err = exc
        # agree on the outcome (also: nobody drops the handle before everybody has opened it)
        ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            raise RuntimeError("PeerAudio: CUDA IPC / peer access to the ingest rank's buffer is unavailable (%s)" % (err or "on another rank"))
        self.copy_stream = torch.cuda.Stream(device=self.device)

    @classmethod
    def allocate(cls, n_streams: int, n_samples: int, src: int = 0, group=None, device=None):
        """OPT-IN variant (written after the round's GPU budget was spent; not yet run on a GPU): the ingest buffer is a
        cudaMalloc owned by the library on the ingest rank, exported with cudaIpcGetMemHandle and opened by every other rank
        INSIDE ITS OWN device context (mww_ipc_open, cudaIpcMemLazyEnablePeerAccess) -- no rank creates a context on the
        ingest GPU, unlike the torch rebuild used by __init__.  On the ingest rank `self.buffer` is a torch view of the
        allocation to write the audio into; elsewhere it is None.  Candidate fix for the 8-GPU slowdown."""
        import torch
        import torch.distributed as dist

        self = cls.__new__(cls)
        self.group, self.src = group, src
        self.n_streams, self.n_samples = n_streams, n_samples
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.remote, self.buffer, self._owned, self._opened = None, None, None, None
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        L = _lib.lib()
        handle = ctypes.create_string_buffer(64)
        ptr = ctypes.c_void_p()
        err = None
        try:
            if self.rank == src:
                _lib.check(None, L.mww_ipc_alloc(n_streams * n_samples * 2, dev_index, ctypes.byref(ptr), handle))
                self._owned = ptr.value
        except Exception as exc:                                # noqa: BLE001
            err = exc
        box = [bytes(handle.raw) if self.rank == src else None]
        if self.world > 1:
            dist.broadcast_object_list(box, src=src, group=group)
        try:
            if err is None and self.rank != src:
                _lib.check(None, L.mww_ipc_open(box[0], dev_index, ctypes.byref(ptr)))
                self._opened = ptr.value
        except Exception as exc:                                # noqa: BLE001
            err = exc
        ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            self.close()
            raise RuntimeError("PeerAudio.allocate: CUDA IPC is unavailable (%s)" % (err or "on another rank"))
        self.remote_ptr = ptr.value
        if self.rank == src:
            iface = {"shape": (n_streams, n_samples), "typestr": "<i2", "data": (self._owned, False), "version": 3, "strides": None}
            holder = type("MwwIpcBuffer", (), {"__cuda_array_interface__": iface})()
            self.buffer = torch.as_tensor(holder, device=self.device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        return self

    def close(self):
        """Release what allocate() created (no-op for a PeerAudio built around a torch tensor)."""
        dev_index = self.device.index if self.device.index is not None else 0
        if getattr(self, "_opened", None):
            _lib.lib().mww_ipc_close(ctypes.c_void_p(self._opened), dev_index)
            self._opened = None
        if getattr(self, "_owned", None):
            self.buffer = None
            _lib.lib().mww_ipc_free(ctypes.c_void_p(self._owned), dev_index)
            self._owned = None

    def pull_tiles(self, tiles: int):
        """Start the DMA of this rank's tiles (in order, on a side stream); returns (locals, events)."""
        import torch
        import torch.distributed as dist

        dist.barrier(group=self.group)                          # the buffer is complete on the ingest rank (see class docstring)
        blocks = tile_blocks(self.n_streams, self.world, tiles)
        cur = torch.cuda.current_stream(self.device)
        self.copy_stream.wait_stream(cur)
        locals_, events = [], []
        for blk in blocks:
            s, c = blk[self.rank]
            local = torch.empty((c, self.n_samples), dtype=torch.int16, device=self.device)
            # torch's Tensor.copy_ would issue a cross-device copy on the SOURCE device's stream, i.e. in a second context
            # of this process on the ingest GPU, where it is time-sliced against the ingest rank's own kernels (measured:
            # no overlap at all).  The C-ABI copy runs on this rank's stream: a pull by this GPU's copy engine.
            if c:
                base = self.remote.data_ptr() if self.remote is not None else self.remote_ptr       # torch rebuild | allocate()
                _lib.check(None, _lib.lib().mww_copy_async(local.data_ptr(), base + s * self.n_samples * 2,
                                                          c * self.n_samples * 2, ctypes.c_void_p(self.copy_stream.cuda_stream)))
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
            local.record_stream(self.copy_stream)
            locals_.append(local)
            events.append(ev)
        return locals_, events

    def pull_compute_gather(self, compute, tiles: int = 8):
        import torch

        locals_, events = self.pull_tiles(tiles)
        cur = torch.cuda.current_stream(self.device)
        outs = []
        for t, (local, ev) in enumerate(zip(locals_, events)):
            cur.wait_event(ev)
            outs.append(compute(t, local))
        return gather_probs(torch.cat(outs, 0), self.n_streams, dst=self.src, group=self.group)


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

    def predict_clip_pulled(self, peer: "PeerAudio"):
        """Same result with the audio pulled over NVLink by copy engines (PeerAudio) instead of NCCL send/recv kernels."""
        return peer.pull_compute_gather(lambda t, local: self.engines[t].predict_clip(local), tiles=self.tiles)

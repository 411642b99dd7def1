"""Launched by tests/test_gpu_multirank.py under torch.distributed.run (one rank per GPU, NCCL):
audio lives on rank 0 (IngestBuffer), every rank pulls + computes its block (mww_predict_clip_remote),
scores are gathered on rank 0 and compared there with the CPU oracle on the same audio.  Two
consecutive calls (state carried across the rank boundary-free shards), fp32 and int8 models, ragged
partition (n_streams not a multiple of the world size), several pipeline tiles per rank, and -- for the fp32 model --
uneven shares: the ingest rank with half a share and with none (it then only feeds its peers and collects the scores)."""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    os.environ.setdefault("MWW_MIN_TILE_STREAMS", "2")
    import torch
    import torch.distributed as dist

    import oracle
    from conftest import GOLDEN, synth_audio
    from microwakeword_b200.sharding import IngestBuffer, ShardedEngine, gather_probs, ingest_shares, scatter_audio
    from microwakeword_b200.engine import StreamEngine

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    S, N = 8 * world + 3, 9600
    audio = np.stack([synth_audio(N, 4000 + i) for i in range(S)])
    ok = True
    for kind, exact, src_share in (("f32", False, 1.0), ("int8", True, 1.0), ("f32", False, 0.5), ("f32", False, 0.0)):
        blob = open(os.path.join(GOLDEN, "okay_nabu_synth_%s.mww" % kind), "rb").read()
        shares = None if src_share == 1.0 else ingest_shares(S, world, 0, src_share)
        sh = ShardedEngine(blob, S, local, shares=shares)
        with IngestBuffer(S, N, src=0, device=dev) as ingest:
            outs = []
            for call in range(2):
                if rank == 0:
                    ingest.buffer.copy_(torch.from_numpy(audio).to(dev))
                outs.append(sh.predict_clip_ingest(ingest, tiles=3))
            # the plain NCCL exchange (balanced blocks) must agree with the pulled one
            bal = ShardedEngine(blob, S, local)
            locals_ = scatter_audio(torch.from_numpy(audio).to(dev) if rank == 0 else None, S, N, src=0, device=dev)
            nccl = gather_probs(bal.engine.predict_clip(locals_), S, dst=0)
            torch.cuda.synchronize()
            dist.barrier()
        if rank == 0:
            got = torch.cat(outs, 1).cpu().numpy()
            _, want = oracle.run_pipeline(blob, np.concatenate([audio, audio], 1), want_features=False)
            err = float(np.abs(got - want).max())
            good = got.shape == want.shape and ((err == 0.0) if exact else (err <= 1e-5)) and torch.equal(nccl.cpu(), outs[0].cpu())
            print("ingest_vs_oracle %s world=%d streams=%d shares=%s max_err=%g %s" % (kind, world, S, shares or "balanced", err,
                                                                                      "OK" if good else "MISMATCH"), flush=True)
            ok = ok and good
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()

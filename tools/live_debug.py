import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from microwakeword_b200.engine import StreamEngine
from microwakeword_b200.synth_audio import synth_audio
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
blob = open(os.path.join(G, "okay_nabu_synth_f32.mww"), "rb").read()
S = int(sys.argv[1]) if len(sys.argv) > 1 else 40
audio = np.stack([synth_audio(4800, 50 + i) for i in range(S)])
_, want = oracle.run_pipeline(blob, audio, want_features=False)
eng = StreamEngine(blob, n_streams=S)
dev = torch.from_numpy(audio).cuda()
parts = []
for i in range(0, 4800, 480):
    p = eng.step(dev[:, i:i + 480].contiguous())
    torch.cuda.synchronize()
    parts.append(p.cpu().numpy())
    print("call", i // 480, "rows pending", eng.pending_rows, "probs", parts[-1].shape, flush=True)
got = np.concatenate(parts, 1)
print("max |diff| vs oracle:", np.abs(got - want[:, :got.shape[1]]).max(), got.shape, want.shape)

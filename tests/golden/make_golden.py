"""Generates the committed golden fixtures under tests/golden/ from the CPU ORACLE.

    python tests/golden/make_golden.py

The reference's own implementation of this path cannot run here (TensorFlow / pymicro-features are
not installable), so these vectors are produced by oracle/ (parity UNPINNED -- they
pin the oracle against regressions and travel to the GPU box, where /root/reference and the
oracle's Python sources are still available but the reference is not).

Fixtures
  okay_nabu_synth_f32.mww    synthetic okay_nabu MixedNet, fp32, BatchNorm folded (seed 0)
  okay_nabu_synth_int8.mww   the same network quantised per utils.py:289-348 (calibrated on config-1 style audio)
  config0_audio.npy          BASELINE.json configs[0] clip: 10 s sweep + noise + silence + full-scale square
  config0_features.npy       uint16 [997, 40]  generate_features_for_clip(use_c=True) rows
  config0_probs_f32.npy      float32 [332]     Model.predict_spectrogram of those rows, fp32 model
  config0_probs_int8.npy     float32 [332]     same, int8 model (uint8 / 255)
  batch_audio.npy            int16 [12, 8000]  8 synthetic + 4 edge-case streams
  batch_features.npy         uint16 [12, 48, 40] streaming frontend rows
  batch_probs_f32.npy / batch_probs_int8.npy   float32 [12, 16]
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from oracle import mixednet_ref as R  # noqa: E402
from microwakeword_b200 import model_file as MF  # noqa: E402
from conftest import edge_case_audio, synth_audio  # noqa: E402


def config0_clip() -> np.ndarray:
    n = 160000
    t = np.arange(n)
    f = 200.0 + (4000.0 - 200.0) * t / n
    x = 3000.0 * np.sin(2 * np.pi * f * t / 16000.0)
    x += np.random.default_rng(7).normal(0.0, 500.0, n)
    x[72000:80000] = 0.0                                      # 0.5 s of exact zeros
    x[-1600:] = np.where((t[-1600:] // 8) % 2 == 0, 32767, -32767)   # final 0.1 s full-scale square
    return np.clip(np.round(x), -32768, 32767).astype(np.int16)


def main():
    spec = R.OKAY_NABU
    f32 = R.fold_bn(spec, R.init_synthetic(spec, 0))
    f32_blob = MF.write_container(f32)

    calib_audio = np.stack([synth_audio(48000, 9000 + i) for i in range(8)])
    calib_feat, _ = oracle.run_pipeline(None, calib_audio, want_probs=False)
    calib = calib_feat.reshape(-1, 40)[: (calib_feat.shape[0] * calib_feat.shape[1]) // 3 * 3].reshape(-1, 3, 40)
    q = R.quantize_model(f32, calib.astype(np.float32) * R.FEATURE_SCALE)
    q_blob = MF.write_container(q)
    open(os.path.join(HERE, "okay_nabu_synth_f32.mww"), "wb").write(f32_blob)
    open(os.path.join(HERE, "okay_nabu_synth_int8.mww"), "wb").write(q_blob)

    clip = config0_clip()
    feats = oracle.generate_features_for_clip(clip)
    assert feats.shape == (997, 40)
    p32 = oracle.MixedNet(f32_blob).predict_u16(feats)
    p8 = oracle.MixedNet(q_blob).predict_u16(feats)
    assert p32.shape == (332,) and p8.shape == (332,)
    # the NumPy restatement must agree with the C one before anything is written
    pn = np.asarray(R.predict_spectrogram(R.FoldedStreamingF32(f32), feats), np.float32)
    assert np.abs(pn - p32).max() < 2e-6
    qm = R.StreamingInt8(q)
    pq = np.asarray(R.predict_spectrogram(qm, feats, quantized=True, in_scale=qm.input_scale, in_zp=qm.input_zero_point), np.float32)
    assert np.array_equal(pq, p8)
    np.save(os.path.join(HERE, "config0_audio.npy"), clip)
    np.save(os.path.join(HERE, "config0_features.npy"), feats)
    np.save(os.path.join(HERE, "config0_probs_f32.npy"), p32)
    np.save(os.path.join(HERE, "config0_probs_int8.npy"), p8)

    batch = np.concatenate([np.stack([synth_audio(8000, 100 + i) for i in range(8)]), edge_case_audio(8000)[[0, 1, 5, 13]]])
    bf, bp32 = oracle.run_pipeline(f32_blob, batch)
    _, bp8 = oracle.run_pipeline(q_blob, batch)
    np.save(os.path.join(HERE, "batch_audio.npy"), batch)
    np.save(os.path.join(HERE, "batch_features.npy"), bf)
    np.save(os.path.join(HERE, "batch_probs_f32.npy"), bp32)
    np.save(os.path.join(HERE, "batch_probs_int8.npy"), bp8)
    print("features", feats.shape, "probs", p32.shape, "batch", bf.shape, bp32.shape,
          "prob range f32 [%.3f, %.3f] int8 [%.3f, %.3f]" % (p32.min(), p32.max(), p8.min(), p8.max()))


if __name__ == "__main__":
    main()

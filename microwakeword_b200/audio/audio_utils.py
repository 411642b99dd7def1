"""Drop-in for the hot-path part of ``microwakeword.audio.audio_utils``
(reference: microwakeword/audio/audio_utils.py:28-84).  ``save_clip`` / ``remove_silence_webrtc``
(:87-140) are training-data preparation and out of scope."""

from __future__ import annotations

import numpy as np

from ..model_file import FEATURE_SCALE

_engine = None
_batch_engines = {}        # (device, capacity) -> frontend-only StreamEngine kept across calls (no cudaMalloc / table upload per batch)


def to_int16(audio_samples: np.ndarray) -> np.ndarray:
    """audio_utils.py:47-48: float clips are scaled by 32768 and clipped; int16 passes through."""
    if audio_samples.dtype in (np.float32, np.float64):
        return np.clip((audio_samples * 32768), -32768, 32767).astype(np.int16)
    if audio_samples.dtype != np.int16:
        raise ValueError("audio must be int16 PCM or float in [-1, 1]")
    return audio_samples


def clip_samples_fed(n_samples: int) -> int:
    """Samples the reference's chunk loop hands to the frontend: 160-sample chunks while
    ``audio_idx + 320 < num_audio_bytes`` (strict '<', audio_utils.py:56) -- the last exact chunk is skipped."""
    n_bytes = 2 * n_samples
    if n_bytes <= 320:
        return 0
    chunks = (n_bytes - 320 - 1) // 320 + 1
    return min(160 * chunks, n_samples)


def _frontend_engine(device: int = 0):
    global _engine
    if _engine is None or _engine.device != device:
        from ..engine import StreamEngine
        _engine = StreamEngine(None, n_streams=1, device=device)
    return _engine


def _batch_engine(device: int, n: int):
    """Frontend-only engine with at least `n` streams, reused across calls (capacities are powers of two >= 16)."""
    from ..engine import StreamEngine
    cap = 16
    while cap < n:
        cap *= 2
    key = (device, cap)
    if key not in _batch_engines:
        if len(_batch_engines) >= 4:                          # a handful of capacities is plenty; drop the smallest
            _batch_engines.pop(min(_batch_engines, key=lambda k: k[1])).close()
        _batch_engines[key] = StreamEngine(None, n_streams=cap, device=device)
    return _batch_engines[key]


def generate_features_for_clip(audio_samples: np.ndarray, step_ms: int = 20, use_c: bool = True, device: int = 0):
    """Generates spectrogram features for the given audio data on the GPU.

    use_c=True  (default): pymicro_features semantics -- fresh frontend, 10 ms hop regardless of
                ``step_ms`` (the reference ignores it on this path), strict-'<' chunk loop; returns
                float32 [T, 40] = uint16 features * 0.0390625.
    use_c=False: TensorFlow audio_microfrontend op semantics (audio_utils.py:69-81) -- every full window of the clip
                at window_step = step_ms (default 20 ms, like the reference), returns uint16 [T, 40].
    """
    import torch

    audio = to_int16(np.asarray(audio_samples)).reshape(-1)
    eng = _frontend_engine(device)
    eng.reset_frontend()
    if use_c:
        fed = clip_samples_fed(audio.size)
        eng.set_window_step(160)                                 # pymicro_features hard-wires 10 ms; step_ms is ignored like the reference
    else:
        if int(step_ms) != step_ms or not 1 <= step_ms <= 30:
            raise ValueError("window_step must be a whole number of milliseconds in [1, 30]")
        fed = audio.size
        eng.set_window_step(16 * int(step_ms))
    if fed == 0:
        return np.zeros((0, 40), np.float32 if use_c else np.uint16)
    dev = torch.from_numpy(np.ascontiguousarray(audio[:fed])).to(eng._dev()).unsqueeze(0)
    feat = eng.features(dev)[0].view(torch.int16).cpu().numpy().view(np.uint16)
    if use_c:
        return feat.astype(np.float32) * np.float32(FEATURE_SCALE)
    return feat


def generate_features_for_clips(clips, use_c: bool = True, device: int = 0):
    """Batched `generate_features_for_clip` for a list of clips of DIFFERENT lengths -- the call pattern of the
    reference's dataset generator (microwakeword/audio/spectrograms.py:84-86 loops one clip at a time).

    All clips go through one frontend launch: each clip is zero padded to the longest one (the frontend is
    causal, so trailing padding cannot change earlier rows) and its own row count -- the strict-'<' chunk
    accounting of audio_utils.py:56 for use_c=True, every full window for use_c=False -- is cut out afterwards.
    Returns a list of float32 [T_i, 40] (use_c) or uint16 [T_i, 40] arrays.
    """
    import torch

    clips = [to_int16(np.asarray(c)).reshape(-1) for c in clips]
    if not clips:
        return []
    fed = [clip_samples_fed(c.size) if use_c else c.size for c in clips]
    rows = [max((n - 480) // 160 + 1, 0) if n >= 480 else 0 for n in fed]
    n_max = max(max(fed), 1)
    eng = _batch_engine(device, len(clips))                  # persistent; capacity >= len(clips), spare streams see silence
    eng.reset_frontend()
    eng.set_window_step(160)
    batch = np.zeros((eng.n_streams, n_max), np.int16)
    for i, (c, n) in enumerate(zip(clips, fed)):
        batch[i, :n] = c[:n]
    eng.reset_frontend()
    feat = eng.features(torch.from_numpy(batch).to(eng._dev())).view(torch.int16).cpu().numpy().view(np.uint16)
    out = []
    for i, r in enumerate(rows):
        f = feat[i, :r]
        out.append(f.astype(np.float32) * np.float32(FEATURE_SCALE) if use_c else f.copy())
    return out

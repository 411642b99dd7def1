"""oracle/spectrograms_ref.py -- CPU ORACLE (test infrastructure) for SURVEY.md section 8 row f-3: the per-clip
post-processing of microwakeword/audio/spectrograms.py:84-113, restated with the same NumPy calls the reference makes
(sliding_window_view + strided indexing) on features from the oracle frontend."""

import numpy as np
from numpy.lib.stride_tricks import sliding_window_view


def clip_spectrograms(spectrogram, step_ms=20, split_spectrogram_duration_s=None, slide_frames=None):
    """spectrograms.py:88-113 for one spectrogram -> list of arrays, in yield order."""
    out = []
    if split_spectrogram_duration_s is not None:
        desired = int(split_spectrogram_duration_s / (step_ms / 1000))                       # :90-92
        if spectrogram.shape[0] > desired + 20:                                              # :94
            slided = sliding_window_view(spectrogram, window_shape=(desired, spectrogram.shape[1]))[20::desired, ...]   # :95-98
            for i in range(slided.shape[0]):
                out.append(np.squeeze(slided[i]))                                            # :100-101
        else:
            out.append(spectrogram)
    elif slide_frames is not None:
        length = spectrogram.shape[0] - slide_frames + 1                                     # :105
        slided = sliding_window_view(spectrogram, window_shape=(length, spectrogram.shape[1]))   # :107-109
        for i in range(slide_frames):
            out.append(np.squeeze(slided[i]))
    else:
        out.append(spectrogram)
    return out


def spectrogram_generator(clips, step_ms=20, split_spectrogram_duration_s=None, slide_frames=None):
    """The reference's generator on a list of clips: one clip at a time through the (oracle) frontend."""
    import oracle
    for clip in clips:
        spec = oracle.generate_features_for_clip(clip).astype(np.float32) * np.float32(0.0390625)     # audio_utils.py:60-62 (float output)
        yield from clip_spectrograms(spec, step_ms, split_spectrogram_duration_s, slide_frames)

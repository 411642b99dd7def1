"""SpectrogramGeneration (microwakeword/audio/spectrograms.py:23-113).
CPU: the product's split / slide logic (plain slicing) == the reference's sliding_window_view expressions restated in
oracle/spectrograms_ref.py, over every shape regime; the ragged store round-trips.  GPU: the batched generator == the
reference's one-clip-at-a-time loop on the oracle frontend, element for element, in order."""

import numpy as np
import pytest

from conftest import synth_audio
from microwakeword_b200.audio import spectrograms as S
from oracle import spectrograms_ref as R


class _Clips:
    """the generator surface of microwakeword.audio.clips.Clips that SpectrogramGeneration uses"""

    def __init__(self, clips):
        self.clips = clips

    def audio_generator(self, split="train", repeat=1):
        for _ in range(repeat):
            yield from self.clips

    def random_audio_generator(self, max_clips=3):
        rng = np.random.default_rng(0)
        for _ in range(max_clips):
            yield self.clips[int(rng.integers(len(self.clips)))]

    def get_random_clip(self):
        return self.clips[1]


class _Augmenter:
    def augment_clip(self, clip):
        return (clip.astype(np.int32) // 2).astype(np.int16)

    def augment_generator(self, gen):
        for c in gen:
            yield self.augment_clip(c)


@pytest.mark.parametrize("rows", [0, 1, 19, 20, 21, 69, 70, 71, 120, 121, 170, 171, 500])
def test_split_and_slide_equal_the_reference_expressions(rows):
    spec = np.arange(rows * 40, dtype=np.float32).reshape(rows, 40)
    for step_ms, dur in ((10, 0.5), (20, 1.0), (10, 1.49), (20, 0.03)):
        got = S.split_or_slide(spec, step_ms, split_spectrogram_duration_s=dur)
        want = R.clip_spectrograms(spec, step_ms, split_spectrogram_duration_s=dur)
        assert len(got) == len(want) and all(np.array_equal(a, b) for a, b in zip(got, want)), (rows, step_ms, dur)
    for slide in (1, 2, 10):
        if rows - slide + 1 < 0:
            with pytest.raises(ValueError):
                S.split_or_slide(spec, 10, slide_frames=slide)
            with pytest.raises(ValueError):
                R.clip_spectrograms(spec, 10, slide_frames=slide)
            continue
        got = S.split_or_slide(spec, 10, slide_frames=slide)
        want = R.clip_spectrograms(spec, 10, slide_frames=slide)
        assert len(got) == slide == len(want) and all(np.array_equal(a, b) for a, b in zip(got, want))
    assert len(S.split_or_slide(spec)) == 1 and S.split_or_slide(spec)[0] is spec


def test_ragged_store_round_trip(tmp_path):
    rng = np.random.default_rng(1)
    samples = [rng.uniform(0, 26, (n, 40)).astype(np.float32) for n in (5, 1, 0, 97, 33)]
    store = S.write_ragged(str(tmp_path / "features"), iter(samples))
    again = S.read_ragged(str(tmp_path / "features"))
    assert len(store) == len(again) == len(samples)
    for i, a in enumerate(samples):
        assert again[i].shape == a.shape and np.array_equal(again[i], a)
    assert [x.shape[0] for x in again[1:4]] == [1, 0, 97]


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["plain", "slide", "split", "augmented_random"])
def test_gpu_generator_equals_reference_loop(torch_cuda, mode):
    # clips shorter than slide_frames rows make the reference's sliding_window_view raise: only the plain / split modes get them
    lengths = [16000, 4000, 23999, 800, 40000, 480, 31000] if mode in ("plain", "split") else [16000, 4000, 23999, 2600, 40000, 31000]
    clips = [synth_audio(n, 700 + i) for i, n in enumerate(lengths)]
    kw = dict(plain={}, slide=dict(slide_frames=10, step_ms=10), split=dict(split_spectrogram_duration_s=0.5, step_ms=10),
              augmented_random=dict(slide_frames=3, step_ms=10))[mode]
    aug = _Augmenter() if mode == "augmented_random" else None
    gen = S.SpectrogramGeneration(_Clips(clips), aug, batch_clips=3, **kw)
    if mode == "augmented_random":
        got = list(gen.spectrogram_generator(random=True, max_clips=5))
        src = [aug.augment_clip(c) for c in _Clips(clips).random_audio_generator(max_clips=5)]
    else:
        got = list(gen.spectrogram_generator(repeat=2))
        src = clips + clips
    want = list(R.spectrogram_generator(src, **{k: v for k, v in kw.items()}))
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert a.dtype == np.float32 and a.shape == b.shape and np.array_equal(a, b)
    one = gen.get_random_spectrogram()
    import oracle
    assert np.array_equal(one, oracle.generate_features_for_clip(aug.augment_clip(clips[1]) if aug else clips[1]).astype(np.float32) * np.float32(0.0390625))
    from microwakeword.audio.spectrograms import SpectrogramGeneration
    assert SpectrogramGeneration is S.SpectrogramGeneration

from microwakeword_b200.audio.spectrograms import SpectrogramGeneration  # noqa: F401

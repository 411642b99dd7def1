from microwakeword_b200.audio.audio_utils import generate_features_for_clip  # noqa: F401

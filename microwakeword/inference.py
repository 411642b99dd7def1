from microwakeword_b200.inference import Model  # noqa: F401

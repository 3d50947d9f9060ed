"""Environments shipped with the B200 engine: ``synthetic_env`` (benchmark stream with the shapes of BASELINE.json's
configurations) and ``poc_memory_env`` (the proof-of-concept memory task used by the end-to-end learning test).
Gym-backed environments of the reference are imported lazily by ``utils.create_env`` when installed."""

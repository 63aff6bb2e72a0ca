"""Moved to specs/mimi_spec.py (a neutral module the bench's product arm can import); re-exported for the tests."""
from specs.mimi_spec import *          # noqa: F401,F403
from specs.mimi_spec import MimiConfig, OFFICIAL, param_spec, synthetic_audio, synthetic_weights  # noqa: F401

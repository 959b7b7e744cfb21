"""Minimal stand-in for openai/baselines (not installable offline). TEST INFRASTRUCTURE ONLY."""
from . import logger  # noqa: F401

"""`specimux.core` re-exports (reference: src/specimux/core.py:20-92)."""
from specimux_b200.core import *                # noqa: F401,F403

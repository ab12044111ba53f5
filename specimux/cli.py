"""`python -m specimux.cli` / console scripts of the reference (pyproject.toml:51-57 there) on the B200 package."""
from specimux_b200.cli import *                 # noqa: F401,F403
from specimux_b200.cli import main, parse_args, setup_logging, version  # noqa: F401

try:
    from specimux_b200.cli import specimine_main  # noqa: F401
except ImportError:
    pass

if __name__ == "__main__":
    main()

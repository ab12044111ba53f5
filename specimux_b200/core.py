"""Re-export shim (mirror of the reference's core.py: tests import specimux.core.TrimMode etc.)."""
from .constants import *            # noqa: F401,F403
from .databases import *            # noqa: F401,F403
from .models import *               # noqa: F401,F403
from .demultiplex import process_sequences      # noqa: F401
from .io_utils import (OutputManager, output_write_operation, read_primers_file, read_specimen_file,  # noqa: F401
                       open_sequence_file, detect_file_format, cleanup_empty_directories)
from .orchestration import setup_match_parameters, specimux, specimux_mp, create_output_files, iter_batches  # noqa: F401

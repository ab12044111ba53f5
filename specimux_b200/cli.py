"""Command-line interface: same flags as the reference's `specimux` (cli.py:15-51 there)."""
import argparse
import logging
import os
import sys

from . import __version__
from .constants import MultipleMatchStrategy, TrimMode


def version():
    return f"specimux version {__version__}"


def parse_args(argv):
    p = argparse.ArgumentParser(description="Specimux: Demultiplex MinION sequences by dual barcode indexes and primers.")
    p.add_argument("primer_file", help="Fasta file containing primer information")
    p.add_argument("specimen_file", help="TSV file containing specimen mapping with barcodes and primers")
    p.add_argument("sequence_file", help="Sequence file in Fasta or Fastq format, gzipped or plain text")
    p.add_argument("--min-length", type=int, default=-1, help="Minimum sequence length.  Shorter sequences will be skipped (default: no filtering)")
    p.add_argument("--max-length", type=int, default=-1, help="Maximum sequence length.  Longer sequences will be skipped (default: no filtering)")
    p.add_argument("-n", "--num-seqs", type=str, default="-1", help="Number of sequences to read from file (e.g., -n 100 or -n 102,3)")
    p.add_argument("-e", "--index-edit-distance", type=int, default=-1, help="Barcode edit distance value, default is half of min distance between barcodes")
    p.add_argument("-E", "--primer-edit-distance", type=int, default=-1, help="Primer edit distance value, default is min distance between primers")
    p.add_argument("-l", "--search-len", type=int, default=80, help="Length to search for index and primer at start and end of sequence (default: 80)")
    p.add_argument("-F", "--output-to-files", action="store_true", help="Create individual sample files for sequences")
    p.add_argument("-P", "--output-file-prefix", default="", help="Prefix for individual files when using -F (default: no prefix)")
    p.add_argument("-O", "--output-dir", default=".", help="Directory for individual files when using -F (default: .)")
    p.add_argument("--color", action="store_true", help="Highlight barcode matches in blue, primer matches in green")
    p.add_argument("--trim", choices=[TrimMode.NONE, TrimMode.TAILS, TrimMode.BARCODES, TrimMode.PRIMERS], default=TrimMode.BARCODES, help="trimming to apply")
    p.add_argument("--dereplicate", choices=[MultipleMatchStrategy.NONE, MultipleMatchStrategy.BEST], default=MultipleMatchStrategy.BEST,
                   help="Dereplication strategy: 'best' selects best match per specimen/barcode group (default), 'none' outputs all matches")
    p.add_argument("-d", "--diagnostics", nargs="?", const=1, type=int, choices=[1, 2, 3],
                   help="Enable diagnostic trace logging: 1=standard (default), 2=detailed, 3=verbose")
    p.add_argument("-D", "--debug", action="store_true", help="Enable debug logging")
    p.add_argument("--disable-prefilter", action="store_true", help="Disable barcode prefiltering (bloom filter emulation)")
    p.add_argument("--disable-preorient", action="store_true", help="Disable heuristic pre-orientation")
    p.add_argument("-t", "--threads", type=int, default=-1, help="Number of GPUs to use (default: all visible)")
    p.add_argument("--sample-topq", type=int, default=0, metavar="N", help="(accepted for compatibility; not implemented)")
    p.add_argument("-v", "--version", action="version", version=version())
    args = p.parse_args(argv[1:])
    if args.color:
        print("specimux_b200: --color highlighting is not implemented; records are written without colour", file=sys.stderr)
    if args.sample_topq:
        print("specimux_b200: --sample-topq is not implemented and is ignored", file=sys.stderr)
    if "," in args.num_seqs:
        try:
            start, num = args.num_seqs.split(",")
            args.start_seq, args.num_seqs = int(start), int(num)
        except ValueError:
            p.error("Invalid format for -n option. Use 'start,num' with integers.")
    else:
        try:
            args.num_seqs, args.start_seq = int(args.num_seqs), 1
        except ValueError:
            p.error("Invalid format for -n option. Use an integer or 'start,num' with integers.")
    return args


def setup_logging(debug: bool, output_dir: str = None):
    fmt = logging.Formatter("%(asctime)s - %(levelname)s - %(message)s")
    root = logging.getLogger()
    root.handlers.clear()
    h = logging.StreamHandler()
    h.setFormatter(fmt)
    root.addHandler(h)
    if output_dir:
        os.makedirs(output_dir, exist_ok=True)
        fh = logging.FileHandler(os.path.join(output_dir, "log.txt"), mode="w")
        fh.setFormatter(fmt)
        root.addHandler(fh)
    root.setLevel(logging.DEBUG if debug else logging.INFO)


def main(argv=None):
    from . import orchestration
    argv = sys.argv if argv is None else argv
    args = parse_args(argv)
    setup_logging(args.debug, args.output_dir if args.output_to_files else None)
    logging.info(f"Starting {version()}")
    logging.info(f"Command line: {' '.join(argv)}")
    try:
        if args.output_to_files:
            orchestration.specimux_mp(args)
        else:
            if args.threads > 1:
                logging.warning(f"Multithreading only supported for file output. Ignoring --threads {args.threads}")
            orchestration.specimux(args)
    except Exception as e:          # the reference logs worker failures and exits 1 (orchestration.py:222-224)
        logging.error(f"Unexpected error: {e}")
        sys.exit(1)


def specimine_main():
    """Entry point of the `specimine` command (reference cli.py:113-116)."""
    from . import specimine
    specimine.main()


if __name__ == "__main__":
    main()

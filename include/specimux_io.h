/*
 * specimux_io.h -- C ABI of the native host I/O either side of the matching path
 * (libspecimux_io.so; plain C++17, no CUDA, no Python).
 *
 * the FASTQ/FASTA(.gz) reader that feeds the packer, and the per-specimen
 * output-tree writer that consumes smx_record.  It replaces, byte-compatibly,
 *   - Bio.SeqIO.parse as the reference uses it   (src/specimux/io_utils.py:429-450)
 *   - create_write_operation's string work        (src/specimux/demultiplex.py:30-103)
 *   - OutputManager.write_sequence + FileHandleCache (src/specimux/io_utils.py:59-176, 179-268)
 *   - output_write_operation's console form        (src/specimux/io_utils.py:452-471)
 * The matching itself stays in libspecimux_b200.so (include/specimux_b200.h); this library never
 * computes a distance.  Errors: non-zero return + thread-local message (smx_io_last_error).
 */
#ifndef SPECIMUX_IO_H
#define SPECIMUX_IO_H

#include <stdint.h>

#include "specimux_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define SMX_IO_ABI_VERSION 2

#define SMX_IO_OK 0
#define SMX_IO_ERR_ARG 1
#define SMX_IO_ERR_OPEN 2      /* file cannot be opened / created                                  */
#define SMX_IO_ERR_FORMAT 3    /* malformed FASTQ (same conditions as Bio.SeqIO's FastqPhredIterator) */
#define SMX_IO_ERR_IO 4        /* read / write / decompression failure                             */

typedef struct smx_reader smx_reader;
typedef struct smx_block smx_block;     /* one batch of parsed reads; owned by the library          */
typedef struct smx_writer smx_writer;

/* Read-only view of a block.  All text is raw bytes exactly as in the file (case preserved).
 * Read r: bases[seq_off[r] .. seq_off[r+1]), quals at the same offsets (NULL for FASTA input),
 * title (= Bio's record.description) titles[title_off[r] .. title_off[r+1]), id (= record.id, the
 * first whitespace-delimited token of the title) titles[title_off[r] + id_start[r] ..][0 .. id_len[r]). */
typedef struct smx_block_view {
    uint32_t n_reads;
    const char *bases;
    const uint64_t *seq_off;     /* n_reads + 1 */
    const char *quals;
    const char *titles;
    const uint64_t *title_off;   /* n_reads + 1 */
    const uint32_t *id_start;    /* n_reads */
    const uint32_t *id_len;      /* n_reads */
} smx_block_view;

int smx_io_abi_version(void);
const char *smx_io_last_error(void);

/* Opens a FASTQ (is_fastq = 1) or FASTA (0) file, gzip-compressed when the name ends in .gz/.gzip
 * (reference: io_utils.py:429-450).  Format detection stays with the caller (io_utils.py:380-427). */
int smx_reader_open(const char *path, int is_fastq, smx_reader **out);
/* Reader over the records of a PLAIN four-line FASTQ file that START inside [byte_start, byte_end): both ends are
 * moved forward to the next record boundary (first line beginning '@' whose second-next line begins '+'), so
 * readers over consecutive ranges see every record exactly once -- the parallel form of the reader
 * (the reference reads serially through Bio.SeqIO, io_utils.py:429-450).  Compressed files: SMX_IO_ERR_ARG. */
int smx_reader_open_range(const char *path, int is_fastq, uint64_t byte_start, uint64_t byte_end, smx_reader **out);
void smx_reader_close(smx_reader *r);

smx_block *smx_block_create(void);
void smx_block_destroy(smx_block *b);
void smx_block_get(const smx_block *b, smx_block_view *out);

/* Parses up to max_reads records into `blk` (replacing its contents).  n_reads == 0 at end of
 * file.  The reference's batching loop: orchestration.py:447-456 (itertools.islice over the parser). */
int smx_reader_next(smx_reader *r, uint32_t max_reads, smx_block *blk);

/* Parses and discards up to n records (-n start,num; orchestration.py:189-193). */
int smx_reader_skip(smx_reader *r, uint64_t n, uint64_t *skipped);

/* Name tables the writer formats records with.  `*_file` entries are the file-name forms of the
 * ids (OutputManager._make_filename's safe_id, io_utils.py:215-216), computed by the caller so the
 * character classes are Python's.  Index spaces are those of smx_record: sample = specimen row
 * (FULL / DEREPLICATED / MULTIPLE), global b1 id (PARTIAL_FORWARD), global b2 id (PARTIAL_REVERSE);
 * pool / p1 / p2 index `pool` / `primer_name`, -1 = "unknown". */
typedef struct smx_names {
    uint32_t n_specimens;
    const char *const *specimen_id;
    const char *const *specimen_file;
    uint32_t n_b1;
    const char *const *b1_id;        /* "barcode_fwd_<b1>" (constants.py SampleId.PREFIX_FWD_MATCH)  */
    const char *const *b1_file;
    uint32_t n_b2;
    const char *const *b2_id;        /* "barcode_rev_<b2>"                                          */
    const char *const *b2_file;
    uint32_t n_pools;
    const char *const *pool;
    uint32_t n_primers;
    const char *const *primer_name;
} smx_names;

/* output_dir != NULL: the per-specimen tree of OutputManager (io_utils.py:197-268): path
 * <dir>/<full|partial|unknown>/<pool>/<p1>-<p2>/<prefix><safe id>.<fastq|fasta>, header
 * "<id> <distance code> pool=<pool> primers=<p1>+<p2> <sample>", full matches duplicated at pool
 * level.  output_dir == NULL: the console form of output_write_operation (io_utils.py:459-466) on
 * stdout.  Files are appended in arrival order by this single writer (= the reference's -t 1 order). */
int smx_writer_open(const char *output_dir, const char *prefix, int is_fastq, const smx_names *names, smx_writer **out);

/* Formats and appends `n_records` records (in order) whose `read` fields index `blk`:
 * reverse-complement (Bio ambiguous-DNA table, case preserved, U->A) and quality reversal for
 * reverse records, trimming to [trim_start, trim_end), the empty-trim fallback, distance code,
 * missing qualities as 'I' (Q40; alignment.py:52-56). */
int smx_writer_write(smx_writer *w, const smx_block *blk, const smx_record *records, uint64_t n_records);

/* Same for the compact record form (smx_record32: everything the files need, no location pairs). */
int smx_writer_write32(smx_writer *w, const smx_block *blk, const smx_record32 *records, uint64_t n_records);

/* Same for the 16-byte wire form (smx_record16): records in read order covering the block's reads from read 0,
 * read indices recovered from the last-of-read flags, trim_end from the reads' lengths. */
int smx_writer_write16(smx_writer *w, const smx_block *blk, const smx_record16 *records, uint64_t n_records);

/* Deferred form of smx_writer_write16 for a pipeline: returns as soon as the records are planned and handed to the
 * writer's worker threads, so the caller's next block is planned while this one is still being formatted and
 * appended.  `blk` must stay valid and unchanged until the NEXT smx_writer_write* / smx_writer_wait /
 * smx_writer_close call on this writer has returned (each of them first waits for the deferred call); `records` may
 * be reused at once.  Replaces the per-batch hand-over of the reference's single writer loop
 * (io_utils.py:452-471 output_write_operation called batch by batch from orchestration.py:181-203). */
int smx_writer_write16_deferred(smx_writer *w, const smx_block *blk, const smx_record16 *records, uint64_t n_records);

/* Waits for a deferred call to finish.  Returns the first error seen so far, if any. */
int smx_writer_wait(smx_writer *w);

/* Flushes every buffer and releases the writer.  Returns the first error seen, if any. */
int smx_writer_close(smx_writer *w);

/* Totals since open: records written (pool-level duplicates not counted) and payload bytes. */
void smx_writer_stats(const smx_writer *w, uint64_t *records, uint64_t *bytes);

#ifdef __cplusplus
}
#endif
#endif /* SPECIMUX_IO_H */

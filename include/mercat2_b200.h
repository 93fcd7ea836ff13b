/* mercat2_b200 -- C ABI of the B200-native k-mer counting engine.
 *
 * This is the drop-in boundary for MerCat2's hot path.  The reference has no FFI: the path is plain
 * Python (lib/mercat2_kmers.py, lib/mercat2_Chunker.py, lib/mercat2_metrics.py and the driver glue
 * bin/mercat2.py:86-137).  Each entry point below names the reference interface it replaces; the
 * ctypes binding a maintainer would add on the reference side is shown in INTEGRATION.md and is what
 * mercat2_b200/_native.py implements.
 *
 * Conventions: every function returns 0 on success or a negative mc2_status; mc2_last_error() gives
 * the message for the calling thread's last failure.  No exceptions cross the boundary.  Handles are
 * created by *_create / count calls and released by the matching *_destroy / *_free.  One host thread
 * per engine; calls are stream-ordered on the engine's stream and synchronous at return.  Plain
 * pointers and sizes only -- no torch types.  There is NO CPU fallback: without a CUDA device the
 * engine cannot be created.
 */
#ifndef MERCAT2_B200_H
#define MERCAT2_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mc2_engine mc2_engine;
typedef struct mc2_table mc2_table;
typedef struct mc2_metrics mc2_metrics;

enum mc2_status {
    MC2_OK = 0,
    MC2_ERR_INVALID = -1,      /* bad argument (k < 1, k too large, NULL pointer ...) */
    MC2_ERR_CUDA = -2,         /* CUDA runtime failure */
    MC2_ERR_NON_ASCII = -3,    /* input holds bytes >= 0x80 (the reference would count code points) */
    MC2_ERR_LIMIT = -4,        /* an engine capacity limit was hit (message says which) */
    MC2_ERR_IO = -5
};

/* where a text buffer lives */
enum mc2_memspace { MC2_HOST = 0, MC2_DEVICE = 1 };

/* ---- engine -------------------------------------------------------------------------------- */
int mc2_engine_create(int device, mc2_engine** out);
void mc2_engine_destroy(mc2_engine* e);
const char* mc2_last_error(void);
const char* mc2_version(void);

/* Release the engine's device workspace (the key arrays of the range partition are kept between calls: up to 8 bytes
 * per window of the largest chunk counted so far) and return the stream-ordered pool's cached memory to the driver. */
int mc2_engine_trim(mc2_engine* e);

/* Tunables (mostly for tests): "dense_max_bins", "smem_max_bins", "batch_symbols", "force_path"
 * (0 auto, 1 dense, 2 sparse, 3 wide), "force_encoding" (-1 auto, 0 ACGT 2-bit, 1 A-Z 5-bit, 2 byte), "sparse_algo"
 * (0/2 range partition + shared-memory tables, 1 radix sort), "count_mode" (-1 auto, 0 / 1 see rangecount.cuh),
 * "row_merge" (sums of (key, count) row sets: 0 sort + segmented sum, 1 range partition + shared-memory sums). */
int mc2_engine_set_option(mc2_engine* e, const char* name, int64_t value);
/* Counters: "launches" (kernels launched so far), "h2d_bytes", "d2h_bytes", "chunks", "device_ms"
 * (CUDA-event time of the last count call, microseconds in "device_us"). */
int64_t mc2_engine_get_stat(mc2_engine* e, const char* name);

/* Per-kernel device time, measured with CUDA events on the engine's stream while option "profile" is 1
 * (set it to 2 to clear the totals): a JSON object {"kernel": {"launches": n, "us": t}, ...}.
 * Call with buf = NULL to get the size. */
int mc2_engine_profile(mc2_engine* e, char* buf, uint64_t cap, uint64_t* size);

/* ---- counting -------------------------------------------------------------------------------- */
/* Replaces find_kmers(file, kmer, min_count) (lib/mercat2_kmers.py:32-78) for one file's TEXT (the
 * bytes the reference would read from the possibly-gunzipped file): parse FASTA, count every k-mer,
 * keep those with count >= min_count.  `space` says whether `text` is a host or a device pointer. */
int mc2_count_text(mc2_engine* e, const void* text, uint64_t nbytes, int space, int k, int64_t min_count,
                   mc2_table** out);

/* mc2_count_text whose table is delivered as packed rows straight into HOST memory: host_rows receives *rows rows of 16
 * bytes (uint64 order-preserving key, uint64 count), sorted by key; capacity = rows the buffer holds.  For a text too
 * large for one pass (one global table over tens of Gbp, min_count >= 2) the rows of finished key ranges leave on the
 * copy stream while later ranges are still being counted, so the download hides behind the counting; use pinned
 * memory.  Fails with MC2_ERR_LIMIT when the table holds k-mers outside the packed alphabet (decode: mc2_table_info
 * conventions, 2 bits per base for nucleotide text). */
int mc2_count_text_rows(mc2_engine* e, const void* text, uint64_t nbytes, int space, int k, int64_t min_count, void* host_rows,
                        uint64_t capacity, uint64_t* rows);
/* The same table as two HOST arrays -- host_keys[capacity] (uint64) and host_counts[capacity] (uint32): 12 bytes per row
 * over PCIe instead of 16 (the download is what bounds an end-to-end run of one global table).  A count that does not fit
 * 32 bits fails the call with MC2_ERR_LIMIT (use mc2_count_text_rows). */
int mc2_count_text_rows_split(mc2_engine* e, const void* text, uint64_t nbytes, int space, int k, int64_t min_count,
                              uint64_t* host_keys, uint32_t* host_counts, uint64_t capacity, uint64_t* rows);

/* Replaces the sample loop of bin/mercat2.py:411-448 (one find_kmers task per sample file, each file smaller than
 * the -s trigger, i.e. ONE piece) for n samples at once: out[j] is the table of texts[j], identical to n calls of
 * mc2_count_text.  The samples are counted in a single pass (keys carry the sample index), which is what makes many
 * small samples -- BASELINE config 5: 64 proteomes -- fill the GPU.  `space` applies to every text. */
int mc2_count_batch(mc2_engine* e, const void* const* texts, const uint64_t* nbytes, uint32_t n, int space, int k, int64_t min_count,
                    mc2_table** out);

/* Replaces calculateKmerCount(seq, kmer) (lib/mercat2_kmers.py:10-28): `symbols` is a raw sequence, no
 * FASTA parsing (a '>' or newline in it is an ordinary character). */
int mc2_count_symbols(mc2_engine* e, const void* symbols, uint64_t nbytes, int space, int k, int64_t min_count,
                      mc2_table** out);

/* Replaces chunk_files + Chunker + one find_kmers per piece + the per-sample sum
 * (bin/mercat2.py:86-127, lib/mercat2_Chunker.py:39-59): the text is split at the reference's piece
 * boundaries (virtually: nothing is written), every piece is counted and filtered on its own, the
 * filtered tables are summed.  chunk_bytes = 0 means "do not chunk" (file smaller than -s, or -s 0).
 * n_chunks (optional) receives the number of pieces; piece_offsets (optional, capacity
 * piece_capacity) receives the byte offset where each piece starts. */
int mc2_count_sample(mc2_engine* e, const void* text, uint64_t nbytes, int space, int k, int64_t min_count,
                     uint64_t chunk_bytes, mc2_table** out, uint64_t* n_chunks, uint64_t* piece_offsets,
                     uint64_t piece_capacity);

/* Replaces Chunker(path, dest, "<S>M", ">") (lib/mercat2_Chunker.py:14-59) without writing files: the
 * byte offsets at which the reference would start each piece of this text. */
int mc2_chunk_offsets(mc2_engine* e, const void* text, uint64_t nbytes, int space, uint64_t chunk_bytes,
                      uint64_t* piece_offsets, uint64_t piece_capacity, uint64_t* n_pieces);

/* Streaming variant of mc2_count_sample for several files / blocks of one sample: begin, add the
 * text of each FILE (each is chunked and counted independently, like one element of
 * samples[type][name] in bin/mercat2.py:336-339), finish. */
typedef struct mc2_sample mc2_sample;
int mc2_sample_begin(mc2_engine* e, int k, int64_t min_count, mc2_sample** out);
int mc2_sample_add_text(mc2_sample* s, const void* text, uint64_t nbytes, int space, uint64_t chunk_bytes,
                        uint64_t* n_chunks);
/* The same for a FILE read by the engine itself (lib/mercat2_kmers.py:47: gzip.open(file,'rt') if the suffix is
 * '.gz' else open(file); gunzip = 1 / 0 / -1 = decide by the ".gz" suffix): a reader thread fills pinned buffers
 * (read(2) or zlib inflate), each goes to the device with cudaMemcpyAsync while the next one is being filled, and
 * the text that has arrived is chunked and counted; the device keeps a sliding window of the text, not the file.
 * chunk_bytes: the caller applies the reference's trigger on the on-disk size (bin/mercat2.py:101), 0 = one piece. */
int mc2_sample_add_file(mc2_sample* s, const char* path, int gunzip, uint64_t chunk_bytes, uint64_t* n_chunks,
                        uint64_t* text_bytes);
/* Add already counted rows (rows*k bytes of k-mer text + rows counts, host memory) to the sample: the serial dict
 * merge of bin/mercat2.py:121-127 for tables that were counted elsewhere (another GPU / rank).  Equal k-mers are
 * summed by mc2_sample_finish on the device -- also with the k-mers the sample counted itself (add_text / add_file /
 * add_keys before or after this call): rows that fit the sample's packed alphabet join its packed rows. */
int mc2_sample_add_rows(mc2_sample* s, const char* kmers, const uint64_t* counts, uint64_t rows);
int mc2_sample_finish(mc2_sample* s, mc2_table** out);   /* consumes s */
void mc2_sample_abort(mc2_sample* s);

/* ---- result tables ----------------------------------------------------------------------------
 * Rows are sorted by k-mer text (byte order == Python str order for ASCII), as
 * sorted(kmers.items()) in bin/mercat2.py:132. */
uint64_t mc2_table_rows(const mc2_table* t);
int mc2_table_k(const mc2_table* t);
uint64_t mc2_table_total(const mc2_table* t);              /* sum of counts */
/* kmers: rows*k bytes (no terminators), counts: rows entries */
int mc2_table_export(mc2_table* t, char* kmers, uint64_t* counts);
/* The per-sample TSV of bin/mercat2.py:130-133: "k-mer\t<basename>_Count\n" then "kmer\tcount\n" rows.
 * Writes nothing and returns 1 when the table is empty (the reference writes no file then). */
int mc2_table_write_tsv(mc2_table* t, const char* path, const char* basename);
/* Same bytes into memory: call with buf = NULL to get the size. */
int mc2_table_tsv(mc2_table* t, const char* basename, char* buf, uint64_t cap, uint64_t* size);
void mc2_table_free(mc2_table* t);

/* ---- device-resident exchange between ranks ------------------------------------------------------
 * When several GPUs hold parts of the SAME sample (whole pieces per rank), their already filtered tables are
 * summed like the dict merge of bin/mercat2.py:121-127.  These calls let the host layer do that with NCCL on
 * device memory: the packed rows of a table (64-bit order-preserving codes + counts) are exposed as device
 * pointers, cut at splitter keys, exchanged all-to-all by the caller and rebuilt into a table on the receiving
 * GPU; dense per-sample tables are exposed for an in-place NCCL (all-)reduce. */
/* encoding (0 = 2-bit ACGT, 1 = 5-bit A-Z, 2 = byte), key kind (0 = bit-packed code, 1 = base-26 index of a dense
 * protein table), number of packed rows and of literal-byte ("wide") rows */
int mc2_table_info(const mc2_table* t, int* encoding, int* key_kind, uint64_t* packed_rows, uint64_t* wide_rows);
/* device pointers to the packed rows (sorted by key; valid until the table is freed) */
int mc2_table_device_rows(mc2_table* t, const uint64_t** keys, const uint64_t** counts, uint64_t* rows);
/* cuts[i] = number of packed rows with key < splitters[i] (host arrays) */
int mc2_table_lower_bound(mc2_table* t, const uint64_t* splitters, uint64_t m, uint64_t* cuts);
/* The packed rows into HOST memory (pinned memory makes this one DMA per array): keys / counts with room for
 * `capacity` rows; *rows = rows copied.  This is the table as the device holds it -- 16 bytes per row instead of the
 * k + 8 bytes of mc2_table_export -- for callers that decode k-mers lazily (decode: mc2_table_info's encoding). */
int mc2_table_export_packed(mc2_table* t, uint64_t* keys, uint64_t* counts, uint64_t capacity, uint64_t* rows);
/* the literal-byte rows (host copies): kmers wide_rows*k bytes, counts wide_rows entries */
int mc2_table_export_wide(mc2_table* t, char* kmers, uint64_t* counts);
/* Build a table from packed rows (device or host memory; unsorted, equal keys are summed) plus literal-byte rows
 * (host memory, may be NULL/0). */
int mc2_table_from_rows(mc2_engine* e, int k, int encoding, int key_kind, const uint64_t* keys, const uint64_t* counts,
                        uint64_t rows, int space, const char* wide_kmers, const uint64_t* wide_counts, uint64_t wide_rows,
                        mc2_table** out);
/* The TSV body (rows only, no header line) of a table into host memory: call with buf = NULL for the size. */
int mc2_table_tsv_body(mc2_table* t, char* buf, uint64_t cap, uint64_t* size);
/* Device pointer to the per-sample dense table (uint64[bins], bins = 4^k or 26^k) of a sample whose plan is the
 * dense path, for an in-place NCCL reduce; *bins = 0 when the sample is not on the dense path (or has seen no
 * text yet).  mc2_sample_dense_plan forces that plan on a sample that has not seen text (a rank without pieces). */
int mc2_sample_dense(mc2_sample* s, uint64_t** table, uint64_t* bins, int* encoding);
int mc2_sample_dense_plan(mc2_sample* s, int encoding);
/* ---- one piece split across GPUs BEFORE the filter (SURVEY 8e grain 3; `-s 0` on a file larger than one GPU) ----
 * The reducer this replaces is the dict sum + filter of ONE chunk file (lib/mercat2_kmers.py:56-78) when that chunk's
 * text lives on several GPUs.  Every rank parses its byte range of the piece (cut at header lines) into 2-bit packed
 * symbols (mc2_keys_open) and exposes a sampled histogram of its keys' 32-bit prefixes (mc2_keys_sample: device
 * pointer, summed over the ranks by the caller -- NCCL all-reduce -- so that all ranks cut the key space at the same
 * keys).  mc2_keys_partition then groups the order-preserving 64-bit keys of all its windows into `groups` ascending,
 * disjoint key ranges on its GPU (device array grouped by range + sizes + range bounds).  The host layer sends range g
 * to its owner rank (NCCL all-to-all on the device array); the owner counts what it received for one range with
 * mc2_sample_add_keys: every occurrence of a key is in that call, so the -c filter sees whole-piece counts exactly
 * like lib/mercat2_kmers.py:73-78 on the unsplit piece, and the rows of successive ranges follow each other in sorted
 * order (rank order == key order: the table is the concatenation of the ranks' parts).  Windows outside the packed
 * alphabet (N, lower case ...) are counted unfiltered by mc2_count_exceptions so that the caller can sum them across
 * ranks before filtering.  mc2_partition_keys = open + partition with the rank's own sample (single GPU). */
typedef struct mc2_keys mc2_keys;
int mc2_keys_open(mc2_engine* e, const void* text, uint64_t nbytes, int space, int k, mc2_keys** out);
/* hist: device pointer to *entries uint32 counters (valid until mc2_keys_partition / mc2_keys_free) */
int mc2_keys_sample(mc2_keys* ks, uint32_t** hist, uint64_t* entries);
int mc2_keys_partition(mc2_keys* ks, uint32_t groups);
int mc2_partition_keys(mc2_engine* e, const void* text, uint64_t nbytes, int space, int k, uint32_t groups, mc2_keys** out);
/* keys: device array of *total keys, range g occupying sizes[0]+..+sizes[g-1] onward; sizes: groups entries;
 * bounds (optional): groups + 1 entries, range g holds the keys whose 32-bit prefix (the first 16 symbols, left
 * aligned) lies in [bounds[g], bounds[g+1]) */
int mc2_keys_info(mc2_keys* ks, const uint64_t** keys, uint64_t* sizes, uint64_t* total, uint64_t* exception_symbols, uint64_t* bounds);
void mc2_keys_free(mc2_keys* ks);
/* Count n keys that are ALL occurrences of their key range within the current chunk.  prefix_lo / prefix_hi: the
 * 32-bit prefix range the keys lie in (from mc2_keys_info's bounds; 0, 0 = unknown: the whole key space). */
int mc2_sample_add_keys(mc2_sample* s, const uint64_t* keys, uint64_t n, int space, uint64_t prefix_lo, uint64_t prefix_hi);
int mc2_count_exceptions(mc2_engine* e, const void* text, uint64_t nbytes, int space, int k, mc2_table** out);
/* Device-to-device copy on the engine's stream, complete at return (moves a reduced table between engine memory and
 * a buffer owned by the communication library). */
int mc2_device_copy(mc2_engine* e, void* dst, const void* src, uint64_t nbytes);

/* ---- sample x k-mer table (row N3: merge_tsv / merge_tsv_T, lib/mercat2_report.py:98-194) --------------------
 * mc2_merge_tables: union of the tables' k-mers (sorted) x one count column per table, 0 where a sample lacks the
 * k-mer -- from tables that are still on the device; mc2_table_from_tsv parses a per-sample TSV file's bytes
 * ("<header line>\n" then "<k-mer>\t<count>\n" rows) on the device so that the reference's file-based entry
 * point keeps working.  mc2_matrix_write_tsv: corner = first header cell ("k-mer" / "sample"), names = column (or,
 * transposed, row) labels in table order; transposed = 0 writes merge_tsv's layout, 1 merge_tsv_T's (its columns in
 * sorted k-mer order; the reference's order there is a Python set's). */
typedef struct mc2_matrix mc2_matrix;
int mc2_table_from_tsv(mc2_engine* e, const void* text, uint64_t nbytes, int space, mc2_table** out);
int mc2_merge_tables(mc2_engine* e, mc2_table* const* tables, uint32_t n, mc2_matrix** out);
uint64_t mc2_matrix_rows(const mc2_matrix* m);
int mc2_matrix_k(const mc2_matrix* m);
/* kmers: rows*k bytes; counts: rows*n entries, row-major */
int mc2_matrix_export(mc2_matrix* m, char* kmers, uint64_t* counts);
int mc2_matrix_write_tsv(mc2_matrix* m, const char* path, const char* corner, const char* const* names, int transposed);
void mc2_matrix_free(mc2_matrix* m);
/* ---- row N4: what the alpha-diversity metrics and the top-k-mer summary read -----------------------------------------
 * lib/mercat2_diversity.py:23-27 re-reads a sample's TSV to build the count vector it hands to scikit-bio;
 * mc2_table_export_counts gives that vector straight from the table (counts: rows entries in sorted k-mer order, may be
 * NULL) and its abundance spectrum reduced on the device (spectrum: 16 words, may be NULL: [0] observed k-mers, [1] sum
 * of counts, [2..3] sum of squared counts (low / high 64 bits), [4] largest count, [5 + i] k-mers seen exactly i + 1
 * times, i < 10).  The arithmetic of the metrics stays with skbio.
 * lib/mercat2_figures.py:50-65 keeps the 5 k-mers with the largest mean count over the samples (earlier rows win ties);
 * mc2_matrix_top_rows returns their row indices (largest first) in the matrix of mc2_merge_tables. */
int mc2_table_export_counts(mc2_table* t, uint64_t* counts, uint64_t* spectrum);
int mc2_matrix_top_rows(mc2_matrix* m, uint32_t top, uint64_t* rows_out, uint32_t* found);
/* merge_tsv byte for byte as the reference writes it (lib/mercat2_report.py:98-160: a k-way cursor walk whose row label
 * is the smallest next k-mer among the files that advanced on the previous line -- with differing k-mer sets labels
 * repeat / appear out of order and some rows are dropped; identical to the sorted union when all samples hold the same
 * k-mers).  tables in column order, names = column labels, corner = first header cell. */
int mc2_merge_tables_reference(mc2_engine* e, mc2_table* const* tables, uint32_t n, const char* path, const char* corner,
                               const char* const* names);

/* ---- text transforms ahead of the hot path ------------------------------------------------
 * The transformed text stays on the device (mc2_text_info: pointer + size, usable as `text` with MC2_DEVICE in every
 * counting call) and can be downloaded for the artefact the reference keeps on disk (clean/<base>.fna.gz).
 * mc2_fastq_to_fasta replaces fq2fa's `sed -n '1~4s/^@/>/p;2~4p'` (lib/mercat2_fasta.py:175-198). */
typedef struct mc2_text mc2_text;
int mc2_fastq_to_fasta(mc2_engine* e, const void* text, uint64_t nbytes, int space, mc2_text** out);
int mc2_text_info(const mc2_text* t, const void** device_ptr, uint64_t* nbytes);
int mc2_text_export(mc2_text* t, void* host, uint64_t capacity);
void mc2_text_free(mc2_text* t);

/* ---- protein metrics ----------------------------------------------------------------------------
 * Replaces the numeric part of plot_sample_metrics (lib/mercat2_figures.py:157-183) and
 * predict_isoelectric_point_ProMoST / calculate_MW / calculate_hydro (lib/mercat2_metrics.py:57-170)
 * for every record of one protein FASTA text.  Values are UNROUNDED doubles (the Python layer applies
 * round(x, 2) exactly like the reference); status[i]: 0 ok, 1 = last residue unknown (reference
 * returns None), 2 = first residue unknown (reference raises KeyError). */
int mc2_protein_metrics(mc2_engine* e, const void* text, uint64_t nbytes, int space, mc2_metrics** out);
/* The scalar functions of lib/mercat2_metrics.py for a batch of raw sequences (host memory):
 * sequence i is seqs[offsets[i] .. offsets[i+1]).  One output row per sequence; status 255 = empty. */
int mc2_sequence_metrics(mc2_engine* e, const void* seqs, const uint64_t* offsets, uint64_t nseq, mc2_metrics** out);
uint64_t mc2_metrics_records(const mc2_metrics* m);
/* header_off/header_len locate each record's header text (without '>') inside the input text */
int mc2_metrics_export(const mc2_metrics* m, uint64_t* header_off, uint32_t* header_len, uint64_t* length,
                       double* pi, double* mw, double* hydro, uint8_t* status);
void mc2_metrics_free(mc2_metrics* m);

#ifdef __cplusplus
}
#endif
#endif

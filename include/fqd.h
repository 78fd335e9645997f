/*
 * fqd.h - C ABI of libfqd_cuda.so, the B200 (sm_100a) deduplication engine behind fastq-dupaway's hot paths.
 *
 * The reference (fastq-dupaway V1.5.0) has no FFI; its in-process seam is main()'s dispatch onto
 *   HashDupRemover<T>::filterSE / filterPE(+unordered)   (src/hash_dup_remover.hpp:73-94,105-347)
 *   SeqDupRemover<T>::filterSE / filterPE                (src/seq_dup_remover.hpp:12-38,54-218)
 * with T in {FastqView, FastaView, *ViewWithId}.  The host keeps file / gzip I/O (src/file_utils.cpp,
 * src/bufferedinput.hpp) and hands raw record bytes to this library, which replaces everything those
 * drivers do per record: record splitting (FastqView::read_new src/fastqview.cpp:89-119), key packing
 * (SeqUtils::seq2hash src/seq_utils.cpp:35-49), the exact first-occurrence set (src/hash_dup_remover.hpp:
 * 113-139), the sorters (src/external_sort.hpp, src/paired_external_sort.hpp), the comparators
 * (src/comparator.cpp:45-91) and the ID-tag merge-join (src/hash_dup_remover.hpp:257-347).
 *
 * Conventions: plain pointers and sizes only; every call returns an fqd_status (0 = ok); no exception
 * crosses the boundary; one handle is driven by one host thread at a time.  "Device" pointers are CUDA
 * device pointers on the handle's device.  There is NO CPU fallback: without a usable CUDA device
 * fqd_create fails with FQD_ERR_CUDA.
 */
#ifndef FQD_H
#define FQD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FQD_ABI_VERSION 1

typedef enum {
    FQD_OK = 0,
    FQD_ERR_INVALID = 1,        /* bad argument / call sequence                                          */
    FQD_ERR_CUDA = 2,           /* CUDA runtime failure (fqd_last_error has the text)                    */
    FQD_ERR_EMPTY = 3,          /* "Not enough memory to read a single object!"  src/bufferedinput.hpp:82-85 */
    FQD_ERR_BAD_START = 4,      /* record does not start with '@' / '>'          src/fastqview.cpp:121-126   */
    FQD_ERR_LEN_MISMATCH = 5,   /* FASTQ: len(seq) != len(qual)                  src/fastqview.cpp:117       */
    FQD_ERR_BAD_BASE = 6,       /* fast mode: byte outside {A,C,G,T,N}           src/seq_utils.cpp:17-19     */
    FQD_ERR_CAPACITY = 7,       /* key store / tables full and no device memory left to grow them in place */
    FQD_ERR_SEQ_TOO_LONG = 8,   /* sequence longer than fqd_config.max_seq_len                           */
    FQD_ERR_TAG_TOO_LONG = 10,  /* --unordered: an ID tag is longer than fqd_config.max_tag_len          */
    FQD_ERR_UNSUPPORTED_BYTE = 9 /* sequence mode with 3-bit rows: a sequence byte outside {A,C,G,T,N}; run the job again
                                    with fqd_config.byte_keys = 1 (the reference orders any byte there)       */
} fqd_status;

typedef enum { FQD_FORMAT_FASTQ = 0, FQD_FORMAT_FASTA = 1 } fqd_format;      /* --format           src/main.cpp:111-120 */
typedef enum {
    FQD_MODE_FAST = 0,          /* --fast                      HashDupRemover   src/main.cpp:147-150 */
    FQD_MODE_SEQ_TIGHT = 1,     /* --compare-seq tight         TightComparator  src/comparator.cpp:45-58 */
    FQD_MODE_SEQ_LOOSE = 2,     /* --compare-seq loose         LooseComparator  src/comparator.cpp:60-74 */
    FQD_MODE_SEQ_HAMMING = 3    /* --compare-seq tail-hamming  HammingComparator src/comparator.cpp:76-91 */
} fqd_mode;

typedef struct {
    uint32_t abi_version;       /* FQD_ABI_VERSION                                                       */
    int32_t  device;            /* CUDA device ordinal                                                   */
    int32_t  mode;              /* fqd_mode                                                              */
    int32_t  format;            /* fqd_format                                                            */
    int32_t  paired;            /* 0 single-end, 1 paired-end (-u/-p)                                    */
    int32_t  unordered;         /* --unordered (FAST + paired only)       src/main.cpp:158-164           */
    uint32_t hamming_dist;      /* --distance                             src/main.cpp:34                */
    uint32_t max_seq_len;       /* longest sequence line (bases, without '\n') the key rows must hold    */
    uint64_t max_records;       /* capacity of the key store: records (pairs) over the whole run         */
    uint64_t max_chunk_bytes;   /* largest buffer ever passed to one fqd_push* call (per mate), < 4 GiB  */
    uint64_t max_chunk_records; /* records per chunk the per-chunk tables hold; 0 = max_chunk_bytes/64   */
    uint32_t max_tag_len;       /* --unordered: longest ID tag in bytes the tag keys must hold; 0 = 32    */
    uint32_t byte_keys;         /* sequence-based modes: 1 = key rows hold the raw bytes of sequence + '\n' (8 bits per
                                   symbol, any byte is legal, src/fastqview.cpp:56-67 orders arbitrary bytes) instead of
                                   3-bit codes of {A,C,G,T,N}; the host sets it after FQD_ERR_UNSUPPORTED_BYTE         */
} fqd_config;

typedef struct {
    uint64_t total;             /* records / pairs processed   (-v line, src/hash_dup_remover.hpp:147,254)  */
    uint64_t dups;              /* duplicates removed                                                     */
    uint64_t unmatched;         /* --unordered: non-matching entries skipped  src/hash_dup_remover.hpp:345 */
    int32_t  err;               /* sticky fqd_status of the first data error in input order, or FQD_OK    */
    int32_t  err_char;          /* offending byte for BAD_START / BAD_BASE                                */
    uint64_t err_record;        /* global index of the record that raised it                              */
    int32_t  err_mate;          /* 0 / 1: which input file the offending record belongs to                */
    int32_t  reserved;
} fqd_stats_t;

/* Result of one pushed chunk (FAST ordered mode) - host pointers owned by the handle, valid until the next
 * push on the same handle. */
typedef struct {
    uint64_t n_records;         /* complete records (pairs) parsed from this chunk                        */
    uint64_t consumed[2];       /* bytes of each mate's buffer covered by those records (carry the rest)  */
    uint64_t first_record;      /* global index of record 0 of this chunk                                 */
    uint64_t n_survivors;       /* records of this chunk that are WRITTEN                                 */
    const uint32_t* rec_start[2]; /* n_records+1 offsets into each mate's buffer                          */
    const uint8_t*  dup;        /* n_records flags: 1 = duplicate (dropped), 0 = written                  */
} fqd_chunk_result;

typedef struct fqd_handle fqd_handle;

/* Lifecycle. */
int fqd_create(const fqd_config* cfg, fqd_handle** out);
void fqd_destroy(fqd_handle* h);
const char* fqd_last_error(const fqd_handle* h);   /* also valid with h == NULL for fqd_create failures */
int fqd_abi_version(void);

/* Pinned host staging memory for the reader threads (cudaHostAlloc / cudaFreeHost). */
int fqd_host_alloc(void** p, size_t bytes);
int fqd_host_free(void* p);

/*
 * FAST ordered mode, streaming (replaces the per-record loop of impl_filterSE / impl_filterPE,
 * src/hash_dup_remover.hpp:126-144,228-251).  Each push holds whole records from the start of the buffer;
 * an incomplete trailing record is left to the caller (consumed[] says where it begins), exactly as
 * BufferedInput::refresh carries the partial tail (src/bufferedinput.hpp:66-74).  Paired: the two buffers
 * advance in lock-step, min(n1, n2) pairs are processed.  Records are numbered in input order across pushes;
 * the first occurrence of a key in that order is the one written.
 *   fqd_push        : host buffers (H2D copy inside)          fqd_push_device : buffers already in HBM
 * r2 / n2 are NULL / 0 for single-end.  res may be NULL (statistics only).
 */
int fqd_push(fqd_handle* h, const char* r1, size_t n1, const char* r2, size_t n2, fqd_chunk_result* res);
/* fqd_push in two halves, so that the host-to-device copy of chunk c+1 overlaps the processing of chunk c (the successor
 * of BufferedInput's single buffer, src/bufferedinput.hpp:57-88, is a pair of device buffers): fqd_push_prefetch stages a
 * chunk (copy on its own stream, returns at once; at most two staged), fqd_push_staged processes the oldest staged chunk
 * and returns what fqd_push would have returned for it.  The host buffers must stay valid until that call returns. */
int fqd_push_prefetch(fqd_handle* h, const char* r1, size_t n1, const char* r2, size_t n2);
int fqd_push_staged(fqd_handle* h, fqd_chunk_result* res);
int fqd_push_device(fqd_handle* h, const void* d_r1, size_t n1, const void* d_r2, size_t n2, fqd_chunk_result* res);
/* Same as fqd_push_device but returns after enqueueing; nothing is copied to the host.  fqd_sync() waits and
 * folds the chunk counters into the statistics; the first chunk that raised a data error is remembered on the device,
 * whichever chunk it was, and reported then (fqd_stats).  A chunk that does not fit the key store is refused as a whole
 * (FQD_ERR_CAPACITY at fqd_sync; the synchronous pushes grow the key store and the table in place instead and run the
 * chunk again).  Used for device-resident throughput measurement. */
int fqd_push_device_async(fqd_handle* h, const void* d_r1, size_t n1, const void* d_r2, size_t n2);
int fqd_sync(fqd_handle* h);
/* Survivor list of a whole ordered --fast run, kept on the device: the global input indices (record 0 = first record pushed)
 * of the records that are WRITTEN, ascending - what the reference's write loop produces one record at a time
 * (src/hash_dup_remover.hpp:136-138,240-244).  fqd_keep_survivors(h, 1) before the first chunk (or right after fqd_reset)
 * makes every chunk append its survivors (8 bytes each, capacity = fqd_config.max_records, grown with the key store);
 * fqd_survivors waits for the chunks enqueued so far, reports the count, copies entries [first, first + cap) to dst (may be
 * NULL) and hands out the device pointer of the list (may be NULL).  For callers that keep the input on the device or
 * fetch the written records themselves; the streaming host path gets per-chunk flags from fqd_push instead. */
int fqd_keep_survivors(fqd_handle* h, int on);
int fqd_survivors(fqd_handle* h, uint64_t first, uint64_t* dst, uint64_t cap, uint64_t* n_total, const uint64_t** d_list);
/* Forget every key seen so far (empty set, counters and sticky error cleared): start a new job on the same handle. */
int fqd_reset(fqd_handle* h);

/*
 * Whole-input modes: sequence-based (sort + comparator scan) and --fast --unordered (tag join).  The caller
 * appends the raw bytes of the whole input (any number of fqd_append calls per mate; buffers are concatenated
 * in call order and may cut records anywhere), then fqd_finish runs the sort / scan / join on the device.
 * fqd_emission returns, in EMISSION order (sorted order for sequence mode, tag order for unordered -
 * SURVEY.md F2), the byte span of every written record inside the concatenated input of each mate.
 */
int fqd_append(fqd_handle* h, int mate, const char* buf, size_t n);
int fqd_append_device(fqd_handle* h, int mate, const void* d_buf, size_t n);
/* Zero-copy variant for input that is ALREADY in device memory as one buffer per mate (GPUDirect-style ingest, the
 * receive buffer of the multi-GPU exchange): the engine parses the buffer in place (views of at most max_chunk_bytes,
 * cut at record boundaries) and gathers the output from it.  One call per mate instead of fqd_append*; d_buf must be
 * 16-byte aligned and stay valid and unchanged until fqd_reset / fqd_destroy. */
int fqd_adopt_device(fqd_handle* h, int mate, const void* d_buf, size_t n);
int fqd_finish(fqd_handle* h);
typedef struct {
    uint64_t n_out;             /* records (pairs) written                                               */
    const uint64_t* off[2];     /* per mate: byte offset of each written record in the concatenated input */
    const uint32_t* len[2];     /* per mate: its byte length                                              */
    const uint64_t* head;       /* sequence mode: for every INPUT-sorted position, unused unless clusters  */
} fqd_emission_t;
int fqd_emission(fqd_handle* h, fqd_emission_t* out);
/* Streams the OUTPUT BYTES of one mate in emission order: the device gathers the written records out of the raw
 * input it still holds, so the host never needs the input again.  Each call fills dst with whole records
 * (up to cap bytes, cap >= the longest record), sets *n_bytes, and sets *done when the mate's output is complete.
 * (Replaces the `output_file.write(obj.start(), obj.size())` calls of src/seq_dup_remover.hpp:74,88,163-186 and
 * src/hash_dup_remover.hpp:295-298.) */
int fqd_emit(fqd_handle* h, int mate, void* dst, size_t cap, size_t* n_bytes, int* done);
/* --write-clusters (sequence-based modes): streams the text of `<out>.clusters` for one mate the same way - one line
 * per input record in sorted order, the ID line of a written record (cluster head) or "--" + the ID line of a removed
 * one (src/seq_dup_remover.hpp:60-62,75-76,89-101,142-146,165-169,187-208; src/file_utils.cpp:98-112). */
int fqd_emit_clusters(fqd_handle* h, int mate, void* dst, size_t cap, size_t* n_bytes, int* done);
/*
 * Whole-input modes on inputs LARGER THAN DEVICE MEMORY (the successor of the reference's bounded-memory external sort:
 * chunks of 2/3 memlimit sorted and written to disk, then a k-way merge, src/external_sort.hpp:88-207,
 * src/paired_external_sort.hpp:112-257).  Sorting, the comparator scans and the tag join only ever look at the packed key
 * rows and the per-record tables, so the raw bytes need not stay: after fqd_discard_input(h, 1) (before the first
 * fqd_append; it survives fqd_reset) every input segment is freed as soon as it has been split and packed.  What stays
 * resident per record is its key row (3 bits per base) + 16 bytes of (offset, lengths) per mate - about a quarter of
 * the record - and the caller, which still has the input (a file it can map, or a spool it wrote while reading a pipe
 * or a .gz), fetches the written records itself:
 *   fqd_emission_read   entries [first, first + count) of one mate's emission list (emission order, as fqd_emission):
 *                       byte offset in the concatenated input and byte length of every WRITTEN record;
 *   fqd_cluster_read    --write-clusters: for the sorted positions [first, first + count) (count <= records processed)
 *                       offset and length of the record standing there and head[i] = 1 when it is written (cluster head;
 *                       its ID line goes to <out>.clusters as it is, a removed record's with "--" in front).
 * Both work on any finished whole-input handle; fqd_emit / fqd_emit_clusters / fqd_partition_gather need the raw bytes
 * and fail with FQD_ERR_INVALID after fqd_discard_input.
 */
int fqd_discard_input(fqd_handle* h, int on);
/* n_written: length of the emission lists; n_sorted: sorted positions fqd_cluster_read can be asked for (0 outside the
 * sequence-based modes or after a data error).  Either pointer may be NULL. */
int fqd_emission_count(fqd_handle* h, uint64_t* n_written, uint64_t* n_sorted);
int fqd_emission_read(fqd_handle* h, int mate, uint64_t first, uint64_t count, uint64_t* off, uint32_t* len);
int fqd_cluster_read(fqd_handle* h, int mate, uint64_t first, uint64_t count, uint64_t* off, uint32_t* len, uint8_t* head);
/* Free and total memory of a device in bytes (cudaMemGetInfo): how the host decides whether an input can stay resident. */
int fqd_device_memory(int device, size_t* free_bytes, size_t* total_bytes);

/*
 * Multi-GPU sequence-based mode (one process per GPU, SURVEY.md 8e: sampled splitters + all-to-all).  The reference has
 * no counterpart (it is single-threaded; its only "sharding" is the chunk + k-way merge of src/external_sort.hpp:88-207).
 * Every rank appends ITS contiguous slice of the input to an "origin" handle, then
 *   fqd_partition_sample   parses what was appended and returns n_samples evenly spaced (word 0, word 1) key pairs
 *                          (2 x n_samples words; all-ones when the slice is empty) and the slice's record count;
 *   the ranks all-gather the samples, sort them and agree on n_ranges - 1 ascending splitters;
 *   fqd_partition_plan     owner of a record = number of splitters <= its (word 0, word 1); returns the record count
 *                          per owner and, per mate and owner, the raw bytes that go there (bytes[mate * n_ranges + o]);
 *   fqd_partition_gather   writes the raw records of one mate to a device buffer, grouped by owner, input order kept;
 *   one all-to-all per mate moves the bytes; the receiver appends what it got, in rank order, to a second, ordinary
 *   handle of the same mode (fqd_append_device) - the records of one key range, still in global input order.
 * Exact duplicates share their leading key words and therefore their owner; prefix (loose) and Hamming neighbours may
 * straddle two ranges.  So the finish is staged: fqd_finish_scan (sort + comparator scan, as if nothing preceded this
 * range), then rank by rank fqd_boundary_get on range k-1 -> fqd_boundary_fix on range k (re-evaluates the first sorted
 * records against the last record / last cluster head of the range before), then fqd_finish_emit.  The output of the
 * job is the concatenation of the ranges' outputs in range order.  fqd_finish == fqd_finish_scan + fqd_finish_emit.
 */
int fqd_partition_sample(fqd_handle* h, uint32_t n_samples, uint64_t* samples, uint64_t* n_records);
int fqd_partition_plan(fqd_handle* h, const uint64_t* splitters, uint32_t n_ranges, uint64_t* counts, uint64_t* bytes);
int fqd_partition_gather(fqd_handle* h, int mate, void* d_out);
int fqd_finish_scan(fqd_handle* h);
size_t fqd_boundary_bytes(fqd_handle* h);
int fqd_boundary_get(fqd_handle* h, void* state);
int fqd_boundary_fix(fqd_handle* h, const void* prev_state);
int fqd_finish_emit(fqd_handle* h);

/*
 * Multi-GPU --fast --unordered (one process per GPU; the reference's counterpart is the single-threaded
 * src/hash_dup_remover.hpp:150-192,257-347).  Both files are range-partitioned by ID TAG with shared splitters, so the
 * ranges in rank order are the job's tag order and equal tags always meet in one range.  On an --unordered handle
 *   fqd_partition_sample / _plan / _gather   work per file: samples come half from each file's tags, every file is
 *                          partitioned by its own tags (counts[o] = records of both files, bytes[mate * n_ranges + o]);
 *   the receiver appends what it got to a second --unordered handle, then runs the stages of fqd_finish one by one:
 *   fqd_unordered_prepare  parse + sort both tag lists + partner of every record; returns the two list lengths.  An
 *                          empty list is fine here.
 *   -- all-gather of the lengths: every rank now knows where each range starts in the job's sorted lists L and R --
 *   fqd_unordered_enter    (side 0) position in this range's R list when the walk first stands on element i of its L
 *                          list, i.e. after element i - 1 has been consumed (side 1: the roles swapped).  The walk of
 *                          src/hash_dup_remover.hpp:279-315 stops when either side has fetched its LAST record, so the
 *                          range holding L[n-2] and the range holding R[m-2] answer this once each and the job's stop
 *                          state (is, js) follows: (n-1, ja) if ja < m-1, else (ia, m-1).
 *   fqd_unordered_join     limit_i / limit_j: the stop state clamped to this range's lists; final_i / final_j: its
 *                          local position when the stop state's record lives here, else UINT64_MAX.  Emits the matched
 *                          pairs before the limits, plus the stop state when both of its records are here and their
 *                          tags are equal.  out[0] pairs emitted, out[1] unmatched records counted here, out[2] emission
 *                          index of the first pair holding a byte outside {A,C,G,T,N} or UINT64_MAX, out[3] 1 when the
 *                          job's last comparison happened here and matched.
 *   -- all-gather of out[]: emission offsets, the job's abort point (first bad pair), unmatched total (+1 when the
 *      last comparison matched nowhere) --
 *   fqd_unordered_rows     rows (fqd_unordered_row_bytes() each: pair key + hash) of this range's pairs before `limit`,
 *                          grouped by the GPU that owns their hash range, emission order kept; counts[k] rows for k
 *   -- all-to-all of the rows: what arrives is in rank order = the job's emission order --
 *   fqd_unordered_insert   on the owner: d_flags[i] = 1 when an earlier row holds the same pair key
 *   -- all-to-all of the flags back --
 *   fqd_unordered_apply    survivors + emission lists of this range; fqd_emit / fqd_emission work as after fqd_finish.
 *                          report_bad: this range holds the pair the job aborts on (FQD_ERR_BAD_BASE in fqd_stats).
 * The job's output is the concatenation of the ranges' outputs in rank order.
 */
int fqd_unordered_prepare(fqd_handle* h, uint64_t* n_left, uint64_t* n_right);
int fqd_unordered_enter(fqd_handle* h, int side, uint64_t i, uint64_t* pos);
int fqd_unordered_join(fqd_handle* h, uint64_t limit_i, uint64_t limit_j, uint64_t final_i, uint64_t final_j, uint64_t out[4]);
size_t fqd_unordered_row_bytes(fqd_handle* h);
int fqd_unordered_rows(fqd_handle* h, uint64_t limit, uint32_t n_shards, void* d_send, uint64_t* counts);
int fqd_unordered_insert(fqd_handle* h, const void* d_recv, uint64_t n_recv, uint32_t n_shards, void* d_flags);
int fqd_unordered_apply(fqd_handle* h, const void* d_flags_back, uint64_t limit, int report_bad);

/*
 * Multi-GPU --fast mode (one handle per GPU / process).  The all-to-all exchanges themselves are
 * done by the caller (torch.distributed / NCCL) on the stream given to fqd_set_stream, between these calls:
 *   fqd_shard_pack     split + pack one chunk of THIS rank's input and group the packed keys by owning shard;
 *                      d_send receives n_records rows of fqd_shard_row_bytes() bytes, counts[k] (host) rows for shard k
 *   -- all-to-all of the rows --
 *   fqd_shard_insert   insert n_recv received rows (already ordered by global input position: shard 0's rows first)
 *                      into this rank's set; d_flags receives one byte per row, 1 = duplicate
 *   -- all-to-all of the flags back --
 *   fqd_shard_apply    flags, in the order the rows were sent, are stored against this rank's records; the chunk's
 *                      duplicate count is added to the statistics
 * Chunks must be fed in global input order: chunk c of rank r holds the records that follow chunk c of rank r-1.
 */
int fqd_set_stream(fqd_handle* h, void* cuda_stream);
size_t fqd_shard_row_bytes(fqd_handle* h);
int fqd_shard_pack(fqd_handle* h, const void* d_raw, size_t n, uint32_t n_shards, void* d_send, uint64_t* counts, uint64_t* n_records);
/* paired-end handle: the two chunks hold the same records (cut at the same record index); the rows carry both mates'
 * keys and are owned by the hash of the pair (setRecordPair, src/hash_dup_remover.hpp:30-41,55-68) */
int fqd_shard_pack_pe(fqd_handle* h, const void* d_r1, size_t n1, const void* d_r2, size_t n2, uint32_t n_shards, void* d_send,
                      uint64_t* counts, uint64_t* n_records);
int fqd_shard_insert(fqd_handle* h, const void* d_recv, uint64_t n_recv, uint32_t n_shards, void* d_flags);
int fqd_shard_apply(fqd_handle* h, const void* d_flags_back, uint64_t* chunk_dups);
/* duplicate flags (one byte per record, record order) of the chunk last given to fqd_shard_apply, copied to the host */
int fqd_shard_read_flags(fqd_handle* h, void* dst, size_t n);

/*
 * Multi-GPU --fast mode, round 2: the same hash-range ownership without a host round trip or a staging copy per chunk.
 * The key store of an owner is a sequence of fixed REGIONS, one per (chunk, source rank), region_rows rows each, so a source
 * knows the slot of every row it sends before anything is exchanged: slot = (chunk * n_shards + source) * region_rows +
 * position among the source's rows for that owner - still global input order (chunks dealt round-robin in input order, sources
 * in rank order), which is what makes "smallest slot wins" keep the record the reference keeps (src/hash_dup_remover.hpp:
 * 133-139).  fqd_config.max_records of every handle must cover n_chunks * n_shards * region_rows.
 *   fqd_shard2_init      allocates the per-chunk regions (hashes, counts, flags; double-buffered by chunk parity) and the
 *                        interprocess events of this rank;
 *   fqd_shard2_export    fqd_shard2_blob_bytes() bytes (CUDA IPC handles of the key store, the regions and the events); the
 *                        ranks all-gather them and fqd_shard2_import every peer's;
 *   fqd_shard2_pack      split + pack chunk number `chunk` of THIS rank's input, partition its rows stably by owner into a
 *                        staging area, let the copy engines move every owner's part straight into that owner's key store
 *                        over mapped peer memory (NVLink; fixed-size copies, nothing crosses the host), record "scattered";
 *   fqd_shard2_insert    owner side: waits (on the device) for every source's "scattered" of that chunk, inserts region by
 *                        region, writes one flag byte per row into the SOURCE's flag regions, records "flags sent";
 *   fqd_shard2_apply     source side: waits (on the device) for every owner's "flags sent", stores the flags against this rank's
 *                        records (fqd_shard2_read_flags) and counts the duplicates;
 *   fqd_shard2_finish    waits for this rank's streams; totals of this rank's records; data errors / a region overflow are
 *                        reported through fqd_stats.
 * Every call but finish only enqueues.  The caller's loop is  pack(0) pack(1) | for c: insert(c) | apply(c) [pack(c+2)]  with a host
 * barrier at every "|": it makes sure an event has been RECORDED by its owner before a peer enqueues the wait for it.
 * Chunks must be numbered alike on every rank (a rank whose slice is shorter passes empty chunks).
 */
size_t fqd_shard2_blob_bytes(void);
int fqd_shard2_init(fqd_handle* h, uint32_t n_shards, uint32_t me, uint32_t region_rows);
int fqd_shard2_export(fqd_handle* h, void* blob);
int fqd_shard2_import(fqd_handle* h, uint32_t rank, const void* blob);
int fqd_shard2_pack(fqd_handle* h, uint64_t chunk, const void* d_r1, size_t n1, const void* d_r2, size_t n2);
int fqd_shard2_insert(fqd_handle* h, uint64_t chunk);
int fqd_shard2_apply(fqd_handle* h, uint64_t chunk);
int fqd_shard2_finish(fqd_handle* h, uint64_t* n_records, uint64_t* n_dups);
int fqd_shard2_reset(fqd_handle* h);
int fqd_shard2_read_flags(fqd_handle* h, void* dst, size_t n);
/* The same protocol inside ONE process that drives several GPUs (the drop-in binary): fqd_shard2_link instead of export /
 * import (peer access between the devices of one process, ordinary events), no barriers (one thread enqueues in order).
 * fqd_shard2_push_host = pack for a chunk in pinned host memory; it waits until the chunk is split and reports the records
 * (pairs) it holds and where the incomplete tail of each mate begins.  fqd_shard2_result, after insert(chunk) on EVERY handle
 * and apply(chunk) on this one, returns what fqd_push would have returned for the chunk. */
int fqd_shard2_link(fqd_handle* h, uint32_t rank, fqd_handle* peer);
int fqd_shard2_push_host(fqd_handle* h, uint64_t chunk, const char* r1, size_t n1, const char* r2, size_t n2,
                         uint64_t* n_records, uint64_t* consumed);
int fqd_shard2_result(fqd_handle* h, uint64_t first_record, size_t n1, size_t n2, fqd_chunk_result* res);
int fqd_shard2_timer_start(fqd_handle* h);
int fqd_shard2_timer_stop(fqd_handle* h, double* ms);

/* Host barrier between the rank processes of one box (POSIX shared memory, a few microseconds): the "|" of the loop above.
 * Every rank opens the same name (unique per job), rank 0 unlinks it at close. */
int fqd_hostbar_open(const char* name, uint32_t n_ranks, void** out);
int fqd_hostbar_wait(void* bar);
int fqd_hostbar_close(void* bar, const char* unlink_name);

/* Statistics / sticky data error (feeds the -v lines and the reference's error messages). */
int fqd_stats(fqd_handle* h, fqd_stats_t* out);

/* Device time (ms, CUDA events on the handle's stream) spent in kernels since creation, and kernel launches. */
int fqd_device_time_ms(fqd_handle* h, double* ms, uint64_t* launches);
/* Stopwatch on the handle's stream: start records a CUDA event, stop records another, waits for it and returns
 * the elapsed device time between the two (everything enqueued on the handle in between). */
int fqd_timer_start(fqd_handle* h);
int fqd_timer_stop(fqd_handle* h, double* ms);
/* Per-kernel-class profile (CUDA events around every launch of the parse+pack kernel and of the insert kernel
 * while enabled): accumulated milliseconds and launch counts since the last fqd_profile_enable(h, 1). */
typedef struct {
    double   parse_ms;   uint64_t parse_launches;   uint64_t parse_bytes;    uint64_t parse_records;
    double   insert_ms;  uint64_t insert_launches;
    double   scatter_ms; uint64_t scatter_launches;   /* multi-GPU --fast: partition by owner + scatter over peer memory */
} fqd_profile_t;
int fqd_profile_enable(fqd_handle* h, int on);
int fqd_profile_get(fqd_handle* h, fqd_profile_t* out);

/*
 * Counter-based synthetic FASTQ generator: fills a DEVICE buffer with records
 * [first, first+count) of the stream (seed, dup_permille, n_permille): "@SYN.%010u <mate>\n" + read_len bases
 * + "\n+\n" + read_len quals + "\n".  Record size is 22 + 2*read_len bytes (322 for 150 bp).  mate is 1 or 2.
 * variant: 0 exact duplicates; 1 = loose (some duplicates truncated); 2 = hamming (tail substitutions).
 */
int fqd_synth_fastq(int device, void* d_out, uint64_t first, uint64_t count, uint32_t read_len, int mate,
                    uint64_t seed, uint32_t dup_permille, uint32_t n_permille, int variant);
size_t fqd_synth_record_bytes(uint32_t read_len);

/* Raw device memory helpers so that bindings need no CUDA runtime of their own. */
/* Peer-memory exchange between the ranks of one box (one process per GPU): export a device allocation made with
 * fqd_device_alloc as a 64-byte handle, map a peer's handle, and copy into mapped peer memory over NVLink. */
int fqd_ipc_export(int device, void* d_ptr, void* handle64);
int fqd_ipc_open(int device, const void* handle64, void** d_ptr);
int fqd_ipc_close(int device, void* d_ptr);
int fqd_peer_copy_async(int device, void* d_dst, const void* d_src, size_t bytes, void* cuda_stream);
int fqd_device_alloc(int device, void** d_ptr, size_t bytes);
int fqd_device_free(int device, void* d_ptr);
int fqd_memcpy_d2h(int device, void* dst, const void* d_src, size_t bytes);
int fqd_memcpy_h2d(int device, void* d_dst, const void* src, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* FQD_H */

/*
 * kpal_b200.h -- C ABI of libkpal_b200.so: the B200 (sm_100a) implementation of
 * kPAL's data-parallel hot path (k-mer counting + reverse-complement balance,
 * and the N x N profile distance matrix).
 *
 * kPAL (LUMC/kPAL) is pure Python and has no FFI of its own; the boundary this
 * library replaces is the body of a handful of Python functions.  Each entry
 * point below names the reference function (file:line, relative to the
 * reference tree) whose arithmetic it replaces.  INTEGRATION.md shows the
 * ctypes stub a kPAL maintainer would add to call them from kpal/klib.py and
 * kpal/kdistlib.py.
 *
 * Conventions
 *   - plain C types only; the caller owns every buffer; the library never
 *     keeps a host pointer after a call returns;
 *   - every function returns 0 on success or a KPAL_E* code; the message is
 *     available (per thread) from kpal_last_error();
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point fails with KPAL_ECUDA;
 *   - "host" entry points take host pointers and do their own H2D/D2H copies;
 *     "dev" entry points take CUDA device pointers plus a cudaStream_t (passed
 *     as void*) and never synchronise unless stated.
 *
 * Profile layout (kpal/klib.py:160-170): int64[4^k], index = base-4 number of
 * the k-mer, first base most significant, A=0 C=1 G=2 T=3.
 */
#ifndef KPAL_B200_H
#define KPAL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KPAL_B200_ABI_VERSION 1

/* error codes */
#define KPAL_OK        0
#define KPAL_EINVAL    1   /* bad argument (Python wrapper raises ValueError)   */
#define KPAL_ECUDA     2   /* CUDA runtime / no device (RuntimeError)           */
#define KPAL_ENOMEM    3   /* host or device allocation failed (MemoryError)    */
#define KPAL_EOVERFLOW 4   /* a 32-bit device counter would overflow            */

/* metric / pairwise selectors (kpal/metrics.py:151-162) */
#define KPAL_METRIC_MULTISET  0   /* metrics.multiset, kpal/metrics.py:101-123  */
#define KPAL_METRIC_EUCLIDEAN 1   /* metrics.euclidean, kpal/metrics.py:126-135 */
#define KPAL_METRIC_COSINE    2   /* metrics.cosine_similarity, 138-147         */
#define KPAL_PAIRWISE_PROD    0   /* |x-y| / ((x+1)(y+1)), kpal/metrics.py:160  */
#define KPAL_PAIRWISE_SUM     1   /* |x-y| / (x+y+1),      kpal/metrics.py:161  */

#define KPAL_MAX_K 15

/* ------------------------------------------------------------------ misc */

int         kpal_abi_version(void);
const char *kpal_last_error(void);
/* number of visible CUDA devices (0 when there is none; never fails) */
int         kpal_device_count(void);
/* bind the calling thread's library context to `device` (default 0) */
int         kpal_set_device(int device);
/* the calling thread's device (-1 without one); a helper thread passes it to kpal_set_device */
int         kpal_get_device(void);
/* pinned (page-locked) host memory for fast H2D/D2H; caller frees */
void       *kpal_host_alloc(size_t bytes);
void        kpal_host_free(void *p);
/* device memory helpers for hosts that do not link the CUDA runtime */
void       *kpal_dev_alloc(size_t bytes);
void        kpal_dev_free(void *p);
int         kpal_memcpy_h2d(void *dst_dev, const void *src_host, size_t bytes, void *stream);
int         kpal_memcpy_d2h(void *dst_host, const void *src_dev, size_t bytes, void *stream);
int         kpal_dev_memset(void *dst_dev, int value, size_t bytes, void *stream);
int         kpal_stream_sync(void *stream);

/* ------------------------------------------------- packed sequence format
 *
 * The device kernels read sequence as two bit streams over a single run of
 * "bases" in which records are separated by ONE invalid base:
 *   codes: 2 bits per base, 16 bases per uint32, first base in the MOST
 *          significant bits (so a k-mer index is a funnel shift of 2 words);
 *   valid: 1 bit per base, 32 bases per uint32, same order; 0 for every byte
 *          that is not ACGTacgt (kpal/klib.py:152) and for record separators.
 * "window valid <=> its k valid bits are all 1" then reproduces both the
 * N-split (kpal/klib.py:152-156) and the no-window-across-records rule
 * (kpal/klib.py:154) in one predicate.  Buffers are padded with zero words:
 * kpal_packed_words() gives the required uint32 counts.
 */
void kpal_packed_words(uint64_t n_bases, uint64_t *code_words, uint64_t *valid_words);

/*
 * Host packer for a list of sequences (the iterable handed to
 * Profile.from_sequences, kpal/klib.py:135): record r is
 * text[offsets[r] .. offsets[r+1]).  Emits n_bases = sum(len) + n_records.
 * rec_starts (optional, n_records+1 entries) receives the base index at which
 * each record starts in the packed stream.
 */
int kpal_pack_sequences(const char *text, const uint64_t *offsets, uint64_t n_records,
                        uint32_t *codes, uint32_t *valid, uint64_t *rec_starts,
                        uint64_t *n_bases_out);

/*
 * Host FASTA scanner + packer, replacing Bio.SeqIO.parse + str(record.seq)
 * at kpal/klib.py:111 / 131 (FastaIterator rules: text before the first '>'
 * skipped; sequence lines right-stripped, spaces and '\r' removed, joined).
 * Call kpal_fasta_scan first to size the buffers.
 *   n_records   number of '>' records
 *   n_bases     packed length (sequence bytes + one separator per record)
 *   name_bytes  bytes needed for the '\0'-separated record names
 */
int kpal_fasta_scan(const char *fasta, uint64_t n_bytes, uint64_t *n_records,
                    uint64_t *n_bases, uint64_t *name_bytes);
int kpal_fasta_pack(const char *fasta, uint64_t n_bytes, uint32_t *codes, uint32_t *valid,
                    uint64_t *rec_starts /* n_records+1 */, char *names /* name_bytes */);

/*
 * One SEGMENT of a FASTA text (text[begin, end), begin at a line start; bytes before
 * the segment's first header line are skipped) packed into a slot of its own, same rules as kpal_fasta_pack (kpal/klib.py:111):
 * codes / valid point at the slot's first words, cap_bases is the slot's length (a
 * multiple of 64, >= end - begin); one invalid base is emitted at every header (the
 * record separator) and the slot is invalid behind the emitted bases.  This is what
 * idle host threads do to the tail of a large text while kpal_count_fasta uploads
 * its head raw (hybrid upload); 32 bytes per step with AVX2 + BMI2, scalar otherwise.
 */
int kpal_fasta_pack_segment(const char *fasta, uint64_t n_bytes, uint64_t begin, uint64_t end,
                            uint32_t *codes, uint32_t *valid, uint64_t cap_bases,
                            uint64_t *n_bases_out);

/*
 * The whole text as a SLOTTED stream (csrc/slotted.h): cut at line starts into segments
 * of about seg_bytes, each packed independently into a slot of its own, plus one 64-base
 * junction record per cut holding the k - 1 windows that cross it.  Counting the windows
 * of this stream gives Profile.from_fasta's counts (kpal/klib.py:97-112) for that k.  It
 * is the layout kpal_count_fasta builds on the device when host threads pack the tail of
 * a large text while its head uploads raw; this entry point does all of it on the host
 * (tests, and callers that want to pack ahead of time).  codes / valid must hold
 * kpal_packed_words(kpal_fasta_slotted_bases(n_bytes, seg_bytes)) words; *stream_bases_out
 * receives the stream's length.  KPAL_EINVAL when the text has no header line within its
 * first MB or lines above 64 KB (such a text cannot be cut).
 */
uint64_t kpal_fasta_slotted_bases(uint64_t n_bytes, uint64_t seg_bytes);
int kpal_fasta_pack_slotted(const char *fasta, uint64_t n_bytes, int k, uint64_t seg_bytes,
                            uint32_t *codes, uint32_t *valid, uint64_t *stream_bases_out);

/* ------------------------------------------------------ counting: host API */

/*
 * Replaces Profile.from_sequences (kpal/klib.py:135-170) [+ Profile.balance,
 * kpal/klib.py:285-298, when balance != 0]: counts_out[4^k] int64.
 */
int kpal_count_sequences(const char *text, const uint64_t *offsets, uint64_t n_records,
                         int k, int balance, int64_t *counts_out);

/* Replaces Profile.from_fasta (kpal/klib.py:97-112) [+ balance]. */
int kpal_count_fasta(const char *fasta, uint64_t n_bytes, int k, int balance,
                     int64_t *counts_out);

/*
 * Replaces the per-record loop of Profile.from_fasta_by_record
 * (kpal/klib.py:114-133) for records [first, first+n) of an already packed
 * stream: rows_out is row-major [n][4^k] int64.
 */
int kpal_count_by_record(const uint32_t *codes, const uint32_t *valid, uint64_t n_bases,
                         const uint64_t *rec_starts, uint64_t first, uint64_t n,
                         int k, int balance, int64_t *rows_out);

/* Replaces Profile.balance (kpal/klib.py:285-298) on a host vector, in place. */
int kpal_balance(int64_t *counts_inout, int k);

/* ----------------------------------------------------- distances: host API */

/*
 * Replaces kdistlib.distance_matrix's pair loop (kpal/kdistlib.py:179-184) with
 * ProfileDistance.distance (kpal/kdistlib.py:126-161) restricted to the
 * device fast path (do_positive = do_smooth = False, built-in pairwise):
 * profiles is row-major [n_profiles][4^k] int64; out is row-major
 * [n_profiles][n_profiles] float64, symmetric, diagonal = d(p,p).
 */
int kpal_distance_matrix(const int64_t *profiles, uint64_t n_profiles, int k,
                         int metric, int pairwise, int do_balance, int do_scale, int down,
                         double *out);

/*
 * The same matrix with the profile set handed over in slabs, replacing the
 * load-everything loop of kpal/kmer.py:694-698 (a list of N Profile objects,
 * 34 GB at N = 4096, k = 10): open a session for n_profiles, push the profiles
 * in order in any number of calls (rows is row-major [m][4^k] int64, ideally
 * in kpal_host_alloc memory; it may be reused as soon as push returns -- the
 * upload of one slab overlaps the preparation of the previous one), then
 * finish, which runs the distance kernels and fills out[n][n] as above.
 * A session belongs to the device current at open and to one thread at a time.
 */
int  kpal_matrix_open(uint64_t n_profiles, int k, int metric, int pairwise, int do_balance,
                      int do_scale, int down, void **session);
int  kpal_matrix_push(void *session, const int64_t *rows, uint64_t m);
int  kpal_matrix_finish(void *session, double *out);
void kpal_matrix_close(void *session);

/*
 * Replaces the text loop of kdistlib.distance_matrix (kpal/kdistlib.py:179-186):
 * rows i = 1 .. n-1 of the lower triangle of values (row-major, leading
 * dimension ld >= n), every value as Python's '{0:.{precision}f}', separated
 * by one space, one '\n' per row.  Host only (multi-threaded C++), no GPU.
 * *length receives the text length; when it exceeds capacity (or text is NULL)
 * nothing is written and KPAL_EOVERFLOW is returned, so a caller can size the
 * buffer with a first call.
 */
int kpal_format_matrix(const double *values, uint64_t n, uint64_t ld, int precision,
                       char *text, uint64_t capacity, uint64_t *length);

/*
 * Host stage of the profile's device->host copy.  Profile.counts is int64[4^k]
 * (kpal/klib.py:170); from k = 10 on, kpal_count_fasta / kpal_count_sequences move
 * the counts over PCIe in the narrowest of uint8 / uint16 that holds every count
 * (exact: the device reports the widths that fit; a count above 65535 sends the
 * call down the int64 copy instead) and a pool of host threads widens chunk c into
 * the caller's array while chunk c+1 is in flight.  These entry points are that
 * widening stage alone (host only, no GPU): narrow[0..n) -> counts_out[0..n),
 * consumed in chunks of `chunk` elements (0 = one chunk).
 * kpal_set_option("narrow_d2h", 0) turns the narrow copy off, 2 limits it to uint16.
 */
int kpal_widen_u16(const uint16_t *narrow, uint64_t n, uint64_t chunk, int64_t *counts_out);
int kpal_widen_u8(const uint8_t *narrow, uint64_t n, uint64_t chunk, int64_t *counts_out);

/*
 * Host stages of writing many per-record profiles (Profile.save, kpal/klib.py:227-256, called
 * once per record by kpal count --by-record, kpal/kmer.py:137-146); host only, multi-threaded.
 * kpal_row_stats: stats_out[r][0..4] = total, non_zero, mean, median, std of row r
 * (kpal/klib.py:192-225), bit-identical to the NumPy calls of the reference (same operations
 * in the same order, NumPy's pairwise summation included).  kpal_deflate_chunks: chunk c of
 * `data` (chunk_bytes each) -> a zlib stream of sizes[c] bytes at out + c * slot_bytes
 * (slot_bytes >= kpal_deflate_bound(chunk_bytes)): what the HDF5 deflate filter stores.
 */
int      kpal_row_stats(const int64_t *rows, uint64_t n_rows, uint64_t n_cols, double *stats_out);
uint64_t kpal_deflate_bound(uint64_t chunk_bytes);
int      kpal_deflate_chunks(const void *data, uint64_t n_chunks, uint64_t chunk_bytes, int level,
                             void *out, uint64_t slot_bytes, uint32_t *sizes);
/* The same container with a one-pass encoder for sparse data (a per-record count row is almost
 * all zero bytes: zero runs as distance-1 matches in one dynamic-Huffman block with a code fixed
 * in advance): valid zlib streams that inflate to the same bytes, ~20x faster than zlib on such
 * rows; dense chunks go through zlib.  What klib.save_profiles / h5lite use. */
int      kpal_deflate_chunks_sparse(const void *data, uint64_t n_chunks, uint64_t chunk_bytes, int level,
                                    void *out, uint64_t slot_bytes, uint32_t *sizes);
/* Straight to the packed form (no slot array): begin compresses (sparse != 0: the one-pass encoder
 * where it applies) and reports sizes[c] and their total, finish copies the streams back to back
 * into out (NULL: discard) and frees the job. */
int      kpal_deflate_packed_begin(const void *data, uint64_t n_chunks, uint64_t chunk_bytes, int level,
                                   int sparse, uint32_t *sizes, void **handle_out, uint64_t *total_out);
int      kpal_deflate_packed_finish(void *handle, void *out);
/* kpal_row_stats and the packed streams of the rows' chunks (chunk_bytes divides a row) in one
 * pass over the rows; finish with kpal_deflate_packed_finish. */
int      kpal_rows_stats_deflate_begin(const int64_t *rows, uint64_t n_rows, uint64_t n_cols, uint64_t chunk_bytes,
                                       int level, int sparse, double *stats_out, uint32_t *sizes,
                                       void **handle_out, uint64_t *total_out);
/* the streams packed back to back into out (NULL: only their total size, which is returned) */
uint64_t kpal_compact_slots(const void *slots, uint64_t slot_bytes, const uint32_t *sizes,
                            uint64_t n_chunks, void *out);

/* ProfileDistance.distance for one pair (kpal/kdistlib.py:126-161). */
int kpal_pair_distance(const int64_t *left, const int64_t *right, int k,
                       int metric, int pairwise, int do_balance, int do_scale, int down,
                       double *out);

/*
 * ProfileDistance.distance with do_positive (kpal/kdistlib.py:139-157): after the
 * optional balance, both profiles keep only the positions that are non-zero
 * in both (metrics.positive, kpal/metrics.py:89-98); the scale factors then
 * come from the MASKED totals, which makes them pair-dependent -- hence a pair
 * entry point and no matrix form.
 */
int kpal_pair_distance_positive(const int64_t *left, const int64_t *right, int k,
                                int metric, int pairwise, int do_balance, int do_scale, int down,
                                double *out);

/* ------------------------------------- split / showbalance: host API
 *
 * Replaces Profile.split (kpal/klib.py:300-327): forward / reverse receive
 * kpal_split_length(k) = (4^k + #palindromes) / 2 entries each -- for every
 * index i <= rc(i), in index order, (2 c[i], 2 c[rc(i)]), or (c[i], c[i]) when
 * i is its own reverse complement.
 */
uint64_t kpal_split_length(int k);
int kpal_split(const int64_t *counts, int k, int64_t *forward, int64_t *reverse);

/*
 * The figure kmer.get_balance prints (kpal/kmer.py:240-245):
 * metrics.multiset(forward, reverse, pairwise['prod']) of the two split lists,
 * computed in one pass without materialising them, in the reference's
 * arithmetic (int64 numerator / denominator, one IEEE division per term).
 */
int kpal_show_balance(const int64_t *counts, int k, double *out);

/* ---------------------------------------------------- counting: device API
 * (bench harness, multi-GPU sharding: every pointer is a device pointer)    */

/*
 * Accumulate the windows of a packed stream into a device table of 4^k
 * counters (counter_bits = 32 or 64; the caller zeroes the table).
 * 32-bit counters are exact while n_bases < 2^32.
 */
int kpal_dev_count_packed(const uint32_t *d_codes, const uint32_t *d_valid, uint64_t n_bases,
                          int k, void *d_table, int counter_bits, void *stream);

/* The same for a table in ANY state: the call zeroes it first (a memset on the stream). */
int kpal_dev_count_packed_fresh(const uint32_t *d_codes, const uint32_t *d_valid, uint64_t n_bases,
                                int k, void *d_table, int counter_bits, void *stream);

/*
 * Host FASTA bytes -> windows accumulated into a caller-owned DEVICE table
 * (Profile.from_fasta, kpal/klib.py:97-112, split for the multi-GPU driver:
 * every rank counts its shard of records, the tables are then summed with an
 * NCCL reduce and finalised once).  Synchronises `stream` before returning.
 */
int kpal_count_fasta_to_dev(const char *fasta, uint64_t n_bytes, int k, void *d_table,
                            int counter_bits, void *stream, uint64_t *n_bases_out);

/*
 * The same count into the library's own table (no copy into a caller-owned one): *d_table_out
 * stays valid until the next host-level counting call on this device.  For the multi-GPU
 * driver, which hands the table to kpal_dev_slice_push.  Synchronises `stream`.
 */
int kpal_count_fasta_dev_table(const char *fasta, uint64_t n_bytes, int k, void **d_table_out,
                               int *counter_bits_out, void *stream);

/* table (u32/u64) -> int64 counts, optionally fused with balance
 * (out[i] = t[i] + t[rc(i)], kpal/klib.py:290-298). */
int kpal_dev_finalize_counts(const void *d_table, int counter_bits, int k, int balance,
                             int64_t *d_counts, void *stream);

/* out[i] = in[i] + in[rc(i)] on int64 device vectors (in != out). */
/*
 * Device counter table -> int64 profile in HOST memory: kpal_dev_finalize_counts
 * followed by the device->host copy, in the narrow form described at
 * kpal_widen_u16 (uint8 or uint16 over PCIe from k = 10 on, widened by host threads
 * while the copy runs; int64 copy when a count exceeds 65535).  The table is this
 * GPU's or, on the root of a multi-GPU count, the sum of all ranks' tables.
 * Synchronises `stream`.
 */
int kpal_dev_table_to_host(const void *d_table, int counter_bits, int k, int balance,
                           int64_t *counts_out, void *stream);

int kpal_dev_balance(const int64_t *d_in, int64_t *d_out, int k, void *stream);

/* rows [first, first+n) of the per-record count matrix, device resident. */
int kpal_dev_count_by_record(const uint32_t *d_codes, const uint32_t *d_valid,
                             const uint64_t *d_rec_starts, uint64_t first, uint64_t n,
                             int k, int balance, int64_t *d_rows, void *stream);

/* ------------------------------- multi-GPU: peer-memory reduce of count tables
 *
 * One process per GPU.  The per-rank tables of a record-sharded count
 * (the reference's only parallel workflow is one process per
 * file, kpal/Makefile:3) are summed over NVLink peer memory instead of a
 * library reduce:
 *   1. every rank allocates an inbox of kpal_peer_inbox_bytes() with
 *      kpal_dev_alloc, exports it (kpal_ipc_export -> 64 opaque bytes, exchanged
 *      by the host, e.g. torch.distributed.all_gather_object) and opens its
 *      peers' (kpal_ipc_open); the root does the same for its result table;
 *   2. kpal_dev_reduce_push: slice o of the local table -> slot `rank` of rank
 *      o's inbox (an all-to-all of 16-byte peer stores);
 *   3. a cross-GPU barrier on the same stream (caller's: 1-element all-reduce);
 *   4. kpal_dev_reduce_collect: sum of the `world` slots of the local inbox ->
 *      the root's table (peer store);  5. barrier;  6. the root finalizes.
 * inbox_ptrs is a HOST array of `world` device pointers (entry `rank` = the
 * local inbox).  counter_bits as in kpal_dev_count_packed.
 */
int      kpal_ipc_export(const void *d_ptr, void *handle64);
int      kpal_ipc_open(const void *handle64, void **d_peer_ptr);
int      kpal_ipc_close(void *d_peer_ptr);
uint64_t kpal_peer_inbox_bytes(int k, int counter_bits, int world);
int      kpal_dev_reduce_push(const void *d_table, int counter_bits, int k, int rank, int world,
                              void *const *inbox_ptrs, void *stream);
int      kpal_dev_reduce_collect(const void *d_inbox, int counter_bits, int k, int rank, int world,
                                 void *d_root_table, void *stream);
/*
 * kpal_dev_count_packed + kpal_dev_reduce_push in one call: on the radix path
 * (large inputs, k >= 9, world dividing the bucket count) the second pass of
 * the count stores every bucket's slice (table + histogram) directly into its
 * owner's inbox, so the all-to-all overlaps the histogram work and no separate
 * push kernel runs (*fused_out = 1); otherwise it counts, then pushes (0).
 * d_table is scratch the caller zeroes first, as for kpal_dev_count_packed.
 */
int      kpal_dev_count_packed_push(const uint32_t *d_codes, const uint32_t *d_valid,
                                    uint64_t n_bases, int k, void *d_table, int counter_bits,
                                    int rank, int world, void *const *inbox_ptrs, void *stream,
                                    int *fused_out);

/*
 * The fused form of the multi-GPU table sum (csrc/peer_reduce.cu): balance is linear
 * (kpal/klib.py:285-298), so every rank balances ITS table and sends the result narrow --
 *   kpal_dev_slice_push     balanced counts of the local table, as 1 byte per bin, straight into
 *                           the inbox of the rank that owns the bin's slice (a count >= 255 is sent
 *                           as the escape byte 255 plus a 4-byte store of its value; wide_rows != 0
 *                           sends u32 rows instead -- for shards whose MEAN balanced count, about
 *                           2 * bases / 4^k, is not small).  One kernel launch, peer stores only;
 *   kpal_dev_slice_signal   one release store per peer: "the rows of `epoch` from `rank` have
 *                           landed" (wide_rows as given to the push).  After the push, on its stream;
 *   kpal_dev_slice_collect  on every rank: waits for the world's signals in its own inbox, sums
 *                           the senders' rows and writes the int64 slice
 *                           [kpal_slice_begin(k, rank, world), kpal_slice_begin(k, rank + 1, world))
 *                           of the final balanced profile.  signal = 0 / 1: the kernel first sends
 *                           this rank's signal itself (the value = the wide_rows of its push, which
 *                           must be the previous work on `stream`) -- no separate signal launch;
 *                           signal = -1: kpal_dev_slice_signal has been called;
 *   kpal_dev_slice_collect_to_host  the same followed by the narrow device->host copy of the
 *                           slice (slice_out: host memory, e.g. this rank's part of a profile in
 *                           memory shared between the processes).  Synchronises the stream.
 * The profile stays sharded by slice, so no NVLink or PCIe link carries more than its share.
 * Inboxes: kpal_slice_inbox_bytes() of kpal_dev_alloc memory per rank, ZEROED once, exchanged with
 * kpal_ipc_export / kpal_ipc_open as above; inbox_ptrs[rank] is the local one.  `epoch` counts the
 * steps from 1, the same on every rank (the two parities of an inbox alternate).  k >= 6.
 */
uint64_t kpal_slice_inbox_bytes(int k, int world);
uint64_t kpal_slice_begin(int k, int rank, int world);
int      kpal_dev_slice_push(const void *d_table, int counter_bits, int k, int rank, int world,
                             void *const *inbox_ptrs, uint64_t epoch, int wide_rows, void *stream);
int      kpal_dev_slice_signal(int k, int rank, int world, void *const *inbox_ptrs, uint64_t epoch,
                               int wide_rows, void *stream);
int      kpal_dev_slice_collect(void *const *inbox_ptrs, int k, int rank, int world, uint64_t epoch,
                                int signal, int64_t *d_slice_out, void *stream);
int      kpal_dev_slice_collect_to_host(void *const *inbox_ptrs, int k, int rank, int world, uint64_t epoch,
                                        int signal, int64_t *slice_out, void *stream);

/* --------------------------------------------------- distances: device API */

/* stride (in doubles) of one prepared profile row for a given k */
uint64_t kpal_prepared_stride(int k);

/*
 * Per-profile pre-pass (done once per profile instead of once per pair as in
 * kpal/kdistlib.py:136-157): optional balance, total S (kpal/metrics.py:64-65),
 * F = x/S (x when do_scale == 0), P = x + 1, non-zero bitmap, sum(F^2).
 *   d_counts  [n][4^k] int64
 *   d_F, d_P  [n][stride] float64: F = x/S, P = x + 1 (d_P may be NULL unless multiset/prod)
 *   d_bitmap  [n][stride/32] uint32
 *   d_totals  [n] float64, d_norm2 [n] float64
 */
int kpal_dev_profiles_prepare(const int64_t *d_counts, uint64_t n, int k, int do_balance,
                              int do_scale, double *d_F, double *d_P, uint32_t *d_bitmap,
                              double *d_totals, double *d_norm2, void *stream);

/*
 * Profile order by total for the scaled metrics: ascending (the scaled profile
 * of a pair is the one with the smaller total, kpal/metrics.py:67-70), or
 * descending when down != 0 (kpal/metrics.py:84-86: then the larger one is
 * scaled).  Tiny host-side stable sort of n doubles; synchronises the stream.
 */
int kpal_dev_order_by_total(const double *d_totals, uint64_t n, int down, int32_t *d_order,
                            void *stream);

/* number of tiles of the (sorted) upper triangle the tile kernel walks */
uint64_t kpal_distance_num_tiles(uint64_t n_profiles);

/*
 * Distances for tiles [tile_begin, tile_end) (multi-GPU: each rank takes a
 * slice; tiles are independent).  d_order[n] is the profile order by total
 * (ascending; descending when down != 0) or NULL for identity (do_scale == 0).
 * Writes both out[i][j] and out[j][i] (row-major [n][n]) for every pair of
 * the tiles; other entries are left untouched.
 */
int kpal_dev_distance_tiles(const double *d_F, const double *d_P, const uint32_t *d_bitmap,
                            const double *d_totals, const double *d_norm2,
                            const int32_t *d_order, uint64_t n, int k,
                            int metric, int pairwise, int do_scale, int down,
                            uint64_t tile_begin, uint64_t tile_end,
                            double *d_out, void *stream);

/*
 * Multi-GPU matrix (tiles of the upper triangle sharded over the GPUs
 * after an all-gather of the prepared profile set; kpal/kdistlib.py:179-184 is the loop that
 * is being cut up).  kpal_dev_distance_tiles_packed is kpal_dev_distance_tiles with the
 * finished values of tiles [tile_begin, tile_end) written as a compact
 * [tile_end - tile_begin][kpal_distance_tile_elems()] array (entries outside the triangle
 * are 0), so that the ranks' results travel in ONE gather of N^2 / 2 doubles;
 * kpal_dev_distance_unpack_tiles scatters such an array into the symmetric out[n][n] on the
 * root (diagonal != 0: also writes d(p, p)).
 */
uint64_t kpal_distance_tile_elems(void);
int kpal_dev_distance_tiles_packed(const double *d_F, const double *d_P, const uint32_t *d_bitmap,
                                   const double *d_totals, const double *d_norm2,
                                   const int32_t *d_order, uint64_t n, int k,
                                   int metric, int pairwise, int do_scale, int down,
                                   uint64_t tile_begin, uint64_t tile_end,
                                   double *d_packed, void *stream);
int kpal_dev_distance_unpack_tiles(const double *d_packed, const double *d_totals,
                                   const double *d_norm2, const int32_t *d_order, uint64_t n,
                                   int metric, int pairwise, int do_scale,
                                   uint64_t tile_begin, uint64_t tile_end, int diagonal,
                                   double *d_out, void *stream);

/*
 * Euclidean distance / cosine similarity (kpal/metrics.py:126-147, after the scale step of
 * kpal/kdistlib.py:150-157) through an EXACT integer Gram matrix on the tensor cores
 * (tcgen05.mma kind::i8, accumulators in tensor memory; csrc/distance_gram.cu).  Usable when
 * every (balanced) count fits 8 bits: kpal_dev_gram_prepare raises d_flags[0] otherwise and
 * the caller takes kpal_dev_distance_tiles.  The host entry points (kpal_distance_matrix,
 * matrix sessions) do this by themselves unless kpal_set_option("gram", 0).
 *   d_rows_u8  [n][kpal_gram_row_stride(k)] uint8 (zero padded rows)
 *   d_totals   [n] exact sums of the counts, d_norms [n] exact sums of their squares
 *   d_flags    [1] zeroed by the caller; bit 0 <- a count above 255
 *   norm_max   max of d_norms (host value): one accumulation is exact below 2^31, else the
 *              profile is accumulated in chunks of 32768 elements
 *   d_gram     [n][n] int64 scratch, d_out [n][n] float64 symmetric result (diagonal included)
 * Tolerance: the Gram matrix and the numerators are exact integers; the result carries the
 * rounding of one conversion, one division and one square root (tests state 1e-9 relative,
 * the same as for the fp64 form).
 */
uint64_t kpal_gram_row_stride(int k);
int kpal_dev_gram_prepare(const int64_t *d_counts, uint64_t n, int k, int do_balance, uint8_t *d_rows_u8,
                          uint64_t *d_totals, uint64_t *d_norms, uint32_t *d_flags, void *stream);
int kpal_dev_gram_distances(const uint8_t *d_rows_u8, const uint64_t *d_totals, const uint64_t *d_norms,
                            uint64_t norm_max, uint64_t n, int k, int metric, int do_scale, int down,
                            int64_t *d_gram, double *d_out, void *stream);

/* ------------------------------------------------ FASTA scan/pack: device API
 *
 * GPU version of kpal_fasta_scan + kpal_fasta_pack for whole-file counting
 * (Bio.SeqIO.parse + str(record.seq), kpal/klib.py:111): raw FASTA bytes in
 * device memory -> packed stream.  d_text must be readable up to the next
 * multiple of 16 bytes; d_codes / d_valid must hold kpal_packed_words(n_bytes)
 * words (one base per input byte is the upper bound) and the count kernel can
 * be run with n_bases = n_bytes (the tail is invalid padding).  One invalid
 * base is emitted per header line.  d_scratch (kpal_fasta_scratch_bytes) starts
 * with { uint64 first_header; uint64 n_bases; uint32 flags; } -- flags bit 0 =
 * the text has tabs / VT / FF / FS..US on sequence lines, whose rstrip()
 * semantics need the host packer.
 */
uint64_t kpal_fasta_scratch_bytes(uint64_t n_bytes);
int kpal_dev_fasta_pack(const void *d_text, uint64_t n_bytes, uint32_t *d_codes, uint32_t *d_valid,
                        void *d_scratch, void *stream);

/* run-time switches: "host_fasta" (1 = kpal_count_fasta uses the C++ packer
 * instead of the GPU one), "exact_div" (1 = IEEE division in the distance
 * kernels instead of MUFU.RCP64H + Newton), "gram" (0 = euclidean / cosine matrices
 * always take the element-wise fp64 kernel), "narrow_d2h" (profile copy of the
 * host entry points: 1 = uint8 / uint16, 2 = uint16 only, 0 = int64),
 * "dma_share" (0..8 sixteenths of a narrow-copied profile that the copy engine
 * moves as int64 into a pinned destination; default 0), "fasta_chunks" (0 = auto
 * .. 32 chunks of the pipelined FASTA upload), "fasta_hybrid" (1 = host threads
 * pack segments from the end of a text of >= 32 MB while its head uploads raw,
 * 0 = the whole text uploads raw, 2..64 = at most that many host packers),
 * "fasta_hybrid_share" (percent of the text the host packs; 0 = adapted from call
 * to call to the measured host and bus rates), and the tuning switches of the
 * count path ("count_path", "radix_shape", "radix_payload_bits", "tiled_finalize"). */
int kpal_set_option(const char *name, int value);

/*
 * The text upload of the last kpal_count_fasta / kpal_count_fasta_to_dev /
 * kpal_count_fasta_dev_table call on this process: bytes that crossed the bus host ->
 * device (raw text + packed segments) and how many text bytes the host threads packed
 * themselves (hybrid upload; 0 when the whole text went up raw).  For reporting.
 */
void kpal_last_upload(uint64_t *h2d_bytes, uint64_t *host_packed_text_bytes);

/* counters for bench.py's "gpu_launches": kernels launched by this library
 * in the calling process since load / since the last reset. */
uint64_t kpal_kernel_launches(void);
void     kpal_reset_kernel_launches(void);

#ifdef __cplusplus
}
#endif
#endif /* KPAL_B200_H */

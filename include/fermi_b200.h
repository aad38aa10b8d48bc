/* fermi_b200.h -- C-ABI of libfermi_b200.so: the B200 (sm_100a) drop-in for the FMD-index hot
 * path of lh3/fermi.  Plain C types only (pointers + sizes); no torch / C++ types cross this
 * boundary.  Every entry point names the reference interface it replaces (file:line relative to
 * the reference tree, lh3/fermi 1.1-r751-beta).  Host code (fermi's cmd.c mains, or any FFI)
 * binds exactly these symbols; INTEGRATION.md shows the patch.
 *
 * Conventions (same as the reference):
 *   - every function that can fail returns int, 0 = ok; on failure it prints "[E::fmg_*] ..." to
 *     stderr (honouring fmg_verbose like fm_verbose, utils.c:8) and returns non-zero.  There is
 *     NO CPU fallback: without a CUDA device every device entry point fails loudly.
 *   - result arrays returned through `T **out` are malloc'd and owned by the caller
 *     (release with fmg_free), like the kvec results of fm6_smem (fermi.h:21,102-104).
 *   - handles are read-only and re-entrant during queries; one host thread per GPU.
 */
#ifndef FERMI_B200_H
#define FERMI_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FMG_VERSION "0.2-r1"

extern int fmg_verbose;                                   /* fm_verbose, utils.c:8 */

/* bi-interval, layout-identical to fmintv_t (fermi.h:13-16) */
typedef struct { uint64_t x[3]; uint64_t info; } fmg_intv_t;

/* ------------------------------------------------------------------ host .fmd container
 * Mirror of rld_t (rld.h:20-39) restricted to the Makefile flavour: asize=6, sbits=3, Elias-delta
 * codec.  The bit stream is one flat array instead of 64 MB chunks. */
typedef struct fmg_fmd_s fmg_fmd_t;

fmg_fmd_t *fmg_fmd_restore(const char *fn);               /* rld_restore, rld.c:288-325: "RLD\2" or raw "RLE\6" */
fmg_fmd_t *fmg_fmd_from_bwt(int64_t n, const uint8_t *bwt);   /* fm_bwtenc, build.c:11-31 (nt6 BWT string) */
fmg_fmd_t *fmg_fmd_from_rle6(int64_t n, const uint8_t *rle);  /* rld.c:295-309 (bytes len<<3|sym, bcr.c:20-126) */
/* adopt the fields of an existing rld_t without copying semantics changes: z = n_chunks pointers to
 * 2^23-word chunks (rld.h:9-11,31), frame = n_frames*7 words (rld.h:35-36), mcnt[0..6] (rld.h:33) */
fmg_fmd_t *fmg_fmd_from_rld(int asize, int sbits, uint64_t n_bytes, int n_chunks, const uint64_t *const *z,
                            const uint64_t *mcnt, uint64_t n_frames, const uint64_t *frame);
int        fmg_fmd_dump(const fmg_fmd_t *e, const char *fn);  /* rld_dump, rld.c:242-263 (byte-identical .fmd) */
void       fmg_fmd_destroy(fmg_fmd_t *e);                     /* rld_destroy, rld.c:81-94 */
/* out[0..6]=mcnt, out[7..13]=cnt, out[14]=n_bytes, out[15]=n_frames, out[16]=ibits */
void       fmg_fmd_info(const fmg_fmd_t *e, uint64_t out[17]);
int64_t    fmg_fmd_decode_bwt(const fmg_fmd_t *e, uint8_t *out); /* main_chkbwt -p / unpack view, cmd.c:106 */

/* ------------------------------------------------------------------ device index
 * fmg_index_upload copies the .fmd image to the HBM of `device` and builds the query layout
 * there ("occ blocks": 64-byte blocks of 128 symbols = mid-block cumulative counts + 3 bit planes).
 * Replaces holding an rld_t for queries: rld_restore + rld_rank_index (rld.c:186-224,288). */
typedef struct fmg_index_s fmg_index_t;

fmg_index_t *fmg_index_upload(const fmg_fmd_t *e, int device);
void         fmg_index_free(fmg_index_t *idx);
uint64_t     fmg_index_bytes(const fmg_index_t *idx);         /* HBM bytes of the query layout */
int          fmg_index_device(const fmg_index_t *idx);
/* copy the query layout back to the host: blocks = n_blocks x 16 u32, cs = n_super x 8 u64 (either may be NULL to query sizes) */
int          fmg_index_export(const fmg_index_t *idx, uint32_t *blocks, uint64_t *cs, uint64_t *n_blocks, uint64_t *n_super);

/* ------------------------------------------------------------------ rank on the .fmd stream itself (memory-lean mode)
 * fmg_rldx_upload keeps the run-length / Elias-delta blocks of the .fmd as they are in HBM plus a dense block directory: one 64-byte
 * line per 64-byte block with its first BWT coordinate and the six cumulative counts (together about twice the .fmd file: 1.6-4.4
 * bits per symbol for read sets, against 4 for the occ blocks), and a coarse position -> block table.  One WARP serves a query: ballot search of the block,
 * the block and its directory line staged by bulk asynchronous copies (TMA) on an mbarrier, the codes located by pointer doubling
 * over the payload bit offsets and summed with warp prefix sums / reductions (fermi_b200/csrc/rldx.cu).  Same results as
 * rld_rank2a / fm6_extend; far fewer ranks per second than the occ-block path -- the layout north_star sketches, kept as the
 * alternative for highly repetitive (deep-coverage) indexes that must fit a smaller share of HBM. */
typedef struct fmg_rldx_s fmg_rldx_t;
fmg_rldx_t *fmg_rldx_upload(const fmg_fmd_t *e, int device);
void        fmg_rldx_free(fmg_rldx_t *x);
uint64_t    fmg_rldx_bytes(const fmg_rldx_t *x);
int fmg_rldx_rank2a_batch(const fmg_rldx_t *x, int64_t n, const uint64_t *k, const uint64_t *l, uint64_t *ok, uint64_t *ol);   /* rld_rank2a, rld.c:457-492 */
int fmg_rldx_extend_batch(const fmg_rldx_t *x, int64_t n, const fmg_intv_t *ik, const uint8_t *is_back, fmg_intv_t *ok6);     /* fm6_extend, exact.c:72-88 */

/* ------------------------------------------------------------------ batched queries, HOST buffers
 * (host->device and device->host copies happen inside the call) */

/* n x rld_rank2a (rld.c:457-492): ok/ol = n*6 counts; k = (uint64_t)-1 allowed */
int fmg_rank2a_batch(const fmg_index_t *idx, int64_t n, const uint64_t *k, const uint64_t *l, uint64_t *ok, uint64_t *ol);
/* n x rld_rank1a (rld.c:424-446): ok = n*6 counts of BWT[0..k]; sym[i] (may be NULL) = the value rld_rank1a returns, BWT[k]
 * (-1 for k = (uint64_t)-1, whose counts are 0) */
int fmg_rank1a_batch(const fmg_index_t *idx, int64_t n, const uint64_t *k, uint64_t *ok, int32_t *sym);
/* the self-check of `fermi chkbwt -r` (cmd.c:90-116) for every position of the BWT at once, on the device layout: rank1a(k) -
 * rank1a(k-1) must be the one-hot vector of BWT[k] and rank1a(n-1) the marginal counts; *n_bad = violations (0 = the rank
 * function is consistent with the BWT), *first_bad = the first offending position */
int fmg_check_rank(const fmg_index_t *idx, uint64_t *n_bad, uint64_t *first_bad);
/* n x fm6_extend (exact.c:72-88): ok6 = n*6 intervals; ok6[].info is set to 0 */
int fmg_extend_batch(const fmg_index_t *idx, int64_t n, const fmg_intv_t *ik, const uint8_t *is_back, fmg_intv_t *ok6);
/* n x fm_backward_search (exact.c:7-23): reads are nt6 bytes, read i = seq[off[i]..off[i+1]) */
int fmg_backward_search_batch(const fmg_index_t *idx, int64_t n, const uint8_t *seq, const uint64_t *off,
                              uint64_t *sa_beg, uint64_t *sa_end, uint64_t *size);
/* n x fm6_smem (smem.c:397-410): *mem = all records ordered by read, then exactly as fm6_smem
 * orders them; mem_off[n+1] = first record of each read.  self_match as in fm6_smem1 (smem.c:104). */
int fmg_smem_batch(const fmg_index_t *idx, int64_t n, const uint8_t *seq, const uint64_t *off, int self_match,
                   fmg_intv_t **mem, uint64_t *mem_off);
/* same, into caller-owned host buffers (pinned memory makes the copies asynchronous): mem holds
 * mem_cap records.  Reads are processed in batches of batch_reads (0 = default) with the host->device
 * copy, the kernels and the device->host copy of consecutive batches overlapped on three streams.
 * Returns 0, or 1 when mem_cap is too small (*n_records then holds the required capacity). */
int fmg_smem_batch_into(const fmg_index_t *idx, int64_t n, const uint8_t *seq, const uint64_t *off, int self_match,
                        fmg_intv_t *mem, uint64_t mem_cap, uint64_t *mem_off, uint64_t *n_records, int64_t batch_reads);
/* The same with packed 16-byte records, for indexes of < 2^32 symbols and reads shorter than 32768 bases (else -3): half the bytes
 * on the way back to the host, which is what bounds the call when several GPUs share the host's memory system.  Records are
 * packed on the device after the kernels; fmg_intv16_expand turns them into fmintv_t where the caller needs that layout. */
typedef struct { uint32_t x[3]; uint32_t info; } fmg_intv16_t;    /* info = end | start << 16 | left_closed << 31 (smem.c:63) */
int fmg_smem_batch_into16(const fmg_index_t *idx, int64_t n, const uint8_t *seq, const uint64_t *off, int self_match,
                          fmg_intv16_t *mem, uint64_t mem_cap, uint64_t *mem_off, uint64_t *n_records, int64_t batch_reads);
void fmg_intv16_expand(uint64_t n, const fmg_intv16_t *in, fmg_intv_t *out);
void fmg_free(void *p);

/* ------------------------------------------------------------------ device-resident SMEM session
 * For callers that keep reads and results in HBM (bench.py's kernel-only figure, pipelines).
 * A session owns the scratch for batches of up to max_reads reads of up to max_len bases.
 * `stream` is a cudaStream_t passed as void* (NULL = default stream). */
typedef struct fmg_smem_session_s fmg_smem_session_t;

fmg_smem_session_t *fmg_smem_session_create(const fmg_index_t *idx, int64_t max_reads, int max_len);
void                fmg_smem_session_destroy(fmg_smem_session_t *s);
/* d_seq/d_off are DEVICE pointers. Runs the SMEM kernel + compaction on `stream`; no host sync
 * unless a read overflowed its record slots (then it is re-run with larger slots). */
int fmg_smem_session_run(fmg_smem_session_t *s, int64_t n, const uint8_t *d_seq, const uint64_t *d_off,
                         int self_match, void *stream);
/* after the stream has been synchronised: device pointers to the compacted result of the last run */
int fmg_smem_session_result(fmg_smem_session_t *s, uint64_t *n_records, const fmg_intv_t **d_mem, const uint64_t **d_mem_off);
/* CUDA-event timing of the SMEM kernel alone, on the launching stream (bench.py's roofline figure):
 * enable, run, then read the summed duration of the k_smem launches since the last query */
void   fmg_smem_session_set_timing(fmg_smem_session_t *s, int on);
double fmg_smem_session_kernel_ms(fmg_smem_session_t *s, int *n_launches);
/* return the device scratch the library keeps between calls (overlap batches) to the driver */
void fmg_release_cache(void);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches) */
uint64_t fmg_launch_count(void);

/* ------------------------------------------------------------------ overlap / unitig (unitig.c)
 * Per-sequence overlap record: everything unitig_unidir/unitig1 (unitig.c:227-317) ask the index about one
 * read.  For sequence ids[t] (or first + t*step when ids == NULL):
 *   rec[10*t + 0] value returned by fm_retrieve (exact.c:59)      rec[.. + 1] length
 *   rec[.. + 2]  fm6_is_contained result (unitig.c:77): 0, -1 contained, -9 length <= min_match
 *   rec[.. + 3..5] intv0.x[0..2]                                    rec[.. + 6] rbeg of fm6_get_nei (unitig.c:93), -1 = no neighbour
 *   rec[.. + 7]  number of right neighbours                       rec[.. + 8] length of the grown consensus
 *   rec[.. + 9]  check_left_simple (unitig.c:186) of a unique neighbour: 0 ok, -1 fork, 1 not evaluated
 * nei (malloc'd, fmg_free) = neighbour intervals with info = overlap length, nei_off[n+1];
 * seq / ext (n x max_len, may be NULL) = the sequence itself and the bases fm6_get_nei appended to it.
 * Returns 0, or 2 when a sequence is longer than max_len. */
int fmg_overlap_batch(const fmg_index_t *idx, int min_match, int64_t n, const uint64_t *ids, uint64_t first, uint64_t step,
                      int max_len, int64_t *rec, fmg_intv_t **nei, uint64_t *nei_off, uint8_t *seq, int32_t *len, uint8_t *ext);
/* unitig walk over the records of ALL n_seq sequences (host code; this is what runs after the all-gather of
 * per-GPU record shards): MAG text as written by mag_v_write (mag.c:149-174) to out_path ("-" = stdout) */
int fmg_unitig_assemble(uint64_t n_seq, int max_len, int min_match, const int64_t *rec, const fmg_intv_t *nei,
                        const uint64_t *nei_off, const uint8_t *seq, const uint8_t *ext, const char *out_path, uint64_t *n_unitigs);
/* fm6_unitig (unitig.c:378-407) / `fermi unitig -l min_match`: overlap records of every sequence and the unitigs themselves
 * (link graph + list ranking) on the GPU; an irregular link graph (cycle, one-sided link) or FMG_UNITIG_HOST=1 takes
 * the records to the host and walks them in the reference's seed order (fmg_unitig_assemble).  The set of MAG records
 * equals that of the reference after canonicalisation. max_len 0 = estimate. */
int fmg_unitig(const fmg_index_t *idx, int min_match, int max_len, const char *out_path, uint64_t *n_unitigs);

/* Multi-GPU `fermi unitig`: the sequences (BWT rows) are sharded over the GPUs the way fm6_unitig stripes its threads
 * (unitig.c:394-404), every GPU holding the whole index.  All pointers are DEVICE pointers owned by the caller, who also runs
 * the exchanges of the path (fermi_b200/parallel.py does it with NCCL through torch.distributed):
 *   fmg_overlap_shard   records of rows [row_lo, row_hi) (row_lo even), in ROW order: d_rec = (row_hi - row_lo) x 64-byte records,
 *                       d_rank[row - row_lo] = rank of the row; d_ext / d_spill (32-byte entries) = appended bases / neighbour lists of
 *                       forks, addressed by the records with shard-local offsets; totals = {ext bytes, spill entries} used.  Returns 1
 *                       when ext_cap / spill_cap are too small (totals = the need).
 *   -- ONE all-gather of the four arrays, every shard padded to row_pad rows / ext_pad bytes / spill_pad entries --
 *   fmg_overlap_merge   gathered shards (n_shards <= 64 consecutive row ranges of rows[s] rows) -> d_pack = n_seq x 64-byte records
 *                       indexed by sequence rank (the layout the assembly chases), d_rank_of_row = n_seq ranks; offsets are rebased onto
 *                       the padded gathered ext / spill arrays, which stay as they are
 *   fmg_overlap_left_fix  check_left_simple (unitig.c:186-204) for the records where it decides a link (the reverse complement of the
 *                       unique neighbour has several neighbours): needs the merged array, patches OV_LEFT in it; *n_left = rows evaluated.
 *                       fmg_overlap_left_fix_rows does it for rows [row_lo, row_hi) only, so that N GPUs share the work: each fixes its
 *                       own rows, fmg_overlap_left_flags(apply = 0) reads the row_hi - row_lo flags out, one all-gather of those bytes,
 *                       fmg_overlap_left_flags(apply = 1) writes the flags of the other shards into the local array
 *   fmg_unitig_part     link graph + list ranking over all records, then emission + MAG text (mag_v_write, mag.c:149-174) of the chains
 *                       with head rank % n_parts == part; 0 = ok, 1 = irregular link graph (run fmg_unitig on one GPU).  The text
 *                       travels to a pinned buffer of the index handle in slices on a stream of its own while the caller exchanges
 *                       the sizes (*n_bytes is known at return): ONE part per index handle may be outstanding, until fmg_magpart_free
 *   -- all-gather of the text sizes --
 *   fmg_magpart_write   the text of this part (each slice as it lands) at `offset` of the output file, which must have its final size: ONE caller passes
 *                       total_bytes != 0 first (the file is created / resized), the others 0 after a barrier; the parts are copied
 *                       into a shared mapping, so the ranks fill the file side by side without serialising on it
 *   fmg_unitig_from_device  the whole assembly on one GPU from merged arrays; 0 = MAG written, 1 = irregular link graph */
typedef struct fmg_magpart_s fmg_magpart_t;
int fmg_overlap_shard(const fmg_index_t *idx, int min_match, int max_len, uint64_t row_lo, uint64_t row_hi, void *d_rec, int64_t *d_rank,
                      uint8_t *d_ext, uint64_t ext_cap, void *d_spill, uint64_t spill_cap, uint64_t totals[2]);
int fmg_overlap_merge(const fmg_index_t *idx, int n_shards, const uint64_t *rows, uint64_t row_pad, uint64_t ext_pad, uint64_t spill_pad,
                      const void *d_rec_all, const int64_t *d_rank_all, void *d_pack, int64_t *d_rank_of_row);
int fmg_overlap_left_fix(const fmg_index_t *idx, int min_match, int max_len, void *d_pack, const int64_t *d_rank_of_row, uint64_t *n_left);
int fmg_overlap_left_fix_rows(const fmg_index_t *idx, int min_match, int max_len, void *d_pack, const int64_t *d_rank_of_row, uint64_t row_lo, uint64_t row_hi,
                              uint64_t *n_left);
int fmg_overlap_left_flags(const fmg_index_t *idx, void *d_pack, const int64_t *d_rank_of_row, uint64_t row_lo, uint64_t row_hi, int8_t *d_flags, int apply);
int fmg_unitig_part(const fmg_index_t *idx, int min_match, const void *d_pack, const int64_t *d_rank_of_row, const uint8_t *d_ext, const void *d_spill,
                    int part, int n_parts, fmg_magpart_t **out, uint64_t *n_unitigs, uint64_t *n_bytes);
int fmg_magpart_write(const fmg_magpart_t *p, const char *path, uint64_t offset, uint64_t total_bytes);
void fmg_magpart_free(fmg_magpart_t *p);
int fmg_unitig_from_device(const fmg_index_t *idx, int min_match, const void *d_pack, const int64_t *d_rank, const uint8_t *d_ext, uint64_t ext_total,
                           const void *d_spill, uint64_t spill_total, const char *out_path, uint64_t *n_unitigs);

/* fm6_seqsort (seqsort.c:37-70) / `fermi seqrank`: sorted[mcnt[1]] exactly as the reference fills it (fm6_retrieve, exact.c:100-127,
 * of every even BWT row on the GPU: sorted[rank] = row << 2 | contained << 1 | duplicate); stats (may be NULL) = #zeros,
 * #contained, #duplicates as fm6_seqsort reports them.  The array is what `fermi unitig -r` / `fermi remap -r` load. */
int fmg_seqsort(const fmg_index_t *idx, uint64_t *sorted, int64_t stats[3]);

/* CUDA-event durations (ms, summed over the batches) of the kernels of the last overlap pass on this process:
 * ms[1] fm_retrieve + fm6_is_contained chain, ms[2] fm6_get_nei, ms[3] the deferred check_left_simple pass (selection, all four phases
 * for the selected rows, patch), ms[4] = rows it evaluated, ms[5] record packing, ms[6] seed rows, ms[7] = number of batches */
void fmg_overlap_stats(double ms[8]);

/* ------------------------------------------------------------------ `fermi correct`: k-mer collection
 * fm6_traverse (exact.c:141-171) + ec_collect (correct.c:35-87) for all 4^SUF_LEN suffixes, as a breadth-first
 * frontier expansion on the GPU.  w < 0 selects the k-mer length like fm6_ec_correct (correct.c:313-318).
 * triples (malloc'd, fmg_free), sorted: suffix << 40 | key << 8 | val  -- exactly the (key, val) pairs ec_collect
 * stores in solid[suffix] (correct.c:71-75); cnt[0], cnt[1] as in correct.c:59,66.  The host fills khash from them
 * and runs ec_fix (correct.c:121-300) unchanged. */
int fmg_ec_collect(const fmg_index_t *idx, int w, int min_occ, uint64_t **triples, uint64_t *n_triples, int64_t cnt[2]);
/* the share of one GPU: only the subtrees of the suffixes s with s % n_parts == part (the unit of work the reference hands to
 * its threads, correct.c:346-350); the union over all parts is the result of fmg_ec_collect and cnt adds up.  Needs w >= 17. */
int fmg_ec_collect_part(const fmg_index_t *idx, int w, int min_occ, int part, int n_parts, uint64_t **triples, uint64_t *n_triples, int64_t cnt[2]);
int fmg_ec_kmer_length(uint64_t n_symbols);

/* ------------------------------------------------------------------ index construction
 * BWT of the FMD text  r0 $ rc(r0) $ r1 $ rc(r1) $ ...  (cmd.c:457-469) by prefix-doubling suffix
 * sorting on the GPU; replaces fm_build / ksa_bwt (build.c:33-50, ksa.c:231-242) for texts that fit
 * one GPU.  `text` = nt6 bytes with 0 sentinels (n < 2^32); `bwt` receives n symbols. */
int fmg_build_bwt(int device, int64_t n, const uint8_t *text, uint8_t *bwt);
/* fm_build (build.c:33-50) entirely on the device: suffix sort, BWT and the RLD encoder (rld_enc / enc_next_block /
 * rld_enc_finish, rld.c:111-236, as parallel segment chasing + one thread per 64-byte block); the image is byte-identical
 * to what fmg_fmd_from_bwt / the reference produce.  NULL on failure. */
fmg_fmd_t *fmg_build_fmd(int device, int64_t n, const uint8_t *text);
/* fm_bwtenc (build.c:11-31) on the device for a BWT held by the host (n nt6 symbols) */
fmg_fmd_t *fmg_fmd_from_bwt_device(int device, int64_t n, const uint8_t *bwt);

/* `fermi merge` (cmd.c:335-373): fm_compute_gap_bits (merge.c:31-94) -- bits = (n0 + n1 + 63) / 64 words, bit q set when symbol q of
 * the merged BWT comes from the second index -- as one LF chain per sequence of the second index on the GPU, and fm_merge
 * (merge.c:100-137): the two BWTs interleaved by the gap vector and RLD-encoded on the device; the image is byte-identical to
 * what `fermi merge` writes.  Both indexes of fmg_gap_bits live on the same device. */
int fmg_gap_bits(const fmg_index_t *idx0, const fmg_index_t *idx1, uint64_t *bits);
fmg_fmd_t *fmg_merge(const fmg_fmd_t *e0, const fmg_fmd_t *e1, int device);

/* `fermi contrast` (cmd.c:567-638): fm6_contrast (cmp.c:94-126) -- the lock-step walk of the backward-extension tries of two indexes
 * down to k-mers, children with < min_occ occurrences in both pruned -- as a breadth-first expansion on the GPU.  sub0 / sub1 =
 * (mcnt[1] + 63) / 64 words each: bit x set when sequence rank x of that index starts with a string absent from the other index
 * (collect_tips, cmp.c:22-43).  The bitmaps equal the reference's (a set: the order of the walk does not matter).  k > 4. */
int fmg_contrast(const fmg_index_t *idx0, const fmg_index_t *idx1, int k, int min_occ, uint64_t *sub0, uint64_t *sub1);

/* BCR construction (bcr.h:43-49: bcr_init / bcr_append / bcr_build / bcr_itr_next / bcr_destroy), for collections of
 * any total size that fits HBM.  Sequences are nt6 codes 1..4 (no N, like bcr_append, ropebwt.c:98); they are
 * indexed in the order appended, so `fermi ropebwt` semantics = append the read, then its reverse complement
 * (ropebwt.c:22-45).  The BWT equals the one of fmg_build_bwt / fermi build on the same text. */
typedef struct fmg_bcr_s fmg_bcr_t;
fmg_bcr_t *fmg_bcr_init(int device);                                    /* bcr_init, bcr.c:330 */
int        fmg_bcr_append(fmg_bcr_t *b, int len, const uint8_t *seq);   /* bcr_append, bcr.c:358 */
int        fmg_bcr_append_batch(fmg_bcr_t *b, int64_t n, int len, const uint8_t *seqs);  /* n equal-length sequences */
int        fmg_bcr_build(fmg_bcr_t *b);                                 /* bcr_build, bcr.c:462 (all cycles on the GPU) */
int64_t    fmg_bcr_size(const fmg_bcr_t *b);                            /* symbols in the BWT, -1 before the build */
int        fmg_bcr_bwt(const fmg_bcr_t *b, uint8_t *bwt);               /* one nt6 byte per symbol */
int        fmg_bcr_rle(const fmg_bcr_t *b, uint8_t **rle, int64_t *n);  /* bcr_itr_next stream: bytes len<<3|sym (ropebwt.c:127-144) */
/* `fermi ropebwt | fermi recode` in one step: the BWT stays in HBM with the handle after fmg_bcr_build; fmg_bcr_fmd runs the
 * RLD encoder on it there and hands the image to the caller (fmg_bcr_bwt / fmg_bcr_rle copy it out instead). */
fmg_fmd_t *fmg_bcr_fmd(fmg_bcr_t *b);
void       fmg_bcr_destroy(fmg_bcr_t *b);                               /* bcr_destroy, bcr.c:342 */

/* The library also exports the reference's own BCR symbols with the reference's signatures (bcr.h:43-49: bcr_init, bcr_append,
 * bcr_build, bcr_itr_init, bcr_itr_next, bcr_destroy, bcr_verbose; fermi_b200/csrc/bcr_compat.cpp) so that ropebwt.c links against
 * it unmodified in place of bcr.c: `make -C oracle drop` builds the reference with its bcr.c left out and cmd.c patched by
 * integration/main_exact.patch, and tests/test_gpu_parity.py compares that binary's output with the reference's own. */

/* ------------------------------------------------------------------ synthetic data
 * Deterministic, seed-addressed generators shared by the GPU run, the CPU baseline and the tests. */
void fmg_synth_genome(uint64_t seed, int64_t n, uint8_t *nt6);
void fmg_synth_reads(uint64_t seed, int64_t genome_len, const uint8_t *genome, int64_t n_reads, int len,
                     double err, uint8_t *reads);
/* the text fermi indexes for a set of equal-length sequences (cmd.c:457-469); returns its length,
 * text may be NULL to query the size */
int64_t fmg_fmd_text(int64_t n_seq, int len, const uint8_t *seqs, uint8_t *text);

#ifdef __cplusplus
}
#endif
#endif

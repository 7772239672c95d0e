/*
 * hsa_b200.h -- C ABI of the B200-native replacement for HSA's inexact-search hot path.
 *
 * Drop-in boundary.  Every entry point below names the reference interface it
 * replaces (file:line into the HSA tree).  Plain pointers and sizes only; no torch / CUDA types.
 * All functions return 0 on success and a negative HSA_E_* code on failure; hsa_last_error() returns a
 * human-readable description of the last failure on the calling thread.  Nothing here ever falls back
 * to a CPU implementation: without a CUDA device every compute entry point fails with HSA_E_CUDA.
 *
 * Struct mirrors are byte-identical to the reference's (x86-64, gcc bit-field layout):
 *   hsa_gap_opt_t == gap_opt_t    bwtaln.h:133-143   (64 bytes)
 *   hsa_aln1_t    == bwt_aln1_t   bwtaln.h:41-50     (36 bytes)
 *   hsa_width_t   == bwt_width_t  bwtaln.h:35-38     ( 8 bytes)
 */
#ifndef HSA_B200_H
#define HSA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HSA_B200_ABI_VERSION 1

/* error codes */
#define HSA_OK            0
#define HSA_E_ARG        -1   /* bad argument / unsupported option combination                    */
#define HSA_E_CUDA       -2   /* CUDA runtime error (no device, OOM, launch failure)               */
#define HSA_E_CAPACITY   -3   /* a task exceeded every device stack / hit capacity (never silent)  */
#define HSA_E_NOMEM      -4   /* host allocation failed                                            */

/* bwtaln.h:124-132 mode bits used on the path */
#define HSA_MODE_GAPE     0x01
#define HSA_MODE_COMPREAD 0x02
#define HSA_MODE_LOGGAP   0x04
#define HSA_MODE_NONSTOP  0x10

typedef struct hsa_gap_opt_t {          /* == gap_opt_t, bwtaln.h:133-143; defaults bwtaln.c:21-44 */
    int s_mm, s_gapo, s_gape;
    int mode;
    int indel_end_skip, max_del_occ, max_entries;
    float fnr;
    int max_diff, max_gapo, max_gape;
    int max_seed_diff, seed_len;
    int n_threads;
    int max_top2;
    int trim_qual;
} hsa_gap_opt_t;

typedef struct hsa_aln1_t {             /* == bwt_aln1_t, bwtaln.h:41-50 */
    uint32_t n_mm:16, n_gapo:8, n_gape:8;
    uint32_t k, l;
    uint32_t rev_k, rev_l;
    uint32_t type:30, strand:2;
    int start, end;
    int score;
} hsa_aln1_t;

typedef struct hsa_width_t {            /* == bwt_width_t, bwtaln.h:35-38 */
    uint32_t w;
    int bid;
} hsa_width_t;

/* The fields of the reference's `BWT` (BWT.h:61-83) that the path reads, as loaded by BWTLoad
 * (BWT.c:107-223).  Host pointers; the arrays keep the reference's own layout (MSB-first 2-bit codes,
 * 16-bit minor / 32-bit major bidirectional occ samples every 256 / 65536 symbols). */
typedef struct hsa_bwt_view_t {
    uint32_t textLength;
    uint32_t inverseSa0;
    uint32_t cumulativeFreq[5];
    const uint32_t *bwtCode;       uint32_t bwtSizeInWord;
    const uint32_t *occValue;      uint32_t occSizeInWord;
    const uint32_t *occValueMajor; uint32_t occMajorSizeInWord;
} hsa_bwt_view_t;

typedef struct hsa_index hsa_index_t;   /* opaque: device-resident 2BWT search arrays (one GPU) */

int         hsa_b200_abi_version(void);
const char *hsa_last_error(void);
void        hsa_gap_opt_default(hsa_gap_opt_t *opt);            /* gap_init_opt, bwtaln.c:21-44 */
int         hsa_cal_maxdiff(int l, double err, double thres);   /* bwa_cal_maxdiff, bwtaln.c:46-58 (host) */

/* ---- index -------------------------------------------------------------------------------------
 * Replaces the in-memory result of BWTLoad2BWT (2BWT-Interface.c:13-66) for the search arrays: the two
 * BWTs are copied to `device` once and re-packed there into the device layout.   */
int  hsa_index_upload(int device, const hsa_bwt_view_t *fwd, const hsa_bwt_view_t *rev, hsa_index_t **out);
/* Adopt arrays that already live on `device` in the REFERENCE layout (built there, or received by an
 * NCCL broadcast); pointers are device pointers, the view's scalar fields are host values. */
int  hsa_index_from_device(int device, const hsa_bwt_view_t *fwd_dev, const hsa_bwt_view_t *rev_dev, hsa_index_t **out);
/* Device pointer + byte size of the packed device-layout blocks of one direction (0 fwd, 1 rev), so a
 * caller can broadcast a built index to peer GPUs; and the inverse: wrap received blocks. */
int  hsa_index_blocks(const hsa_index_t *idx, int which, void **dev_ptr, size_t *bytes);
int  hsa_index_from_blocks(int device, const uint32_t meta_fwd[7], const uint32_t meta_rev[7],
                           void *blocks_fwd_dev, void *blocks_rev_dev, int take_ownership, hsa_index_t **out);
int  hsa_index_meta(const hsa_index_t *idx, int which, uint32_t meta[7]); /* textLength, inverseSa0, C[0..4] */
void hsa_index_free(hsa_index_t *idx);

/* ---- rank: BWTAllOccValue (BWT.c:793-837) / BWTOccValue (BWT.c:682-719) ---------------------------
 * which: 0 = forward BWT, 1 = reverse BWT.  layout: 0 = evaluate on the reference-layout arrays,
 * 1 = on the re-packed device layout.  occ4_out[n*4]; occ1_out[n*4] (one BWTOccValue per character). */
int  hsa_occ_batch(const hsa_index_t *idx, int which, int layout, const uint32_t *indices, size_t n,
                   uint32_t *occ4_out, uint32_t *occ1_out);

/* ---- bwt_cal_width (bwtaln.c:73-116), type 1 (forward search on rev_bwt, :85-97) -------------------
 * Type 0 (the backward variant, :99-111, used only by the splice extension) is refused with HSA_E_ARG.
 * codes: concatenated base codes (0..3, N = 4); read r is codes[off[r] .. off[r]+len[r]).
 * width_out: sum(len[r]+1) entries, read r at off[r]+r; bid_out[r] = return value. */
int  hsa_cal_width_batch(const hsa_index_t *idx, const uint8_t *codes, const uint64_t *off,
                         const uint32_t *len, size_t n, int type, hsa_width_t *width_out, int *bid_out);

/* ---- bwt_match_gap (bwtgap.c:118-331), batched --------------------------------------------------- */
#define HSA_SEED_NONE   0   /* aux->width_seed == NULL                                   (bwtgap.c:919,1192) */
#define HSA_SEED_TAIL   1   /* width_seed over the last opt->seed_len bases              (bwtaln.c:344-346)  */
#define HSA_SEED_ALIAS  2   /* width_seed aliases width_back                             (bwtgap.c:809)      */

typedef struct hsa_task_t {
    uint64_t read_off;    /* offset of the READ (forward strand, as given) in `codes`                         */
    uint32_t read_len;    /* length of the read                                                               */
    uint32_t strand;      /* aux->strand: 1 = search the reverse complement of the read, 0 = the read itself  */
    uint32_t sub_off;     /* the searched sequence starts here inside the strand-resolved read                */
    uint32_t len;         /* aux->len: number of bases searched                                               */
    uint32_t wsrc_off;    /* width_back is computed on [wsrc_off, wsrc_off+len) of the strand-resolved read
                             (== sub_off everywhere except the splice seeds, bwtgap.c:807-808, where it is 0) */
    uint32_t seed_mode;   /* HSA_SEED_*                                                                       */
    uint32_t opt_idx;     /* index into opts[]: the gap_opt_t the reference would pass in aux->opt            */
    uint32_t reserved;
} hsa_task_t;

typedef struct hsa_result_t {     /* flat result of a batch.  ZERO-INITIALISE before first use; the library
                                     (re)allocates the three arrays as pinned host memory and re-uses them
                                     across calls when large enough; release with hsa_result_free          */
    size_t      n_items;          /* tasks (hsa_match_gap_batch), reads (whole) or 6 x reads (seeds)     */
    int32_t    *n_aln;            /* [n_items]                                                          */
    uint64_t   *aln_off;          /* [n_items] first hit of the item in aln[]                           */
    hsa_aln1_t *aln;              /* all hits, per item in discovery order (== reference order)         */
    size_t      n_aln_total;
    uint64_t    occ_lookups;      /* BWTAllOccValue + BWTOccValue calls the reference would have issued */
    uint64_t    n_strict;         /* searches the fast kernel handed on to the cooperative / large kernels */
    uint64_t    pops, steps;      /* diagnostics: stack pops and worker iterations                      */
    float       kernel_ms;        /* device time of the search kernel(s), CUDA events on the stream     */
    uint32_t    kernel_launches;
    size_t      cap_items, cap_aln; /* library-managed capacities                                       */
} hsa_result_t;

/* One reference bwt_cal_width (+ seed width) + bwt_match_gap call per task.  Host buffers in, host
 * buffers out (pageable or pinned). */
int  hsa_match_gap_batch(const hsa_index_t *idx, const uint8_t *codes, size_t codes_bytes,
                         const hsa_task_t *tasks, size_t n_tasks,
                         const hsa_gap_opt_t *opts, size_t n_opts, hsa_result_t *res);

/* bwt_match_gap (bwtgap.h:26; bwtgap.c:118-331) itself, ONE call, with the arguments bwt_aux_t carries: the
 * strand-resolved sequence (aux->strand ? aux->rc_seq : aux->seq), aux->len, aux->width_back (len + 1 entries, IN/OUT:
 * gap_shadow rewrites it in place, bwtgap.c:217, and bwt_splice_match reads it again afterwards), aux->width_seed (NULL,
 * == width_back as at bwtgap.c:809, or the opt->seed_len + 1 entries of bwtaln.c:344-346) and aux->opt.  *aln_out is a
 * malloc-family array of max(n_aln, 10) entries owned by the caller (free()), zero-filled beyond n_aln like the
 * reference's calloc (bwtgap.c:137-138); `strand` (aux->strand) is stamped on every hit (bwtgap.c:235), start / end /
 * type stay 0 as bwt_match_gap leaves them.
 * A batch of one: every call costs a host round trip (tens of microseconds); the batched entry points below are
 * the ones to build pipelines on.  Serves the callers at bwtaln.c:350 and bwtgap.c:812, 919, 1192. */
int  hsa_match_gap_call(const hsa_index_t *idx, const uint8_t *seq, uint32_t len, int strand, hsa_width_t *width_back,
                        hsa_width_t *width_seed, const hsa_gap_opt_t *opt, int *n_aln_out, hsa_aln1_t **aln_out);

/* The whole-read part of bwa_cal_sa_reg_gap (bwtaln.c:303-360, 371-372) for n reads with ONE caller
 * gap_opt_t: per-read filters (too many N, poly-A/T prefix), per-read max_diff from opt->fnr
 * (bwtaln.c:330-331), seed_len clamp (:332), reverse-complement strand first, forward strand only if
 * that found nothing, strand stamped on every hit and start/end on the first.  mode & GAPE is cleared as
 * bwtaln.c:261 does unless keep_gape != 0.  No splice fallback: reads with n_aln == 0 are what the
 * caller hands to bwt_splice_match. */
int  hsa_whole_reads(const hsa_index_t *idx, const uint8_t *codes, const uint64_t *off, const uint32_t *len,
                     size_t n_reads, const hsa_gap_opt_t *opt, int keep_gape, hsa_result_t *res);

/* Asynchronous form of hsa_whole_reads for double-buffered pipelines: submit enqueues the H2D copies and the
 * kernels and returns; hsa_job_wait finishes the batch, copies the results into `res` and releases the job.
 * The host buffers must stay valid (pinned memory recommended) until the job has been waited for; at most three
 * jobs may be in flight per index; jobs complete in submission order. */
typedef struct hsa_job hsa_job_t;
int  hsa_whole_reads_submit(const hsa_index_t *idx, const uint8_t *codes, const uint64_t *off, const uint32_t *len,
                            size_t n_reads, const hsa_gap_opt_t *opt, int keep_gape, hsa_job_t **job);
int  hsa_job_wait(hsa_job_t *job, hsa_result_t *res);

/* The six seed searches of bwt_splice_match (bwtgap.c:797-820) for each read: item 6*r+s is seed s%3 of
 * strand s/3, with the reference's prefix-width quirk; hits carry start/end as bwtgap.c:816-819. */
int  hsa_splice_seeds(const hsa_index_t *idx, const uint8_t *codes, const uint64_t *off, const uint32_t *len,
                      size_t n_reads, const hsa_gap_opt_t *opt, hsa_result_t *res);
/* asynchronous form (same job rules as hsa_whole_reads_submit; finish with hsa_job_wait) */
int  hsa_splice_seeds_submit(const hsa_index_t *idx, const uint8_t *codes, const uint64_t *off, const uint32_t *len,
                             size_t n_reads, const hsa_gap_opt_t *opt, hsa_job_t **job);

void hsa_result_free(hsa_result_t *res);

/* ---- device-resident variants for pipelines that keep reads / results in HBM (bench `value`) ------
 * All pointers are device pointers on the index's device; `stream` is a cudaStream_t passed as void*.
 * Results stay on the device: n_aln_dev[n_reads], aln_off_dev[n_reads], aln_dev[aln_capacity] and an
 * 8 x uint64 stats block {-, hits, lookups, heavy, bad, pops, steps, unprocessed}.  Nothing is synchronised:
 * the fast kernel and, behind it, the warp-cooperative kernel for the `heavy` searches it handed on are
 * queued on `stream` (in as many rounds as it takes to cover the batch, whatever share of it is heavy).  Completion is
 * verified with hsa_workspace_check(): it waits for the call's stream and
 * returns HSA_E_CAPACITY unless every read was searched to the end and every hit fits aln_capacity (searches
 * even the cooperative kernel could not hold leave n_aln = 0 behind; such batches go through hsa_whole_reads,
 * which finishes them with the large-capacity kernel).  One workspace serves one stream at a time. */
/* codes_dev is read in aligned 32-bit words: it must be readable up to the next 4-byte boundary past its last base
 * (any cudaMalloc / framework allocation is; a sub-allocation ending exactly on a page end is not). */
typedef struct hsa_workspace hsa_workspace_t;
int  hsa_workspace_create(const hsa_index_t *idx, size_t max_reads, uint32_t max_len, size_t aln_capacity,
                          hsa_workspace_t **out);
void hsa_workspace_free(hsa_workspace_t *ws);
int  hsa_whole_reads_device(const hsa_index_t *idx, hsa_workspace_t *ws, const uint8_t *codes_dev,
                            const uint64_t *off_dev, const uint32_t *len_dev, size_t n_reads,
                            const uint32_t *lens_present, size_t n_lens_present,   /* host: distinct read lengths */
                            const hsa_gap_opt_t *opt, int keep_gape, int32_t *n_aln_dev, uint64_t *aln_off_dev,
                            hsa_aln1_t *aln_dev, size_t aln_capacity, uint64_t *stats_dev, void *stream);
/* waits for the last hsa_whole_reads_device call of this workspace and checks that its results are complete;
 * stats_out (may be NULL) receives the statistics block with word 7 = searches left unprocessed */
int  hsa_workspace_check(hsa_workspace_t *ws, uint64_t stats_out[8]);
/* number of kernels the last call on this workspace launched (for bench's gpu_launches) */
uint32_t hsa_workspace_last_launches(const hsa_workspace_t *ws);
/* how the per-lane search kernel of the last call was configured (reporting only): out[0] = launch bound = resident blocks per SM
 * the register allocation allows, out[1] = score buckets per lane (a search that files a record of a higher score is handed to
 * the cooperative kernel), out[2] = shared memory per block in bytes, out[3] = 1 if the bound bytes sit in shared memory */
void hsa_workspace_last_config(const hsa_workspace_t *ws, uint32_t out[4]);
/* per-launch timing of the workspace's next calls (bench.py's roofline of the dominant kernel): enable records one
 * CUDA event behind every kernel launch; launch_times waits for the device and returns, for the last call, the
 * launches' names ("width1;search1;width2;search2;...": passes 1/2 of the width and search kernels, suffix C for the
 * warp-cooperative stage) with their durations in ms, and the occ lookups of the searches that the per-lane search
 * kernel's launches completed (the algorithmic work of that kernel; the width passes' lookups are not in it). */
int  hsa_workspace_launch_timing(hsa_workspace_t *ws, int enable);
int  hsa_workspace_launch_times(hsa_workspace_t *ws, char *names, size_t names_cap, float *ms, size_t ms_cap,
                                size_t *n_out, uint64_t *fast_search_lookups);

/* ---- SA index -> text position ---------------------------------------------------
 * Replaces BWTSaValue (BWT.c:1195-1225), the kernel of BWTRetrievePositionFromSAIndex (2BWT-Interface.c:329-362) as
 * called by bwa_cal_pac_pos (bwtse.c:350-369) and bwt_aln_corelate_check (bwtgap.c:669-742): walk BWTPsiMinusValue
 * (BWT.c:1142-1165) from the SA index to a sampled one and add the steps walked.
 * attach: `sa_value` = the reference's loaded BWT::saValue of the FORWARD bwt ((textLength + saInterval) / saInterval
 * words, BWT.c:219; entry 0 is forced to -1 as BWT.c:222 does), copied to the device once.
 * hsa_sa_values: host buffers, n SA indices (each <= textLength, else HSA_E_ARG) -> n values, bit-identical to
 * BWTSaValue's, including its SA[0] = -1 convention; *steps_total (may be NULL) = PsiMinus steps walked in total.
 * hsa_sa_values_device: the same with device buffers on `stream` (cudaStream_t as void*), nothing synchronised;
 * indices are not range-checked.  Each call takes its own work cursor; at most 8 calls may be in flight per index. */
int  hsa_index_attach_sa(hsa_index_t *idx, const uint32_t *sa_value, size_t n_words, uint32_t sa_interval);
int  hsa_sa_values(const hsa_index_t *idx, const uint32_t *sa_index, size_t n, uint32_t *sa_value_out, uint64_t *steps_total);
int  hsa_sa_values_device(const hsa_index_t *idx, const uint32_t *sa_index_dev, size_t n, uint32_t *sa_value_out_dev,
                          uint64_t *steps_total_dev, void *stream);
/* The whole of BWTRetrievePositionFromSAIndex (2BWT-Interface.c:329-362): BWTSaValue, then the block search over
 * HSP::blockList (HSP.h:41-46, read from <prefix>.index.ann by HSPLoad, HSP.c:85-106).  attach: n_blocks rows
 * {chrID, blockStart, blockEnd, ori} as the reference holds them (ascending, disjoint; else HSA_E_ARG).
 * hsa_sa_locate: per SA index *occ_pos = BWTSaValue, *seq_id = chrID and *ori_pos = occ_pos - blockStart + ori + 1 of the
 * block holding it; where no block holds the position (only SA[0] = -1) seq_id and ori_pos are 0xFFFFFFFF (the reference
 * leaves its outputs untouched there). */
int  hsa_index_attach_blocks(hsa_index_t *idx, const uint32_t *blocks4, uint32_t n_blocks);
int  hsa_sa_locate(const hsa_index_t *idx, const uint32_t *sa_index, size_t n, uint32_t *occ_pos_out, uint32_t *seq_id_out,
                   uint32_t *ori_pos_out);

/* ---- the spliced-read fallback ----------------------------------------------------------
 * Replaces bwt_splice_match (bwtgap.h:31; bwtgap.c:748-1332) as bwa_cal_sa_reg_gap calls it for every read that found
 * nothing on either strand (bwtaln.c:362-369), with everything under it: the six seed searches, bwt_aln_corelate_check
 * (bwtgap.c:669-742), splice_site_search_from_pos / check_site_by_intron_end on the packed text (:523-635),
 * bwt_extend_backward / _foreward = bwt_backtracing_search (:346-511, 640-663), bwt_cal_width types 1 and 0
 * (bwtaln.c:73-116), the 12-base anchor searches (:919, :1192).  One read per GPU thread; results are bit-identical to the
 * reference's, quirks included (hsa_b200/csrc/hsa_splice.cuh lists them).
 * Needs hsa_index_attach_sa, hsa_index_attach_blocks and hsa_index_attach_packed_dna (HSP::packedDNA as DNALoadPacked
 * leaves it -- 16 symbols per 32-bit word, first symbol in the two MSBs -- and HSP::dnaLength; HSP.c:67-82).
 * opts[opt_idx[r]] (opt_idx NULL: opts[0]) = the gap_opt_t the driver holds in aux->opt when it calls bwt_splice_match
 * for read r, i.e. local_opt after its per-read writes (bwtaln.c:254, 273-276, 330-332, 363); reads need >= 36 bases.
 * n_aln_out[r] in {0, 1, 2}; aln_out[2 r], aln_out[2 r + 1] = the reference's res_aln[0..1] (type = BWA_TYPE_SPLICING,
 * start / end = the read span of each part).  *occ_lookups (may be NULL) = rank lookups issued. */
int  hsa_index_attach_packed_dna(hsa_index_t *idx, const uint32_t *packed_dna, uint32_t dna_length);
int  hsa_splice_match_batch(const hsa_index_t *idx, const uint8_t *codes, const uint64_t *off, const uint32_t *len,
                            size_t n_reads, const hsa_gap_opt_t *opts, size_t n_opts, const uint32_t *opt_idx,
                            int32_t *n_aln_out, hsa_aln1_t *aln_out, uint64_t *occ_lookups);

/* ---- from hits to SAM fields -----------------------------------------------------------------
 * Replaces what generate_sam_se_core (bwtse.h:27; bwtse.c:884-931) computes for a batch before it prints: the hit selection
 * of bwt_aln2seq_core (bwtse.c:21-113; n_occ = 3 at bwtaln.c:514), bwa_cal_pac_pos (bwtse.c:350-369) with
 * bwt_aln2pos_splicing / bwt_combine_segment_splice for spliced hits (:197-348), bwa_refine_gapped (:536-638) with
 * refine_gapped_core (:380-440) = aln_global_core (stdaln.c:345-524) + bwa_aln_path2cigar (bwtaln.c:624-634), and
 * bwa_cal_md1 (bwtse.c:442-494).  Needs hsa_index_attach_sa / _blocks / _packed_dna.
 * The selection consumes the reference's process-wide drand48 stream in read order; *rng48_state is that stream's 48-bit
 * state, in and out -- 0 for a process that has not drawn yet (glibc).  Everything runs on the GPU: the selection (reads with
 * one best hit jump to their place in the stream, reads with several are walked by one thread; hsa_sam.cuh), positions,
 * pairing of spliced parts, the banded dynamic programme, CIGAR, MD and NM.
 * Inputs: the reads as for hsa_whole_reads, and per read its hits as the driver leaves them in bwa_seq_t::n_aln / aln
 * (hsa_whole_reads' result, with hsa_splice_match_batch's two parts for the reads it rescued); opt: the caller's gap_opt_t
 * as generate_sam_se_core sees it (fnr > 0: max_diff per read length as bwa_cal_pac_pos_core does, else opt->max_diff).
 * Results (library-managed host arrays, released by hsa_sam_result_free; zero-initialise the struct before first use):
 * rec[r] = the bwa_seq_t fields of read r; cigar words are bwa_cigar_t (op << 28 | len); md strings are not terminated. */
typedef struct hsa_sam1_t {            /* bwa_seq_t (bwtaln.h:92-122), the fields generate_sam_se_core writes */
    uint32_t type, strand, n_mm, n_gapo, n_gape, mapQ;
    int32_t  score;
    uint32_t sa, seq_id, ori_pos, occ_pos, c1, c2;
    int32_t  start, end;
    uint32_t n_cigar, cigar_off;        /* cigar[cigar_off .. + n_cigar); n_cigar == 0: no CIGAR array (prints "<len>M") */
    uint32_t nm, md_len, md_off;        /* md[md_off .. + md_len); md_len == 0: no MD                                   */
    uint32_t n_multi, multi_off;        /* multi[multi_off .. + n_multi): the alternative hits (XA)                      */
} hsa_sam1_t;
typedef struct hsa_multi1_t {          /* bwt_multi1_t (bwtaln.h:82-90) */
    uint32_t n_cigar, cigar_off, gap, mm, strand, sa, ori_pos, occ_pos, seq_id, aln_id;
    int32_t  start, end;
} hsa_multi1_t;
typedef struct hsa_sam_result_t {
    size_t        n_reads;
    hsa_sam1_t   *rec;
    hsa_multi1_t *multi;  size_t n_multi;
    uint32_t     *cigar;  size_t n_cigar;
    char         *md;     size_t md_bytes;
    uint64_t      n_refined;            /* reads that went through the dynamic programme */
    float         kernel_ms;            /* device time of the two kernels */
    size_t        cap_rec, cap_multi, cap_cigar, cap_md;
} hsa_sam_result_t;
int  hsa_sam_se_batch(const hsa_index_t *idx, const uint8_t *codes, const uint64_t *off, const uint32_t *len, size_t n_reads,
                      const int32_t *n_aln, const uint64_t *aln_off, const hsa_aln1_t *aln, const hsa_gap_opt_t *opt,
                      int n_occ, uint64_t *rng48_state, hsa_sam_result_t *res);
void hsa_sam_result_free(hsa_sam_result_t *res);
/* The same stage with reads AND hits resident in HBM (device pointers; e.g. the n_aln / aln_off / aln arrays
 * hsa_whole_reads_device wrote): nothing crosses PCIe but a few counters.  max_len = the longest read of the batch.  The call
 * waits for `stream` (cudaStream_t as void*) where it has to size arrays; on return the results are complete and stay on the
 * device, in library-managed arrays that remain valid until the next SAM call on this index. */
typedef struct hsa_sam_device_t {
    const hsa_sam1_t   *rec_dev;        /* [n_reads]  */
    const hsa_multi1_t *multi_dev;      /* [n_multi]  */
    const uint32_t     *cigar_dev;      /* [n_cigar]  */
    const char         *md_dev;         /* [md_bytes] */
    size_t   n_multi, n_cigar, md_bytes;
    uint64_t n_refined, n_several_best; /* reads through the dynamic programme; reads whose best score is held by several hits */
    float    kernel_ms;                 /* first kernel to last kernel, CUDA events on `stream` (sizing syncs included) */
} hsa_sam_device_t;
/* plain cudaMemcpy device -> host on the index's device, for callers without a CUDA runtime of their own (ctypes / cgo) that
 * want to look at device-resident results */
int  hsa_copy_from_device(const hsa_index_t *idx, void *dst_host, const void *src_dev, size_t bytes);
int  hsa_sam_se_device(const hsa_index_t *idx, const uint8_t *codes_dev, const uint64_t *off_dev, const uint32_t *len_dev, size_t n_reads,
                       uint32_t max_len, const int32_t *n_aln_dev, const uint64_t *aln_off_dev, const hsa_aln1_t *aln_dev,
                       const hsa_gap_opt_t *opt, int n_occ, uint64_t *rng48_state, void *stream, hsa_sam_device_t *out);
/* bwa_print_sam1 (bwtse.c:677-835) for single-end reads without qualities, read groups or barcodes: the SAM lines of reads
 * [first, first + count) whose type is not BWA_TYPE_NO_MATCH, as generate_sam_se_core prints them (bwtse.c:922-926).
 * names[r] (NULL: "r<r>"), chr_names[seq_id] = HSP::chrName; mode / max_top2 from the caller's gap_opt_t.  *text_out is a
 * malloc'ed buffer of *bytes_out bytes (not terminated) owned by the caller.  Host-only formatting. */
int  hsa_sam_format(const hsa_sam_result_t *res, size_t first, size_t count, const uint8_t *codes, const uint64_t *off,
                    const uint32_t *len, const char *const *names, const char *const *chr_names, size_t n_chr,
                    const hsa_gap_opt_t *opt, char **text_out, size_t *bytes_out);

/* ---- roofline probe: random 32-byte-sector loads over `footprint_bytes` ----
 * The random-access denominator on this GPU: achieved GB/s (sectors * 32 B / time) at full occupancy of six access
 * shapes -- [0] four dependent chains per thread with two 16-byte loads per sector, [1..3] 4 / 8 / 16 independent 256-bit
 * loads in flight per thread (the kernels' own load shape, multiply-shift sector choice), [4] the SA kernel's mix (a
 * 256-bit load plus a 4-byte load from another sector), [5] the LF-walk shape (one data-dependent chain per thread of geometric
 * length, ended by a 4-byte load from a second array; 9/8 sectors per step).  hsa_random_sector_probe returns the best of them,
 * _ex all of them (gbs_out[0..min(n_out,6))).  Used only by bench.py / tools/ to establish the roofline. */
int  hsa_random_sector_probe(int device, size_t footprint_bytes, int iters, double *gbs_out);
int  hsa_random_sector_probe_ex(int device, size_t footprint_bytes, int iters, double *gbs_out, int n_out);

#ifdef __cplusplus
}
#endif
#endif /* HSA_B200_H */

/*
 * hsa_gpu_shim.c -- the reference-side binding of libhsa_b200.so, as real code.
 *
 * This is the file a maintainer adds to HSA: it includes the reference's own headers (bwtaln.h, bwtgap.h) and
 * include/hsa_b200.h and provides
 *     bwa_cal_sa_reg_gap_gpu()  -- same signature and same observable results as bwa_cal_sa_reg_gap
 *                                  (bwtaln.c:246-417), with the whole-read searches done on the GPU.
 *     hsa_gpu_sa_values()       -- BWTSaValue (BWT.c:1195) for a batch of SA indices.
 *     generate_sam_se_core_gpu() -- same signature and same printed SAM as generate_sam_se_core (bwtse.c:884-931): hit
 *                                  selection, positions, gapped refinement (CIGAR) and MD / NM come from hsa_sam_se_batch;
 *                                  the lines are printed by the reference's own bwa_print_sam1.
 * bwt_splice_match for the reads that found nothing (bwtgap.c:748) runs on the GPU when the full index is loaded
 * (HSA_GPU_SPLICE=0 keeps it on the host).
 *
 * It is compiled only where the reference tree exists (oracle/Makefile target `shim`, outputs under oracle/_ref/)
 * and is exercised by tests/test_gpu_shim.py, which compares its per-read output with the stock driver's.
 *
 * Reference behaviours mirrored on purpose:
 *   - local_opt is copied BEFORE MODE_GAPE is cleared in the caller's opt (bwtaln.c:254, 260-261);
 *   - after the first read of a batch that falls through to the splice path, aux->opt points at local_opt for
 *     the rest of the batch (:363): later reads are searched with local_opt's mode / max_gapo;
 *   - aux->opt->max_diff / seed_len are written through aux->opt for every read (:330-332), i.e. into the
 *     caller's opt before the switch and into local_opt after it;
 *   - the N filter compares with local_opt.max_diff (:314-317), which drifts with those writes after the switch.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bwtaln.h"
#include "bwtgap.h"
#include "BWT.h"
#include "bwtse.h"
#include "hsa_b200.h"
#include <time.h>
#include <pthread.h>
#include <unistd.h>

/* HSA_GPU_SHIM_TIMING=1: seconds per phase of bwa_cal_sa_reg_gap_gpu, summed over the calls, printed by hsa_gpu_close */
static double g_t[12]; static int g_timing = -1;
static double now_s_(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
#define TICK(k) do { if (g_timing > 0) { double t_ = now_s_(); g_t[k] += t_ - t_last; t_last = t_; } } while (0)

static hsa_index_t *g_idx = NULL;
static hsa_result_t g_res_a, g_res_b;
static void g_sa_attached_flag(void);
static void pool_stop(void);

static void view_of(const BWT *b, hsa_bwt_view_t *v)          /* BWT.h:61-83 */
{
    int i;
    v->textLength = b->textLength; v->inverseSa0 = b->inverseSa0;
    for (i = 0; i < 5; ++i) v->cumulativeFreq[i] = b->cumulativeFreq[i];
    v->bwtCode = b->bwtCode;             v->bwtSizeInWord = b->bwtSizeInWord;
    v->occValue = b->occValue;           v->occSizeInWord = b->occSizeInWord;
    v->occValueMajor = b->occValueMajor; v->occMajorSizeInWord = b->occMajorSizeInWord;
}

static int g_splice_gpu = 0;    /* bwt_splice_match on the GPU too (needs the full index: SA samples, annotation, packed text) */

int hsa_gpu_open(const Idx2BWT *bi, int device)                /* call once after BWTLoad2BWT, bwtaln.c:469 */
{
    hsa_bwt_view_t f, r;
    const char *env = getenv("HSA_GPU_SPLICE");
    view_of(bi->bwt, &f); view_of(bi->rev_bwt, &r);
    if (hsa_index_upload(device, &f, &r, &g_idx)) { fprintf(stderr, "[hsa_gpu] %s\n", hsa_last_error()); return -1; }
    g_splice_gpu = 0;
    if (bi->hsp && bi->hsp->packedDNA && bi->bwt->saValue && (!env || atoi(env) != 0)) {
        /* what bwt_splice_match reads besides the two BWTs: BWT::saValue (BWT.c:205-223), HSP::blockList (HSP.c:85-106),
         * HSP::packedDNA / dnaLength (HSP.c:74-77) */
        const HSP *h = bi->hsp;
        uint32_t *b4 = (uint32_t *)malloc(sizeof(uint32_t) * 4 * (size_t)(h->numOfBlock > 0 ? h->numOfBlock : 1));
        int i;
        for (i = 0; i < h->numOfBlock; ++i) {
            b4[4 * i] = (uint32_t)h->blockList[i].chrID; b4[4 * i + 1] = h->blockList[i].blockStart;
            b4[4 * i + 2] = h->blockList[i].blockEnd; b4[4 * i + 3] = h->blockList[i].ori;
        }
        if (hsa_index_attach_sa(g_idx, bi->bwt->saValue, bi->bwt->saValueSizeInWord, bi->bwt->saInterval) ||
            hsa_index_attach_blocks(g_idx, b4, (uint32_t)h->numOfBlock) ||
            hsa_index_attach_packed_dna(g_idx, h->packedDNA, h->dnaLength)) {
            fprintf(stderr, "[hsa_gpu] splice path stays on the host: %s\n", hsa_last_error());
        } else { g_splice_gpu = 1; g_sa_attached_flag(); }
        free(b4);
    }
    return 0;
}

/* BWTSaValue (BWT.c:1195-1225) for a batch of SA indices: what BWTRetrievePositionFromSAIndex (2BWT-Interface.c:339)
 * looks up for bwa_cal_pac_pos (bwtse.c:350-369) and bwt_aln_corelate_check (bwtgap.c:669-742).  The SA samples are
 * handed to the library on first use (they need the full index, which hsa_gpu_open's search-only callers may not load). */
static int g_sa_attached = 0;
static void g_sa_attached_flag(void) { g_sa_attached = 1; }
void hsa_gpu_sa_values(const Idx2BWT *bi, const unsigned int *sa_index, size_t n, unsigned int *occ_pos)
{
    if (!g_sa_attached) {
        const BWT *b = bi->bwt;
        if (!b->saValue || hsa_index_attach_sa(g_idx, b->saValue, b->saValueSizeInWord, b->saInterval)) {
            fprintf(stderr, "[hsa_gpu] SA samples: %s\n", b->saValue ? hsa_last_error() : "index loaded without them");
            exit(1);
        }
        g_sa_attached = 1;
    }
    if (hsa_sa_values(g_idx, sa_index, n, occ_pos, NULL)) { fprintf(stderr, "[hsa_gpu] %s\n", hsa_last_error()); exit(1); }
}

void hsa_gpu_close(void)
{
    if (g_timing > 0)
        fprintf(stderr, "[hsa_gpu] seconds: pack %.4f | pass A (hsa_whole_reads) %.4f | first leak + pass B %.4f | order-dependent pass %.4f | "
                        "hit arrays (splice batch in flight) %.4f | wait for the splice batch %.4f | frees %.4f\n",
                g_t[0], g_t[1], g_t[2], g_t[3], g_t[4], g_t[5], g_t[6]);
    if (g_timing > 0 && g_t[9] > 0)
        fprintf(stderr, "[hsa_gpu] SAM stage seconds: pack %.4f | hsa_sam_se_batch %.4f | fields into bwa_seq_t %.4f | SAM text %.4f\n",
                g_t[8], g_t[9], g_t[10], g_t[11]);
    pool_stop();
    g_sa_attached = 0;
    hsa_result_free(&g_res_a); hsa_result_free(&g_res_b);
    hsa_index_free(g_idx); g_idx = NULL;
}

static void die_gpu(void) { fprintf(stderr, "[hsa_gpu] %s\n", hsa_last_error()); exit(1); }

/* bwa_cal_maxdiff (bwtaln.c:46-58) is a pure function of (length, fnr): one table per fnr instead of an exp() per read */
static int maxdiff_of(int len, float fnr)
{
    static int tab[4096]; static float tab_fnr = -1.0f;
    if (len < 0 || len >= 4096) return bwa_cal_maxdiff(len, BWA_AVG_ERR, fnr);
    if (tab_fnr != fnr) { memset(tab, 0xff, sizeof(tab)); tab_fnr = fnr; }
    if (tab[len] < 0) tab[len] = bwa_cal_maxdiff(len, BWA_AVG_ERR, fnr);
    return tab[len];
}

/* same signature and contract as bwt_match_gap (bwtgap.h:26; bwtgap.c:118-331): one search on the GPU with the frame the
 * caller built in `aux` -- sequence by aux->strand, aux->len, aux->width_back (rewritten in place by gap_shadow, as the
 * reference does), aux->width_seed, aux->opt.  Returns a malloc-family array the caller frees (never NULL).  Serves the
 * callers at bwtaln.c:350 and bwtgap.c:812, 919, 1192; a batch of one, so use it where call-for-call substitution
 * matters, not for throughput. */
bwt_aln1_t *bwt_match_gap_gpu(bwt_aux_t *aux, int *_n_aln)
{
    const ubyte_t *seq = aux->strand == 0 ? aux->seq : aux->rc_seq;
    hsa_aln1_t *aln = NULL;
    int n = 0;
    if (hsa_match_gap_call(g_idx, seq, (uint32_t)aux->len, aux->strand, (hsa_width_t *)aux->width_back, (hsa_width_t *)aux->width_seed,
                           (const hsa_gap_opt_t *)aux->opt, &n, &aln)) die_gpu();
    *_n_aln = n;
    return (bwt_aln1_t *)aln;
}

/* ---- helper threads ---------------------------------------------------------------------------------------------
 * What is left on the host per batch is byte shuffling: packing the reads for the GPU call, counting Ns, and giving every
 * read its own calloc'ed hit array (the contract of bwa_seq_t.aln).  None of it depends on the order of the reads, so it
 * is spread over a few helper threads (HSA_GPU_SHIM_THREADS, default min(8, online CPUs); 1 = none).  Everything whose
 * outcome depends on the order of the reads -- the option switch -- stays one sequential pass. */
typedef void (*range_fn)(void *ctx, int lo, int hi);
static struct {
    pthread_t th[32]; int n_helpers, started, stop;
    pthread_mutex_t mu; pthread_cond_t cv_go, cv_done;
    range_fn fn; void *ctx; int total, chunk, next, active; unsigned gen;
} g_pool = { .mu = PTHREAD_MUTEX_INITIALIZER, .cv_go = PTHREAD_COND_INITIALIZER, .cv_done = PTHREAD_COND_INITIALIZER };

static void pool_work(void)
{
    for (;;) {
        int lo = __atomic_fetch_add(&g_pool.next, g_pool.chunk, __ATOMIC_RELAXED), hi;
        if (lo >= g_pool.total) break;
        hi = lo + g_pool.chunk < g_pool.total ? lo + g_pool.chunk : g_pool.total;
        g_pool.fn(g_pool.ctx, lo, hi);
    }
}
static void *pool_main(void *arg)
{
    unsigned seen = 0;
    (void)arg;
    for (;;) {
        pthread_mutex_lock(&g_pool.mu);
        while (g_pool.gen == seen && !g_pool.stop) pthread_cond_wait(&g_pool.cv_go, &g_pool.mu);
        if (g_pool.stop) { pthread_mutex_unlock(&g_pool.mu); return NULL; }
        seen = g_pool.gen;
        pthread_mutex_unlock(&g_pool.mu);
        pool_work();
        pthread_mutex_lock(&g_pool.mu);
        if (--g_pool.active == 0) pthread_cond_signal(&g_pool.cv_done);
        pthread_mutex_unlock(&g_pool.mu);
    }
}
static void pool_start(void)
{
    const char *e = getenv("HSA_GPU_SHIM_THREADS");
    long want = e ? atol(e) : 8, cpus = sysconf(_SC_NPROCESSORS_ONLN);
    int i;
    g_pool.started = 1;
    if (want > cpus) want = cpus;
    if (want > 32) want = 32;
    for (i = 0; i + 1 < want; ++i) {
        if (pthread_create(&g_pool.th[g_pool.n_helpers], NULL, pool_main, NULL) != 0) break;
        ++g_pool.n_helpers;
    }
}
static void pool_stop(void)
{
    int i;
    if (!g_pool.started) return;
    pthread_mutex_lock(&g_pool.mu); g_pool.stop = 1; pthread_cond_broadcast(&g_pool.cv_go); pthread_mutex_unlock(&g_pool.mu);
    for (i = 0; i < g_pool.n_helpers; ++i) pthread_join(g_pool.th[i], NULL);
    g_pool.n_helpers = 0; g_pool.started = 0; g_pool.stop = 0;
}
/* fn over [0, total) in chunks, on the helpers and the calling thread; returns when all of it is done */
static void pool_run(range_fn fn, void *ctx, int total, int chunk)
{
    if (!g_pool.started) pool_start();
    if (g_pool.n_helpers == 0 || total <= chunk) { if (total > 0) fn(ctx, 0, total); return; }
    pthread_mutex_lock(&g_pool.mu);
    g_pool.fn = fn; g_pool.ctx = ctx; g_pool.total = total; g_pool.chunk = chunk; g_pool.next = 0;
    g_pool.active = g_pool.n_helpers; ++g_pool.gen;
    pthread_cond_broadcast(&g_pool.cv_go);
    pthread_mutex_unlock(&g_pool.mu);
    pool_work();
    pthread_mutex_lock(&g_pool.mu);
    while (g_pool.active > 0) pthread_cond_wait(&g_pool.cv_done, &g_pool.mu);
    pthread_mutex_unlock(&g_pool.mu);
}

/* per-read facts that do not depend on the option state: number of Ns (bwtaln.c:314-316) and the poly-A / poly-T test on
 * the first 15 bases (:324-325) */
enum { RD_POLY = 1 };
typedef struct {
    bwa_seq_t *seqs; uint8_t *codes; const uint64_t *off; uint32_t *nn; uint8_t *flag;
} pack_ctx_t;
static void pack_range(void *c_, int lo, int hi)
{
    pack_ctx_t *c = (pack_ctx_t *)c_;
    int i, j;
    for (i = lo; i < hi; ++i) {
        const bwa_seq_t *p = c->seqs + i;
        uint32_t nn = 0; int pa = 1, pt = 1;
        memcpy(c->codes + c->off[i], p->seq, p->len);
        for (j = 0; j < (int)p->len; ++j) nn += p->seq[j] > 3;
        for (j = 0; j < 15; ++j) { if (p->seq[j] != 0) pa = 0; if (p->seq[j] != 3) pt = 0; }
        c->nn[i] = nn; c->flag[i] = (pa || pt) ? RD_POLY : 0;
    }
}
/* the GPU's hits of one read as the array bwa_seq_t.aln owns (hits in the reference's order, strand stamped on every hit,
 * start / end on the first) */
enum { ST_SKIP = 0, ST_TAKE_A = 1, ST_TAKE_B = 2 };
typedef struct {
    bwa_seq_t *seqs; const uint8_t *state; const hsa_result_t *res_a, *res_b; int first_leak;
} take_ctx_t;
static void take_range(void *c_, int lo, int hi)
{
    take_ctx_t *c = (take_ctx_t *)c_;
    int i;
    for (i = lo; i < hi; ++i) {
        bwa_seq_t *p = c->seqs + i;
        const hsa_result_t *res; size_t ri;
        if (c->state[i] == ST_SKIP) continue;
        res = c->state[i] == ST_TAKE_B ? c->res_b : c->res_a;
        ri = c->state[i] == ST_TAKE_B ? (size_t)(i - c->first_leak - 1) : (size_t)i;
        p->n_aln = res->n_aln[ri];
        p->aln = (bwt_aln1_t *)calloc(p->n_aln < 10 ? 10 : p->n_aln, sizeof(bwt_aln1_t));
        memcpy(p->aln, res->aln + res->aln_off[ri], sizeof(bwt_aln1_t) * (size_t)p->n_aln);
    }
}
/* bwt_splice_match for the batch's fallback reads on its own thread, so that the hit arrays of the other reads are built
 * while the GPU works on these */
typedef struct {
    const uint8_t *pc; const uint64_t *po; const uint32_t *pl, *pi; size_t n; const gap_opt_t *tab; size_t n_tab;
    int32_t *pn; hsa_aln1_t *pa; int rc;
} splice_job_t;
static void *splice_main(void *a_)
{
    splice_job_t *a = (splice_job_t *)a_;
    a->rc = hsa_splice_match_batch(g_idx, a->pc, a->po, a->pl, a->n, (const hsa_gap_opt_t *)a->tab, a->n_tab, a->pi, a->pn, a->pa, NULL);
    return NULL;
}

/* same signature as bwa_cal_sa_reg_gap (bwtaln.h:199-200) */
void bwa_cal_sa_reg_gap_gpu(int tid, const Idx2BWT *bi_bwt, int n_seqs, bwa_seq_t *seqs, const gap_opt_t *opt_c, bwt_array_t *arr)
{
    gap_opt_t *opt = (gap_opt_t *)opt_c;                       /* the reference casts const away too (:260) */
    gap_opt_t local_opt = *opt;                                /* :254, BEFORE the clear */
    gap_opt_t opt_a, opt_b;
    bwt_aux_t *aux = (bwt_aux_t *)calloc(1, sizeof(bwt_aux_t));
    int i, max_len = 0, first_leak = -1, pass_b = 0;
    uint8_t *codes, *state, *flag;                             /* state: what the last phase does with the read (ST_*) */
    uint64_t *off, total = 0;
    uint32_t *len, *nn;
    int *pend_read = NULL; gap_opt_t *pend_opt = NULL; int n_pend = 0;   /* reads for the GPU splice batch + their aux->opt */
    double t_last = 0;
    (void)tid;
    if (g_timing < 0) { const char *e = getenv("HSA_GPU_SHIM_TIMING"); g_timing = e && atoi(e) ? 1 : 0; }
    if (g_timing > 0) t_last = now_s_();

    opt->mode &= ~BWA_MODE_GAPE;                               /* :261, sticks in the caller's struct */
    for (i = 0; i < n_seqs; ++i) if ((int)seqs[i].len > max_len) max_len = seqs[i].len;
    if (opt->fnr > 0.0) local_opt.max_diff = bwa_cal_maxdiff(max_len, BWA_AVG_ERR, opt->fnr);   /* :273-274 */
    if (local_opt.max_diff < local_opt.max_gapo) local_opt.max_gapo = local_opt.max_diff;        /* :275-276 */

    /* the options of the two GPU passes: before the switch (caller's opt, GAPE cleared) and after it (local_opt) */
    opt_a = *opt; opt_b = local_opt;
    opt_b.seed_len = opt->seed_len;

    /* ---- pass A on every read: GPU whole-read search with the pre-switch options --------------------------------
     * The per-read filters depend on state that drifts after the switch, so they are applied on the host further
     * down; the GPU's own N filter uses bwa_cal_maxdiff(max_len), never stricter than the drifting one. */
    off = (uint64_t *)malloc(sizeof(uint64_t) * (n_seqs + 1)); len = (uint32_t *)malloc(sizeof(uint32_t) * (n_seqs + 1));
    nn = (uint32_t *)malloc(sizeof(uint32_t) * (n_seqs + 1)); flag = (uint8_t *)malloc(n_seqs + 1);
    state = (uint8_t *)calloc(n_seqs + 1, 1);
    for (i = 0; i < n_seqs; ++i) { off[i] = total; len[i] = seqs[i].len; total += seqs[i].len; }
    codes = (uint8_t *)malloc(total + 16);
    if (g_splice_gpu) {
        pend_read = (int *)malloc(sizeof(int) * (n_seqs + 1));
        pend_opt = (gap_opt_t *)malloc(sizeof(gap_opt_t) * (n_seqs + 1));
    }
    { pack_ctx_t pc = { seqs, codes, off, nn, flag }; pool_run(pack_range, &pc, n_seqs, 4096); }
    TICK(0);
    if (n_seqs && hsa_whole_reads(g_idx, codes, off, len, (size_t)n_seqs, (const hsa_gap_opt_t *)&opt_a, 0, &g_res_a)) die_gpu();
    TICK(1);

    /* ---- sequential host part: filters, the option switch, splice fallback -------------------------------------- */
    aux->bi_bwt = (Idx2BWT *)bi_bwt; aux->arr = arr; aux->max_len = max_len;
    aux->stack = gap_init_stack(local_opt.max_diff, local_opt.max_gapo, local_opt.max_gape, &local_opt);   /* :279 */
    aux->width_back = (bwt_width_t *)calloc(max_len + 1, sizeof(bwt_width_t));
    aux->width_fore = (bwt_width_t *)calloc(max_len + 1, sizeof(bwt_width_t));
    aux->width_seed = (bwt_width_t *)calloc(max_len + 1, sizeof(bwt_width_t));
    aux->rc_seq = (ubyte_t *)calloc(max_len, sizeof(ubyte_t));
    aux->opt = opt;

    /* find the first read that falls through (it decides where the option switch happens) using pass-A results and
     * the pre-switch filters; then, if the two option sets differ, re-search the reads after it with local_opt */
    {
        int need_b = (local_opt.mode != opt->mode) || (local_opt.max_gapo != opt->max_gapo) ||
                     (opt->fnr <= 0.0 && local_opt.max_diff != opt->max_diff);
        for (i = 0; i < n_seqs && first_leak < 0; ++i) {
            if ((int)nn[i] > local_opt.max_diff) continue;
            if (flag[i] & RD_POLY) continue;
            if (g_res_a.n_aln[i] == 0) first_leak = i;
        }
        if (first_leak >= 0 && need_b && first_leak + 1 < n_seqs) {
            size_t k0 = (size_t)first_leak + 1, nb = (size_t)n_seqs - k0;
            uint64_t *off_b = (uint64_t *)malloc(sizeof(uint64_t) * nb);
            size_t q;
            for (q = 0; q < nb; ++q) off_b[q] = off[k0 + q] - off[k0];
            if (hsa_whole_reads(g_idx, codes + off[k0], off_b, len + k0, nb, (const hsa_gap_opt_t *)&opt_b,
                                (opt_b.mode & BWA_MODE_GAPE) ? 1 : 0, &g_res_b)) die_gpu();
            free(off_b);
            pass_b = 1;
        }
    }
    TICK(2);

    /* the order-dependent pass: what happens to every read (the hit arrays themselves are built afterwards, in parallel) */
    for (i = 0; i < n_seqs; ++i) {
        bwa_seq_t *p = seqs + i;
        const int use_b = pass_b && i > first_leak;
        const int32_t n_hits = use_b ? g_res_b.n_aln[i - first_leak - 1] : g_res_a.n_aln[i];
        if ((int)nn[i] > local_opt.max_diff) continue;         /* :314-317 (fields stay as bwa_read_seq left them) */
        p->sa = 0; p->type = BWA_TYPE_NO_MATCH; p->c1 = p->c2 = 0; p->n_aln = 0; p->aln = 0;    /* :319-323 */
        if (flag[i] & RD_POLY) continue;                       /* :324-325 */
        aux->seq = p->seq; aux->len = p->len;
        if (opt->fnr > 0.0) aux->opt->max_diff = maxdiff_of(p->len, opt->fnr);                    /* :330-331, through aux->opt */
        aux->opt->seed_len = opt->seed_len < (int)p->len ? opt->seed_len : 0x7fffffff;            /* :332 */
        if (n_hits > 0) { state[i] = use_b ? ST_TAKE_B : ST_TAKE_A; continue; }
        /* nothing on either strand: the splice path (bwtaln.c:362-369) with aux->opt = &local_opt as it stands NOW */
        if (g_splice_gpu && p->len >= 36) {
            aux->opt = &local_opt;
            pend_read[n_pend] = i; pend_opt[n_pend] = local_opt; ++n_pend;     /* searched in one GPU batch below */
            continue;
        }
        memset(aux->rc_seq, 0, max_len * sizeof(ubyte_t));
        memcpy(aux->rc_seq, p->seq, p->len * sizeof(ubyte_t));
        seq_reverse(p->len, aux->rc_seq, 1);
        aux->strand = 0;
        aux->opt = &local_opt;
        p->aln = bwt_splice_match(aux, &p->n_aln);
        if (p->n_aln == 0) { free(p->aln); p->aln = NULL; }
    }
    TICK(3);
    {
        /* bwt_splice_match for all fallback reads at once (hsa_splice_match_batch): distinct option states become an
         * option table, reads keep their order.  It runs on its own thread while the hit arrays are built. */
        uint8_t *pc = NULL; uint64_t *po = NULL; uint32_t *pl = NULL, *pi = NULL; int32_t *pn = NULL; hsa_aln1_t *pa = NULL;
        gap_opt_t *tab = NULL; int n_tab = 0, q, t, threaded = 0;
        splice_job_t job; pthread_t th;
        if (n_pend) {
            uint64_t tot = 0;
            for (q = 0; q < n_pend; ++q) tot += seqs[pend_read[q]].len;
            pc = (uint8_t *)malloc(tot + 16); po = (uint64_t *)malloc(sizeof(uint64_t) * n_pend);
            pl = (uint32_t *)malloc(sizeof(uint32_t) * n_pend); pi = (uint32_t *)malloc(sizeof(uint32_t) * n_pend);
            pn = (int32_t *)calloc(n_pend, sizeof(int32_t)); pa = (hsa_aln1_t *)calloc((size_t)n_pend * 2, sizeof(hsa_aln1_t));
            tab = (gap_opt_t *)malloc(sizeof(gap_opt_t) * n_pend);
            tot = 0;
            for (q = 0; q < n_pend; ++q) {
                bwa_seq_t *p = seqs + pend_read[q];
                po[q] = tot; pl[q] = p->len; memcpy(pc + tot, p->seq, p->len); tot += p->len;
                for (t = 0; t < n_tab; ++t) if (memcmp(tab + t, pend_opt + q, sizeof(gap_opt_t)) == 0) break;
                if (t == n_tab) tab[n_tab++] = pend_opt[q];
                pi[q] = (uint32_t)t;
            }
            job.pc = pc; job.po = po; job.pl = pl; job.pi = pi; job.n = (size_t)n_pend; job.tab = tab; job.n_tab = (size_t)n_tab;
            job.pn = pn; job.pa = pa; job.rc = 0;
            threaded = pthread_create(&th, NULL, splice_main, &job) == 0;
            if (!threaded) splice_main(&job);
        }
        { take_ctx_t tc = { seqs, state, &g_res_a, &g_res_b, first_leak }; pool_run(take_range, &tc, n_seqs, 2048); }
        TICK(4);
        if (n_pend) {
            if (threaded) pthread_join(th, NULL);
            if (job.rc) die_gpu();
            for (q = 0; q < n_pend; ++q) {
                bwa_seq_t *p = seqs + pend_read[q];
                p->n_aln = pn[q];
                if (pn[q]) {                                   /* res_aln: calloc(2, ...) in the reference (bwtgap.c:853) */
                    p->aln = (bwt_aln1_t *)calloc(2, sizeof(bwt_aln1_t));
                    memcpy(p->aln, pa + 2 * (size_t)q, sizeof(bwt_aln1_t) * 2);
                } else p->aln = NULL;                          /* bwtaln.c:366-369 */
            }
            free(pc); free(po); free(pl); free(pi); free(pn); free(pa); free(tab);
        }
    }
    TICK(5);
    free(pend_read); free(pend_opt);
    free(codes); free(off); free(len); free(nn); free(flag); free(state);
    free(aux->width_seed); free(aux->width_fore); free(aux->width_back); free(aux->rc_seq);
    gap_destroy_stack(aux->stack);
    free(aux);
    TICK(6);
}


/* same signature as generate_sam_se_core (bwtse.h:27; bwtse.c:884-931).  The fields it leaves in every bwa_seq_t come from
 * hsa_sam_se_batch; the printing loop (:922-926) is the reference's own bwa_print_sam1.  The reference draws its random
 * numbers from the process-wide drand48 stream and nothing else in HSA uses that stream, so the shim carries the stream's
 * 48-bit state itself (0 = a process that has not drawn yet). */
static uint64_t g_rng48 = 0;
static hsa_sam_result_t g_sam;
void bwa_print_sam1(const HSP *hsp, bwa_seq_t *p, const bwa_seq_t *mate, int mode, int max_top2);

static bwa_cigar_t *cigar_copy(const hsa_sam_result_t *r, uint32_t off, uint32_t n)
{
    bwa_cigar_t *c = (bwa_cigar_t *)malloc(sizeof(bwa_cigar_t) * (n ? n : 1));
    memcpy(c, r->cigar + off, sizeof(bwa_cigar_t) * n);
    return c;
}

/* ---- bwa_print_sam1 (bwtse.c:677-835) for the lines generate_sam_se_core prints (mate == NULL, type != NO_MATCH), into a
 * buffer.  The reference formats every line with a dozen printf calls and a putchar per base -- 2 us per read, ten times the
 * rest of the stage -- and nothing in a line depends on another read, so the shim formats chunks of reads on the helper threads
 * and writes the chunks to stdout in order.  Reads whose CIGAR holds an operation beyond the reference's letter tables (it indexes
 * past "MIDNSHP=X" / "MIDNS" there, bwtse.c:713, 811) are printed by the reference's own function, byte for byte what it prints. */
extern char *bwt_rg_id;
typedef struct { char *p; size_t n, cap; } tbuf_t;
static void tb_need(tbuf_t *b, size_t m)
{
    if (b->n + m > b->cap) {
        b->cap = (b->n + m) * 2 + 4096;
        b->p = (char *)realloc(b->p, b->cap);
        if (!b->p) { fprintf(stderr, "[hsa_gpu] out of memory formatting SAM text\n"); exit(1); }
    }
}
static void tb_str(tbuf_t *b, const char *s) { size_t m = strlen(s); tb_need(b, m); memcpy(b->p + b->n, s, m); b->n += m; }
static void tb_put(tbuf_t *b, char c) { tb_need(b, 1); b->p[b->n++] = c; }
static void tb_int(tbuf_t *b, long long v)                     /* %d / %lld */
{
    char t[24]; int k = 0; unsigned long long u = v < 0 ? 0ull - (unsigned long long)v : (unsigned long long)v;
    tb_need(b, 24);
    if (v < 0) b->p[b->n++] = '-';
    do { t[k++] = (char)('0' + u % 10); u /= 10; } while (u);
    while (k) b->p[b->n++] = t[--k];
}
static int sam_line_is_plain(const bwa_seq_t *p)
{
    int j, k;
    if (p->cigar) for (j = 0; j < p->n_cigar; ++j) if (__cigar_op(p->cigar[j]) > 8) return 0;
    for (j = 0; j < p->n_multi; ++j) {
        const bwt_multi1_t *q = p->multi + j;
        if (q->cigar) for (k = 0; k < (int)q->n_cigar; ++k) if (__cigar_op(q->cigar[k]) > 4) return 0;
    }
    for (j = 0; j < (int)p->full_len; ++j) if (p->seq[j] > 4) return 0;
    return 1;
}
static void sam_line(tbuf_t *b, const HSP *hsp, bwa_seq_t *p, int mode, int max_top2)
{
    int j, flag = p->extra_flag;
    if (p->strand) flag |= SAM_FSR;
    tb_str(b, p->name); tb_put(b, '\t'); tb_int(b, flag); tb_put(b, '\t'); tb_str(b, hsp->chrName[p->seq_id]); tb_put(b, '\t');
    tb_int(b, (int)(p->ori_pos)); tb_put(b, '\t'); tb_int(b, p->mapQ); tb_put(b, '\t');
    if (p->cigar) {
        for (j = 0; j != p->n_cigar; ++j) { tb_int(b, (int)__cigar_len(p->cigar[j])); tb_put(b, "MIDNSHP=X"[__cigar_op(p->cigar[j])]); }
    } else { tb_int(b, p->len); tb_put(b, 'M'); }
    tb_str(b, "\t*\t0\t0\t");
    tb_need(b, (size_t)p->full_len + 2);
    if (p->strand == 0) for (j = 0; j != (int)p->full_len; ++j) b->p[b->n++] = "ACGTN"[(int)p->seq[j]];
    else for (j = 0; j != (int)p->full_len; ++j) b->p[b->n++] = "TGCAN"[p->seq[p->full_len - 1 - j]];
    b->p[b->n++] = '\t';
    if (p->qual) {
        if (p->strand) seq_reverse(p->len, p->qual, 0);        /* in place, as the reference leaves it (:745-746) */
        tb_str(b, (const char *)p->qual);
    } else tb_put(b, '*');
    if (bwt_rg_id) { tb_str(b, "\tRG:Z:"); tb_str(b, bwt_rg_id); }
    if (p->bc[0]) { tb_str(b, "\tBC:Z:"); tb_str(b, p->bc); }
    if (p->clip_len < (int)p->full_len) { tb_str(b, "\tXC:i:"); tb_int(b, p->clip_len); }
    tb_str(b, "\tXT:A:"); tb_put(b, "NURMS"[p->type]);
    tb_str(b, (mode & BWA_MODE_COMPREAD) ? "\tNM:i:" : "\tCM:i:"); tb_int(b, p->nm);
    if (p->type != BWA_TYPE_MATESW) {
        tb_str(b, "\tX0:i:"); tb_int(b, (int)p->c1);
        if ((long long)p->c1 <= (long long)max_top2) { tb_str(b, "\tX1:i:"); tb_int(b, (int)p->c2); }
    }
    tb_str(b, "\tXM:i:"); tb_int(b, p->n_mm); tb_str(b, "\tXO:i:"); tb_int(b, p->n_gapo); tb_str(b, "\tXG:i:"); tb_int(b, p->n_gapo + p->n_gape);
    if (p->md) { tb_str(b, "\tMD:Z:"); tb_str(b, p->md); }
    if (p->n_multi) {
        int i, k;
        tb_str(b, "\tXA:Z:");
        for (i = 0; i < p->n_multi; ++i) {
            bwt_multi1_t *q = p->multi + i;
            tb_str(b, hsp->chrName[q->seq_id]); tb_put(b, ','); tb_put(b, q->strand ? '-' : '+'); tb_int(b, (int)(q->ori_pos));
            if (q->cigar) {
                for (k = 0; k < (int)q->n_cigar; ++k) { tb_int(b, (int)__cigar_len(q->cigar[k])); tb_put(b, "MIDNS"[__cigar_op(q->cigar[k])]); }
            } else { tb_put(b, ','); tb_int(b, q->gap + q->mm); tb_put(b, ';'); }
        }
    }
    tb_put(b, '\n');
}
enum { SAM_CHUNK = 2048 };
typedef struct { const HSP *hsp; bwa_seq_t *seqs; int mode, max_top2; tbuf_t *bufs; } fmt_ctx_t;
static void fmt_range(void *c_, int lo, int hi)
{
    fmt_ctx_t *c = (fmt_ctx_t *)c_;
    tbuf_t *b = c->bufs + lo / SAM_CHUNK;
    int i;
    for (i = lo; i < hi; ++i) if (c->seqs[i].type != BWA_TYPE_NO_MATCH) sam_line(b, c->hsp, c->seqs + i, c->mode, c->max_top2);
}
static void sam_print_batch(const HSP *hsp, int n_seqs, bwa_seq_t *seqs, int mode, int max_top2)
{
    int i, plain = 1;
    const char *e = getenv("HSA_GPU_SHIM_PRINT");               /* =ref: every line through the reference's bwa_print_sam1 */
    if (e && strcmp(e, "ref") == 0) plain = -1;
    for (i = 0; i < n_seqs && plain > 0; ++i) if (seqs[i].type != BWA_TYPE_NO_MATCH && !sam_line_is_plain(seqs + i)) plain = 0;
    if (plain > 0) {
        const int n_chunks = (n_seqs + SAM_CHUNK - 1) / SAM_CHUNK;
        fmt_ctx_t c = { hsp, seqs, mode, max_top2, (tbuf_t *)calloc(n_chunks ? n_chunks : 1, sizeof(tbuf_t)) };
        pool_run(fmt_range, &c, n_seqs, SAM_CHUNK);
        for (i = 0; i < n_chunks; ++i) { if (c.bufs[i].n) fwrite(c.bufs[i].p, 1, c.bufs[i].n, stdout); free(c.bufs[i].p); }
        free(c.bufs);
        return;
    }
    {   /* in read order on this thread: plain lines into one buffer, the others through the reference's function */
        tbuf_t b = { NULL, 0, 0 };
        for (i = 0; i < n_seqs; ++i) {                         /* bwtse.c:922-926 */
            bwa_seq_t *p = seqs + i;
            if (p->type == BWA_TYPE_NO_MATCH) continue;
            if (plain == 0 && sam_line_is_plain(p)) { sam_line(&b, hsp, p, mode, max_top2); continue; }
            if (b.n) { fwrite(b.p, 1, b.n, stdout); b.n = 0; }
            bwa_print_sam1(hsp, p, 0, mode, max_top2);
        }
        if (b.n) fwrite(b.p, 1, b.n, stdout);
        free(b.p);
    }
}

/* the print loop of generate_sam_se_core (bwtse.c:922-926) on its own: used by generate_sam_se_core_gpu below, and by the
 * harness mode `samfmt`, which checks the formatter against bwa_print_sam1 on fields computed by the reference (no GPU needed) */
void hsa_gpu_sam_print(const HSP *hsp, int n_seqs, bwa_seq_t *seqs, int mode, int max_top2)
{
    sam_print_batch(hsp, n_seqs, seqs, mode, max_top2);
}

/* the two per-read loops of the SAM stage that do not depend on the order of the reads, for the helper threads */
typedef struct {
    bwa_seq_t *seqs; uint8_t *codes; const uint64_t *off, *aoff; const int32_t *n_aln; hsa_aln1_t *aln; const hsa_sam_result_t *res;
} sam_ctx_t;
static void sam_pack_range(void *c_, int lo, int hi)
{
    sam_ctx_t *c = (sam_ctx_t *)c_;
    int i;
    for (i = lo; i < hi; ++i) {
        const bwa_seq_t *p = c->seqs + i;
        memcpy(c->codes + c->off[i], p->seq, p->len);
        if (c->n_aln[i]) memcpy(c->aln + c->aoff[i], p->aln, sizeof(bwt_aln1_t) * (size_t)c->n_aln[i]);
    }
}
static void sam_fields_range(void *c_, int lo, int hi)
{
    sam_ctx_t *c = (sam_ctx_t *)c_;
    const hsa_sam_result_t *R = c->res;
    int i; uint32_t j;
    for (i = lo; i < hi; ++i) {
        bwa_seq_t *p = c->seqs + i; const hsa_sam1_t *r = R->rec + i;
        p->type = r->type;
        if (p->multi) { free(p->multi); p->multi = NULL; }
        p->n_multi = 0;
        if (r->type == BWA_TYPE_NO_MATCH) { if (p->n_aln == 0) p->c1 = p->c2 = 0; continue; }
        p->strand = r->strand; p->n_mm = r->n_mm; p->n_gapo = r->n_gapo; p->n_gape = r->n_gape; p->score = r->score;
        p->mapQ = p->seQ = r->mapQ; p->sa = r->sa; p->seq_id = r->seq_id; p->ori_pos = r->ori_pos; p->occ_pos = r->occ_pos;
        p->c1 = r->c1; p->c2 = r->c2; p->start = r->start; p->end = r->end;
        if (r->n_cigar) { p->n_cigar = (int)r->n_cigar; p->cigar = cigar_copy(R, r->cigar_off, r->n_cigar); }
        p->nm = r->nm;
        if (r->type != BWA_TYPE_SPLICING) {                    /* bwa_cal_md1 returns strdup(str->s) (bwtse.c:493) */
            p->md = (char *)malloc(r->md_len + 1); memcpy(p->md, R->md + r->md_off, r->md_len); p->md[r->md_len] = 0;
        }
        if (r->n_multi) {
            p->n_multi = (int)r->n_multi;
            p->multi = (bwt_multi1_t *)calloc(r->n_multi, sizeof(bwt_multi1_t));
            for (j = 0; j < r->n_multi; ++j) {
                const hsa_multi1_t *q = R->multi + r->multi_off + j; bwt_multi1_t *m = p->multi + j;
                m->gap = q->gap; m->mm = q->mm; m->strand = q->strand; m->sa = q->sa; m->ori_pos = q->ori_pos; m->occ_pos = q->occ_pos;
                m->seq_id = q->seq_id; m->aln_id = q->aln_id; m->start = q->start; m->end = q->end;
                if (q->n_cigar) { m->n_cigar = q->n_cigar; m->cigar = cigar_copy(R, q->cigar_off, q->n_cigar); }
            }
        }
    }
}

void generate_sam_se_core_gpu(Idx2BWT *bi_bwt, int n_seqs, bwa_seq_t *seqs, gap_opt_t *opt, int n_occ)
{
    uint8_t *codes; uint64_t *off, *aoff, total = 0, hits = 0; uint32_t *len; int32_t *n_aln; hsa_aln1_t *aln;
    int i;
    double t_last = 0;
    if (g_timing < 0) { const char *e = getenv("HSA_GPU_SHIM_TIMING"); g_timing = e && atoi(e) ? 1 : 0; }
    if (g_timing > 0) t_last = now_s_();
    if (!g_splice_gpu) { fprintf(stderr, "[hsa_gpu] the SAM stage needs the full index (SA samples, annotation, packed text)\n"); exit(1); }
    off = (uint64_t *)malloc(sizeof(uint64_t) * (n_seqs + 1)); aoff = (uint64_t *)malloc(sizeof(uint64_t) * (n_seqs + 1));
    len = (uint32_t *)malloc(sizeof(uint32_t) * (n_seqs + 1)); n_aln = (int32_t *)malloc(sizeof(int32_t) * (n_seqs + 1));
    for (i = 0; i < n_seqs; ++i) {
        const bwa_seq_t *p = seqs + i;
        off[i] = total; len[i] = p->len; total += p->len;
        n_aln[i] = p->n_aln > 0 ? p->n_aln : 0; aoff[i] = hits; hits += (uint64_t)n_aln[i];
    }
    codes = (uint8_t *)malloc(total + 16); aln = (hsa_aln1_t *)malloc(sizeof(hsa_aln1_t) * (hits + 1));
    {
        sam_ctx_t c = { seqs, codes, off, aoff, n_aln, aln, NULL };
        pool_run(sam_pack_range, &c, n_seqs, 4096);
    }
    TICK(8);
    if (hsa_sam_se_batch(g_idx, codes, off, len, (size_t)n_seqs, n_aln, aoff, aln, (const hsa_gap_opt_t *)opt, n_occ, &g_rng48, &g_sam)) die_gpu();
    TICK(9);
    {
        sam_ctx_t c = { seqs, NULL, NULL, NULL, NULL, NULL, &g_sam };
        pool_run(sam_fields_range, &c, n_seqs, 2048);
    }
    TICK(10);
    sam_print_batch(bi_bwt->hsp, n_seqs, seqs, opt->mode, opt->max_top2);
    free(codes); free(off); free(aoff); free(len); free(n_aln); free(aln);
    TICK(11);
}

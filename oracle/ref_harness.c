/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY (oracle).  Never part of the product path.
 *
 * A small driver of OUR OWN that is linked against the UNMODIFIED reference objects compiled in place
 * from /root/reference by oracle/Makefile (`make ref`).  It exposes the reference's hot path
 *  through a file-in / file-out command line so that Python tests and bench.py's
 * cpu_baseline / --impl reference legs can (a) generate golden vectors, (b) pin oracle/hsa_oracle.c and
 * the CUDA path against the real thing, and (c) time the reference CPU path on the host cores.
 *
 * Reference entry points called (all unmodified):
 *   bwa_index_main          2BWT-Builder.c:215     index construction
 *   BWTLoad2BWT             2BWT-Interface.c:13    index load
 *   BWTAllOccValue/OccValue BWT.c:793 / BWT.c:682  rank
 *   bwt_cal_width           bwtaln.c:73            lower-bound widths
 *   bwt_match_gap           bwtgap.c:118           inexact search
 *   bwa_cal_sa_reg_gap      bwtaln.c:246           stock per-batch driver
 *   gap_init_opt/gap_init_stack/gap_destroy_stack/seq_reverse/bwa_cal_maxdiff
 *
 * File formats (little-endian uint32 words):
 *   reads  : 'HSAR' n len[n] then sum(len) bytes of base codes (A,C,G,T = 0..3, N = 4)
 *   alns   : 'HSAA' n_items then per item: n_aln, n_aln x 12 words
 *            {n_mm,n_gapo,n_gape,k,l,rev_k,rev_l,type,strand,start,end,score}
 *   widths : 'HSAW' n then per read: bid, len+1, (len+1) x {w,bid}
 *   occ in : n idx[n];  occ out: n x 16 words {fwd occ4[4], rev occ4[4], fwd occ1[c=0..3], rev occ1[c=0..3]}
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <time.h>
#include <unistd.h>
#include <sys/wait.h>
#include "bwtaln.h"
#include "bwtgap.h"
#include "BWT.h"

int bwa_index_main(int argc, char **argv);
#ifdef HSA_WITH_GPU_SHIM
/* shim/hsa_gpu_shim.c: the reference-side binding of libhsa_b200.so (same signature as bwa_cal_sa_reg_gap) */
int hsa_gpu_open(const Idx2BWT *bi, int device);
void hsa_gpu_close(void);
void hsa_gpu_sa_values(const Idx2BWT *bi, const unsigned int *sa_index, size_t n, unsigned int *occ_pos);
void bwa_cal_sa_reg_gap_gpu(int tid, const Idx2BWT *bwt, int n_seqs, bwa_seq_t *seqs, const gap_opt_t *opt, bwt_array_t *arr);
bwt_aln1_t *bwt_match_gap_gpu(bwt_aux_t *aux, int *_n_aln);
void generate_sam_se_core_gpu(Idx2BWT *bi_bwt, int n_seqs, bwa_seq_t *seqs, gap_opt_t *opt, int n_occ);
#endif
typedef bwt_aln1_t *(*match_fn)(bwt_aux_t *, int *);
static match_fn g_match = bwt_match_gap;            /* gpupercall: shim/hsa_gpu_shim.c's bwt_match_gap_gpu */
static unsigned long long g_call_mismatch = 0, g_width_mismatch = 0;
typedef void (*driver_fn)(int, const Idx2BWT *, int, bwa_seq_t *, const gap_opt_t *, bwt_array_t *);
static driver_fn g_driver = bwa_cal_sa_reg_gap;
void generate_sam_se_core(Idx2BWT *bi_bwt, int n_seqs, bwa_seq_t *seqs, gap_opt_t *opt, int n_occ);
typedef void (*sam_fn)(Idx2BWT *, int, bwa_seq_t *, gap_opt_t *, int);
static sam_fn g_sam = generate_sam_se_core;            /* gpusam: shim/hsa_gpu_shim.c's generate_sam_se_core_gpu */

#ifdef HSA_COUNT_OCC
static unsigned long long g_occ4 = 0, g_occ1 = 0;
void __real_BWTAllOccValue(const BWT *bwt, unsigned int index, unsigned int *occValue);
unsigned int __real_BWTOccValue(const BWT *bwt, unsigned int index, const unsigned int character);
void __wrap_BWTAllOccValue(const BWT *bwt, unsigned int index, unsigned int *occValue)
{ ++g_occ4; __real_BWTAllOccValue(bwt, index, occValue); }
unsigned int __wrap_BWTOccValue(const BWT *bwt, unsigned int index, const unsigned int character)
{ ++g_occ1; return __real_BWTOccValue(bwt, index, character); }
#else
static unsigned long long g_occ4 = 0, g_occ1 = 0;
#endif

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void die(const char *msg) { fprintf(stderr, "ref_harness: %s\n", msg); exit(2); }

/* bwt_match_gap through g_match.  With the per-call GPU symbol installed (gpupercall / gpuseeds) the reference runs
 * first on a copy of the frame's width_back; hits AND the in-place rewrite of width_back by gap_shadow (bwtgap.c:217)
 * must agree between the two. */
static bwt_aln1_t *match_checked(bwt_aux_t *aux, int *n_aln)
{
    bwt_width_t *wb_ref, *wb_gpu = aux->width_back, *ws_keep = aux->width_seed;
    int n_ref = 0, j, len = aux->len; bwt_aln1_t *a_ref, *aln;
    if (g_match == bwt_match_gap) return bwt_match_gap(aux, n_aln);
    wb_ref = (bwt_width_t*)malloc((len + 1) * sizeof(bwt_width_t));
    memcpy(wb_ref, wb_gpu, (len + 1) * sizeof(bwt_width_t));
    aux->width_back = wb_ref;
    if (ws_keep == wb_gpu) aux->width_seed = wb_ref;           /* keep the aliasing of bwtgap.c:809 */
    a_ref = bwt_match_gap(aux, &n_ref);
    aux->width_back = wb_gpu; aux->width_seed = ws_keep;
    aln = g_match(aux, n_aln);
    if (n_ref != *n_aln || (*n_aln && memcmp(a_ref, aln, *n_aln * sizeof(bwt_aln1_t)) != 0)) ++g_call_mismatch;
    for (j = 0; j <= len; ++j) if (wb_ref[j].w != wb_gpu[j].w || wb_ref[j].bid != wb_gpu[j].bid) { ++g_width_mismatch; break; }
    free(a_ref); free(wb_ref);
    return aln;
}

/* ---------------------------------------------------------------- reads */
typedef struct { uint32_t n; uint32_t *len; uint64_t *off; ubyte_t *codes; } reads_t;

static reads_t load_reads(const char *fn)
{
    reads_t r; uint32_t hdr[2]; uint64_t tot = 0; uint32_t i;
    FILE *f = fopen(fn, "rb");
    if (!f) die("cannot open reads file");
    if (fread(hdr, 4, 2, f) != 2 || hdr[0] != 0x52415348u) die("bad reads magic");
    r.n = hdr[1];
    r.len = (uint32_t*)malloc(4 * (size_t)(r.n + 1));
    r.off = (uint64_t*)malloc(8 * (size_t)(r.n + 1));
    if (fread(r.len, 4, r.n, f) != r.n) die("short reads file (len)");
    for (i = 0; i < r.n; ++i) { r.off[i] = tot; tot += r.len[i]; }
    r.off[r.n] = tot;
    r.codes = (ubyte_t*)malloc(tot + 64);
    if (fread(r.codes, 1, tot, f) != tot) die("short reads file (codes)");
    fclose(f);
    return r;
}

/* ---------------------------------------------------------------- options */
static void apply_opt(gap_opt_t *o, const char *kv)
{
    const char *eq = strchr(kv, '=');
    char key[64]; size_t kl;
    if (!eq) die("option must be key=value");
    kl = (size_t)(eq - kv); if (kl >= sizeof(key)) die("option key too long");
    memcpy(key, kv, kl); key[kl] = 0;
#define OPT_I(name) if (strcmp(key, #name) == 0) { o->name = atoi(eq + 1); return; }
    OPT_I(s_mm) OPT_I(s_gapo) OPT_I(s_gape) OPT_I(mode) OPT_I(indel_end_skip) OPT_I(max_del_occ)
    OPT_I(max_entries) OPT_I(max_diff) OPT_I(max_gapo) OPT_I(max_gape) OPT_I(max_seed_diff)
    OPT_I(seed_len) OPT_I(max_top2)
#undef OPT_I
    if (strcmp(key, "fnr") == 0) { o->fnr = (float)atof(eq + 1); return; }
    fprintf(stderr, "ref_harness: unknown option %s\n", key); exit(2);
}

typedef struct { int batch; int nout; int procs; int clear_gape; int qual; } hopt_t;

static gap_opt_t *parse_opts(int argc, char **argv, int first, hopt_t *h)
{
    gap_opt_t *o = gap_init_opt();
    int i;
    h->batch = 0x186A0; h->nout = 0; h->procs = 1; h->clear_gape = 1; h->qual = 0;
    for (i = first; i < argc; ++i) {
        if (strncmp(argv[i], "batch=", 6) == 0) h->batch = atoi(argv[i] + 6);
        else if (strncmp(argv[i], "nout=", 5) == 0) h->nout = atoi(argv[i] + 5);
        else if (strncmp(argv[i], "procs=", 6) == 0) h->procs = atoi(argv[i] + 6);
        else if (strncmp(argv[i], "clear_gape=", 11) == 0) h->clear_gape = atoi(argv[i] + 11);
        else if (strncmp(argv[i], "qual=", 5) == 0) h->qual = atoi(argv[i] + 5);
        else apply_opt(o, argv[i]);
    }
    return o;
}

/* ---------------------------------------------------------------- output */
static void put_aln(FILE *f, int n_aln, const bwt_aln1_t *aln)
{
    uint32_t w[12]; int j; uint32_t n = (uint32_t)n_aln;
    fwrite(&n, 4, 1, f);
    for (j = 0; j < n_aln; ++j) {
        const bwt_aln1_t *p = aln + j;
        w[0] = p->n_mm; w[1] = p->n_gapo; w[2] = p->n_gape; w[3] = p->k; w[4] = p->l;
        w[5] = p->rev_k; w[6] = p->rev_l; w[7] = p->type; w[8] = p->strand;
        w[9] = (uint32_t)p->start; w[10] = (uint32_t)p->end; w[11] = (uint32_t)p->score;
        fwrite(w, 4, 12, f);
    }
}

static Idx2BWT *load_index(const char *prefix)
{
    char *str = (char*)calloc(strlen(prefix) + 32, 1);
    Idx2BWT *bi;
    FILE *t;
    strcpy(str, prefix); strcat(str, ".index.pac");  /* a full index (packed DNA, annotation, SA) as `index` writes it */
    t = fopen(str, "rb");
    strcpy(str, prefix); strcat(str, ".index");      /* bwtaln.c:463-469 */
    if (t) {
        fclose(t);
        bi = BWTLoad2BWT(str, ".sa");
    } else {
        /* search arrays only (.bwt/.fmv/.rev.bwt/.rev.fmv), e.g. an index written by the product's own
         * builder: the same BWTLoad calls BWTLoad2BWT makes (2BWT-Interface.c:51-52), no SA / packed DNA.
         * Enough for occ / width / percall / whole / seeds / sa, which never touch hsp. */
        char a[1100], b[1100];
        MMPool *mmPool;
        bi = (Idx2BWT*)calloc(1, sizeof(Idx2BWT));
        MMMasterInitialize(3, 0, FALSE, NULL);
        mmPool = MMPoolCreate(2097152);
        snprintf(a, sizeof(a), "%s.bwt", str); snprintf(b, sizeof(b), "%s.fmv", str);
        {   /* the SA samples too when they are there (mode sa): the same BWTLoad argument BWTLoad2BWT passes */
            char c[1100]; FILE *ts;
            snprintf(c, sizeof(c), "%s.sa", str);
            ts = fopen(c, "rb");
            if (ts) fclose(ts);
            bi->bwt = BWTLoad(mmPool, a, b, ts ? c : NULL, NULL, NULL, NULL);
        }
        snprintf(a, sizeof(a), "%s.rev.bwt", str); snprintf(b, sizeof(b), "%s.rev.fmv", str);
        bi->rev_bwt = BWTLoad(mmPool, a, b, NULL, NULL, NULL, NULL);
        bi->hsp = NULL; bi->mmPool = mmPool;
    }
    free(str);
    return bi;
}

/* per-read option resolution exactly as the driver does it for a read that has NOT yet seen the
 * local_opt switch (bwtaln.c:260-261, 330-332), applied to a private copy so nothing leaks. */
static void resolve_read_opt(gap_opt_t *dst, const gap_opt_t *src, int len, int clear_gape)
{
    *dst = *src;
    if (clear_gape) dst->mode &= ~BWA_MODE_GAPE;
    if (src->fnr > 0.0) dst->max_diff = bwa_cal_maxdiff(len, BWA_AVG_ERR, src->fnr);
    dst->seed_len = src->seed_len < len ? src->seed_len : 0x7fffffff;
}

/* ---------------------------------------------------------------- forked ranges
 * The reference has no working threading (bwtaln.c:307-311, 481-504 are commented out), so a read set is cut into
 * `procs` contiguous shards, one forked process each.  Every worker reports {its own seconds, two counters, occ4,
 * occ1}; with an output path it also writes its items to <out>.part<p>, which the parent concatenates behind the
 * header in shard order -- so a multi-process run produces the same dump as a single-process one. */
typedef struct { double secs; unsigned long long a, b, occ4, occ1; } range_res_t;
typedef void (*range_fn)(void *ctx, uint32_t lo, uint32_t hi, FILE *fo, range_res_t *res);

static void run_ranges(range_fn fn, void *ctx, uint32_t n, int procs, const char *out_path, uint32_t items_per_read,
                       range_res_t *tot)
{
    uint32_t hdr[2];
    memset(tot, 0, sizeof(*tot));
    if (procs <= 1) {
        FILE *fo = NULL;
        if (out_path) { fo = fopen(out_path, "wb"); if (!fo) die("cannot open output"); hdr[0] = 0x41415348u; hdr[1] = items_per_read * n; fwrite(hdr, 4, 2, fo); }
        g_occ4 = g_occ1 = 0;
        fn(ctx, 0, n, fo, tot);
        tot->occ4 = g_occ4; tot->occ1 = g_occ1;
        if (fo) fclose(fo);
        return;
    }
    {
        int p, (*fds)[2] = (int (*)[2])calloc((size_t)procs, sizeof(int[2]));
        if (procs > 1024) die("procs too large");
        fflush(stdout);
        for (p = 0; p < procs; ++p) {
            pid_t pid;
            if (pipe(fds[p]) != 0) die("pipe");
            pid = fork();
            if (pid < 0) die("fork");
            if (pid == 0) {
                uint32_t lo = (uint32_t)((uint64_t)n * p / procs), hi = (uint32_t)((uint64_t)n * (p + 1) / procs);
                range_res_t r; FILE *fo = NULL;
                memset(&r, 0, sizeof(r));
                if (out_path) {
                    char *pp = (char*)malloc(strlen(out_path) + 32);
                    sprintf(pp, "%s.part%d", out_path, p);
                    fo = fopen(pp, "wb"); if (!fo) _exit(4);
                }
                g_occ4 = g_occ1 = 0;
                fn(ctx, lo, hi, fo, &r);
                r.occ4 = g_occ4; r.occ1 = g_occ1;
                if (fo) fclose(fo);
                if (write(fds[p][1], &r, sizeof(r)) != (ssize_t)sizeof(r)) _exit(3);
                _exit(0);
            }
            close(fds[p][1]);
        }
        for (p = 0; p < procs; ++p) {
            range_res_t r;
            memset(&r, 0, sizeof(r));
            if (read(fds[p][0], &r, sizeof(r)) != (ssize_t)sizeof(r)) die("worker failed");
            if (r.secs > tot->secs) tot->secs = r.secs;
            tot->a += r.a; tot->b += r.b; tot->occ4 += r.occ4; tot->occ1 += r.occ1;
            close(fds[p][0]);
        }
        while (wait(NULL) > 0) {}
        free(fds);
        if (out_path) {
            FILE *fo = fopen(out_path, "wb"); char *pp = (char*)malloc(strlen(out_path) + 32), *buf = (char*)malloc(1 << 20);
            if (!fo) die("cannot open output");
            hdr[0] = 0x41415348u; hdr[1] = items_per_read * n; fwrite(hdr, 4, 2, fo);
            for (p = 0; p < procs; ++p) {
                FILE *fi; size_t got;
                sprintf(pp, "%s.part%d", out_path, p);
                fi = fopen(pp, "rb"); if (!fi) die("missing part file");
                while ((got = fread(buf, 1, 1 << 20, fi)) > 0) fwrite(buf, 1, got, fo);
                fclose(fi); remove(pp);
            }
            fclose(fo); free(pp); free(buf);
        }
    }
}

/* ---------------------------------------------------------------- modes */
static int mode_occ(int argc, char **argv)
{
    Idx2BWT *bi; FILE *fi, *fo; uint32_t n, i, c, *idx;
    if (argc < 5) die("usage: occ <prefix> <idx.bin> <out.bin>");
    bi = load_index(argv[2]);
    fi = fopen(argv[3], "rb"); if (!fi) die("cannot open idx");
    if (fread(&n, 4, 1, fi) != 1) die("short idx");
    idx = (uint32_t*)malloc(4 * (size_t)n);
    if (fread(idx, 4, n, fi) != n) die("short idx");
    fclose(fi);
    fo = fopen(argv[4], "wb");
    fwrite(&n, 4, 1, fo);
    for (i = 0; i < n; ++i) {
        unsigned int __attribute__((aligned(16))) o[16];
        BWTAllOccValue(bi->bwt, idx[i], o);
        BWTAllOccValue(bi->rev_bwt, idx[i], o + 4);
        for (c = 0; c < 4; ++c) {
            o[8 + c] = BWTOccValue(bi->bwt, idx[i], c);
            o[12 + c] = BWTOccValue(bi->rev_bwt, idx[i], c);
        }
        fwrite(o, 4, 16, fo);
    }
    fclose(fo);
    printf("{\"mode\":\"occ\",\"n\":%u,\"textLength\":%u,\"inverseSa0\":%u,\"rev_inverseSa0\":%u}\n",
           n, bi->bwt->textLength, bi->bwt->inverseSa0, bi->rev_bwt->inverseSa0);
    return 0;
}

static int mode_width(int argc, char **argv)
{
    Idx2BWT *bi; reads_t r; FILE *fo; uint32_t i, hdr[2]; int type;
    if (argc < 6) die("usage: width <prefix> <reads> <type> <out>");
    bi = load_index(argv[2]); r = load_reads(argv[3]); type = atoi(argv[4]);
    fo = fopen(argv[5], "wb");
    hdr[0] = 0x57415348u; hdr[1] = r.n; fwrite(hdr, 4, 2, fo);
    for (i = 0; i < r.n; ++i) {
        int len = (int)r.len[i];
        bwt_width_t *w = (bwt_width_t*)calloc(len + 1, sizeof(bwt_width_t));
        uint32_t h2[2];
        h2[0] = (uint32_t)bwt_cal_width(bi, len, r.codes + r.off[i], w, type);
        h2[1] = (uint32_t)(len + 1);
        fwrite(h2, 4, 2, fo);
        fwrite(w, sizeof(bwt_width_t), len + 1, fo);
        free(w);
    }
    fclose(fo);
    printf("{\"mode\":\"width\",\"n\":%u}\n", r.n);
    return 0;
}

/* one bwt_match_gap call per (read, strand), strand 1 (revcomp) then 0, BOTH always searched, each with
 * a fresh option copy; widths computed as bwtaln.c:344-348 does. */
static int mode_percall(int argc, char **argv)
{
    Idx2BWT *bi; reads_t r; FILE *fo = NULL; hopt_t h; gap_opt_t *opt, ropt; uint32_t i, hdr[2];
    int max_len = 0; bwt_aux_t aux; double t0, t1; unsigned long long n_hits = 0;
    if (argc < 5) die("usage: percall <prefix> <reads> <out> [opts]");
    bi = load_index(argv[2]); r = load_reads(argv[3]);
    opt = parse_opts(argc, argv, 5, &h);
    for (i = 0; i < r.n; ++i) if ((int)r.len[i] > max_len) max_len = (int)r.len[i];
    if (!h.nout) { fo = fopen(argv[4], "wb"); hdr[0] = 0x41415348u; hdr[1] = 2 * r.n; fwrite(hdr, 4, 2, fo); }
    memset(&aux, 0, sizeof(aux));
    aux.bi_bwt = bi; aux.max_len = max_len;
    aux.width_back = (bwt_width_t*)calloc(max_len + 1, sizeof(bwt_width_t));
    aux.width_seed = (bwt_width_t*)calloc(max_len + 1, sizeof(bwt_width_t));
    aux.rc_seq = (ubyte_t*)calloc(max_len + 1, 1);
    g_occ4 = g_occ1 = 0;
    t0 = now_s();
    for (i = 0; i < r.n; ++i) {
        int len = (int)r.len[i], s;
        ubyte_t *seq = r.codes + r.off[i];
        memcpy(aux.rc_seq, seq, len); seq_reverse(len, aux.rc_seq, 1);
        aux.seq = seq; aux.len = len;
        for (s = 1; s >= 0; --s) {
            int n_aln = 0; bwt_aln1_t *aln; bwt_width_t *wseed = NULL;
            resolve_read_opt(&ropt, opt, len, h.clear_gape);
            aux.opt = &ropt;
            aux.stack = gap_init_stack(ropt.max_diff, ropt.max_gapo, ropt.max_gape, &ropt);
            memset(aux.width_back, 0, (max_len + 1) * sizeof(bwt_width_t));
            memset(aux.width_seed, 0, (max_len + 1) * sizeof(bwt_width_t));
            if (len > ropt.seed_len) {
                bwt_cal_width(bi, ropt.seed_len, (s == 0 ? seq : aux.rc_seq) + (len - ropt.seed_len), aux.width_seed, 1);
                wseed = aux.width_seed;
            }
            /* NOTE: the stock driver always passes a non-NULL width_seed (bwtaln.c:285); for len > seed_len
             * (the only safe case, SURVEY.md 8c hazard 3) that is what is passed here. For len <= seed_len the
             * harness passes NULL (no seeding) instead of reading out of bounds. */
            bwt_cal_width(bi, len, s == 0 ? seq : aux.rc_seq, aux.width_back, 1);
            aux.strand = s;
            {
                bwt_width_t *keep = aux.width_seed;
                aux.width_seed = wseed;
                aln = match_checked(&aux, &n_aln);
                aux.width_seed = keep;
            }
            n_hits += (n_aln != 0);
            if (fo) put_aln(fo, n_aln, aln);
            free(aln);
            gap_destroy_stack(aux.stack);
        }
    }
    t1 = now_s();
    if (fo) fclose(fo);
    printf("{\"mode\":\"percall\",\"reads\":%u,\"calls\":%u,\"calls_with_hits\":%llu,\"secs\":%.6f,\"occ4\":%llu,\"occ1\":%llu,"
           "\"call_mismatches\":%llu,\"width_mismatches\":%llu}\n",
           r.n, 2 * r.n, n_hits, t1 - t0, g_occ4, g_occ1, g_call_mismatch, g_width_mismatch);
    return 0;
}

/* the six seed calls of bwt_splice_match (bwtgap.c:797-820), WITHOUT its early-outs, through the real
 * bwt_cal_width / bwt_match_gap: gaps off, max_diff = max_seed_diff, seed_len = len_align,
 * width computed on the read PREFIX (bwtgap.c:807-808) and aliased as width_back (bwtgap.c:809). */
typedef struct { Idx2BWT *bi; reads_t *r; const gap_opt_t *opt; const hopt_t *h; } range_ctx_t;

static void seeds_range(void *vctx, uint32_t lo, uint32_t hi, FILE *fo, range_res_t *res)
{
    range_ctx_t *c = (range_ctx_t*)vctx;
    Idx2BWT *bi = c->bi; reads_t *r = c->r; const gap_opt_t *opt = c->opt; gap_opt_t sopt;
    int max_len = 0; bwt_aux_t aux; double t0; uint32_t i;
    for (i = lo; i < hi; ++i) if ((int)r->len[i] > max_len) max_len = (int)r->len[i];
    memset(&aux, 0, sizeof(aux));
    aux.bi_bwt = bi; aux.max_len = max_len;
    aux.width_seed = (bwt_width_t*)calloc(max_len + 1, sizeof(bwt_width_t));
    t0 = now_s();
    for (i = lo; i < hi; ++i) {
        int len = (int)r->len[i], s, seed_len = len / 3;
        ubyte_t *seq = r->codes + r->off[i];
        ubyte_t *rc = (ubyte_t*)calloc(max_len + 1, 1);
        memcpy(rc, seq, len); seq_reverse(len, rc, 1);
        for (s = 0; s < 6; ++s) {
            int n_aln = 0, len_align = seed_len + (s % 3 == 2 ? len % 3 : 0);
            bwt_aln1_t *aln;
            sopt = *opt;                                   /* bwtgap.c:769-774 */
            sopt.mode &= ~BWA_MODE_GAPE; sopt.max_gapo = 0; sopt.max_gape = 0;
            sopt.max_diff = opt->max_seed_diff;
            sopt.seed_len = len_align;                     /* bwtgap.c:802 */
            aux.opt = &sopt;
            /* bucket count only has to cover every reachable score; results do not depend on it */
            aux.stack = gap_init_stack(sopt.max_diff + 8, 4, 12, opt);
            aux.strand = s / 3; aux.len = len_align;
            aux.seq = seq; aux.rc_seq = rc;
            if (s < 3) aux.seq = seq + (s % 3) * seed_len; else aux.rc_seq = rc + (s % 3) * seed_len;
            memset(aux.width_seed, 0, sizeof(bwt_width_t) * (max_len + 1));
            bwt_cal_width(bi, len_align, aux.strand == 0 ? seq : rc, aux.width_seed, 1);
            aux.width_back = aux.width_seed;
            aln = match_checked(&aux, &n_aln);
            { int j; for (j = 0; j < n_aln; ++j) { aln[j].start = (s % 3) * seed_len; aln[j].end = aln[j].start + len_align - 1; } } /* bwtgap.c:816-819 */
            res->a += (n_aln != 0);
            if (fo) put_aln(fo, n_aln, aln);
            free(aln);
            gap_destroy_stack(aux.stack);
        }
        free(rc);
    }
    res->secs = now_s() - t0;
    free(aux.width_seed);
}

static int mode_seeds(int argc, char **argv)
{
    Idx2BWT *bi; reads_t r; hopt_t h; gap_opt_t *opt; range_ctx_t ctx; range_res_t tot;
    if (argc < 5) die("usage: seeds <prefix> <reads> <out> [opts] [procs=P] [nout=1]");
    bi = load_index(argv[2]); r = load_reads(argv[3]);
    opt = parse_opts(argc, argv, 5, &h);
    ctx.bi = bi; ctx.r = &r; ctx.opt = opt; ctx.h = &h;
    run_ranges(seeds_range, &ctx, r.n, h.procs, h.nout ? NULL : argv[4], 6, &tot);
    printf("{\"mode\":\"seeds\",\"reads\":%u,\"calls\":%u,\"procs\":%d,\"calls_with_hits\":%llu,\"secs\":%.6f,\"occ4\":%llu,\"occ1\":%llu,"
           "\"call_mismatches\":%llu,\"width_mismatches\":%llu}\n",
           r.n, 6 * r.n, h.procs, tot.a, tot.secs, tot.occ4, tot.occ1, g_call_mismatch, g_width_mismatch);
    return 0;
}

/* the stock batch driver bwa_cal_sa_reg_gap (bwtaln.c:246) over <batch>-read batches, option leak and
 * splice fallback included.  With procs=P the read set is cut into P contiguous shards, one forked
 * process each (the reference has no working threading: bwtaln.c:307-311, 481-504 are commented out);
 * timing covers only the driver calls, max over processes. */
static double run_driver_range(Idx2BWT *bi, reads_t *r, uint32_t lo, uint32_t hi, const gap_opt_t *opt0,
                               const hopt_t *h, FILE *fo, unsigned long long *n_whole, unsigned long long *n_any)
{
    gap_opt_t *opt = (gap_opt_t*)calloc(2, sizeof(gap_opt_t));
    bwt_array_t *arr = bwt_array_init();
    double secs = 0; uint32_t b;
    *opt = *opt0;
    for (b = lo; b < hi; b += (uint32_t)h->batch) {
        uint32_t e = b + (uint32_t)h->batch < hi ? b + (uint32_t)h->batch : hi, i;
        int n = (int)(e - b);
        bwa_seq_t *seqs = (bwa_seq_t*)calloc(n, sizeof(bwa_seq_t));
        double t0;
        for (i = b; i < e; ++i) { seqs[i - b].seq = r->codes + r->off[i]; seqs[i - b].len = r->len[i]; }
        t0 = now_s();
        g_driver(0, bi, n, seqs, opt, arr);
        secs += now_s() - t0;
        for (i = 0; i < (uint32_t)n; ++i) {
            bwa_seq_t *p = seqs + i;
            if (p->n_aln) { ++*n_any; if (p->aln[0].type != BWA_TYPE_SPLICING && p->aln[0].start == 0 && p->aln[0].end == (int)p->len - 1) ++*n_whole; }
            if (fo) put_aln(fo, p->n_aln, p->aln);
            free(p->aln);
        }
        free(seqs);
    }
    return secs;
}

static int mode_driver(int argc, char **argv)
{
    Idx2BWT *bi; reads_t r; FILE *fo = NULL; hopt_t h; gap_opt_t *opt; uint32_t hdr[2];
    unsigned long long n_whole = 0, n_any = 0; double secs;
    if (argc < 5) die("usage: driver <prefix> <reads> <out> [opts] [batch=N] [procs=P] [nout=1]");
    bi = load_index(argv[2]); r = load_reads(argv[3]);
    opt = parse_opts(argc, argv, 5, &h);
    g_occ4 = g_occ1 = 0;
    if (h.procs <= 1) {
        if (!h.nout) { fo = fopen(argv[4], "wb"); hdr[0] = 0x41415348u; hdr[1] = r.n; fwrite(hdr, 4, 2, fo); }
        secs = run_driver_range(bi, &r, 0, r.n, opt, &h, fo, &n_whole, &n_any);
        if (fo) fclose(fo);
    } else {
        /* timing only: fork P workers over contiguous shards, each reports its own driver seconds */
        int p, fds[256][2]; double mx = 0;
        if (h.procs > 256) die("procs too large");
        for (p = 0; p < h.procs; ++p) {
            pid_t pid;
            if (pipe(fds[p]) != 0) die("pipe");
            pid = fork();
            if (pid < 0) die("fork");
            if (pid == 0) {
                uint32_t lo = (uint32_t)((uint64_t)r.n * p / h.procs), hi = (uint32_t)((uint64_t)r.n * (p + 1) / h.procs);
                double res[3]; unsigned long long w = 0, a = 0;
                res[0] = run_driver_range(bi, &r, lo, hi, opt, &h, NULL, &w, &a);
                res[1] = (double)w; res[2] = (double)a;
                if (write(fds[p][1], res, sizeof(res)) != (ssize_t)sizeof(res)) _exit(3);
                _exit(0);
            }
            close(fds[p][1]);
        }
        for (p = 0; p < h.procs; ++p) {
            double res[3] = {0, 0, 0};
            if (read(fds[p][0], res, sizeof(res)) != (ssize_t)sizeof(res)) die("worker failed");
            if (res[0] > mx) mx = res[0];
            n_whole += (unsigned long long)res[1]; n_any += (unsigned long long)res[2];
            close(fds[p][0]);
        }
        while (wait(NULL) > 0) {}
        secs = mx;
    }
    printf("{\"mode\":\"driver\",\"reads\":%u,\"procs\":%d,\"batch\":%d,\"aligned_any\":%llu,\"aligned_whole\":%llu,\"secs\":%.6f,\"occ4\":%llu,\"occ1\":%llu}\n",
           r.n, h.procs, h.batch, n_any, n_whole, secs, g_occ4, g_occ1);
    return 0;
}


/* The whole-read part of bwa_cal_sa_reg_gap (bwtaln.c:303-360) WITHOUT the splice fallback and WITHOUT
 * the option leak: per-read filters (:314-317, :324-325), revcomp strand first, forward strand only when
 * the revcomp search found nothing (:343-359), a private option copy per read.  The loop body calls the
 * real bwt_cal_width / bwt_match_gap.  This is exactly the work the CUDA batch entry point does, so it is
 * the like-for-like CPU baseline; procs=P forks P workers over contiguous shards (timing only). */
static double run_whole_range(Idx2BWT *bi, reads_t *r, uint32_t lo, uint32_t hi, const gap_opt_t *opt,
                              const hopt_t *h, FILE *fo, unsigned long long *n_hit)
{
    int max_len = 0, local_max_diff; uint32_t i; bwt_aux_t aux; gap_opt_t ropt; double t0, secs;
    for (i = lo; i < hi; ++i) if ((int)r->len[i] > max_len) max_len = (int)r->len[i];
    local_max_diff = opt->fnr > 0.0 ? bwa_cal_maxdiff(max_len, BWA_AVG_ERR, opt->fnr) : opt->max_diff; /* :273-274 */
    memset(&aux, 0, sizeof(aux));
    aux.bi_bwt = bi; aux.max_len = max_len;
    aux.width_back = (bwt_width_t*)calloc(max_len + 1, sizeof(bwt_width_t));
    aux.width_seed = (bwt_width_t*)calloc(max_len + 1, sizeof(bwt_width_t));
    aux.rc_seq = (ubyte_t*)calloc(max_len + 1, 1);
    aux.stack = gap_init_stack(local_max_diff + 2, opt->max_gapo + 1, opt->max_gape + 1, opt);
    t0 = now_s();
    for (i = lo; i < hi; ++i) {
        int len = (int)r->len[i], s, j, nn = 0, n_aln = 0; bwt_aln1_t *aln = NULL;
        ubyte_t *seq = r->codes + r->off[i];
        int polya = len >= 15, polyt = len >= 15;
        for (j = 0; j < len; ++j) nn += seq[j] > 3;
        for (j = 0; j < 15 && j < len; ++j) { polya &= seq[j] == 0; polyt &= seq[j] == 3; }
        if (nn > local_max_diff || polya || polyt) { if (fo) put_aln(fo, 0, NULL); continue; }
        memcpy(aux.rc_seq, seq, len); seq_reverse(len, aux.rc_seq, 1);
        aux.seq = seq; aux.len = len;
        resolve_read_opt(&ropt, opt, len, h->clear_gape);
        aux.opt = &ropt;
        for (s = 1; s >= 0; --s) {
            bwt_width_t *keep = aux.width_seed;
            if (len > ropt.seed_len)
                bwt_cal_width(bi, ropt.seed_len, (s == 0 ? seq : aux.rc_seq) + (len - ropt.seed_len), aux.width_seed, 1);
            else aux.width_seed = NULL;
            bwt_cal_width(bi, len, s == 0 ? seq : aux.rc_seq, aux.width_back, 1);
            aux.strand = s;
            aln = bwt_match_gap(&aux, &n_aln);
            aux.width_seed = keep;
            if (n_aln) { for (j = 0; j < n_aln; ++j) aln[j].strand = s; break; }
            free(aln); aln = NULL;
        }
        if (n_aln) { aln[0].start = 0; aln[0].end = len - 1; ++*n_hit; }   /* :371-372 */
        if (fo) put_aln(fo, n_aln, aln);
        free(aln);
    }
    secs = now_s() - t0;
    gap_destroy_stack(aux.stack);
    return secs;
}

static void whole_range(void *vctx, uint32_t lo, uint32_t hi, FILE *fo, range_res_t *res)
{
    range_ctx_t *c = (range_ctx_t*)vctx;
    unsigned long long n_hit = 0;
    res->secs = run_whole_range(c->bi, c->r, lo, hi, c->opt, c->h, fo, &n_hit);
    res->a = n_hit;
}

static int mode_whole(int argc, char **argv)
{
    Idx2BWT *bi; reads_t r; hopt_t h; gap_opt_t *opt; range_ctx_t ctx; range_res_t tot;
    if (argc < 5) die("usage: whole <prefix> <reads> <out> [opts] [procs=P] [nout=1]");
    bi = load_index(argv[2]); r = load_reads(argv[3]);
    opt = parse_opts(argc, argv, 5, &h);
    ctx.bi = bi; ctx.r = &r; ctx.opt = opt; ctx.h = &h;
    run_ranges(whole_range, &ctx, r.n, h.procs, h.nout ? NULL : argv[4], 1, &tot);
    printf("{\"mode\":\"whole\",\"reads\":%u,\"procs\":%d,\"aligned\":%llu,\"secs\":%.6f,\"occ4\":%llu,\"occ1\":%llu}\n",
           r.n, h.procs, tot.a, tot.secs, tot.occ4, tot.occ1);
    return 0;
}

/* splice <prefix> <reads> <out> [opts] [procs=P]: bwt_splice_match (bwtgap.c:748) for EVERY read, with the frame the
 * driver builds for a read that found nothing on either strand (bwtaln.c:279-288, 362-369): aux->opt = the options the
 * driver would hold in local_opt at that point -- the caller's values with max_diff / seed_len resolved per read as
 * bwtaln.c:330-332 writes them through aux->opt, MODE_GAPE kept (clear_gape=0, the first batch of a process) or cleared.
 * Needs a full index (SA samples, packed DNA, annotation).  Output: per read n_aln (0..2) and its bwt_aln1_t entries. */
static void splice_range(void *vctx, uint32_t lo, uint32_t hi, FILE *fo, range_res_t *res)
{
    range_ctx_t *c = (range_ctx_t*)vctx;
    Idx2BWT *bi = c->bi; reads_t *r = c->r; const gap_opt_t *opt = c->opt; gap_opt_t ropt;
    int max_len = 0; uint32_t i; bwt_aux_t aux; double t0; bwt_array_t *arr = bwt_array_init();
    for (i = lo; i < hi; ++i) if ((int)r->len[i] > max_len) max_len = (int)r->len[i];
    memset(&aux, 0, sizeof(aux));
    aux.bi_bwt = bi; aux.arr = arr; aux.max_len = max_len;
    aux.width_back = (bwt_width_t*)calloc(max_len + 1, sizeof(bwt_width_t));
    aux.width_fore = (bwt_width_t*)calloc(max_len + 1, sizeof(bwt_width_t));
    aux.width_seed = (bwt_width_t*)calloc(max_len + 1, sizeof(bwt_width_t));
    aux.rc_seq = (ubyte_t*)calloc(max_len + 1, 1);
    {
        int md = opt->fnr > 0.0 ? bwa_cal_maxdiff(max_len, BWA_AVG_ERR, opt->fnr) : opt->max_diff;
        aux.stack = gap_init_stack(md + 2, opt->max_gapo + 2, opt->max_gape + 2, opt);    /* covers every reachable score */
    }
    t0 = now_s();
    for (i = lo; i < hi; ++i) {
        int len = (int)r->len[i], n_aln = 0; bwt_aln1_t *aln;
        ubyte_t *seq = r->codes + r->off[i];
        resolve_read_opt(&ropt, opt, len, c->h->clear_gape);
        memset(aux.rc_seq, 0, max_len);                          /* bwtaln.c:334-337 */
        memcpy(aux.rc_seq, seq, len); seq_reverse(len, aux.rc_seq, 1);
        aux.seq = seq; aux.len = len; aux.strand = 0; aux.opt = &ropt;
        aln = bwt_splice_match(&aux, &n_aln);
        res->a += (n_aln != 0); res->b += (n_aln == 2);
        if (fo) put_aln(fo, n_aln, aln);
        free(aln);
    }
    res->secs = now_s() - t0;
}

static int mode_splice(int argc, char **argv)
{
    Idx2BWT *bi; reads_t r; hopt_t h; gap_opt_t *opt; range_ctx_t ctx; range_res_t tot;
    if (argc < 5) die("usage: splice <prefix> <reads> <out> [opts] [procs=P] [nout=1] [clear_gape=0|1]");
    bi = load_index(argv[2]); r = load_reads(argv[3]);
    if (!bi->hsp || !bi->bwt->saValue) die("splice needs a full index (SA samples, packed DNA, annotation)");
    opt = parse_opts(argc, argv, 5, &h);
    ctx.bi = bi; ctx.r = &r; ctx.opt = opt; ctx.h = &h;
    run_ranges(splice_range, &ctx, r.n, h.procs, h.nout ? NULL : argv[4], 1, &tot);
    printf("{\"mode\":\"splice\",\"reads\":%u,\"procs\":%d,\"aligned\":%llu,\"two_parts\":%llu,\"secs\":%.6f,\"occ4\":%llu,\"occ1\":%llu}\n",
           r.n, h.procs, tot.a, tot.b, tot.secs, tot.occ4, tot.occ1);
    return 0;
}

/* dump the raw search arrays of both BWTs so that the product's own index builder can be compared bit
 * for bit with the reference builder's output:  per direction
 *   textLength inverseSa0 cumulativeFreq[5] bwtWords occWords occMajorWords, then the three arrays. */
static int mode_dumpindex(int argc, char **argv)
{
    Idx2BWT *bi; FILE *fo; int d;
    if (argc < 4) die("usage: dumpindex <prefix> <out>");
    bi = load_index(argv[2]);
    fo = fopen(argv[3], "wb");
    for (d = 0; d < 2; ++d) {
        BWT *b = d == 0 ? bi->bwt : bi->rev_bwt;
        uint32_t h[10];
        h[0] = b->textLength; h[1] = b->inverseSa0; memcpy(h + 2, b->cumulativeFreq, 20);
        h[7] = b->bwtSizeInWord; h[8] = b->occSizeInWord; h[9] = b->occMajorSizeInWord;
        fwrite(h, 4, 10, fo);
        fwrite(b->bwtCode, 4, b->bwtSizeInWord, fo);
        fwrite(b->occValue, 4, b->occSizeInWord, fo);
        fwrite(b->occValueMajor, 4, b->occMajorSizeInWord, fo);
    }
    fclose(fo);
    printf("{\"mode\":\"dumpindex\",\"textLength\":%u}\n", bi->bwt->textLength);
    return 0;
}

/* sa <prefix> <idx.bin> <out.bin>: BWTSaValue (BWT.c:1195-1225) of every listed SA index on the forward BWT; needs the
 * full index (<prefix>.index.sa) as written by `index`.  Output: n, then per index {SA value, PsiMinus steps walked}. */
static int mode_sa(int argc, char **argv)
{
    Idx2BWT *bi; FILE *fi, *fo; uint32_t n, i, *idx; double secs = 0;
    if (argc < 5) die("usage: sa <prefix> <idx.bin> <out.bin>");
    bi = load_index(argv[2]);
    if (!bi->bwt->saValue) die("index has no SA samples");
    fi = fopen(argv[3], "rb"); if (!fi) die("cannot open idx");
    if (fread(&n, 4, 1, fi) != 1) die("short idx");
    idx = (uint32_t*)malloc(4 * (size_t)n);
    if (fread(idx, 4, n, fi) != n) die("short idx");
    fclose(fi);
    {   /* timed pass: BWTSaValue alone, as bwa_cal_pac_pos / bwt_aln_corelate_check call it */
        double t0 = now_s(); uint32_t acc = 0;
        for (i = 0; i < n; ++i) acc ^= BWTSaValue(bi->bwt, idx[i]);
        secs = now_s() - t0;
        if (acc == 0x12345678u) fprintf(stderr, "\n");
    }
    fo = fopen(argv[4], "wb");
    fwrite(&n, 4, 1, fo);
    for (i = 0; i < n; ++i) {
        uint32_t o[2], s = idx[i];
        o[0] = BWTSaValue(bi->bwt, idx[i]);
        o[1] = 0;
        while (s % bi->bwt->saInterval != 0) { s = BWTPsiMinusValue(bi->bwt, s); ++o[1]; }
        fwrite(o, 4, 2, fo);
    }
    fclose(fo);
    printf("{\"mode\":\"sa\",\"n\":%u,\"textLength\":%u,\"saInterval\":%u,\"secs\":%.6f}\n", n, bi->bwt->textLength, bi->bwt->saInterval, secs);
    return 0;
}

/* locate <prefix> <idx.bin> <out.bin>: BWTRetrievePositionFromSAIndex (2BWT-Interface.c:329-362) for every listed SA
 * index (full index needed: SA samples and the annotation's block list).  Output: n, then {occ_pos, seq_id, ori_pos}
 * per index; seq_id / ori_pos are preset to -1 so that "no block found" is visible. */
static int mode_locate(int argc, char **argv)
{
    Idx2BWT *bi; FILE *fi, *fo; uint32_t n, i, *idx;
    if (argc < 5) die("usage: locate <prefix> <idx.bin> <out.bin>");
    bi = load_index(argv[2]);
    if (!bi->hsp || !bi->bwt->saValue) die("locate needs a full index");
    fi = fopen(argv[3], "rb"); if (!fi) die("cannot open idx");
    if (fread(&n, 4, 1, fi) != 1) die("short idx");
    idx = (uint32_t*)malloc(4 * (size_t)n);
    if (fread(idx, 4, n, fi) != n) die("short idx");
    fclose(fi);
    fo = fopen(argv[4], "wb");
    fwrite(&n, 4, 1, fo);
    for (i = 0; i < n; ++i) {
        unsigned int o[3] = {0, 0xFFFFFFFFu, 0xFFFFFFFFu};
        BWTRetrievePositionFromSAIndex(bi, idx[i], &o[1], &o[2], &o[0]);
        fwrite(o, 4, 3, fo);
    }
    fclose(fo);
    printf("{\"mode\":\"locate\",\"n\":%u,\"blocks\":%d}\n", n, bi->hsp->numOfBlock);
    return 0;
}


/* ---------------------------------------------------------------- sam
 * The stock batch loop of bwa_aln_core (bwtaln.c:477-522): bwa_cal_sa_reg_gap, then generate_sam_se_core (bwtse.c:884-931:
 * bwt_aln2seq_core hit selection with the process-wide drand48 stream, bwa_cal_pac_pos, bwa_refine_gapped, bwa_print_sam1)
 * with n_occ = 3 as bwtaln.c:514 passes it.  The SAM text the reference prints goes to <out.sam>; <out.bin> receives, per
 * read, the bwa_seq_t fields generate_sam_se_core computed:
 *   'HSAM' n, then per read 20 words {type, strand, n_mm, n_gapo, n_gape, mapQ, score, sa, seq_id, ori_pos, occ_pos, c1, c2,
 *   start, end, n_cigar, nm, md_len, n_multi, 0}, n_cigar cigar words, md bytes padded to 4, and per multi record 12 words
 *   {n_cigar, gap, mm, strand, sa, ori_pos, occ_pos, seq_id, aln_id, start, end, 0} + its cigar words.
 * Reads of type NO_MATCH dump zeros after the type word (the reference leaves stale values there and prints nothing). */
#include "bwtse.h"
static void put_sam_rec(FILE *f, const bwa_seq_t *p)
{
    uint32_t w[20]; int j; uint32_t md_len = 0, pad = 0;
    memset(w, 0, sizeof(w));
    w[0] = p->type;
    if (p->type != BWA_TYPE_NO_MATCH) {
        md_len = p->md ? (uint32_t)strlen(p->md) : 0;
        w[1] = p->strand; w[2] = p->n_mm; w[3] = p->n_gapo; w[4] = p->n_gape; w[5] = p->mapQ; w[6] = (uint32_t)p->score;
        w[7] = p->sa; w[8] = p->seq_id; w[9] = p->ori_pos; w[10] = p->occ_pos; w[11] = (uint32_t)p->c1; w[12] = (uint32_t)p->c2;
        w[13] = (uint32_t)p->start; w[14] = (uint32_t)p->end; w[15] = p->cigar ? (uint32_t)p->n_cigar : 0; w[16] = p->nm; w[17] = md_len;
        w[18] = (uint32_t)p->n_multi;
    }
    fwrite(w, 4, 20, f);
    if (p->type == BWA_TYPE_NO_MATCH) return;
    if (p->cigar) fwrite(p->cigar, 4, p->n_cigar, f);
    if (md_len) { fwrite(p->md, 1, md_len, f); if (md_len & 3) fwrite(&pad, 1, 4 - (md_len & 3), f); }
    for (j = 0; j < p->n_multi; ++j) {
        const bwt_multi1_t *q = p->multi + j; uint32_t m[12];
        m[0] = q->cigar ? q->n_cigar : 0; m[1] = q->gap; m[2] = q->mm; m[3] = q->strand; m[4] = q->sa; m[5] = q->ori_pos; m[6] = q->occ_pos;
        m[7] = q->seq_id; m[8] = q->aln_id; m[9] = (uint32_t)q->start; m[10] = (uint32_t)q->end; m[11] = 0;
        fwrite(m, 4, 12, f);
        if (q->cigar) fwrite(q->cigar, 4, q->n_cigar, f);
    }
}

#ifdef HSA_WITH_GPU_SHIM
void hsa_gpu_sam_print(const HSP *hsp, int n_seqs, bwa_seq_t *seqs, int mode, int max_top2);
void bwt_aln2seq_core(bwa_seq_t *s, int set_main, int n_multi);
void bwa_cal_pac_pos(const Idx2BWT *bi_bwt, int n_seqs, bwa_seq_t *seq, int max_mm, float fnr);
static void sam_fields_ref_print_shim(Idx2BWT *bi_bwt, int n_seqs, bwa_seq_t *seqs, gap_opt_t *opt, int n_occ)
{
    int i;
    for (i = 0; i < n_seqs; ++i) {                                 /* bwtse.c:899-908 */
        bwa_seq_t *p = seqs + i;
        if (p->n_aln == 2 && p->aln->type == BWA_TYPE_SPLICING) p->type = BWA_TYPE_SPLICING;
        else bwt_aln2seq_core(p, 1, n_occ);
    }
    bwa_cal_pac_pos(bi_bwt, n_seqs, seqs, opt->max_diff, opt->fnr);  /* :911 */
    bwa_refine_gapped(bi_bwt->hsp, n_seqs, seqs);                    /* :916 */
    hsa_gpu_sam_print(bi_bwt->hsp, n_seqs, seqs, opt->mode, opt->max_top2);
}
#endif

static int mode_sam(int argc, char **argv)
{
    Idx2BWT *bi; reads_t r; FILE *fo; hopt_t h; gap_opt_t *opt0, *opt; uint32_t hdr[2], b; int saved_stdout, n_occ = 3, i;
    bwt_array_t *arr; double secs_search = 0, secs_sam = 0; unsigned long long n_matched = 0, n_gapped = 0, n_splice = 0;
    if (argc < 6) die("usage: sam <prefix> <reads> <out.bin> <out.sam> [opts] [batch=N] [nocc=K]");
    bi = load_index(argv[2]); r = load_reads(argv[3]);
    for (i = 6; i < argc; ++i) if (strncmp(argv[i], "nocc=", 5) == 0) { n_occ = atoi(argv[i] + 5); argv[i] = "batch=100000"; }
    opt0 = parse_opts(argc, argv, 6, &h);
    opt = (gap_opt_t*)calloc(2, sizeof(gap_opt_t)); *opt = *opt0;
    arr = bwt_array_init();
    fo = fopen(argv[4], "wb"); if (!fo) die("cannot open out.bin");
    hdr[0] = 0x4D415348u; hdr[1] = r.n; fwrite(hdr, 4, 2, fo);
    fflush(stdout); saved_stdout = dup(1);
    if (!freopen(argv[5], "w", stdout)) die("cannot open out.sam");
    for (b = 0; b < r.n; b += (uint32_t)h.batch) {
        uint32_t e = b + (uint32_t)h.batch < r.n ? b + (uint32_t)h.batch : r.n, k;
        int n = (int)(e - b);
        bwa_seq_t *seqs = (bwa_seq_t*)calloc(n, sizeof(bwa_seq_t));
        double t0;
        for (k = b; k < e; ++k) {                                  /* what bwa_read_seq leaves (bwaseqio.c:198-224), no qualities */
            bwa_seq_t *p = seqs + (k - b); char nm[32];
            p->tid = -1; p->full_len = p->clip_len = p->len = r.len[k];
            p->seq = (ubyte_t*)calloc(p->len + 1, 1); memcpy(p->seq, r.codes + r.off[k], p->len);
            p->rseq = (ubyte_t*)calloc(p->len + 1, 1); memcpy(p->rseq, p->seq, p->len);
            seq_reverse(p->len, p->rseq, opt->mode & BWA_MODE_COMPREAD);
            sprintf(nm, "r%u", k); p->name = strdup(nm);
            if (h.qual) {                                          /* qual=1: what a FASTQ with barcodes and trimming leaves (bwaseqio.c:150-203) */
                uint32_t j;
                p->qual = (ubyte_t*)calloc(p->len + 1, 1);
                for (j = 0; j < p->len; ++j) p->qual[j] = (ubyte_t)(33 + (k * 7u + j * 13u) % 30u);   /* below 63: the tests drop lines that hold a question mark */
                if (k % 3 == 0) strcpy(p->bc, k % 2 ? "ACgT" : "TTAGGC");
                if (k % 5 == 0) p->clip_len = (int)p->len - 3;     /* only printed (XC:i), bwtse.c:757 */
            }
        }
        t0 = now_s();
        g_driver(0, bi, n, seqs, opt, arr);
        secs_search += now_s() - t0;
        t0 = now_s();
        g_sam(bi, n, seqs, opt, n_occ);
        secs_sam += now_s() - t0;
        for (k = 0; k < (uint32_t)n; ++k) {
            bwa_seq_t *p = seqs + k;
            put_sam_rec(fo, p);
            if (p->type != BWA_TYPE_NO_MATCH) { ++n_matched; if (p->cigar) ++n_gapped; if (p->type == BWA_TYPE_SPLICING) ++n_splice; }
        }
        bwa_free_read_seq(n, seqs);
    }
    fclose(fo);
    fflush(stdout); dup2(saved_stdout, 1); close(saved_stdout);
    printf("{\"mode\":\"sam\",\"reads\":%u,\"batch\":%d,\"n_occ\":%d,\"matched\":%llu,\"with_cigar\":%llu,\"splicing\":%llu,"
           "\"secs_search\":%.6f,\"secs_sam\":%.6f}\n", r.n, h.batch, n_occ, n_matched, n_gapped, n_splice, secs_search, secs_sam);
    return 0;
}


/* ---------------------------------------------------------------- dp
 * aln_global_core (stdaln.c:345) + bwa_aln_path2cigar (bwtaln.c:624) with aln_param_bwa, as refine_gapped_core calls them
 * (bwtse.c:398-399), on arbitrary (reference window, read) pairs: in = n, then per pair {len1, len2, len1 + len2 bytes};
 * out = n, then per pair {score, n_cigar, cigar words}. */
#include "stdaln.h"
bwa_cigar_t *bwa_aln_path2cigar(const path_t *path, int path_len, int *n_cigar);
static int mode_dp(int argc, char **argv)
{
    FILE *fi, *fo; uint32_t n, i;
    if (argc < 4) die("usage: dp <pairs.bin> <out.bin>");
    fi = fopen(argv[2], "rb"); fo = fopen(argv[3], "wb");
    if (!fi || !fo || fread(&n, 4, 1, fi) != 1) die("dp: cannot open files");
    fwrite(&n, 4, 1, fo);
    for (i = 0; i < n; ++i) {
        uint32_t len[2]; ubyte_t *a, *b; path_t *path; int path_len = 0, n_cigar = 0, score; bwa_cigar_t *cigar; AlnParam ap = aln_param_bwa;
        if (fread(len, 4, 2, fi) != 2) die("dp: short input");
        a = (ubyte_t*)calloc(len[0] + 1, 1); b = (ubyte_t*)calloc(len[1] + 1, 1);
        if (fread(a, 1, len[0], fi) != len[0] || fread(b, 1, len[1], fi) != len[1]) die("dp: short input");
        path = (path_t*)calloc(len[0] + len[1] + 2, sizeof(path_t));
        score = aln_global_core(a, (int)len[0], b, (int)len[1], &ap, path, &path_len);
        cigar = path_len ? bwa_aln_path2cigar(path, path_len, &n_cigar) : NULL;
        fwrite(&score, 4, 1, fo); fwrite(&n_cigar, 4, 1, fo);
        if (n_cigar) fwrite(cigar, 4, n_cigar, fo);
        free(cigar); free(path); free(a); free(b);
    }
    fclose(fi); fclose(fo);
    printf("{\"mode\":\"dp\",\"pairs\":%u}\n", n);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc < 2) die("usage: hsa_ref <index|occ|width|percall|seeds|driver|whole|dumpindex|maxdiff> ...");
    if (strcmp(argv[1], "index") == 0) {
        /* argv: index <db> <fasta>; the builder looks for "<argv0>.ini" = "index.ini" in cwd and falls
         * back to compiled defaults (2BWT-Builder.c:242-243, :59-100). */
        return bwa_index_main(argc - 1, argv + 1);
    }
    if (strcmp(argv[1], "occ") == 0) return mode_occ(argc, argv);
    if (strcmp(argv[1], "sa") == 0) return mode_sa(argc, argv);
    if (strcmp(argv[1], "locate") == 0) return mode_locate(argc, argv);
    if (strcmp(argv[1], "width") == 0) return mode_width(argc, argv);
    if (strcmp(argv[1], "percall") == 0) return mode_percall(argc, argv);
    if (strcmp(argv[1], "seeds") == 0) return mode_seeds(argc, argv);
    if (strcmp(argv[1], "driver") == 0) return mode_driver(argc, argv);
    if (strcmp(argv[1], "sam") == 0) return mode_sam(argc, argv);
    if (strcmp(argv[1], "dp") == 0) return mode_dp(argc, argv);
#ifdef HSA_WITH_GPU_SHIM
    if (strcmp(argv[1], "gpudriver") == 0) {
        /* the same batch loop with bwa_cal_sa_reg_gap replaced by the GPU shim (needs a full index: the splice
         * path reads the SA and the packed DNA) */
        Idx2BWT *bi; int rc;
        if (argc < 5) die("usage: gpudriver <prefix> <reads> <out> [opts] [batch=N]");
        bi = load_index(argv[2]);
        if (hsa_gpu_open(bi, 0)) return 3;
        g_driver = bwa_cal_sa_reg_gap_gpu;
        rc = mode_driver(argc, argv);
        hsa_gpu_close();
        return rc;
    }
#endif
#ifdef HSA_WITH_GPU_SHIM
    if (strcmp(argv[1], "samfmt") == 0) {
        /* the stock program with ONLY the print loop replaced: fields by the reference's own generate_sam_se_core steps
         * (bwtse.c:899-918), lines by the shim's formatter (hsa_gpu_sam_print) -- runs without a GPU */
        g_sam = sam_fields_ref_print_shim;
        return mode_sam(argc, argv);
    }
    if (strcmp(argv[1], "gpusam") == 0) {
        /* the stock batch loop with BOTH stages from the shim: bwa_cal_sa_reg_gap_gpu and generate_sam_se_core_gpu */
        Idx2BWT *bi; int rc;
        if (argc < 6) die("usage: gpusam <prefix> <reads> <out.bin> <out.sam> [opts] [batch=N] [nocc=K]");
        bi = load_index(argv[2]);
        if (hsa_gpu_open(bi, 0)) return 1;
        g_driver = bwa_cal_sa_reg_gap_gpu; g_sam = generate_sam_se_core_gpu;
        rc = mode_sam(argc, argv);
        hsa_gpu_close();
        return rc;
    }
    if (strcmp(argv[1], "gpupercall") == 0) {
        /* percall with every bwt_match_gap call ALSO made through the per-call GPU symbol (bwt_match_gap_gpu) on the
         * same bwt_aux_t frame; the dump holds the GPU results, the JSON line counts calls whose hits or rewritten
         * width_back differ from the reference's */
        Idx2BWT *bi; int rc;
        if (argc < 5) die("usage: gpupercall <prefix> <reads> <out> [opts]");
        bi = load_index(argv[2]);
        if (hsa_gpu_open(bi, 0)) return 3;
        g_match = bwt_match_gap_gpu;
        rc = mode_percall(argc, argv);
        hsa_gpu_close();
        return rc ? rc : (g_call_mismatch || g_width_mismatch ? 4 : 0);
    }
    if (strcmp(argv[1], "gpuseeds") == 0) {
        /* the same for the six seed calls of bwt_splice_match (width_seed aliasing width_back, prefix-width quirk);
         * single process: the GPU context does not survive a fork */
        Idx2BWT *bi; int rc;
        if (argc < 5) die("usage: gpuseeds <prefix> <reads> <out> [opts]");
        bi = load_index(argv[2]);
        if (hsa_gpu_open(bi, 0)) return 3;
        g_match = bwt_match_gap_gpu;
        rc = mode_seeds(argc, argv);
        hsa_gpu_close();
        return rc ? rc : (g_call_mismatch || g_width_mismatch ? 4 : 0);
    }
    if (strcmp(argv[1], "gpusa") == 0) {
        /* gpusa <prefix> <idx.bin>: every listed SA index through BWTSaValue on the host and through the shim's GPU
         * batch call; exits 0 only if all values agree (prints the count of mismatches) */
        Idx2BWT *bi; FILE *fi; uint32_t n, i, bad = 0, *idx, *pos;
        if (argc < 4) die("usage: gpusa <prefix> <idx.bin>");
        bi = load_index(argv[2]);
        fi = fopen(argv[3], "rb"); if (!fi) die("cannot open idx");
        if (fread(&n, 4, 1, fi) != 1) die("short idx");
        idx = (uint32_t*)malloc(4 * (size_t)n); pos = (uint32_t*)malloc(4 * (size_t)n);
        if (fread(idx, 4, n, fi) != n) die("short idx");
        fclose(fi);
        if (hsa_gpu_open(bi, 0)) return 3;
        hsa_gpu_sa_values(bi, idx, n, pos);
        for (i = 0; i < n; ++i) bad += pos[i] != BWTSaValue(bi->bwt, idx[i]);
        hsa_gpu_close();
        printf("{\"mode\":\"gpusa\",\"n\":%u,\"mismatches\":%u}\n", n, bad);
        return bad ? 4 : 0;
    }
#endif
    if (strcmp(argv[1], "whole") == 0) return mode_whole(argc, argv);
    if (strcmp(argv[1], "splice") == 0) return mode_splice(argc, argv);
    if (strcmp(argv[1], "dumpindex") == 0) return mode_dumpindex(argc, argv);
    if (strcmp(argv[1], "maxdiff") == 0) {
        int l;
        for (l = 1; l <= (argc > 2 ? atoi(argv[2]) : 300); ++l)
            printf("%d %d\n", l, bwa_cal_maxdiff(l, BWA_AVG_ERR, argc > 3 ? atof(argv[3]) : 0.04));
        return 0;
    }
    die("unknown mode");
    return 1;
}

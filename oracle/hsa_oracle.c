/*
 * hsa_oracle.c -- TEST INFRASTRUCTURE ONLY: plain-C restatement of HSA's inexact-search hot path.
 *
 * Not product code.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg load it.
 * Each function cites the reference file:line it restates.  Parity is PINNED against the unmodified
 * reference (oracle/_ref, built by oracle/Makefile from /root/reference) and the golden vectors under
 * tests/golden/ that the same reference generated (tests/golden/make_golden.py).
 *
 * The restatement is deliberately scalar and literal: same bucketed stack with growable arrays, same
 * push order, same in-place width mutation; only the SSE2 rank kernel is replaced by word popcounts
 * with identical results.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "hsa_oracle.h"

#define OCC_INTERVAL        256u      /* BWT.h:40 */
#define OCC_INTERVAL_MAJOR  65536u    /* BWT.h:42 */

static uint64_t g_occ4_calls = 0, g_occ1_calls = 0;
void     hsao_reset_counters(void) { g_occ4_calls = g_occ1_calls = 0; }
uint64_t hsao_occ4_calls(void) { return g_occ4_calls; }
uint64_t hsao_occ1_calls(void) { return g_occ1_calls; }
void     hsao_free(void *p) { free(p); }

/* ------------------------------------------------------------------ rank (BWT.c) */

/* count the four symbols among pairs [a,b) of one packed word; pair 0 is the two MOST significant bits
 * (BWT.c:954, BWTConstruct.c:1344). */
static void count_pairs(uint32_t w, unsigned a, unsigned b, uint32_t cnt[4])
{
    uint32_t m = 0x55555555u, lo, hi;
    if (a >= b) return;
    if (a > 0)  m &= 0xFFFFFFFFu >> (2 * a);
    if (b < 16) m &= ~(0xFFFFFFFFu >> (2 * b));
    lo = w & m;
    hi = (w >> 1) & m;
    cnt[3] += (uint32_t)__builtin_popcount(lo & hi);
    cnt[2] += (uint32_t)__builtin_popcount(hi & ~lo);
    cnt[1] += (uint32_t)__builtin_popcount(lo & ~hi);
    cnt[0] += (uint32_t)__builtin_popcount(m & ~(lo | hi));
}

/* symbols of bwtCode in [from,to): what BWTDecodeAll (BWT.c:532-679) computes with two 128-bit loads */
static void count_range(const hsa_bwt_view_t *bwt, uint32_t from, uint32_t to, uint32_t cnt[4])
{
    uint32_t p = from;
    cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0;
    while (p < to) {
        uint32_t wi = p >> 4, a = p & 15u, wend = (wi + 1) << 4;
        uint32_t b = to < wend ? (to & 15u) : 16u;
        count_pairs(bwt->bwtCode[wi], a, b, cnt);
        p = to < wend ? to : wend;
    }
}

/* BWTAllOccValueExplicit / BWTOccValueExplicit, BWT.c:1018-1059 */
static uint32_t occ_explicit(const hsa_bwt_view_t *bwt, uint32_t e, uint32_t c)
{
    uint32_t major = bwt->occValueMajor[(e * OCC_INTERVAL / OCC_INTERVAL_MAJOR) * 4 + c];
    uint32_t word = bwt->occValue[(e / 2) * 4 + c];
    return major + ((e % 2 == 0) ? (word >> 16) : (word & 0xFFFFu));
}

/* BWTAllOccValue, BWT.c:793-837 */
void hsao_occ4(const hsa_bwt_view_t *bwt, uint32_t index, uint32_t occ[4])
{
    uint32_t e, base, cnt[4], c;
    ++g_occ4_calls;
    index -= (index > bwt->inverseSa0);                      /* BWT.c:804: '$' is not stored */
    e = (index + OCC_INTERVAL / 2 - 1) / OCC_INTERVAL;       /* BWT.c:813: nearest sample, may be above */
    base = e * OCC_INTERVAL;
    for (c = 0; c < 4; ++c) occ[c] = occ_explicit(bwt, e, c);
    if (base == index) return;
    if (index > base) { count_range(bwt, base, index, cnt); for (c = 0; c < 4; ++c) occ[c] += cnt[c]; }
    else              { count_range(bwt, index, base, cnt); for (c = 0; c < 4; ++c) occ[c] -= cnt[c]; }
}

/* BWTOccValue, BWT.c:682-719 */
uint32_t hsao_occ1(const hsa_bwt_view_t *bwt, uint32_t index, uint32_t c)
{
    uint32_t e, base, cnt[4], v;
    ++g_occ1_calls;
    index -= (index > bwt->inverseSa0);
    e = (index + OCC_INTERVAL / 2 - 1) / OCC_INTERVAL;
    base = e * OCC_INTERVAL;
    v = occ_explicit(bwt, e, c);
    if (base == index) return v;
    if (index > base) { count_range(bwt, base, index, cnt); return v + cnt[c]; }
    count_range(bwt, index, base, cnt);
    return v - cnt[c];
}

/* ------------------------------------------------------------------ SA index -> position (BWT.c) */

/* BWTPsiMinusValue, BWT.c:1142-1165: the SA index of the suffix one position to the left.  BWTOccValueOnSpot(index + 1)
 * (BWT.c:924-965) returns the BWT symbol in front of `index + 1` and its occ count up to and including itself. */
uint32_t hsao_psi_minus(const hsa_bwt_view_t *bwt, uint32_t index)
{
    uint32_t p, c;
    if (index == bwt->inverseSa0) return 0;                  /* BWT.c:1161-1163 */
    p = index + 1;
    p -= (p > bwt->inverseSa0);                              /* BWT.c:948 */
    c = (bwt->bwtCode[(p - 1) / 16] >> (30 - 2 * ((p - 1) % 16))) & 3u;     /* BWT.c:954 */
    return bwt->cumulativeFreq[c] + hsao_occ1(bwt, index + 1, c);
}

/* BWTSaValue, BWT.c:1195-1225 */
uint32_t hsao_sa_value(const hsa_bwt_view_t *bwt, const uint32_t *sa_value, uint32_t sa_interval, uint32_t sa_index,
                       uint32_t *steps)
{
    uint32_t skipped = 0;
    while (sa_index % sa_interval != 0) {
        ++skipped;
        sa_index = hsao_psi_minus(bwt, sa_index);
    }
    if (steps) *steps = skipped;
    return sa_value[sa_index / sa_interval] + skipped;       /* SA[0] is stored as -1 (BWT.c:222, :1221-1222) */
}

/* 2BWT-Interface.c:339-361.  The reference's loop runs `while (l <= h)` from h = numOfBlock with unsigned arithmetic; for
 * a position inside some block it never looks at blockList[numOfBlock] and never takes m - 1 of m = 0, so it is the
 * plain binary search restated here; positions outside every block (only SA[0] = -1 in practice) are "not found". */
int hsao_locate(const uint32_t *b4, uint32_t n_blocks, uint32_t occ_pos, uint32_t *seq_id, uint32_t *ori_pos)
{
    uint32_t l = 0, h = n_blocks;
    while (l < h) {
        uint32_t m = (l + h) >> 1;
        if (b4[4 * m + 1] > occ_pos) h = m;
        else if (b4[4 * m + 2] < occ_pos) l = m + 1;
        else { *seq_id = b4[4 * m]; *ori_pos = occ_pos - b4[4 * m + 1] + b4[4 * m + 3] + 1; return 1; }
    }
    return 0;
}

/* ------------------------------------------------------------------ SA-range stepping (2BWT-Interface.c) */

/* BWTSARangeForeward, 2BWT-Interface.c:121-132: forward extension = backward step on rev_bwt with the
 * FORWARD bwt's cumulativeFreq */
static void sa_forward(const hsao_index_t *ix, uint32_t c, uint32_t *k, uint32_t *l)
{
    uint32_t a = *k, b = *l;
    *k = ix->fwd.cumulativeFreq[c] + hsao_occ1(&ix->rev, a, c) + 1;
    *l = ix->fwd.cumulativeFreq[c] + hsao_occ1(&ix->rev, b + 1, c);
}

/* BWTSARangeBackward, 2BWT-Interface.c:107-118 */
static void sa_backward(const hsao_index_t *ix, uint32_t c, uint32_t *k, uint32_t *l)
{
    uint32_t a = *k, b = *l;
    *k = ix->fwd.cumulativeFreq[c] + hsao_occ1(&ix->fwd, a, c) + 1;
    *l = ix->fwd.cumulativeFreq[c] + hsao_occ1(&ix->fwd, b + 1, c);
}

/* BWTAllSARangesBackward_Bidirection, 2BWT-Interface.c:235-271 */
static void sa_backward_all_bi(const hsao_index_t *ix, uint32_t k, uint32_t l, uint32_t rev_l,
                               uint32_t sk[4], uint32_t sl[4], uint32_t rsk[4], uint32_t rsl[4])
{
    uint32_t oL[4], oR[4], oCount[4];
    int c;
    hsao_occ4(&ix->fwd, k, oL);
    hsao_occ4(&ix->fwd, l + 1, oR);
    oCount[3] = 0;
    for (c = 2; c >= 0; --c) oCount[c] = oCount[c + 1] + oR[c + 1] - oL[c + 1];
    for (c = 0; c < 4; ++c) {
        sk[c] = ix->fwd.cumulativeFreq[c] + oL[c] + 1;
        sl[c] = ix->fwd.cumulativeFreq[c] + oR[c];
        rsl[c] = rev_l - oCount[c];
        rsk[c] = rsl[c] - (sl[c] - sk[c]);
    }
}

/* BWTSARangeBackward_Bidirection, 2BWT-Interface.c:135-168 */
static void sa_backward_bi(const hsao_index_t *ix, uint32_t c, uint32_t *k, uint32_t *l,
                           uint32_t *rev_k, uint32_t *rev_l)
{
    uint32_t sk[4], sl[4], rsk[4], rsl[4];
    sa_backward_all_bi(ix, *k, *l, *rev_l, sk, sl, rsk, rsl);
    *k = sk[c]; *l = sl[c]; *rev_k = rsk[c]; *rev_l = rsl[c];
}

/* bwt_match_exact, 2BWT-Interface.c:365-388 (including the write-back-only-if-non-zero quirk :383-386) */
int hsao_match_exact(const hsao_index_t *ix, const uint8_t *seq, int len,
                     uint32_t *k, uint32_t *l, uint32_t *rev_k, uint32_t *rev_l)
{
    uint32_t a = *k, b = *l, ra = *rev_k, rb = *rev_l;
    int i;
    for (i = len - 1; i >= 0; --i) {
        if (seq[i] > 3) return 0;
        sa_backward_bi(ix, seq[i], &a, &b, &ra, &rb);
        if (a > b) break;
    }
    if (a > b) return 0;
    if (*k) *k = a;
    if (*l) *l = b;
    if (*rev_k) *rev_k = ra;
    if (*rev_l) *rev_l = rb;
    return (int)(b - a + 1);
}

/* ------------------------------------------------------------------ host helpers (bwtaln.c, bwaseqio.c) */

void hsao_gap_opt_default(hsa_gap_opt_t *o)                 /* gap_init_opt, bwtaln.c:21-44 */
{
    memset(o, 0, sizeof(*o));
    o->s_mm = 3; o->s_gapo = 11; o->s_gape = 4;
    o->max_diff = -1; o->max_gapo = 1; o->max_gape = 6;
    o->indel_end_skip = 5; o->max_del_occ = 10; o->max_entries = 2000000;
    o->mode = HSA_MODE_GAPE | HSA_MODE_COMPREAD;
    o->seed_len = 32; o->max_seed_diff = 2;
    o->fnr = 0.04f;
    o->n_threads = 1;
    o->max_top2 = 30;
    o->trim_qual = 0;
}

int hsao_cal_maxdiff(int l, double err, double thres)       /* bwa_cal_maxdiff, bwtaln.c:46-58 */
{
    double elambda = exp(-l * err);
    double sum, y = 1.0;
    int k, x = 1;
    for (k = 1, sum = elambda; k < 1000; ++k) {
        y *= l * err;
        x *= k;                                             /* int overflow for k > 12 is the reference's own */
        sum += elambda * y / x;
        if (1.0 - sum < thres) return k;
    }
    return 2;
}

void hsao_seq_revcomp(int len, uint8_t *seq)                /* seq_reverse(len, seq, 1), bwaseqio.c:73-90 */
{
    int i;
    for (i = 0; i < len >> 1; ++i) {
        uint8_t t = seq[len - 1 - i];
        if (t < 4) t = 3 - t;
        seq[len - 1 - i] = seq[i] >= 4 ? seq[i] : 3 - seq[i];
        seq[i] = t;
    }
    if (len & 1) seq[i] = seq[i] >= 4 ? seq[i] : 3 - seq[i];
}

/* bwt_cal_width, bwtaln.c:73-116 */
int hsao_cal_width(const hsao_index_t *ix, int len, const uint8_t *str, hsa_width_t *width, int type)
{
    uint32_t k = 0, l = ix->fwd.textLength;
    int i, bid = 0;
    if (type == 1) {
        for (i = 0; i < len; ++i) {
            uint8_t c = str[i];
            if (c < 4) sa_forward(ix, c, &k, &l);
            if (k > l || c > 3) { k = 0; l = ix->fwd.textLength; ++bid; }
            width[i].w = l - k + 1;
            width[i].bid = bid;
        }
    } else {
        for (i = len - 1; i > 0; --i) {                     /* bwtaln.c:99: note i > 0, entry 0 untouched */
            uint8_t c = str[i];
            if (c < 4) sa_backward(ix, c, &k, &l);
            if (k > l || c > 3) { k = 0; l = ix->fwd.textLength; ++bid; }
            width[i].w = l - k + 1;
            width[i].bid = bid;
        }
    }
    width[len].w = 0;
    width[len].bid = ++bid;
    return bid;
}

/* ------------------------------------------------------------------ priority stack (bwtgap.c:13-92) */

enum { ST_M = 0, ST_I = 1, ST_D = 2 };

typedef struct {                 /* gap_entry_t, bwtaln.h:52-58 */
    uint32_t info;               /* score<<21 | i */
    uint8_t n_mm, n_gapo, n_gape, state;
    uint32_t k, l, rev_k, rev_l;
    int last_diff_pos;
} node_t;

typedef struct { int n, m; node_t *a; } bucket_t;
typedef struct { int n_buckets, best, n_entries; bucket_t *b; } pstack_t;

static int score_of(int mm, int o, int e, const hsa_gap_opt_t *p) { return mm * p->s_mm + o * p->s_gapo + e * p->s_gape; }

static pstack_t *pstack_new(int n_buckets)
{
    pstack_t *s = (pstack_t*)calloc(1, sizeof(*s));
    int i;
    s->n_buckets = n_buckets;
    s->b = (bucket_t*)calloc((size_t)n_buckets, sizeof(bucket_t));
    for (i = 0; i < n_buckets; ++i) { s->b[i].m = 4; s->b[i].a = (node_t*)calloc(4, sizeof(node_t)); }
    s->best = n_buckets;
    return s;
}

static void pstack_free(pstack_t *s)
{
    int i;
    for (i = 0; i < s->n_buckets; ++i) free(s->b[i].a);
    free(s->b); free(s);
}

/* gap_push, bwtgap.c:46-78 */
static void push(pstack_t *s, int i, uint32_t k, uint32_t l, uint32_t rev_k, uint32_t rev_l,
                 int n_mm, int n_gapo, int n_gape, int state, int is_diff, const hsa_gap_opt_t *opt)
{
    int sc = score_of(n_mm, n_gapo, n_gape, opt);
    bucket_t *q;
    node_t *p;
    if (sc >= s->n_buckets) {               /* the reference would write out of bounds; we size generously */
        int nb = sc + 1, j;
        s->b = (bucket_t*)realloc(s->b, (size_t)nb * sizeof(bucket_t));
        for (j = s->n_buckets; j < nb; ++j) { s->b[j].n = 0; s->b[j].m = 4; s->b[j].a = (node_t*)calloc(4, sizeof(node_t)); }
        if (s->best == s->n_buckets) s->best = nb;
        s->n_buckets = nb;
    }
    q = s->b + sc;
    if (q->n == q->m) { q->m <<= 1; q->a = (node_t*)realloc(q->a, sizeof(node_t) * (size_t)q->m); }
    p = q->a + q->n;
    p->info = (uint32_t)sc << 21 | (uint32_t)i;
    p->k = k; p->l = l; p->rev_k = rev_k; p->rev_l = rev_l;
    p->n_mm = (uint8_t)n_mm; p->n_gapo = (uint8_t)n_gapo; p->n_gape = (uint8_t)n_gape; p->state = (uint8_t)state;
    p->last_diff_pos = is_diff ? i : 0;
    ++q->n; ++s->n_entries;
    if (s->best > sc) s->best = sc;
}

/* gap_pop, bwtgap.c:80-92: LAST entry of the LOWEST non-empty bucket */
static void pop(pstack_t *s, node_t *e)
{
    bucket_t *q = s->b + s->best;
    *e = q->a[q->n - 1];
    --q->n; --s->n_entries;
    if (q->n == 0 && s->n_entries) {
        int i;
        for (i = s->best + 1; i < s->n_buckets; ++i) if (s->b[i].n != 0) break;
        s->best = i;
    } else if (s->n_entries == 0) s->best = s->n_buckets;
}

/* gap_shadow, bwtgap.c:94-105 */
static void shadow(uint32_t x, uint32_t max, int last_diff_pos, hsa_width_t *w)
{
    int i, j;
    for (i = j = 0; i < last_diff_pos; ++i) {
        if (w[i].w > x) w[i].w -= x;
        else if (w[i].w == x) { w[i].bid = 1; w[i].w = max - (uint32_t)(++j); }
    }
}

static int ilog2(uint32_t v)             /* int_log2, bwtgap.c:107-116 */
{
    int c = 0;
    if (v & 0xffff0000u) { v >>= 16; c |= 16; }
    if (v & 0xff00) { v >>= 8; c |= 8; }
    if (v & 0xf0) { v >>= 4; c |= 4; }
    if (v & 0xc) { v >>= 2; c |= 2; }
    if (v & 0x2) c |= 1;
    return c;
}

/* ------------------------------------------------------------------ bwt_match_gap (bwtgap.c:118-331) */
hsa_aln1_t *hsao_match_gap(const hsao_index_t *ix, const hsa_gap_opt_t *opt, const uint8_t *seq, int len,
                           hsa_width_t *width, hsa_width_t *width_seed, int strand,
                           int *_n_aln, int *n_entries_peak)
{
    const uint32_t N = ix->fwd.textLength;
    int best_score = score_of(opt->max_diff + 1, opt->max_gapo + 1, opt->max_gape + 1, opt);   /* :128 */
    int best_diff = opt->max_diff + 1, max_diff = opt->max_diff;
    int best_cnt = 0, peak = 0, j, n_aln = 0, m_aln = 10;
    hsa_aln1_t *aln = (hsa_aln1_t*)calloc((size_t)m_aln, sizeof(hsa_aln1_t));                 /* :138 */
    pstack_t *st = pstack_new(best_score + opt->s_mm + 1);
    node_t e;
    (void)best_diff;

    push(st, len, 0, N, 0, N, 0, 0, 0, ST_M, 0, opt);                                          /* :142 */

    while (st->n_entries) {                                                                    /* :144 */
        int i, m, m_seed = 0, hit_found, allow_diff, allow_M, tmp;
        uint32_t k, l, rev_k, rev_l, sk[4], sl[4], rsk[4], rsl[4], occ;

        if (peak < st->n_entries) peak = st->n_entries;
        if (st->n_entries > opt->max_entries) break;                                           /* :150 */
        pop(st, &e);
        k = e.k; l = e.l; rev_k = e.rev_k; rev_l = e.rev_l;
        i = (int)(e.info & 0xffff);
        if (!(opt->mode & HSA_MODE_NONSTOP) && (e.info >> 21) > (uint32_t)(best_score + opt->s_mm)) break; /* :158 */

        m = max_diff - (e.n_mm + e.n_gapo);                                                    /* :161 */
        if (opt->mode & HSA_MODE_GAPE) m -= e.n_gape;
        if (m < 0) continue;
        if (width_seed) {
            m_seed = opt->max_seed_diff - (e.n_mm + e.n_gapo);
            if (opt->mode & HSA_MODE_GAPE) m_seed -= e.n_gape;
        }
        if (i > 0 && m < width[i - 1].bid) continue;                                           /* :172 */

        hit_found = 0;
        if (i == 0) hit_found = 1;
        else if (m == 0 && (e.state == ST_M || (opt->mode & HSA_MODE_GAPE) || e.n_gape == opt->max_gape)) { /* :180 */
            if (hsao_match_exact(ix, seq, i, &k, &l, &rev_k, &rev_l)) hit_found = 1;
            else continue;
        }

        if (hit_found) {                                                                       /* :188-241 */
            int score = score_of(e.n_mm, e.n_gapo, e.n_gape, opt);
            int do_add = 1;
            if (n_aln == 0) {
                best_score = score;
                best_diff = e.n_mm + e.n_gapo;
                if (opt->mode & HSA_MODE_GAPE) best_diff += e.n_gape;
                if (!(opt->mode & HSA_MODE_NONSTOP))
                    max_diff = (best_diff + 1 > opt->max_diff) ? opt->max_diff : best_diff + 1;
            }
            if (score == best_score) best_cnt = (int)((uint32_t)best_cnt + (l - k + 1));
            else if (best_cnt > opt->max_top2) break;
            if (e.n_gapo) {
                for (j = 0; j != n_aln; ++j) if (aln[j].k == k && aln[j].l == l) break;
                if (j < n_aln) do_add = 0;
            }
            if (do_add) {
                hsa_aln1_t *p;
                shadow(l - k + 1, N, e.last_diff_pos, width);
                if (n_aln == m_aln) {
                    m_aln <<= 1;
                    aln = (hsa_aln1_t*)realloc(aln, (size_t)m_aln * sizeof(hsa_aln1_t));
                    memset(aln + m_aln / 2, 0, (size_t)(m_aln / 2) * sizeof(hsa_aln1_t));
                }
                p = aln + n_aln;
                p->n_mm = e.n_mm; p->n_gapo = e.n_gapo; p->n_gape = e.n_gape;
                p->k = k; p->l = l; p->strand = (uint32_t)strand;
                p->rev_k = rev_k; p->rev_l = rev_l;
                p->score = score;
                ++n_aln;
            }
            continue;
        }

        --i;                                                                                   /* :244 */
        sa_backward_all_bi(ix, k, l, rev_l, sk, sl, rsk, rsl);
        occ = l - k + 1;
        allow_diff = allow_M = 1;
        if (i > 0) {                                                                           /* :252-265 */
            int ii = i - (len - opt->seed_len);
            if (width[i - 1].bid > m - 1) allow_diff = 0;
            else if (width[i - 1].bid == m - 1 && width[i].bid == m - 1 && width[i - 1].w == width[i].w) allow_M = 0;
            if (width_seed && ii > 0) {
                if (width_seed[ii - 1].bid > m_seed - 1) allow_diff = 0;
                else if (width_seed[ii - 1].bid == m_seed - 1 && width_seed[ii].bid == m_seed - 1
                         && width_seed[ii - 1].w == width_seed[ii].w) allow_M = 0;
            }
        }
        tmp = (opt->mode & HSA_MODE_LOGGAP) ? ilog2((uint32_t)(e.n_gape + e.n_gapo)) / 2 + 1 : e.n_gapo + e.n_gape;
        if (allow_diff && i >= opt->indel_end_skip + tmp && len - i >= opt->indel_end_skip + tmp) { /* :268 */
            if (e.state == ST_M) {
                if (e.n_gapo < opt->max_gapo) {
                    push(st, i, k, l, rev_k, rev_l, e.n_mm, e.n_gapo + 1, e.n_gape, ST_I, 1, opt);
                    for (j = 0; j != 4; ++j)
                        if (sk[j] <= sl[j])
                            push(st, i + 1, sk[j], sl[j], rsk[j], rsl[j], e.n_mm, e.n_gapo + 1, e.n_gape, ST_D, 1, opt);
                }
            } else if (e.state == ST_I) {
                if (e.n_gape < opt->max_gape)
                    push(st, i, k, l, rev_k, rev_l, e.n_mm, e.n_gapo, e.n_gape + 1, ST_I, 1, opt);
            } else if (e.state == ST_D) {
                if (e.n_gape < opt->max_gape) {
                    if (e.n_gape + e.n_gapo < max_diff || occ < (uint32_t)opt->max_del_occ) {
                        for (j = 0; j != 4; ++j)
                            if (sk[j] <= sl[j])
                                push(st, i + 1, sk[j], sl[j], rsk[j], rsl[j], e.n_mm, e.n_gapo, e.n_gape + 1, ST_D, 1, opt);
                    }
                }
            }
        }
        if (allow_diff && allow_M) {                                                           /* :302-314 */
            for (j = 1; j <= 4; ++j) {
                int c = (seq[i] + j) & 3;
                int is_mm = (j != 4 || seq[i] > 3);
                if (sk[c] <= sl[c])
                    push(st, i, sk[c], sl[c], rsk[c], rsl[c], e.n_mm + is_mm, e.n_gapo, e.n_gape, ST_M, is_mm, opt);
            }
        } else if (seq[i] < 4) {                                                               /* :315-325 */
            int c = seq[i] & 3;
            if (sk[c] <= sl[c])
                push(st, i, sk[c], sl[c], rsk[c], rsl[c], e.n_mm, e.n_gapo, e.n_gape, ST_M, 0, opt);
        }
    }
    pstack_free(st);
    *_n_aln = n_aln;
    if (n_entries_peak) *n_entries_peak = peak;
    return aln;
}

/* ------------------------------------------------------------------ batch helpers (mirror ref_harness.c) */

typedef struct { uint32_t *w; size_t n, cap; } dump_t;

static void dump_put(dump_t *d, int n_aln, const hsa_aln1_t *aln)
{
    int j;
    if (d->n + (size_t)n_aln * 12 > d->cap) {
        d->cap = (d->n + (size_t)n_aln * 12) * 2 + 1024;
        d->w = (uint32_t*)realloc(d->w, d->cap * 4);
    }
    for (j = 0; j < n_aln; ++j) {
        const hsa_aln1_t *p = aln + j;
        uint32_t *w = d->w + d->n;
        w[0] = p->n_mm; w[1] = p->n_gapo; w[2] = p->n_gape; w[3] = p->k; w[4] = p->l;
        w[5] = p->rev_k; w[6] = p->rev_l; w[7] = p->type; w[8] = p->strand;
        w[9] = (uint32_t)p->start; w[10] = (uint32_t)p->end; w[11] = (uint32_t)p->score;
        d->n += 12;
    }
}

/* per-read option resolution of the driver for a read before any local_opt switch (bwtaln.c:260-261, 330-332) */
static void resolve_read_opt(hsa_gap_opt_t *dst, const hsa_gap_opt_t *src, int len, int clear_gape)
{
    *dst = *src;
    if (clear_gape) dst->mode &= ~HSA_MODE_GAPE;
    if (src->fnr > 0.0) dst->max_diff = hsao_cal_maxdiff(len, 0.02, src->fnr);
    dst->seed_len = src->seed_len < len ? src->seed_len : 0x7fffffff;
}

static size_t max_len_of(const uint32_t *len, size_t n)
{
    size_t i, m = 0;
    for (i = 0; i < n; ++i) if (len[i] > m) m = len[i];
    return m;
}

/* ref_harness.c mode_percall: both strands of every read, strand 1 first */
uint32_t *hsao_percall(const hsao_index_t *ix, const uint8_t *codes, const uint64_t *off, const uint32_t *len,
                       size_t n_reads, const hsa_gap_opt_t *opt, int clear_gape, int32_t *n_aln_out, size_t *total)
{
    dump_t d = {0, 0, 0};
    size_t r, ml = max_len_of(len, n_reads);
    hsa_width_t *wb = (hsa_width_t*)calloc(ml + 1, sizeof(hsa_width_t));
    hsa_width_t *wsd = (hsa_width_t*)calloc(ml + 1, sizeof(hsa_width_t));
    uint8_t *rc = (uint8_t*)calloc(ml + 1, 1);
    for (r = 0; r < n_reads; ++r) {
        int L = (int)len[r], s;
        const uint8_t *seq = codes + off[r];
        memcpy(rc, seq, (size_t)L); hsao_seq_revcomp(L, rc);
        for (s = 1; s >= 0; --s) {
            hsa_gap_opt_t ro; int n_aln = 0; hsa_aln1_t *aln; hsa_width_t *ws = NULL;
            const uint8_t *q = s == 0 ? seq : rc;
            resolve_read_opt(&ro, opt, L, clear_gape);
            memset(wb, 0, (ml + 1) * sizeof(hsa_width_t));
            memset(wsd, 0, (ml + 1) * sizeof(hsa_width_t));
            if (L > ro.seed_len) { hsao_cal_width(ix, ro.seed_len, q + (L - ro.seed_len), wsd, 1); ws = wsd; }
            hsao_cal_width(ix, L, q, wb, 1);
            aln = hsao_match_gap(ix, &ro, q, L, wb, ws, s, &n_aln, NULL);
            n_aln_out[2 * r + (size_t)(1 - s)] = n_aln;
            dump_put(&d, n_aln, aln);
            free(aln);
        }
    }
    free(wb); free(wsd); free(rc);
    *total = d.n / 12;
    return d.w ? d.w : (uint32_t*)calloc(1, 4);
}

/* ref_harness.c mode_whole == whole-read part of bwa_cal_sa_reg_gap, bwtaln.c:303-360, 371-372 */
uint32_t *hsao_whole(const hsao_index_t *ix, const uint8_t *codes, const uint64_t *off, const uint32_t *len,
                     size_t n_reads, const hsa_gap_opt_t *opt, int clear_gape, int32_t *n_aln_out, size_t *total)
{
    dump_t d = {0, 0, 0};
    size_t r, ml = max_len_of(len, n_reads);
    int local_max_diff = opt->fnr > 0.0 ? hsao_cal_maxdiff((int)ml, 0.02, opt->fnr) : opt->max_diff;  /* :273-274 */
    hsa_width_t *wb = (hsa_width_t*)calloc(ml + 1, sizeof(hsa_width_t));
    hsa_width_t *wsd = (hsa_width_t*)calloc(ml + 1, sizeof(hsa_width_t));
    uint8_t *rc = (uint8_t*)calloc(ml + 1, 1);
    for (r = 0; r < n_reads; ++r) {
        int L = (int)len[r], s, j, nn = 0, n_aln = 0, polya = L >= 15, polyt = L >= 15;
        const uint8_t *seq = codes + off[r];
        hsa_aln1_t *aln = NULL; hsa_gap_opt_t ro;
        for (j = 0; j < L; ++j) nn += seq[j] > 3;                                      /* :314-317 */
        for (j = 0; j < 15 && j < L; ++j) { polya &= seq[j] == 0; polyt &= seq[j] == 3; } /* :324-325 */
        if (nn > local_max_diff || polya || polyt) { n_aln_out[r] = 0; continue; }
        memcpy(rc, seq, (size_t)L); hsao_seq_revcomp(L, rc);
        resolve_read_opt(&ro, opt, L, clear_gape);
        for (s = 1; s >= 0; --s) {                                                     /* :343-359 */
            const uint8_t *q = s == 0 ? seq : rc; hsa_width_t *ws = NULL;
            if (L > ro.seed_len) { hsao_cal_width(ix, ro.seed_len, q + (L - ro.seed_len), wsd, 1); ws = wsd; }
            hsao_cal_width(ix, L, q, wb, 1);
            aln = hsao_match_gap(ix, &ro, q, L, wb, ws, s, &n_aln, NULL);
            if (n_aln) { for (j = 0; j < n_aln; ++j) aln[j].strand = (uint32_t)s; break; }
            free(aln); aln = NULL;
        }
        if (n_aln) { aln[0].start = 0; aln[0].end = L - 1; }                           /* :371-372 */
        n_aln_out[r] = n_aln;
        dump_put(&d, n_aln, aln);
        free(aln);
    }
    free(wb); free(wsd); free(rc);
    *total = d.n / 12;
    return d.w ? d.w : (uint32_t*)calloc(1, 4);
}

/* ref_harness.c mode_seeds == the six seed calls of bwt_splice_match, bwtgap.c:797-820, no early-outs;
 * hits get start/end as :816-819. */
uint32_t *hsao_seeds(const hsao_index_t *ix, const uint8_t *codes, const uint64_t *off, const uint32_t *len,
                     size_t n_reads, const hsa_gap_opt_t *opt, int32_t *n_aln_out, size_t *total)
{
    dump_t d = {0, 0, 0};
    size_t r, ml = max_len_of(len, n_reads);
    hsa_width_t *wsd = (hsa_width_t*)calloc(ml + 1, sizeof(hsa_width_t));
    uint8_t *rc = (uint8_t*)calloc(ml + 1, 1);
    for (r = 0; r < n_reads; ++r) {
        int L = (int)len[r], s, seed_len = L / 3;
        const uint8_t *seq = codes + off[r];
        memcpy(rc, seq, (size_t)L); hsao_seq_revcomp(L, rc);
        for (s = 0; s < 6; ++s) {
            hsa_gap_opt_t so = *opt;                                                   /* :769-774 */
            int strand = s / 3, seg = s % 3, n_aln = 0, j;
            int len_align = seed_len + (seg == 2 ? L % 3 : 0);
            const uint8_t *base = strand == 0 ? seq : rc;
            hsa_aln1_t *aln;
            so.mode &= ~HSA_MODE_GAPE; so.max_gapo = 0; so.max_gape = 0;
            so.max_diff = opt->max_seed_diff;
            so.seed_len = len_align;                                                   /* :802 */
            memset(wsd, 0, (ml + 1) * sizeof(hsa_width_t));
            hsao_cal_width(ix, len_align, base, wsd, 1);                               /* :807-808: read PREFIX */
            aln = hsao_match_gap(ix, &so, base + seg * seed_len, len_align, wsd, wsd, strand, &n_aln, NULL);
            for (j = 0; j < n_aln; ++j) { aln[j].start = seg * seed_len; aln[j].end = aln[j].start + len_align - 1; }
            n_aln_out[6 * r + (size_t)s] = n_aln;
            dump_put(&d, n_aln, aln);
            free(aln);
        }
    }
    free(wsd); free(rc);
    *total = d.n / 12;
    return d.w ? d.w : (uint32_t*)calloc(1, 4);
}

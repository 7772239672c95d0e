// hsa_splice.cuh -- the spliced-read fallback bwt_splice_match (bwtgap.c:748-1332) and everything under it, one read per
// CUDA thread.
//
// Compiled by nvcc for sm_100a (splice_kernel in hsa_b200.cu) and by g++ for tests/emu (CPU test-suite only).
//
// What it computes, bit-exact with the reference for every read that found nothing on either strand (bwtaln.c:362-369):
//   bwt_splice_match                     bwtgap.c:748-1332   three seeds x two strands, correlation, motif scan, extension
//   bwt_match_gap (seed / 12-mer calls)  bwtgap.c:118-331    the eager form: every child is pushed as the reference pushes it
//   bwt_backtracing_search               bwtgap.c:346-511    bounded bidirectional extension of a seed hit
//   bwt_extend_backward / _foreward      bwtgap.c:640-663
//   bwt_aln_corelate_check               bwtgap.c:669-742    pairing of two seeds' hits by text distance
//   splice_site_search_from_pos          bwtgap.c:523-594    GT-AG / GC-AG / AT-AC scan on the packed text
//   check_site_by_intron_end             bwtgap.c:602-635
//   bwt_cal_width, types 1 and 0         bwtaln.c:73-116
//   BWTAllSARangesForward_Bidirection    2BWT-Interface.c:274-304, BWTSARangeForward_Bidirection :171-206,
//   bwt_extend_exact                     2BWT-Interface.c:394-439 (incl. its forward branch, which never consumes a base)
//   BWTRetrievePositionFromSAIndex       2BWT-Interface.c:329-362 (sa_value_dev + locate_dev of hsa_core.cuh)
//
// Why one read per thread here, when the whole-read search is a phase-voted state machine: the splice path is a long
// data-dependent control flow (a switch over which seeds hit, loops over candidate sites, five kinds of sub-search that
// share one stack) executed for a fraction of a percent of the reads.  Its parallel axis is reads; its cost is bounded by
// the short extensions (<= ~35 bases).  Each thread keeps the reference's own working set -- three width arrays, six hit
// lists, one score-bucketed stack -- in a private slice of global memory (L2-resident while hot).
//
// The reference's known quirks are reproduced, not repaired (they decide results): widths of a seed are computed on the
// read PREFIX (bwtgap.c:807-808); bwt_extend_exact's forward branch never decrements *leav_len (2BWT-Interface.c:422-436),
// so it extends by one base repeatedly until the interval empties; only element 0 of the 12-mer hit list gets start / end
// (bwtgap.c:925-928, 1197-1200); occ_of_first is compared with 0x3fffffff (:932, :1203), which never matches;
// bwt_array_insert / bwt_find_split_pos_by_record are no-ops (bwt_array.c:34, 77).
#pragma once
#include "hsa_core.cuh"

namespace hsa {

enum : uint32_t { TYPE_SPLICING = 4u };                       // BWA_TYPE_SPLICING, bwtaln.h:13
enum : uint32_t { SPL_OK = 0, SPL_STACK_FULL = 1, SPL_ALN_FULL = 2, SPL_SITE_FULL = 3 };
enum : uint32_t { SPL_NIL = 0xFFFFFFFFu, SPL_POS_CAP = 128, SPL_BUCKETS = 256 };

struct SAln {                   // bwt_aln1_t (bwtaln.h:41-50), fields unpacked
    uint32_t n_mm, n_gapo, n_gape, k, l, rev_k, rev_l, type, strand;
    int32_t start, end, score;
};
struct SWidth { uint32_t w; int32_t bid; };                   // bwt_width_t
struct SEntry {                 // gap_entry_t (bwtaln.h:52-58) + the link of the bucket's LIFO list
    uint32_t info, k, l, rev_k, rev_l;
    uint32_t counts;            // n_mm | n_gapo << 8 | n_gape << 16 | state << 24
    int32_t last_diff_pos;
    uint32_t next;
};
struct SPos { uint32_t ori_pos, occ_pos, seq_id; int32_t r_aln; };   // bwt_pos_t (bwtaln.h:171-178), the fields used

struct SpliceEnv {
    DevIndex ix;
    const uint32_t *sa_value; uint32_t sa_interval;           // forward BWT's SA samples (hsa_index_attach_sa)
    const uint32_t *blocks4; uint32_t n_blocks;               // HSP::blockList rows (hsa_index_attach_blocks)
    const uint32_t *packed_dna; uint32_t dna_length;          // HSP::packedDNA, 16 symbols per word, first in the MSBs
};

struct SpliceScratch {          // one worker's slice of global memory
    SEntry *arena; uint32_t arena_cap;
    uint32_t *heads;            // SPL_BUCKETS
    SWidth *width_back, *width_fore, *width_seed, *width_tmp;   // max_len + 1 entries each (width_tmp: 13)
    SAln *lists; uint32_t aln_cap;                             // six hit lists of aln_cap entries
    uint32_t *site_pos; uint32_t site_cap;
    SPos *pos_info;             // SPL_POS_CAP
};

struct SpliceParams {
    SpliceEnv env;
    const uint8_t *codes; const uint64_t *read_off; const uint32_t *read_len;
    const DevOpt *opts; const uint32_t *opt_idx;               // per read: the gap_opt_t the driver holds in aux->opt
    const uint32_t *work_list;                                 // optional: work index -> read (re-runs)
    uint32_t n_work, max_len;
    // per-worker scratch, worker w at index w of every array group
    SEntry *arena; uint32_t arena_cap; uint32_t *heads; SWidth *widths; SAln *lists; uint32_t aln_cap;
    uint32_t *site_pos; uint32_t site_cap; SPos *pos_info;
    // outputs: n_aln[read] in 0..2, aln[read * 2 + {0,1}] as 9 hsa_aln1_t words each, status[read]
    int32_t *n_aln; uint32_t *aln; uint8_t *status;
    uint32_t *fail_list; unsigned long long *fail_count;       // reads that ran out of scratch capacity
    unsigned long long *cursor;
    unsigned long long *lookups;                               // occ lookups issued (diagnostic)
};

HSA_HD SpliceScratch splice_scratch_of(const SpliceParams &P, size_t w)
{
    SpliceScratch s;
    s.arena = P.arena + w * P.arena_cap; s.arena_cap = P.arena_cap;
    s.heads = P.heads + w * SPL_BUCKETS;
    const size_t wl = (size_t)P.max_len + 1;
    SWidth *wb = P.widths + w * (3 * wl + 16);
    s.width_back = wb; s.width_fore = wb + wl; s.width_seed = wb + 2 * wl; s.width_tmp = wb + 3 * wl;
    s.lists = P.lists + w * 6 * P.aln_cap; s.aln_cap = P.aln_cap;
    s.site_pos = P.site_pos + w * P.site_cap; s.site_cap = P.site_cap;
    s.pos_info = P.pos_info + w * SPL_POS_CAP;
    return s;
}

// ---------------------------------------------------------------------------------------------------------------------
struct Splicer {
    const SpliceEnv &E;
    SpliceScratch S;
    const uint8_t *rd; int32_t len;             // the read as given (aux->seq); aux->rc_seq is read through base()
    uint32_t fail;
    unsigned long long lookups;
    // the stack (gap_stack_t, bwtaln.h:60-68)
    uint32_t st_best, st_n, st_top, st_free;

    HSA_HD Splicer(const SpliceEnv &e, const SpliceScratch &s, const uint8_t *r, int32_t l)
        : E(e), S(s), rd(r), len(l), fail(SPL_OK), lookups(0), st_best(SPL_BUCKETS), st_n(0), st_top(0), st_free(SPL_NIL) {}

    // base p of the strand-resolved read: aux->seq (strand 0) or aux->rc_seq = seq_reverse(len, seq, 1) (bwaseqio.c:73-90)
    HSA_HD uint32_t base(uint32_t strand, int32_t p) const
    {
        if (strand) { const uint32_t c = ld_ro_u8(rd + (len - 1 - p)); return c < 4 ? 3 - c : c; }
        return ld_ro_u8(rd + p);
    }
    static HSA_HD int32_t score_of(int32_t m, int32_t o, int32_t e, const DevOpt &p) { return m * p.s_mm + o * p.s_gapo + e * p.s_gape; }

    // ---- SA range stepping --------------------------------------------------------------------------------------
    // BWTAllSARangesBackward_Bidirection (2BWT-Interface.c:235-271)
    HSA_HD_CALL void back_all(uint32_t k, uint32_t l, uint32_t rev_l, uint32_t sk[4], uint32_t sl[4], uint32_t rsk[4], uint32_t rsl[4])
    {
        uint32_t oL[4], oR[4], oc = 0;
        occ4_dev(E.ix.fwd, k, oL); occ4_dev(E.ix.fwd, l + 1, oR);
        lookups += 2;
        for (int c = 3; c >= 0; --c) {
            sk[c] = E.ix.fwd.cum[c] + oL[c] + 1;
            sl[c] = E.ix.fwd.cum[c] + oR[c];
            rsl[c] = rev_l - oc;
            rsk[c] = rsl[c] - (sl[c] - sk[c]);
            oc += oR[c] - oL[c];
        }
    }
    // BWTAllSARangesForward_Bidirection (2BWT-Interface.c:274-304): the same on rev_bwt with the roles swapped
    HSA_HD_CALL void fore_all(uint32_t l, uint32_t rev_k, uint32_t rev_l, uint32_t sk[4], uint32_t sl[4], uint32_t rsk[4], uint32_t rsl[4])
    {
        uint32_t oL[4], oR[4], oc = 0;
        occ4_dev(E.ix.rev, rev_k, oL); occ4_dev(E.ix.rev, rev_l + 1, oR);
        lookups += 2;
        for (int c = 3; c >= 0; --c) {
            rsk[c] = E.ix.fwd.cum[c] + oL[c] + 1;
            rsl[c] = E.ix.fwd.cum[c] + oR[c];
            sl[c] = l - oc;
            sk[c] = sl[c] - (rsl[c] - rsk[c]);
            oc += oR[c] - oL[c];
        }
    }

    // ---- bwt_cal_width (bwtaln.c:73-116) ----------------------------------------------------------------------------
    // `src`: first base of the n-base string inside the strand-resolved read
    HSA_HD_CALL int32_t cal_width(uint32_t strand, int32_t src, int32_t n, SWidth *width, int type)
    {
        uint32_t k = 0, l = E.ix.fwd.text_length;
        int32_t bid = 0;
        if (type == 1) {                                       // forward search on rev_bwt (:85-97)
            for (int32_t i = 0; i < n; ++i) {
                const uint32_t c = base(strand, src + i);
                if (c < 4) {                                   // BWTSARangeForeward, 2BWT-Interface.c:121-132
                    const uint32_t a = occ1_dev(E.ix.rev, k, c), b = occ1_dev(E.ix.rev, l + 1, c);
                    k = E.ix.fwd.cum[c] + a + 1; l = E.ix.fwd.cum[c] + b;
                    lookups += 2;
                }
                if (k > l || c > 3) { k = 0; l = E.ix.fwd.text_length; ++bid; }
                width[i].w = l - k + 1; width[i].bid = bid;
            }
        } else {                                               // backward search on bwt, entry 0 untouched (:99-111)
            for (int32_t i = n - 1; i > 0; --i) {
                const uint32_t c = base(strand, src + i);
                if (c < 4) {                                   // BWTSARangeBackward, 2BWT-Interface.c:107-118
                    const uint32_t a = occ1_dev(E.ix.fwd, k, c), b = occ1_dev(E.ix.fwd, l + 1, c);
                    k = E.ix.fwd.cum[c] + a + 1; l = E.ix.fwd.cum[c] + b;
                    lookups += 2;
                }
                if (k > l || c > 3) { k = 0; l = E.ix.fwd.text_length; ++bid; }
                width[i].w = l - k + 1; width[i].bid = bid;
            }
        }
        width[n].w = 0; width[n].bid = ++bid;                  // :113-114
        return bid;
    }

    // ---- the score-bucketed stack (gap_reset_stack / gap_push / gap_pop, bwtgap.c:37-92) ---------------------------
    HSA_HD_CALL void st_reset()
    {
        for (uint32_t b = 0; b < SPL_BUCKETS; ++b) S.heads[b] = SPL_NIL;
        st_best = SPL_BUCKETS; st_n = 0; st_top = 0; st_free = SPL_NIL;
    }
    HSA_HD_CALL void st_push(int32_t i, uint32_t k, uint32_t l, uint32_t rev_k, uint32_t rev_l, uint32_t n_mm, uint32_t n_gapo,
                        uint32_t n_gape, uint32_t state, int is_diff, const DevOpt &o)
    {
        const int32_t score = score_of((int32_t)n_mm, (int32_t)n_gapo, (int32_t)n_gape, o);
        uint32_t s;
        if (st_free != SPL_NIL) { s = st_free; st_free = S.arena[s].next; }
        else if (st_top < S.arena_cap) s = st_top++;
        else { if (!fail) fail = SPL_STACK_FULL; return; }
        if ((uint32_t)score >= SPL_BUCKETS) { if (!fail) fail = SPL_STACK_FULL; return; }
        SEntry e;
        e.info = (uint32_t)score << 21 | (uint32_t)i;
        e.k = k; e.l = l; e.rev_k = rev_k; e.rev_l = rev_l;
        e.counts = (n_mm & 255u) | (n_gapo & 255u) << 8 | (n_gape & 255u) << 16 | state << 24;
        e.last_diff_pos = is_diff ? i : 0;
        e.next = S.heads[score];
        S.arena[s] = e;
        S.heads[score] = s;
        ++st_n;
        if (st_best > (uint32_t)score) st_best = (uint32_t)score;
    }
    HSA_HD_CALL SEntry st_pop()
    {
        const uint32_t s = S.heads[st_best];
        const SEntry e = S.arena[s];
        S.heads[st_best] = e.next;
        S.arena[s].next = st_free; st_free = s;
        --st_n;
        if (S.heads[st_best] == SPL_NIL && st_n) {
            uint32_t b = st_best + 1;
            while (b < SPL_BUCKETS && S.heads[b] == SPL_NIL) ++b;
            st_best = b;
        } else if (st_n == 0) st_best = SPL_BUCKETS;
        return e;
    }

    // ---- bwt_match_exact (2BWT-Interface.c:365-388) over bases [src, src + n) of the strand-resolved read ----------
    HSA_HD_CALL bool match_exact(uint32_t strand, int32_t src, int32_t n, uint32_t &k, uint32_t &l, uint32_t &rev_k, uint32_t &rev_l)
    {
        uint32_t ck = k, cl = l, crk = rev_k, crl = rev_l;
        for (int32_t i = n - 1; i >= 0; --i) {
            const uint32_t c = base(strand, src + i);
            if (c > 3) return false;
            uint32_t sk[4], sl[4], rsk[4], rsl[4];
            back_all(ck, cl, crl, sk, sl, rsk, rsl);           // BWTSARangeBackward_Bidirection (:135-168): symbol c of it
            ck = sk[c]; cl = sl[c]; crl = rsl[c]; crk = rsk[c];
            if (ck > cl) break;
        }
        if (ck > cl) return false;
        if (k) k = ck;                                          // written back only where the input was non-zero (:383-386)
        if (l) l = cl;
        if (rev_k) rev_k = crk;
        if (rev_l) rev_l = crl;
        return true;
    }

    // ---- bwt_match_gap (bwtgap.c:118-331), every child pushed as the reference pushes it ---------------------------
    // seq = bases [src, src + n) of the strand-resolved read; width / width_seed as in bwt_aux_t (width_seed may be null
    // or alias width); hits go to out[0..cap) in discovery order with k, l, rev_k, rev_l, counts, strand, score set and
    // everything else zero (the reference's calloc / zero-fill, :137-138, :218-227)
    HSA_HD_CALL int32_t match_gap(uint32_t strand, int32_t src, int32_t n, SWidth *width, SWidth *width_seed, const DevOpt &o, SAln *out)
    {
        const uint32_t N = E.ix.fwd.text_length;
        int32_t best_score = score_of(o.max_diff + 1, o.max_gapo + 1, o.max_gape + 1, o);        // :128
        int32_t max_diff = o.max_diff, best_cnt = 0, n_aln = 0;
        st_reset();                                                                              // :141
        st_push(n, 0, N, 0, N, 0, 0, 0, ST_M, 0, o);                                             // :142
        while (st_n && !fail) {
            if (st_n > (uint32_t)o.max_entries) break;                                           // :150-151
            const SEntry e = st_pop();
            uint32_t k = e.k, l = e.l, rev_k = e.rev_k, rev_l = e.rev_l;
            int32_t i = (int32_t)(e.info & 0xffffu);
            const int32_t e_mm = (int32_t)(e.counts & 255u), e_go = (int32_t)((e.counts >> 8) & 255u),
                          e_ge = (int32_t)((e.counts >> 16) & 255u);
            const uint32_t e_state = e.counts >> 24;
            if (!(o.mode & MODE_NONSTOP) && (int32_t)(e.info >> 21) > best_score + o.s_mm) break;   // :158-159
            int32_t m = max_diff - (e_mm + e_go), m_seed = 0;                                    // :161-171
            if (o.mode & MODE_GAPE) m -= e_ge;
            if (m < 0) continue;
            if (width_seed) {
                m_seed = o.max_seed_diff - (e_mm + e_go);
                if (o.mode & MODE_GAPE) m_seed -= e_ge;
            }
            if (i > 0 && m < width[i - 1].bid) continue;                                         // :172-173
            bool hit = false;
            if (i == 0) hit = true;                                                              // :177-179
            else if (m == 0 && (e_state == ST_M || (o.mode & MODE_GAPE) || e_ge == o.max_gape)) { // :180-186
                if (match_exact(strand, src, i, k, l, rev_k, rev_l)) hit = true;
                else continue;
            }
            if (hit) {                                                                           // :188-241
                const int32_t score = score_of(e_mm, e_go, e_ge, o);
                bool add = true;
                if (n_aln == 0) {
                    best_score = score;
                    int32_t best_diff = e_mm + e_go;
                    if (o.mode & MODE_GAPE) best_diff += e_ge;
                    if (!(o.mode & MODE_NONSTOP)) max_diff = best_diff + 1 > o.max_diff ? o.max_diff : best_diff + 1;
                }
                if (score == best_score) best_cnt = (int32_t)((uint32_t)best_cnt + (l - k + 1));
                else if (best_cnt > o.max_top2) break;
                if (e_go)
                    for (int32_t j = 0; j < n_aln; ++j) if (out[j].k == k && out[j].l == l) { add = false; break; }
                if (add) {
                    const uint32_t x = l - k + 1;                                                // gap_shadow, :94-105
                    for (int32_t q = 0, jj = 0; q < e.last_diff_pos; ++q) {
                        if (width[q].w > x) width[q].w -= x;
                        else if (width[q].w == x) { width[q].bid = 1; width[q].w = N - (uint32_t)(++jj); }
                    }
                    if ((uint32_t)n_aln >= S.aln_cap) { if (!fail) fail = SPL_ALN_FULL; break; }
                    SAln a;
                    a.n_mm = (uint32_t)e_mm; a.n_gapo = (uint32_t)e_go; a.n_gape = (uint32_t)e_ge;
                    a.k = k; a.l = l; a.rev_k = rev_k; a.rev_l = rev_l; a.type = 0; a.strand = strand;
                    a.start = 0; a.end = 0; a.score = score;
                    out[n_aln++] = a;
                }
                continue;
            }
            --i;                                                                                 // :244
            uint32_t sk[4], sl[4], rsk[4], rsl[4];
            back_all(k, l, rev_l, sk, sl, rsk, rsl);
            const uint32_t occ = l - k + 1;
            bool allow_diff = true, allow_M = true;
            if (i > 0) {                                                                         // :252-265
                const int32_t ii = i - (n - o.seed_len);
                if (width[i - 1].bid > m - 1) allow_diff = false;
                else if (width[i - 1].bid == m - 1 && width[i].bid == m - 1 && width[i - 1].w == width[i].w) allow_M = false;
                if (width_seed && ii > 0) {
                    if (width_seed[ii - 1].bid > m_seed - 1) allow_diff = false;
                    else if (width_seed[ii - 1].bid == m_seed - 1 && width_seed[ii].bid == m_seed - 1 &&
                             width_seed[ii - 1].w == width_seed[ii].w) allow_M = false;
                }
            }
            int32_t tmp;                                                                         // :267
            if (o.mode & MODE_LOGGAP) { uint32_t v = (uint32_t)(e_ge + e_go); int32_t lg = 0; while (v > 1) { v >>= 1; ++lg; } tmp = lg / 2 + 1; }
            else tmp = e_go + e_ge;
            if (allow_diff && i >= o.indel_end_skip + tmp && n - i >= o.indel_end_skip + tmp) {  // :268-300
                if (e_state == ST_M) {
                    if (e_go < o.max_gapo) {
                        st_push(i, k, l, rev_k, rev_l, e_mm, e_go + 1, e_ge, ST_I, 1, o);
                        for (int j = 0; j != 4; ++j)
                            if (sk[j] <= sl[j]) st_push(i + 1, sk[j], sl[j], rsk[j], rsl[j], e_mm, e_go + 1, e_ge, ST_D, 1, o);
                    }
                } else if (e_state == ST_I) {
                    if (e_ge < o.max_gape) st_push(i, k, l, rev_k, rev_l, e_mm, e_go, e_ge + 1, ST_I, 1, o);
                } else if (e_ge < o.max_gape && (e_ge + e_go < max_diff || occ < (uint32_t)o.max_del_occ)) {
                    for (int j = 0; j != 4; ++j)
                        if (sk[j] <= sl[j]) st_push(i + 1, sk[j], sl[j], rsk[j], rsl[j], e_mm, e_go, e_ge + 1, ST_D, 1, o);
                }
            }
            const uint32_t sc = base(strand, src + i);
            if (allow_diff && allow_M) {                                                         // :302-314
                for (uint32_t j = 1; j <= 4; ++j) {
                    const uint32_t c = (sc + j) & 3u;
                    const int is_mm = (j != 4 || sc > 3);
                    if (sk[c] <= sl[c]) st_push(i, sk[c], sl[c], rsk[c], rsl[c], e_mm + is_mm, e_go, e_ge, ST_M, is_mm, o);
                }
            } else if (sc < 4) {                                                                 // :315-325
                if (sk[sc] <= sl[sc]) st_push(i, sk[sc], sl[sc], rsk[sc], rsl[sc], e_mm, e_go, e_ge, ST_M, 0, o);
            }
        }
        return n_aln;
    }

    // ---- bwt_extend_exact (2BWT-Interface.c:394-439) on the strand-resolved read ------------------------------------
    HSA_HD_CALL void extend_exact(uint32_t strand, int32_t start, int32_t &leav, int type, uint32_t &k, uint32_t &l, uint32_t &rev_k, uint32_t &rev_l)
    {
        uint32_t ck = k, cl = l, crk = rev_k, crl = rev_l;
        uint32_t sk[4], sl[4], rsk[4], rsl[4];
        if (type == 1) {
            start -= leav;
            while (leav != 0) {
                const int32_t pos = start + leav;
                const uint32_t c = base(strand, pos);
                if (c > 3) break;
                back_all(ck, cl, crl, sk, sl, rsk, rsl);
                ck = sk[c]; cl = sl[c]; crk = rsk[c]; crl = rsl[c];
                if (ck > cl) break;
                k = ck; l = cl; rev_k = crk; rev_l = crl;
                --leav;
            }
        } else {
            start += leav;
            while (leav != 0) {
                // the reference never decrements *leav_len here (:422-436): the SAME base is appended until the interval
                // empties, and the caller's remaining length stays as it was
                const int32_t pos = start - leav;
                const uint32_t c = base(strand, pos);
                if (c > 3) break;
                fore_all(cl, crk, crl, sk, sl, rsk, rsl);      // BWTSARangeForward_Bidirection (:171-206): symbol c of it
                ck = sk[c]; cl = sl[c]; crk = rsk[c]; crl = rsl[c];
                if (ck > cl) break;
                k = ck; l = cl; rev_k = crk; rev_l = crl;
            }
        }
    }

    // ---- bwt_backtracing_search (bwtgap.c:346-511) -------------------------------------------------------------------
    // ext_len = aux->len of the extension frame; returns -1 / 1 / 2 as the reference does
    HSA_HD_CALL int32_t backtrack(uint32_t strand, int32_t ext_len, const DevOpt &o, SAln &aln, int is_backward, int32_t &max_pos_io)
    {
        const int32_t best_score = score_of(o.max_diff + 1, o.max_gapo + 1, o.max_gape + 1, o);
        const int32_t max_diff = o.max_diff, ln = ext_len;
        const int32_t start = aln.start, end = aln.end;
        int32_t max_pos = max_pos_io;
        const SWidth *width = is_backward == 1 ? S.width_back : S.width_fore;
        while (st_n != 0 && !fail) {
            if (st_n > (uint32_t)o.max_entries) break;
            const SEntry e = st_pop();
            uint32_t k = e.k, l = e.l, rev_k = e.rev_k, rev_l = e.rev_l;
            int32_t i = (int32_t)(e.info & 0xffffu);
            const int32_t e_mm = (int32_t)(e.counts & 255u), e_go = (int32_t)((e.counts >> 8) & 255u),
                          e_ge = (int32_t)((e.counts >> 16) & 255u);
            const uint32_t e_state = e.counts >> 24;
            if (!(o.mode & MODE_NONSTOP) && (int32_t)(e.info >> 21) > best_score + o.s_mm) break;   // :381-382
            int32_t m = max_diff - (e_mm + e_go);                                                // :384-387
            if (o.mode & MODE_GAPE) m -= e_ge;
            if (m <= 0 || i == 0) {                                                              // :388-420
                if (m == 0 && i != 0)
                    extend_exact(strand, is_backward == 0 ? end + ln - i + 1 : start - ln + i - 1, i, is_backward, k, l, rev_k, rev_l);
                if (is_backward == 1 && max_pos >= start + i - ln && aln.start > start + i - ln) {
                    aln.start = start + i - ln; max_pos = aln.start;
                } else if (is_backward == 0 && max_pos <= end + ln - i && aln.end < end + ln - i) {
                    aln.end = end + ln - i; max_pos = aln.end;
                } else continue;
                aln.k = k; aln.l = l; aln.type = TYPE_SPLICING; aln.rev_k = rev_k; aln.rev_l = rev_l;
                aln.n_mm = (uint32_t)e_mm; aln.n_gapo = (uint32_t)e_go; aln.n_gape = (uint32_t)e_ge;
                aln.score = (int32_t)(e.info >> 21);
                if (i == 0) { max_pos_io = max_pos; return 1; }
                continue;
            }
            --i;                                                                                 // :422
            const int32_t real_pos = is_backward == 1 ? start - ln + i : ln + end - i;
            uint32_t sk[4], sl[4], rsk[4], rsl[4];
            if (is_backward == 1) back_all(k, l, rev_l, sk, sl, rsk, rsl);
            else fore_all(l, rev_k, rev_l, sk, sl, rsk, rsl);
            const uint32_t occ = l - k + 1;
            bool allow_diff = true;                                                              // :435-447
            if (is_backward == 1 && max_pos < real_pos &&
                (width[real_pos].bid - width[max_pos].bid > m ||
                 (width[real_pos].bid - width[max_pos].bid == m && width[max_pos].bid != width[max_pos + 1].bid))) allow_diff = false;
            if (is_backward == 0 && max_pos > real_pos &&
                (width[real_pos].bid - width[max_pos].bid > m ||
                 (width[real_pos].bid - width[max_pos].bid == m && width[max_pos].bid != width[max_pos - 1].bid))) allow_diff = false;
            int32_t tmp;                                                                         // :451
            if (o.mode & MODE_LOGGAP) { uint32_t v = (uint32_t)(e_ge + e_go); int32_t lg = 0; while (v > 1) { v >>= 1; ++lg; } tmp = lg / 2 + 1; }
            else tmp = e_go + e_ge;
            if (allow_diff && i >= o.indel_end_skip + tmp && ln - i >= o.indel_end_skip + tmp) { // :452-489
                if (e_state == ST_M) {
                    if (e_go < o.max_gapo) {
                        st_push(i, k, l, rev_k, rev_l, e_mm, e_go + 1, e_ge, ST_I, 1, o);
                        for (int j = 0; j != 4; ++j)
                            if ((is_backward == 1 && sk[j] <= sl[j]) || (is_backward == 0 && rsk[j] <= rsl[j]))
                                st_push(i + 1, sk[j], sl[j], rsk[j], rsl[j], e_mm, e_go + 1, e_ge, ST_D, 1, o);
                    }
                } else if (e_state == ST_I) {
                    if (e_ge < o.max_gape) st_push(i, k, l, rev_k, rev_l, e_mm, e_go, e_ge + 1, ST_I, 1, o);
                } else if (e_ge < o.max_gape && (e_ge + e_go < max_diff || occ < (uint32_t)o.max_del_occ)) {
                    for (int j = 0; j != 4; ++j)
                        if (sk[j] <= sl[j]) st_push(i + 1, sk[j], sl[j], rsk[j], rsl[j], e_mm, e_go, e_ge + 1, ST_D, 1, o);
                }
            }
            if (allow_diff) {                                                                    // :491-503
                const uint32_t sc = base(strand, real_pos);
                for (uint32_t j = 1; j <= 4; ++j) {
                    const uint32_t c = (sc + j) & 3u;
                    const int is_mm = (j != 4 || sc > 3);
                    if ((is_backward == 1 && sk[c] <= sl[c]) || (is_backward == 0 && rsk[c] <= rsl[c]))
                        st_push(i, sk[c], sl[c], rsk[c], rsl[c], e_mm + is_mm, e_go, e_ge, ST_M, is_mm, o);
                }
            }
        }
        if (max_pos_io != max_pos) { max_pos_io = max_pos; return 2; }                           // :506-510
        return -1;
    }
    // bwt_extend_backward / bwt_extend_foreward (bwtgap.c:640-663)
    HSA_HD_CALL int32_t extend(uint32_t strand, int32_t ext_len, const DevOpt &o, SAln &aln, int is_backward, int32_t &pos)
    {
        st_reset();
        st_push(ext_len, aln.k, aln.l, aln.rev_k, aln.rev_l, aln.n_mm, aln.n_gapo, aln.n_gape, 0, 0, o);
        return backtrack(strand, ext_len, o, aln, is_backward, pos);
    }

    // ---- SA index -> text position with the block search (BWTRetrievePositionFromSAIndex, 2BWT-Interface.c:329-362) --
    // seq_id / ori_pos keep their previous values when no block holds the position, as in the reference
    HSA_HD uint32_t retrieve(uint32_t sa_index, uint32_t &seq_id, uint32_t &ori_pos)
    {
        uint32_t steps;
        const uint32_t occ_pos = sa_value_dev(E.ix.fwd, E.sa_value, E.sa_interval, sa_index, steps);
        lookups += steps;
        locate_dev(E.blocks4, E.n_blocks, occ_pos, seq_id, ori_pos);
        return occ_pos;
    }
    HSA_HD uint32_t dna_at(uint32_t k) const { return (ld_ro1(E.packed_dna + (k >> 4)) >> ((~k & 15u) << 1)) & 3u; }

    // ---- bwt_aln_corelate_check (bwtgap.c:669-742) on hit lists p, q (n_p, n_q updated) ------------------------------
    HSA_HD_CALL uint32_t corelate(SAln *p, int32_t &n_p, SAln *q, int32_t &n_q)
    {
        int32_t tot_cnt = 0, cur = 0;
        uint32_t min_dist = 0xffffffffu, res_pos = 0xffffffffu, seq_id = 0, ori_pos = 0;
        for (int32_t i = 0; i < n_p && i < 10; ++i) {                                            // :684-708
            const SAln &a = p[i];
            tot_cnt += a.l - a.k + 1 > 10u ? 10 : (int32_t)(a.l - a.k + 1);
            for (uint32_t j = a.k; j <= a.l && j < a.k + 10u; ++j) {
                SPos t;
                t.occ_pos = retrieve(j, seq_id, ori_pos);
                t.ori_pos = ori_pos;
                t.seq_id = seq_id & 0x7fffffffu; t.r_aln = i;
                if (cur < (int32_t)SPL_POS_CAP) S.pos_info[cur] = t;
                ++cur;
            }
        }
        int32_t t_r1 = 0, t_r2 = 0;
        for (int32_t i = 0; i < n_q; ++i) {                                                      // :711-729
            const SAln &a = q[i];
            for (uint32_t j = a.k; j <= a.l && j < a.k + 50u; ++j) {
                const uint32_t occ_pos = retrieve(j, seq_id, ori_pos);
                for (int32_t t = 0; t < tot_cnt && t < (int32_t)SPL_POS_CAP; ++t) {
                    const int32_t dist = (int32_t)(occ_pos - S.pos_info[t].occ_pos);
                    if (S.pos_info[t].seq_id == seq_id && dist > 50 && dist < 50000 && min_dist > (uint32_t)dist) {
                        min_dist = (uint32_t)dist;
                        res_pos = S.pos_info[t].occ_pos;
                        t_r1 = S.pos_info[t].r_aln; t_r2 = i;
                    }
                }
            }
        }
        if (min_dist != 0xffffffffu) {                                                           // :730-737
            const SAln a = p[t_r1], b = q[t_r2];
            p[0] = a; n_p = 1;
            q[0] = b; n_q = 1;
        }
        return res_pos;
    }

    // ---- splice_site_search_from_pos (bwtgap.c:523-594): sites into S.site_pos, returns their number ------------------
    HSA_HD_CALL int32_t site_search(int is_backward, uint32_t strand, uint32_t pos, int32_t ext, int32_t left, int32_t right)
    {
        const int32_t ref_len = right - left - 1;
        // motif tables of :535-536 (positive strand GT AG | GC AG | AT AC, negative strand CT AC | CT GC | GT AT)
        const uint8_t motif_posv[12] = {2, 3, 0, 2, 2, 1, 0, 2, 0, 3, 0, 1};
        const uint8_t motif_neg[12] = {1, 3, 0, 1, 1, 3, 2, 1, 2, 3, 0, 3};
        if (is_backward == 1) pos -= (uint32_t)ref_len; else pos += (uint32_t)(left + 1 + ext);
        // ref_seq[j] = packed text at pos + j while pos + j < dnaLength (computed in 32 bits like the reference's
        // bwtint_t loop, :552-553), zero beyond (calloc)
        const uint32_t lim = pos + (uint32_t)ref_len;
        auto ref_at = [&](int32_t j) -> uint32_t {
            const uint32_t k = pos + (uint32_t)j;
            return (k >= pos && k < lim && k < E.dna_length) ? dna_at(k) : 0u;
        };
        int32_t n_site = 0;
        for (int32_t i = 0; i < 3; ++i) {
            const uint8_t *motif = (strand == 0 ? motif_posv : motif_neg) + (is_backward == 0 ? 4 * i : 4 * i + 2);
            for (int32_t j = 2; j < ref_len - 1; ++j) {
                const int32_t ref_pos = is_backward == 0 ? j : ref_len - 2 - j;
                const int32_t seq_pos = is_backward == 0 ? left + j : right - j;
                if (ref_at(ref_pos) == motif[0] && ref_at(ref_pos + 1) == motif[1]) {
                    if ((is_backward == 0 && ref_at(ref_pos - 1) == base(strand, seq_pos)) ||
                        (is_backward == 1 && ref_at(ref_pos + 2) == base(strand, seq_pos))) {
                        if ((uint32_t)n_site >= S.site_cap) { if (!fail) fail = SPL_SITE_FULL; return n_site; }
                        S.site_pos[n_site++] = (uint32_t)i << 30 | (uint32_t)seq_pos;
                    }
                }
            }
            // the first two motifs share their fore part (:586-588)
            if ((is_backward == 0 && strand == 1 && i == 0) || (is_backward == 1 && strand == 0 && i == 0)) i += 1;
        }
        return n_site;
    }

    // ---- check_site_by_intron_end (bwtgap.c:602-635) -----------------------------------------------------------------
    HSA_HD_CALL int32_t check_site(const SAln &aln, int is_backward, int32_t type, uint32_t strand)
    {
        const uint8_t motif_posv[12] = {2, 3, 0, 2, 2, 1, 0, 2, 0, 3, 0, 1};
        const uint8_t motif_neg[12] = {1, 3, 0, 1, 1, 3, 2, 1, 2, 3, 0, 3};
        const uint8_t *motif = strand == 0 ? motif_posv : motif_neg;
        uint32_t seq_id = 0, ori_pos = 0;
        for (uint32_t m = aln.k; m <= aln.l; ++m) {
            const uint32_t occ_pos = retrieve(m, seq_id, ori_pos);
            uint32_t k = is_backward == 1 ? occ_pos - 2u : occ_pos + (uint32_t)aln.end + 1u;
            // (positions beyond the packed text read whatever follows the array in the reference; zero here)
            const uint32_t r0 = k < E.dna_length ? dna_at(k) : 0u;
            k += 1;
            const uint32_t r1 = k < E.dna_length ? dna_at(k) : 0u;
            const uint8_t *tm = is_backward == 0 ? motif + type * 4 : motif + type * 4 + 2;
            if (((r0 ^ tm[0]) | (r1 ^ tm[1])) == 0) return type;
            else if (type == 0 && ((is_backward == 1 && strand == 1) || (is_backward == 0 && strand == 0))) {
                type = 1;
                tm = is_backward == 0 ? motif + type * 4 : motif + type * 4 + 2;
                if (((r0 ^ tm[0]) | (r1 ^ tm[1])) == 0) return type;
            }
            if (m == 0xffffffffu) break;
        }
        return 3;
    }

    // ---- bwt_splice_match (bwtgap.c:748-1332) ------------------------------------------------------------------------
    // opt = aux->opt as the driver holds it; res[0..2) = the reference's res_aln; returns *_n_aln
    HSA_HD_CALL int32_t splice_match(const DevOpt &opt, SAln res[2])
    {
        const int32_t seed_len = len / 3;
        DevOpt so = opt, xo = opt;                              // aux_seed->opt / aux_ext->opt (:768-783)
        so.mode &= ~(int)MODE_GAPE; so.max_gapo = 0; so.max_gape = 0; so.max_diff = opt.max_seed_diff;
        xo.max_gape = 3;
        SAln zero;
        zero.n_mm = zero.n_gapo = zero.n_gape = zero.k = zero.l = zero.rev_k = zero.rev_l = zero.type = zero.strand = 0;
        zero.start = zero.end = zero.score = 0;
        res[0] = zero; res[1] = zero;
        SAln *L[6]; int32_t nL[6];
        for (int i = 0; i < 6; ++i) { L[i] = S.lists + (size_t)i * S.aln_cap; nL[i] = 0; }
        uint32_t strand = 0, seq_pos = 0;
        int32_t hit_seeds = 0;                                  // _j
        uint32_t seg_mtype = 0;
        for (int32_t i = 0; i < 6 && !fail; ++i) {                                               // :797-848
            const uint32_t st = (uint32_t)(i / 3);
            const int32_t la = seed_len + (i % 3 == 2 ? len % 3 : 0);
            so.seed_len = la;
            for (int32_t q = 0; q <= la; ++q) { S.width_seed[q].w = 0; S.width_seed[q].bid = 0; }
            cal_width(st, 0, la, S.width_seed, 1);             // on the read PREFIX (:807-808)
            nL[i] = match_gap(st, (i % 3) * seed_len, la, S.width_seed, S.width_seed, so, L[i]);
            if (nL[i] != 0) {
                ++hit_seeds; seg_mtype += 1u << (i % 3);
                for (int32_t j = 0; j < nL[i]; ++j) { L[i][j].start = (i % 3) * seed_len; L[i][j].end = L[i][j].start + la - 1; }
            }
            if (i == 1 && hit_seeds == 0) { i += 1; hit_seeds = 0; seg_mtype = 0; continue; }   // :821-824
            if ((i == 2 || i == 5) && hit_seeds > 1) {                                           // :825-843
                strand = i < 3 ? 0u : 1u;
                const int b = (int)strand * 3;
                switch (seg_mtype) {
                case 5: case 7: seq_pos = corelate(L[b], nL[b], L[b + 2], nL[b + 2]); break;
                case 3: seq_pos = corelate(L[b], nL[b], L[b + 1], nL[b + 1]); break;
                case 6: seq_pos = corelate(L[b + 1], nL[b + 1], L[b + 2], nL[b + 2]); break;
                }
                if (seq_pos != 0xffffffffu) break;
            }
            if (i == 2) { hit_seeds = 0; seg_mtype = 0; }
        }
        int32_t n_out = 0;
        if (fail || hit_seeds < 2 || seq_pos == 0xffffffffu) return 0;                           // :857-858

        const int qb = (int)strand * 3;
        SAln *Q0 = L[qb], *Q1 = L[qb + 1], *Q2 = L[qb + 2];
        int32_t &n0 = nL[qb], &n2 = nL[qb + 2];
        cal_width(strand, 0, len, S.width_back, 1);                                              // :865-873
        cal_width(strand, 0, len, S.width_fore, 0);
        int32_t max_pos, b_ext, motif_type, motif_checked, site_num = 0;
        bool mapped = false;                                     // b_mapping_stat

        if (seg_mtype == 3) {                                                                    // :882-1010
            site_num = site_search(0, strand, seq_pos, (int32_t)(Q0[0].n_gapo + Q0[0].n_gape), Q1[0].end, len);
            max_pos = Q1[0].end;
            b_ext = extend(strand, max_pos - Q0[0].end, xo, Q0[0], 0, max_pos);
            if (b_ext != 1 || fail) return 0;
            motif_type = (int32_t)((site_num ? S.site_pos[0] : 0u) >> 30);
            res[0] = Q0[0];
            {   // the last 12 bases of the read (:911-930)
                cal_width(strand, len - 12, 12, S.width_tmp, 1);
                n2 = match_gap(strand, len - 12, 12, S.width_tmp, nullptr, xo, Q2);
                if (n2 == 0 || fail) return 0;
                Q2[0].start = len - 12; Q2[0].end = len - 1;
            }
            corelate(Q0, n0, Q2, n2);                           // occ_of_first == 0x3fffffff never holds (:931-933)
            res[1] = Q2[0];
            for (int32_t m = 0; m < site_num && !fail; ++m) {                                    // :935-984
                max_pos = (int32_t)(S.site_pos[m] & 0x3fffffffu);
                if (motif_type != (int32_t)(S.site_pos[m] >> 30)) { res[0] = Q0[0]; motif_type = (int32_t)(S.site_pos[m] >> 30); }
                b_ext = extend(strand, max_pos - Q0[0].end, xo, Q0[0], 0, max_pos);
                if (b_ext != 1) {
                    while (m + 1 < site_num && (int32_t)(S.site_pos[m + 1] >> 30) == motif_type) m++;
                    continue;
                }
                max_pos += 1;
                const int32_t ext_len = res[1].start - max_pos;
                if (ext_len > 0) {
                    Q2[0] = res[1];
                    b_ext = extend(strand, ext_len, xo, Q2[0], 1, max_pos);
                    if (b_ext != 1) continue;
                    seq_pos = corelate(Q0, n0, Q2, n2);
                    if (seq_pos == 0xffffffffu) continue;
                    motif_checked = check_site(Q2[0], 1, motif_type, strand);
                    if (motif_checked != 3) { n_out = 2; mapped = true; break; }
                    continue;
                } else { n_out = 1; mapped = true; continue; }
            }
            if (!mapped && !fail) {                                                              // :990-1012
                Q2[0] = res[1];
                const int32_t ext_len = Q2[0].start - Q0[0].end - 1;
                max_pos = Q1[0].end + 1;
                b_ext = extend(strand, ext_len, xo, Q0[0], 0, max_pos);
                n_out = 1;
                if (b_ext == 2) {
                    max_pos += 1;
                    const int32_t el2 = Q2[0].start - max_pos;
                    if (el2 > 0) {
                        b_ext = extend(strand, el2, xo, Q2[0], 1, max_pos);
                        if (b_ext == 1) {
                            Q2[0].start = max_pos;
                            seq_pos = corelate(Q0, n0, Q2, n2);
                            if (seq_pos != 0xffffffffu) n_out = 2;
                        }
                    }
                }
            }
            if (n_out != 0) { res[0] = Q0[0]; res[0].type = TYPE_SPLICING; }                     // :1014-1022
            if (n_out == 2) { res[1] = Q2[0]; res[1].type = TYPE_SPLICING; }
        } else if (seg_mtype == 5 || seg_mtype == 7) {                                           // :1023-1139
            site_num = site_search(0, strand, seq_pos, (int32_t)(Q0[0].n_gapo + Q0[0].n_gape), Q0[0].end, Q2[0].start);
            xo.mode &= ~(int)MODE_GAPE;
            motif_type = (int32_t)((site_num ? S.site_pos[0] : 0u) >> 30);
            res[0] = Q0[0]; res[1] = Q2[0];
            for (int32_t m = 0; m < site_num && !fail; ++m) {                                    // :1055-1098
                max_pos = (int32_t)(S.site_pos[m] & 0x3fffffffu);
                if (motif_type != (int32_t)(S.site_pos[m] >> 30)) { Q0[0] = res[0]; motif_type = (int32_t)(S.site_pos[m] >> 30); }
                Q2[0] = res[1];
                b_ext = extend(strand, max_pos - Q0[0].end, xo, Q0[0], 0, max_pos);
                if (b_ext != 1) {
                    while (m + 1 < site_num && (int32_t)(S.site_pos[m + 1] >> 30) == motif_type) m++;
                    continue;
                }
                max_pos = max_pos + 1;
                b_ext = extend(strand, Q2[0].start - max_pos, xo, Q2[0], 1, max_pos);
                if (b_ext != 1) continue;
                motif_checked = check_site(Q2[0], 1, motif_type, strand);
                if (motif_checked != 3) { n_out = 2; mapped = true; break; }
            }
            if (!mapped && !fail) {                                                              // :1111-1138
                xo.mode &= ~(int)MODE_GAPE;
                max_pos = Q0[0].end + 1;
                b_ext = extend(strand, Q2[0].start - Q0[0].end - 1, xo, Q0[0], 0, max_pos);
                if (b_ext == -1) n_out = 0;
                else if (b_ext == 1) n_out = 2;
                else {
                    max_pos = Q0[0].end + 1;
                    xo.mode |= (int)MODE_GAPE;
                    b_ext = extend(strand, Q2[0].start - max_pos, xo, Q2[0], 1, max_pos);
                    seq_pos = corelate(Q0, n0, Q2, n2);
                    n_out = (b_ext == -1 || seq_pos == 0xffffffffu) ? 0 : 2;
                }
            }
            if (n_out != 0) {                                                                    // :1140-1146
                res[0] = Q0[0]; res[1] = Q2[0];
                res[0].type = TYPE_SPLICING; res[1].type = TYPE_SPLICING;
            }
        } else if (seg_mtype == 6) {                                                             // :1147-1301
            site_num = site_search(1, strand, seq_pos - (uint32_t)seed_len, 0, 0, Q1[0].start);
            {
                max_pos = Q1[0].start;
                b_ext = extend(strand, Q2[0].start - Q1[0].start, xo, Q2[0], 1, max_pos);
                if (b_ext != 1 || fail) return 0;
            }
            // the first 12 bases with the whole read's width_back (:1183-1201); gap_shadow rewrites it for what follows
            n0 = match_gap(strand, 0, 12, S.width_back, nullptr, xo, Q0);
            if (n0 == 0 || fail) return 0;
            Q0[0].start = 0; Q0[0].end = 11;
            corelate(Q0, n0, Q2, n2);
            res[0] = Q0[0]; res[1] = Q2[0];
            motif_type = (int32_t)((site_num ? S.site_pos[0] : 0u) >> 30);
            for (int32_t m = 0; m < site_num && !fail; ++m) {                                    // :1208-1263
                max_pos = (int32_t)(S.site_pos[m] & 0x3fffffffu);
                const int32_t ext_len = Q2[0].start - max_pos;
                if (motif_type != (int32_t)(S.site_pos[m] >> 30)) { Q2[0] = res[1]; motif_type = (int32_t)(S.site_pos[m] >> 30); }
                b_ext = extend(strand, ext_len, xo, Q2[0], 1, max_pos);
                if (b_ext != 1) {
                    while (m + 1 < site_num && (int32_t)(S.site_pos[m + 1] >> 30) == motif_type) m++;
                    continue;
                }
                max_pos -= 1;
                const int32_t el2 = max_pos - Q0[0].end;
                if (el2 > 0) {
                    Q0[0] = res[0];
                    b_ext = extend(strand, el2, xo, Q0[0], 0, max_pos);
                    if (b_ext != 1) continue;
                    seq_pos = corelate(Q0, n0, Q2, n2);
                    if (seq_pos == 0xffffffffu) continue;
                    motif_checked = check_site(Q0[0], 0, motif_type, strand);
                    if (motif_checked != 3) { mapped = true; n_out = 2; break; }
                    continue;
                } else { n_out = 1; mapped = true; break; }
            }
            if (!mapped && !fail) {                                                              // :1269-1298
                const int32_t ext_len = Q2[0].start - Q0[0].end - 1;
                max_pos = Q2[0].start - 1;
                b_ext = extend(strand, ext_len, xo, Q2[0], 1, max_pos);
                if (b_ext == 2) n_out = 1;
                max_pos -= 1;
                const int32_t el2 = max_pos - Q0[0].end;
                if (el2 > 0) {
                    b_ext = extend(strand, el2, xo, Q0[0], 0, max_pos);
                    if (b_ext == 1) {
                        seq_pos = corelate(Q0, n0, Q2, n2);
                        if (seq_pos != 0xffffffffu) n_out = 2;
                    }
                }
            }
            if (n_out == 2) {                                                                    // :1299-1309
                res[0] = Q0[0]; res[1] = Q2[0];
                res[0].type = TYPE_SPLICING; res[1].type = TYPE_SPLICING;
            } else if (n_out > 0) { res[0] = Q2[0]; res[0].type = TYPE_SPLICING; }
        }
        return fail ? 0 : n_out;
    }
};

// bwt_aln1_t words of one result entry (bit-fields packed as gcc packs them, bwtaln.h:41-50)
HSA_HD void splice_store(uint32_t *w, const SAln &a)
{
    w[0] = (a.n_mm & 0xFFFFu) | (a.n_gapo & 0xFFu) << 16 | (a.n_gape & 0xFFu) << 24;
    w[1] = a.k; w[2] = a.l; w[3] = a.rev_k; w[4] = a.rev_l;
    w[5] = (a.type & 0x3FFFFFFFu) | (a.strand & 3u) << 30;
    w[6] = (uint32_t)a.start; w[7] = (uint32_t)a.end; w[8] = (uint32_t)a.score;
}

// one read of the batch on worker `worker`
HSA_HD void splice_item(const SpliceParams &P, uint32_t work, size_t worker)
{
    const uint32_t r = P.work_list ? P.work_list[work] : work;
    Splicer sp(P.env, splice_scratch_of(P, worker), P.codes + P.read_off[r], (int32_t)P.read_len[r]);
    SAln res[2];
    const int32_t n = sp.splice_match(P.opts[P.opt_idx ? P.opt_idx[r] : 0], res);
    if (sp.fail) {
        P.n_aln[r] = 0; P.status[r] = (uint8_t)sp.fail;
#if defined(__CUDA_ARCH__)
        const unsigned long long idx = atomicAdd(P.fail_count, 1ull);
#else
        const unsigned long long idx = (*P.fail_count)++;
#endif
        if (P.fail_list) P.fail_list[idx] = r;
    } else {
        P.n_aln[r] = n; P.status[r] = 0;
        splice_store(P.aln + (size_t)r * 18, res[0]);
        splice_store(P.aln + (size_t)r * 18 + 9, res[1]);
    }
#if defined(__CUDA_ARCH__)
    atomicAdd(P.lookups, sp.lookups);
#else
    *P.lookups += sp.lookups;
#endif
}

} // namespace hsa

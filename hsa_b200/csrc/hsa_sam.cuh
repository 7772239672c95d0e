// hsa_sam.cuh -- what generate_sam_se_core (bwtse.c:884-931) computes between the search and the printed SAM line
// (SURVEY.md section 8f item 3), one read per CUDA thread.
//
// Compiled by nvcc for sm_100a (sam_pos_kernel / sam_dp_kernel in hsa_b200.cu) and by g++ for tests/emu (CPU suite only).
//
//   bwt_aln2seq_core                 bwtse.c:21-113     hit selection on the process-wide drand48 stream, cut where the stream
//                                                       allows it (sel_* below; sam_select is the sequential statement)
//   bwa_approx_mapQ                  bwtse.c:122-131    with g_log_n all zero (bwase_initialize is never called, :892)
//   bwa_cal_pac_pos[_core]           bwtse.c:139-149, 350-369   SA index -> text position of the chosen hit and of the
//                                                       alternative hits, which are dropped where they coincide with it
//   bwt_aln2pos_splicing             bwtse.c:295-348    positions of both parts of a spliced hit, their pairing:
//   qsort_for_bwt / bwt_combine_segment_splice   bwtse.c:169-243   (the reference's own unstable quicksort, restated
//                                                       because the order of equal keys decides which pairs survive)
//   refine_gapped_core               bwtse.c:380-440    banded global alignment of the read against len + gaps reference
//   aln_global_core                  stdaln.c:345-524   bases (affine gaps 26/9, end gaps 5, band 50, aln_sm_maq scores),
//   aln_path2cigar32 / bwa_aln_path2cigar   stdaln.c:1010-1041, bwtaln.c:624-634   path -> CIGAR, end clean-up
//   bwa_refine_gapped                bwtse.c:536-638    which hits are refined; merging the two parts of a spliced hit
//   bwa_cal_md1                      bwtse.c:442-494    MD string and NM
//
// Work split: sel_* (the selection), sam_pos (every read: positions, mapQ, pairing; MD directly when no alignment has to be refined; the reads
// that need the dynamic programme are appended to a list) and sam_dp (one listed read per thread: every refinement of the
// read, then its MD).  The DP keeps ONE score row in place (the reference's curr/last pair collapses to a row plus the
// carried diagonal) -- in shared memory for reads up to ~250 bases -- and one byte of trace-back per cell in global memory;
// a thread's scratch is interleaved with its neighbours' (element e of worker t at e * T + t) so that the 32 lanes of a
// warp, which walk cells of the same index, touch one sector (or 32 banks).
#pragma once
#include "hsa_splice.cuh"
#include <cstring>

namespace hsa {

enum : uint32_t { TYPE_NO_MATCH = 0u, TYPE_UNIQUE = 1u, TYPE_REPEAT = 2u, TYPE_MATESW = 3u };     // bwtaln.h:9-13
enum : uint32_t { CIG_M = 0u, CIG_I = 1u, CIG_D = 2u, CIG_N = 3u, CIG_S = 4u };                     // stdaln.h:70-78
enum : int32_t { DP_INF = -1073741823, DP_GAP_OPEN = 26, DP_GAP_EXT = 9, DP_GAP_END = 5, DP_BAND = 50 };   // aln_param_bwa, stdaln.c:227
enum : uint32_t { SAM_OK = 0u, SAM_CIGAR_FULL = 1u, SAM_MD_FULL = 2u, SAM_SCRATCH = 3u };

struct SamRec {                 // == hsa_sam1_t: the bwa_seq_t fields generate_sam_se_core leaves (bwtaln.h:92-122)
    uint32_t type, strand, n_mm, n_gapo, n_gape, mapQ;
    int32_t score;
    uint32_t sa, seq_id, ori_pos, occ_pos, c1, c2;
    int32_t start, end;
    uint32_t n_cigar, cigar_off, nm, md_len, md_off, n_multi, multi_off;
};
struct SamMulti {               // == hsa_multi1_t: bwt_multi1_t (bwtaln.h:82-90)
    uint32_t n_cigar, cigar_off, gap, mm, strand, sa, ori_pos, occ_pos, seq_id, aln_id;
    int32_t start, end;
};

struct SamParams {
    SpliceEnv env;
    const uint8_t *codes; const uint64_t *read_off; const uint32_t *read_len; uint32_t n_reads;
    const int32_t *n_aln; const uint64_t *aln_off; const uint32_t *aln;       // the search's hits (9 words each)
    SamRec *rec; SamMulti *multi;
    const int32_t *maxdiff_by_len; int32_t max_mm; uint32_t max_len;          // bwa_cal_pac_pos_core's max_diff per length
    uint32_t *cigar; unsigned long long cigar_cap; unsigned long long *cigar_used;
    char *md; unsigned long long md_cap; unsigned long long *md_used;
    uint32_t *dp_list; unsigned long long *dp_count;
    unsigned long long *dp_tasks;                                             // refinements + merges the listed reads will make
    unsigned long long *cursor; uint32_t *status;                             // status: first failure kind (0 = none)
    // DP scratch, interleaved over dp_workers workers
    uint8_t *dp_bytes; int32_t *dp_rows; uint32_t dp_workers, dp_w, dp_len1_cap, dp_len2_cap;
    uint32_t dp_rows_smem;      // score row + reference bases of a worker in shared memory (reads up to ~250 bases)
};

// ---- small pieces ---------------------------------------------------------------------------------------------------
HSA_HD uint32_t sam_dna_at(const SpliceEnv &E, uint32_t k) { return (ld_ro1(E.packed_dna + (k >> 4)) >> ((~k & 15u) << 1)) & 3u; }

struct SamRead {                // s->strand ? s->rseq : s->seq, optionally from an offset (bwtse.c:583)
    const uint8_t *rd; uint32_t len, rev, off;
    HSA_HD uint32_t at(uint32_t i) const
    {
        i += off;
        if (!rev) return ld_ro_u8(rd + i);
        const uint32_t c = ld_ro_u8(rd + (len - 1u - i));       // seq_reverse(len, rseq, 1): bwaseqio.c:73-90
        return c < 4u ? 3u - c : c;
    }
};

HSA_HD uint32_t sam_claim(unsigned long long *used, unsigned long long cap, uint32_t n, bool &ok)
{
#if defined(__CUDA_ARCH__)
    const unsigned long long o = atomicAdd(used, (unsigned long long)n);
#else
    const unsigned long long o = *used; *used += n;
#endif
    ok = o + n <= cap;                       // the counter keeps counting past the cap: the host re-runs with that size
    return (uint32_t)o;
}
HSA_HD void sam_fail(const SamParams &P, uint32_t kind)
{
#if defined(__CUDA_ARCH__)
    atomicCAS(P.status, 0u, kind);
#else
    if (*P.status == 0u) *P.status = kind;
#endif
}

// bwa_approx_mapQ (bwtse.c:122-131); g_log_n[] is all zero in the reference as shipped
HSA_HD uint32_t sam_mapq(uint32_t c1, uint32_t c2, uint32_t n_mm, int32_t mm)
{
    if (c1 == 0) return 23;
    if (c1 > 1) return 0;
    if ((int32_t)n_mm == mm) return 25;
    if (c2 == 0) return 37;
    return 23;
}

// BWTRetrievePositionFromSAIndex (2BWT-Interface.c:329-362): occ_pos always, seq_id / ori_pos where a block holds it
HSA_HD void sam_locate(const SpliceEnv &E, uint32_t sa, uint32_t &seq_id, uint32_t &ori_pos, uint32_t &occ_pos)
{
    uint32_t steps;
    occ_pos = sa_value_dev(E.ix.fwd, E.sa_value, E.sa_interval, sa, steps);
    locate_dev(E.blocks4, E.n_blocks, occ_pos, seq_id, ori_pos);
}

// ---- MD / NM (bwa_cal_md1, bwtse.c:442-494) ----------------------------------------------------------------------------
struct MdSink {
    char *out; uint32_t n;          // out == nullptr: count only
    HSA_HD void put(char c) { if (out) out[n] = c; ++n; }
    HSA_HD void num(int32_t u)      // ksprintf(str, "%d", u), u >= 0
    {
        char t[12]; int k = 0;
        do { t[k++] = (char)('0' + u % 10); u /= 10; } while (u);
        while (k) put(t[--k]);
    }
};

HSA_HD uint32_t sam_md(const SpliceEnv &E, uint32_t n_cigar, const uint32_t *cigar, uint32_t len, uint32_t pos, const SamRead &rd, MdSink &s)
{
    uint32_t x = pos, y = 0, nm = 0; int32_t u = 0;
    if (n_cigar) {
        for (uint32_t k = 0; k < n_cigar; ++k) {
            const uint32_t c = ld_ro1(cigar + k), op = c >> 28; const int32_t l = (int32_t)(c & 0x0FFFFFFFu);
            if (op == CIG_M) {
                for (int32_t z = 0; z < l && x + (uint32_t)z < E.dna_length; ++z) {
                    const uint32_t r = sam_dna_at(E, x + (uint32_t)z), q = rd.at(y + (uint32_t)z);
                    if (q > 3u || r != q) { s.num(u); s.put("ACGTN"[r]); ++nm; u = 0; }
                    else ++u;
                }
                x += (uint32_t)l; y += (uint32_t)l;
            } else if (op == CIG_I || op == CIG_S) {
                y += (uint32_t)l;
                if (op == CIG_I) nm += (uint32_t)l;
            } else if (op == CIG_D) {
                s.num(u); s.put('^');
                for (int32_t z = 0; z < l && x + (uint32_t)z < E.dna_length; ++z) s.put("ACGT"[sam_dna_at(E, x + (uint32_t)z)]);
                u = 0; x += (uint32_t)l; nm += (uint32_t)l;
            }
        }
    } else {
        for (uint32_t z = 0; z < len; ++z) {
            const uint32_t r = sam_dna_at(E, x + z), q = rd.at(y + z);
            if (q > 3u || r != q) { s.num(u); s.put("ACGTN"[r]); ++nm; u = 0; }
            else ++u;
        }
    }
    s.num(u);
    return nm;
}

// MD + NM of a read whose CIGAR (if any) is final: count, claim, write
HSA_HD void sam_md_of(const SamParams &P, SamRec &r, const uint8_t *rd, uint32_t len)
{
    const SamRead q{rd, len, r.strand, 0u};
    const uint32_t *cg = r.n_cigar ? P.cigar + r.cigar_off : nullptr;
    MdSink cnt{nullptr, 0u};
    r.nm = sam_md(P.env, r.n_cigar, cg, len, r.occ_pos, q, cnt) & 0xFFFu;            // bwa_seq_t::nm is 12 bits wide
    bool ok;
    r.md_len = cnt.n; r.md_off = sam_claim(P.md_used, P.md_cap, cnt.n, ok);
    if (!ok) { sam_fail(P, SAM_MD_FULL); return; }
    MdSink wr{P.md + r.md_off, 0u};
    sam_md(P.env, r.n_cigar, cg, len, r.occ_pos, q, wr);
}

// ---- the pairing of a spliced hit's two parts (bwtse.c:151-243) ------------------------------------------------------------
HSA_HD void sam_swap(SamMulti *m, int32_t i, int32_t j) { const SamMulti t = m[i]; m[i] = m[j]; m[j] = t; }

// qsort_for_bwt (bwtse.c:169-192) with an explicit range stack: first element as pivot, Hoare-style scans
HSA_HD void sam_qsort(SamMulti *m, int32_t n)
{
    int32_t lo[128], hi[128]; int sp = 0;           // n <= 100, and every pushed range is strictly smaller than its parent
    lo[0] = 0; hi[0] = n - 1; sp = 1;
    while (sp) {
        const int32_t l = lo[--sp], u = hi[sp];
        if (l >= u) continue;
        const uint32_t t = m[l].occ_pos;
        int32_t i = l, j = u + 1;
        for (;;) {
            do ++i; while (i <= u && m[i].occ_pos < t);
            do --j; while (m[j].occ_pos > t);
            if (i > j) break;
            sam_swap(m, i, j);
        }
        sam_swap(m, l, j);
        if (sp + 2 > 128) return;                   // cannot happen for n <= 100
        lo[sp] = j + 1; hi[sp] = u; ++sp;
        lo[sp] = l; hi[sp] = j - 1; ++sp;
    }
}

// bwt_combine_segment_splice (bwtse.c:197-243)
HSA_HD int32_t sam_combine(SamMulti *m, int32_t n_multi)
{
    if (n_multi == 2) return 1;
    sam_qsort(m, n_multi);
    int32_t n_res = 0; uint32_t shortest = 0xFFFFFFFFu;
    for (int32_t i = 0; i < n_multi - 1; ++i) {
        SamMulti *t = m + i;
        if (t->aln_id != 0) continue;
        if (t->aln_id == t[1].aln_id || t->seq_id != t[1].seq_id || t[1].occ_pos - t->occ_pos > shortest) continue;
        const uint32_t d = t[1].occ_pos - t->occ_pos;
        if (d < 50u || d > 50000u) continue;
        if (d < shortest) { shortest = d; n_res = 1; if (i == 0) continue; }
        else if (d == shortest) n_res += 1;
        SamMulti *res = m + (n_res - 1) * 2;
        if (res != t) { const SamMulti a = t[0], b = t[1]; res[0] = a; res[1] = b; }       // memmove of two entries
    }
    return n_res;
}

// ---- stage 1: every read (bwa_cal_pac_pos, bwtse.c:350-369; the decision of bwa_refine_gapped, :544-560) ---------------
HSA_HD void sam_pos_item(const SamParams &P, uint32_t rid)
{
    SamRec r = P.rec[rid];
    const uint8_t *rd = P.codes + P.read_off[rid]; const uint32_t len = P.read_len[rid];
    bool need_dp = false;
    if (r.type == TYPE_SPLICING) {
        // bwt_aln2pos_splicing (bwtse.c:295-348); n_aln == 2 by the caller's test (bwtse.c:901)
        const uint32_t *a0 = P.aln + 9 * P.aln_off[rid];
        SamMulti *m = P.multi + r.multi_off; uint32_t cnt = 0;
        r.strand = a0[5] >> 30;
        uint32_t nm0 = (a0[2] - a0[1]) + (a0[9 + 2] - a0[9 + 1]) + 2u;
        if (nm0 >= 100u) nm0 = 100u;
        r.c1 = r.c2 = nm0;
        for (uint32_t id = 0; id < 2; ++id) {
            const uint32_t *a = a0 + 9 * id;
            const uint32_t mm = ((a[0] & 0xFFFFu) + ((a[0] >> 16) & 0xFFu) + (a[0] >> 24)) & 0xFFu;
            for (uint32_t sa = a[1]; sa <= a[2] && sa < a[1] + 50u; ++sa, ++cnt) {
                SamMulti q; q.n_cigar = q.cigar_off = q.gap = 0; q.sa = 0; q.seq_id = q.ori_pos = 0;
                sam_locate(P.env, sa, q.seq_id, q.ori_pos, q.occ_pos);
                q.strand = a[5] >> 30; q.start = (int32_t)a[6]; q.end = (int32_t)a[7]; q.aln_id = id; q.mm = mm;
                m[cnt] = q;
            }
        }
        const int32_t n_res = sam_combine(m, (int32_t)cnt);
        r.n_multi = (uint32_t)n_res;
        if (n_res == 0) r.type = TYPE_NO_MATCH;
        else need_dp = true;
    } else {
        if (r.type == TYPE_UNIQUE || r.type == TYPE_REPEAT) {                               // bwa_cal_pac_pos_core
            const int32_t mm = P.maxdiff_by_len ? P.maxdiff_by_len[len] : P.max_mm;
            sam_locate(P.env, r.sa, r.seq_id, r.ori_pos, r.occ_pos);
            r.mapQ = sam_mapq(r.c1, r.c2, r.n_mm, mm);
        }
        SamMulti *m = P.multi + r.multi_off; uint32_t keep = 0;
        for (uint32_t j = 0; j < r.n_multi; ++j) {
            SamMulti q = m[j];
            sam_locate(P.env, q.sa, q.seq_id, q.ori_pos, q.occ_pos);
            if (q.occ_pos != r.occ_pos) { m[keep++] = q; if (q.gap) need_dp = true; }
        }
        r.n_multi = keep;
        if (r.type != TYPE_NO_MATCH && r.type != TYPE_MATESW && r.n_gapo != 0) need_dp = true;
        if (!need_dp && r.type != TYPE_NO_MATCH) sam_md_of(P, r, rd, len);
    }
    P.rec[rid] = r;
    if (need_dp) {
        bool ok;
        const uint32_t slot = sam_claim(P.dp_count, (unsigned long long)P.n_reads, 1u, ok);
        if (ok) P.dp_list[slot] = rid;
        uint32_t tasks = r.type == TYPE_SPLICING ? 3u * r.n_multi : (r.n_gapo != 0 ? 1u : 0u);
        if (r.type != TYPE_SPLICING) for (uint32_t j = 0; j < r.n_multi; ++j) tasks += P.multi[r.multi_off + j].gap ? 1u : 0u;
        sam_claim(P.dp_tasks, ~0ull, tasks, ok);
    }
}

// ---- stage 2: the dynamic programme ---------------------------------------------------------------------------------------
struct DpScratch {              // interleaved arrays: element e of worker t lives at e * stride + t
    uint8_t *cells, *ref; int32_t *rows;
    uint32_t T;                 // stride of the trace-back cells (global memory: all workers of the launch)
    uint32_t RT;                // stride of the score row and the reference bases (global: T; shared memory: the block's threads)
    uint32_t W, len1_cap, len2_cap, run_cap;
    HSA_HD uint8_t &cell(uint32_t row, uint32_t col) const { return cells[((size_t)row * W + col) * T]; }
    HSA_HD int32_t &row(uint32_t i, uint32_t c) const { return rows[((size_t)i * 3u + c) * RT]; }
    HSA_HD uint8_t &refb(uint32_t i) const { return ref[(size_t)i * RT]; }
    HSA_HD int32_t &run(uint32_t i) const { return rows[(size_t)i * RT]; }     // the run list re-uses the score row
};

// everything in global memory (long reads, the host build)
HSA_HD DpScratch dp_scratch_of(const SamParams &P, uint32_t w)
{
    DpScratch s;
    s.T = s.RT = P.dp_workers; s.W = P.dp_w; s.len1_cap = P.dp_len1_cap; s.len2_cap = P.dp_len2_cap;
    s.cells = P.dp_bytes + w;
    s.ref = P.dp_bytes + (size_t)(P.dp_len2_cap + 1u) * P.dp_w * P.dp_workers + w;
    s.rows = P.dp_rows + w;
    s.run_cap = 3u * (P.dp_len1_cap + 1u) > P.dp_len1_cap + P.dp_len2_cap + 2u ? 3u * (P.dp_len1_cap + 1u) : P.dp_len1_cap + P.dp_len2_cap + 2u;
    return s;
}

struct DpCell { int32_t M, I, D; };

// stdaln.c:258-318: the three transitions and their end-gap forms (gap_end >= 0)
HSA_HD int32_t dp_set_M(const DpCell &p, int32_t sc, uint32_t &t)
{
    if (p.M >= p.I) { if (p.M >= p.D) { t = CIG_M; return p.M + sc; } t = CIG_D; return p.D + sc; }
    if (p.I > p.D) { t = CIG_I; return p.I + sc; }
    t = CIG_D; return p.D + sc;
}
HSA_HD int32_t dp_set_gap(int32_t m, int32_t g, int32_t ext, uint32_t self, uint32_t &t)
{
    if (m - DP_GAP_OPEN > g) { t = CIG_M; return m - DP_GAP_OPEN - ext; }
    t = self; return g - ext;
}
HSA_HD int32_t dp_score(uint32_t r, uint32_t q)       // aln_sm_maq (stdaln.c:205-211): mat[q * 5 + r]
{
    if (q > 3u || r > 3u) return -13;
    return r == q ? 11 : -19;
}

// One row of aln_global_core.  lo: first column of the band (col0: the matrix edge with an end-gap insertion,
// stdaln.c:396-398; else an all-infinite cell, :448); hi: last column; hi_end_i: the last cell's insertion comes from the
// matrix edge (:411-413, :474) or is infinite (:413, :458); d_end: last row, deletions are end gaps (:417-433, :480-492).
HSA_HD void dp_row(const DpScratch &S, uint32_t j, uint32_t lo, bool col0, uint32_t hi, bool hi_end_i, bool d_end, uint32_t qj,
                   uint32_t cell_base)
{
    DpCell diag{S.row(lo, 0), S.row(lo, 1), S.row(lo, 2)}, left{DP_INF, DP_INF, DP_INF};
    if (col0) {
        uint32_t it; left.I = dp_set_gap(diag.M, diag.I, DP_GAP_END, CIG_I, it);
        S.cell(j, 0) = (uint8_t)(it << 3);
    }
    S.row(lo, 0) = left.M; S.row(lo, 1) = left.I; S.row(lo, 2) = left.D;
    const int32_t d_ext = d_end ? DP_GAP_END : DP_GAP_EXT;
    for (uint32_t i = lo + 1u; i <= hi; ++i) {
        const DpCell up{S.row(i, 0), S.row(i, 1), S.row(i, 2)};
        DpCell cur; uint32_t mt, it = 0, dt;
        cur.M = dp_set_M(diag, dp_score(S.refb(i), qj), mt);
        if (i < hi) cur.I = dp_set_gap(up.M, up.I, DP_GAP_EXT, CIG_I, it);
        else if (hi_end_i) cur.I = dp_set_gap(up.M, up.I, DP_GAP_END, CIG_I, it);
        else cur.I = DP_INF;
        cur.D = dp_set_gap(left.M, left.D, d_ext, CIG_D, dt);
        S.row(i, 0) = cur.M; S.row(i, 1) = cur.I; S.row(i, 2) = cur.D;
        S.cell(j, i - cell_base) = (uint8_t)(mt | it << 3 | dt << 5);
        diag = up; left = cur;
    }
}

// aln_global_core (stdaln.c:345-524) + aln_path2cigar32 (:1010-1041): ref bases S.refb(1..len1), read bases q.at(0..len2-1).
// Leaves the CIGAR's runs in S.run(0..n_runs) in path order REVERSED (last operation first) as op << 28 | length.
HSA_HD int32_t dp_global(const DpScratch &S, int32_t len1, int32_t len2, const SamRead &q, int32_t &n_runs)
{
    n_runs = 0;
    if (len1 == 0 || len2 == 0) return 0;
    int32_t b1, b2;
    if (len1 > len2) { b1 = len1 - len2 + DP_BAND; b2 = DP_BAND; } else { b1 = DP_BAND; b2 = len2 - len1 + DP_BAND; }
    if (b1 > len1) b1 = len1;
    if (b2 > len2) b2 = len2;
    // the trace-back rows hold b1 + b2 + 1 cells (:380); a reference window clipped at the end of the text (len1 < len2) makes
    // the band wider than the batch's scratch was sized for: the caller runs the batch again with full-width rows
    if ((uint32_t)((b1 + b2 <= len1) ? b1 + b2 + 1 : len1 + 1) > S.W) { n_runs = -1; return 0; }
    // first row (:386-391)
    S.row(0, 0) = 0; S.row(0, 1) = DP_INF; S.row(0, 2) = DP_INF;
    {
        DpCell left{0, DP_INF, DP_INF};
        for (int32_t i = 1; i < b1; ++i) {
            uint32_t dt; DpCell cur{DP_INF, DP_INF, 0};
            cur.D = dp_set_gap(left.M, left.D, DP_GAP_END, CIG_D, dt);
            S.row(i, 0) = cur.M; S.row(i, 1) = cur.I; S.row(i, 2) = cur.D;
            S.cell(0, i) = (uint8_t)(dt << 5);
            left = cur;
        }
    }
    int32_t j = 1;
    const int32_t tmp_end = b2 < len2 ? b2 : len2 - 1;
    for (; j <= tmp_end; ++j) {                                                     // part 1 (:394-415)
        const int32_t end = j + b1 <= len1 + 1 ? j + b1 - 1 : len1;
        dp_row(S, j, 0, true, end, j + b1 - 1 > len1, false, q.at(j - 1), 0);
    }
    if (j == len2 && b2 != len2 - 1) {                                              // last row of part 1 (:417-435)
        const int32_t end = j + b1 <= len1 + 1 ? j + b1 - 1 : len1;
        dp_row(S, j, 0, true, end, j + b1 - 1 > len1, true, q.at(j - 1), 0);
        ++j;
    }
    for (; j <= len2 - b2 + 1; ++j)                                                 // part 2 (:438-451)
        dp_row(S, j, j - b2, false, j + b1 - 1, false, false, q.at(j - 1), j > b2 ? j - b2 : 0);
    for (; j < len2; ++j)                                                           // part 3 (:454-466)
        dp_row(S, j, j - b2, false, len1, true, false, q.at(j - 1), j > b2 ? j - b2 : 0);
    if (j == len2)                                                                  // last row (:468-481)
        dp_row(S, j, j - b2, false, len1, true, true, q.at(j - 1), j > b2 ? j - b2 : 0);

    // backtrace (:484-510); the path's entries are the operations in reverse order, run-length coded on the fly
    int32_t i = len1; j = len2;
    const int32_t mx_M = S.row(len1, 0), mx_I = S.row(len1, 1), mx_D = S.row(len1, 2);
    uint32_t c = S.cell(j, i - (j > b2 ? j - b2 : 0));
    int32_t mx = mx_M; uint32_t type = c & 7u, ctype = CIG_M;
    if (mx_I > mx) { mx = mx_I; type = (c >> 3) & 3u; ctype = CIG_I; }
    if (mx_D > mx) { mx = mx_D; type = (c >> 5) & 3u; ctype = CIG_D; }
    uint32_t run_op = ctype, run_len = 0;
    do {
        if (ctype == run_op) ++run_len;
        else {
            if ((uint32_t)n_runs + 2u > S.run_cap) { n_runs = -1; return 0; }       // (shared-memory rows of a clipped window: full-width retry)
            S.run(n_runs++) = (int32_t)(run_op << 28 | run_len); run_op = ctype; run_len = 1;
        }
        if (ctype == CIG_M) { --i; --j; } else if (ctype == CIG_I) --j; else --i;
        c = S.cell(j, i - (j > b2 ? j - b2 : 0));
        ctype = type;
        type = ctype == CIG_M ? (c & 7u) : ctype == CIG_I ? ((c >> 3) & 3u) : ((c >> 5) & 3u);
    } while (i || j);
    S.run(n_runs++) = (int32_t)(run_op << 28 | run_len);
    return mx;
}

// refine_gapped_core (bwtse.c:380-440): the CIGAR goes to the arena; pos / ori_pos / start / end are adjusted
HSA_HD bool sam_refine(const SamParams &P, const DpScratch &S, int32_t len, const SamRead &q, uint32_t &pos, uint32_t &ori_pos,
                       int32_t ext, int32_t &start, int32_t &end, uint32_t &n_cigar_out, uint32_t &cigar_off_out)
{
    const SpliceEnv &E = P.env;
    const int32_t ref_len = len + (ext < 0 ? -ext : ext);
    int32_t l = 0;
    if ((uint32_t)ref_len > S.len1_cap || (uint32_t)len > S.len2_cap) { sam_fail(P, SAM_SCRATCH); return false; }
    for (uint32_t k = pos; k < pos + (uint32_t)ref_len && k < E.dna_length; ++k) S.refb(++l) = (uint8_t)sam_dna_at(E, k);
    int32_t n_runs;
    dp_global(S, l, len, q, n_runs);
    if (n_runs < 0) { sam_fail(P, SAM_SCRATCH); return false; }
    if (n_runs == 0) { n_cigar_out = 0; cigar_off_out = 0; return true; }           // (the reference dereferences NULL here)
    // runs are stored last-first: run(n_runs - 1 - k) is cigar[k]
    int32_t first = 0, n = n_runs;                    // cigar[k] = run(n_runs - 1 - first - k), k < n
    uint32_t c0 = (uint32_t)S.run(n_runs - 1);
    if ((c0 >> 28) == CIG_D) {                                                      // deletion at the 5' end (:413-421)
        const uint32_t dl = c0 & 0x0FFFFFFFu;
        pos += dl; ori_pos += dl; start += (int32_t)dl;
        ++first; --n;
    }
    if (n > 0 && ((uint32_t)S.run(n_runs - 1 - first - (n - 1)) >> 28) == CIG_D) {   // deletion at the 3' end (:422-426)
        end -= (int32_t)((uint32_t)S.run(n_runs - 1 - first - (n - 1)) & 0x0FFFFFFFu);
        --n;
    }
    if (n <= 0) { n_cigar_out = 0; cigar_off_out = 0; return true; }
    bool ok;
    const uint32_t off = sam_claim(P.cigar_used, P.cigar_cap, (uint32_t)n, ok);
    n_cigar_out = (uint32_t)n; cigar_off_out = off;
    if (!ok) { sam_fail(P, SAM_CIGAR_FULL); return false; }
    for (int32_t k = 0; k < n; ++k) {
        uint32_t c = (uint32_t)S.run(n_runs - 1 - first - k);
        if ((k == n - 1 || k == 0) && (c >> 28) == CIG_I) c = CIG_S << 28 | (c & 0x0FFFFFFFu);      // I at either end -> S (:430-434)
        P.cigar[off + (uint32_t)k] = c;
    }
    return true;
}

// two refined parts -> one CIGAR with the intron between them (bwtse.c:590-595, 620-624)
HSA_HD bool sam_merge(const SamParams &P, const SamMulti &a, const SamMulti &b, uint32_t &n_cigar, uint32_t &cigar_off)
{
    bool ok;
    n_cigar = a.n_cigar + b.n_cigar + 1u;
    cigar_off = sam_claim(P.cigar_used, P.cigar_cap, n_cigar, ok);
    if (!ok) { sam_fail(P, SAM_CIGAR_FULL); return false; }
    for (uint32_t k = 0; k < a.n_cigar; ++k) P.cigar[cigar_off + k] = P.cigar[a.cigar_off + k];
    P.cigar[cigar_off + a.n_cigar] = CIG_N << 28 | (b.ori_pos - a.ori_pos - (uint32_t)a.end);
    for (uint32_t k = 0; k < b.n_cigar; ++k) P.cigar[cigar_off + a.n_cigar + 1u + k] = P.cigar[b.cigar_off + k];
    return true;
}

// bwa_refine_gapped for one read that needs it (bwtse.c:536-638), then its MD
HSA_HD void sam_dp_item(const SamParams &P, uint32_t rid, const DpScratch &S)
{
    SamRec r = P.rec[rid];
    const uint8_t *rd = P.codes + P.read_off[rid]; const uint32_t len = P.read_len[rid];
    SamMulti *m = P.multi + r.multi_off;
    if (r.type != TYPE_SPLICING) {
        for (uint32_t j = 0; j < r.n_multi; ++j) {
            SamMulti q = m[j];
            if (q.gap == 0) continue;
            const SamRead rq{rd, len, q.strand, 0u};
            if (!sam_refine(P, S, (int32_t)len, rq, q.occ_pos, q.ori_pos, (q.strand ? 1 : -1) * (int32_t)q.gap, q.start, q.end, q.n_cigar, q.cigar_off)) return;
            m[j] = q;
        }
        if (r.type != TYPE_NO_MATCH && r.type != TYPE_MATESW && r.n_gapo != 0) {
            const SamRead rq{rd, len, r.strand, 0u};
            if (!sam_refine(P, S, (int32_t)len, rq, r.occ_pos, r.ori_pos, (r.strand ? 1 : -1) * (int32_t)(r.n_gapo + r.n_gape), r.start, r.end,
                            r.n_cigar, r.cigar_off)) return;
        }
        if (r.type != TYPE_NO_MATCH) sam_md_of(P, r, rd, len);
    } else {
        const uint32_t n_multi = r.n_multi;
        for (uint32_t i = 0; i < 2u * n_multi; ++i) {
            SamMulti q = m[i];
            const int32_t seg_len = q.end - q.start + 1;
            const SamRead rq{rd, len, r.strand == 1u ? 1u : 0u, (uint32_t)q.start};
            if (!sam_refine(P, S, seg_len, rq, q.occ_pos, q.ori_pos, (r.strand ? 1 : -1) * (int32_t)q.gap, q.start, q.end, q.n_cigar, q.cigar_off)) return;
            m[i] = q;
        }
        if (!sam_merge(P, m[0], m[1], r.n_cigar, r.cigar_off)) return;
        r.occ_pos = m[0].occ_pos; r.ori_pos = m[0].ori_pos; r.seq_id = m[0].seq_id; r.sa = m[0].sa; r.strand = m[0].strand;
        r.n_gapo = (m[0].gap + m[1].gap) & 0xFFu; r.n_gape = 0; r.n_mm = (m[0].mm + m[1].mm) & 0xFFu;
        r.n_multi = n_multi - 1u;
        for (uint32_t i = 1; i < n_multi; ++i) {                                    // further pairings (:605-629)
            const SamMulti a = m[2u * i], b = m[2u * i + 1u];
            SamMulti res = a;
            if (!sam_merge(P, a, b, res.n_cigar, res.cigar_off)) return;
            res.gap = (a.gap + b.gap) & 0xFFu; res.mm = (a.mm + b.mm) & 0xFFu;
            m[i - 1u] = res;
        }
    }
    P.rec[rid] = r;
}


// ---- hit selection (generate_sam_se_core's first loop, bwtse.c:898-910, and bwt_aln2seq_core, :21-113) -------------------
// The reference draws from the process-wide drand48 stream, read after read, and how many numbers a read consumes depends
// on the numbers drawn.  drand48 (POSIX): X' = (0x5DEECE66D X + 0xB) mod 2^48, value X' / 2^48; glibc starts an unseeded
// stream at X = 0.
//
// sam_select is the sequential statement (one read after the other).  The GPU path cuts the chain where it can be cut:
//   * a read whose best score is held by ONE hit draws exactly twice (the take test of the first hit is r * w > 0, which only
//     r = 0 fails) -- its position in the stream follows from a prefix sum, and an LCG can jump: X_n = A^n X + C (A^n - 1) / (A - 1);
//   * reads with no hit or a spliced pair draw nothing;
//   * only reads with SEVERAL best-score hits consume a data-dependent number of draws.  They are compacted into a list of
//     32-byte descriptors and ONE thread walks that list (sel_chain), jumping over the fixed draws in between;
//   * every other read then jumps straight to its own position (sel_finish).
// A read that breaks the assumption (r = 0 on a first draw: probability 2^-48; the unreachable sampling branch of
// bwtse.c:93-107) raises a flag and the batch is selected again by sam_select on one device thread -- exact either way.
struct Rng48 {
    uint64_t x;
    HSA_HD double next() { x = (x * 0x5DEECE66Dull + 0xBull) & 0xFFFFFFFFFFFFull; return (double)x * (1.0 / 281474976710656.0); }
    HSA_HD void jump(uint64_t n)                  // n draws ahead, O(log n)
    {
        uint64_t ca = 0x5DEECE66Dull, cc = 0xBull, aa = 1, ac = 0;
        while (n) {
            if (n & 1ull) { aa = aa * ca; ac = ac * ca + cc; }
            cc = (ca + 1ull) * cc; ca = ca * ca;
            n >>= 1;
        }
        x = (aa * x + ac) & 0xFFFFFFFFFFFFull;
    }
};

enum : uint32_t { SEL_NONE = 0u, SEL_SPLICE = 1u, SEL_SINGLE = 2u, SEL_SEVERAL = 3u };
enum : uint32_t { SEL_DESC_W = 5u };
struct SelDesc { uint32_t read, m, fixed_before, w[SEL_DESC_W]; };       // a read with m >= 2 best-score hits; widths of the first five
struct SelPick { uint32_t j; uint32_t taken; double r2; };               // the hit taken last and the draw that places sa inside it

// hit words: {n_mm | n_gapo << 16 | n_gape << 24, k, l, rev_k, rev_l, type | strand << 30, start, end, score}
HSA_HD uint32_t sel_kind(const uint32_t *aln, int32_t n_aln, uint32_t &m)
{
    m = 0;
    if (n_aln == 2 && (aln[5] & 0x3FFFFFFFu) == TYPE_SPLICING) return SEL_SPLICE;                // bwtse.c:901-904
    if (n_aln == 0) return SEL_NONE;
    const int32_t best = (int32_t)aln[8];
    while ((int32_t)m < n_aln && (int32_t)aln[9 * m + 8] <= best) ++m;                            // :40-42
    return m == 1 ? SEL_SINGLE : SEL_SEVERAL;
}
// multi slots the read owns: every occurrence of every hit when there are at most n_occ + 1 of them (:63-92), or room for
// the positions bwt_aln2pos_splicing looks up (<= 50 per part, bwtse.c:322)
HSA_HD uint32_t sel_slots(const uint32_t *aln, int32_t n_aln, int32_t n_occ, uint32_t kind)
{
    if (kind == SEL_SPLICE) {
        const uint32_t w0 = aln[2] - aln[1] + 1u, w1 = aln[9 + 2] - aln[9 + 1] + 1u;
        return (w0 < 50u ? w0 : 50u) + (w1 < 50u ? w1 : 50u);
    }
    if (kind == SEL_NONE || !n_occ) return 0;
    int32_t tot = 0;
    for (int32_t k = 0; k < n_aln; ++k) tot += (int32_t)(aln[9 * k + 2] - aln[9 * k + 1] + 1u);
    return tot > n_occ + 1 ? 0u : (uint32_t)tot;
}
// the draws of one read's set_main loop (:38-53) from the stream at rng; returns the number drawn
HSA_HD uint32_t sel_draw(const uint32_t *aln, uint32_t m, const uint32_t *w_inline, Rng48 &rng, SelPick &pick)
{
    int32_t cnt = 0; uint32_t drawn = 0;
    pick.j = 0; pick.taken = 0; pick.r2 = 0.0;
    for (uint32_t j = 0; j < m; ++j) {
        const uint32_t w = (w_inline && j < SEL_DESC_W) ? w_inline[j] : aln[9 * j + 2] - aln[9 * j + 1] + 1u;
        ++drawn;
        if (rng.next() * (double)(uint32_t)(w + (uint32_t)cnt) > (double)cnt) { pick.j = j; pick.taken = 1; pick.r2 = rng.next(); ++drawn; }
        cnt += (int32_t)w;
    }
    return drawn;
}
// the fields of a matched read once its pick is known (:44-61), and its alternative-hit list (:63-92); false: the read needs
// the sampling branch (:93-107), which only the sequential statement evaluates
HSA_HD bool sel_fill(const uint32_t *aln, int32_t n_aln, int32_t n_occ, uint32_t m, const SelPick &pick, SamRec &s, SamMulti *multi)
{
    if (pick.taken) {
        const uint32_t *p = aln + 9 * pick.j;
        s.n_mm = p[0] & 0xFFu; s.n_gapo = (p[0] >> 16) & 0xFFu; s.n_gape = p[0] >> 24;               // 8-bit fields of bwa_seq_t
        s.score = (int32_t)p[8]; s.start = (int32_t)p[6]; s.end = (int32_t)p[7];
        s.sa = p[1] + (uint32_t)((double)(uint32_t)(p[2] - p[1] + 1u) * pick.r2);
        s.strand = (p[5] >> 30) & 1u;
    }
    int32_t cnt = 0, i;
    for (i = 0; i < (int32_t)m; ++i) cnt += (int32_t)(aln[9 * i + 2] - aln[9 * i + 1] + 1u);
    s.c1 = (uint32_t)cnt & 0x0FFFFFFFu;
    for (; i < n_aln; ++i) cnt += (int32_t)(aln[9 * i + 2] - aln[9 * i + 1] + 1u);
    s.c2 = ((uint32_t)cnt - s.c1) & 0x0FFFFFFFu;
    s.type = s.c1 > 1u ? TYPE_REPEAT : TYPE_UNIQUE;
    s.n_multi = 0;
    if (!n_occ) return true;
    if (cnt > n_occ + 1) return true;                                                   // too many occurrences: no list (:70-75)
    int32_t rest = cnt; uint32_t z = 0;
    for (int32_t k = 0; k < n_aln; ++k) {
        const uint32_t *q = aln + 9 * k;
        const uint32_t w = q[2] - q[1] + 1u;
        if (w > (uint32_t)rest) return false;
        SamMulti t;
        t.n_cigar = t.cigar_off = 0; t.sa = t.ori_pos = t.occ_pos = t.seq_id = t.aln_id = 0;
        t.start = (int32_t)q[6]; t.end = (int32_t)q[7]; t.strand = (q[5] >> 30) & 1u;
        t.gap = (((q[0] >> 16) & 0xFFu) + (q[0] >> 24)) & 0xFFu; t.mm = q[0] & 0xFFu;
        for (uint32_t l = q[1]; l <= q[2]; ++l) { t.sa = l; multi[z++] = t; if (l == 0xFFFFFFFFu) break; }
        rest -= (int32_t)w;
    }
    s.n_multi = z;
    return true;
}

// The sequential statement.  Fills rec (zero-initialised by the caller) and the read's alternative hits in multi[0 ...];
// returns the number of multi slots the read owns.
HSA_HD uint32_t sam_select(const uint32_t *aln, int32_t n_aln, int32_t n_occ, Rng48 &rng, SamRec &s, SamMulti *multi)
{
    uint32_t m;
    const uint32_t kind = sel_kind(aln, n_aln, m);
    if (kind == SEL_SPLICE) { s.type = TYPE_SPLICING; return sel_slots(aln, n_aln, n_occ, kind); }
    if (kind == SEL_NONE) { s.type = TYPE_NO_MATCH; s.c1 = s.c2 = 0; return 0; }
    SelPick pick;
    sel_draw(aln, m, nullptr, rng, pick);
    if (!sel_fill(aln, n_aln, n_occ, m, pick, s, multi)) {
        // the sampling branch (:93-107): unreachable while the list holds every occurrence (rest == the sum of the widths),
        // kept for the record
        int32_t rest = 0; uint32_t z = 0;
        for (int32_t k = 0; k < n_aln; ++k) rest += (int32_t)(aln[9 * k + 2] - aln[9 * k + 1] + 1u);
        for (int32_t k = 0; k < n_aln; ++k) {
            const uint32_t *q = aln + 9 * k;
            const uint32_t w = q[2] - q[1] + 1u;
            SamMulti t;
            t.n_cigar = t.cigar_off = 0; t.sa = t.ori_pos = t.occ_pos = t.seq_id = t.aln_id = 0;
            t.start = (int32_t)q[6]; t.end = (int32_t)q[7]; t.strand = (q[5] >> 30) & 1u;
            t.gap = (((q[0] >> 16) & 0xFFu) + (q[0] >> 24)) & 0xFFu; t.mm = q[0] & 0xFFu;
            if (w <= (uint32_t)rest) {
                for (uint32_t l = q[1]; l <= q[2]; ++l) { t.sa = l; multi[z++] = t; if (l == 0xFFFFFFFFu) break; }
                rest -= (int32_t)w;
            } else {
                int32_t j, i;
                for (j = rest, i = (int32_t)w; j > 0; --j) {
                    double p = 1.0; const double x = rng.next();
                    while (x < p) p -= p * j / (i--);
                    t.sa = q[2] - (uint32_t)i; multi[z++] = t;
                }
                break;
            }
        }
        s.n_multi = z;
    }
    return sel_slots(aln, n_aln, n_occ, kind);
}

// ---- the selection as kernels' bodies ------------------------------------------------------------------------------------
struct SelParams {
    const int32_t *n_aln; const uint64_t *aln_off; const uint32_t *aln; uint32_t n_reads; int32_t n_occ;
    uint32_t *fixed, *dep, *slots;              // per read; after the scans: exclusive prefix sums (n_reads + 1 entries)
    SelDesc *desc; SelPick *pick; uint32_t *vcum;   // per dependent read: descriptor, outcome, draws of the dependent reads so far
    SamRec *rec; SamMulti *multi;
    uint64_t x0; uint64_t *x_end;               // stream state at the start of the batch / at its end
    uint32_t *rare;                             // != 0: the batch must be selected by the sequential statement
};

HSA_HD void sel_classify_item(const SelParams &P, uint32_t r)
{
    uint32_t m;
    const uint32_t *a = P.aln + 9 * P.aln_off[r];
    const uint32_t kind = sel_kind(a, P.n_aln[r], m);
    P.fixed[r] = kind == SEL_SINGLE ? 2u : 0u;
    P.dep[r] = kind == SEL_SEVERAL ? 1u : 0u;
    P.slots[r] = sel_slots(a, P.n_aln[r], P.n_occ, kind);
}
HSA_HD void sel_desc_item(const SelParams &P, uint32_t r)          // after the scans
{
    if (P.dep[r + 1] == P.dep[r]) return;
    uint32_t m;
    const uint32_t *a = P.aln + 9 * P.aln_off[r];
    sel_kind(a, P.n_aln[r], m);
    SelDesc d;
    d.read = r; d.m = m; d.fixed_before = P.fixed[r];
    for (uint32_t j = 0; j < SEL_DESC_W; ++j) d.w[j] = j < m ? a[9 * j + 2] - a[9 * j + 1] + 1u : 0u;
    P.desc[P.dep[r]] = d;
}
HSA_HD void sel_chain(const SelParams &P)                          // one thread
{
    Rng48 rng{P.x0};
    const uint32_t n_dep = P.dep[P.n_reads];
    uint32_t prev_fixed = 0, v = 0;
    for (uint32_t d = 0; d < n_dep; ++d) {
        const SelDesc e = P.desc[d];
        rng.jump(e.fixed_before - prev_fixed); prev_fixed = e.fixed_before;
        SelPick pk;
        v += sel_draw(P.aln + 9 * P.aln_off[e.read], e.m, e.w, rng, pk);
        P.pick[d] = pk; P.vcum[d] = v;
    }
    rng.jump(P.fixed[P.n_reads] - prev_fixed);
    *P.x_end = rng.x;
}
HSA_HD void sel_finish_item(const SelParams &P, uint32_t r)
{
    SamRec s;
    s.type = s.strand = s.n_mm = s.n_gapo = s.n_gape = s.mapQ = 0; s.score = 0; s.sa = s.seq_id = s.ori_pos = s.occ_pos = s.c1 = s.c2 = 0;
    s.start = s.end = 0; s.n_cigar = s.cigar_off = s.nm = s.md_len = s.md_off = s.n_multi = 0;
    s.multi_off = P.slots[r];
    const uint32_t *a = P.aln + 9 * P.aln_off[r];
    const int32_t n_aln = P.n_aln[r];
    uint32_t m;
    const uint32_t kind = sel_kind(a, n_aln, m);
    if (kind == SEL_SPLICE) s.type = TYPE_SPLICING;
    else if (kind == SEL_NONE) s.type = TYPE_NO_MATCH;
    else {
        SelPick pk;
        if (kind == SEL_SINGLE) {
            const uint32_t d = P.dep[r];                            // dependent reads before this one
            Rng48 rng{P.x0};
            rng.jump((uint64_t)P.fixed[r] + (d ? P.vcum[d - 1] : 0u));
            if (sel_draw(a, 1u, nullptr, rng, pk) != 2u) *P.rare = 1u;      // r = 0 on the first draw: the prefix sums are off
        } else pk = P.pick[P.dep[r]];
        if (!sel_fill(a, n_aln, P.n_occ, m, pk, s, P.multi + s.multi_off)) *P.rare = 1u;
    }
    P.rec[r] = s;
}
HSA_HD void sel_sequential(const SelParams &P)                      // one thread: the whole batch by the sequential statement
{
    Rng48 rng{P.x0};
    for (uint32_t r = 0; r < P.n_reads; ++r) {
        SamRec s;
        s.type = s.strand = s.n_mm = s.n_gapo = s.n_gape = s.mapQ = 0; s.score = 0; s.sa = s.seq_id = s.ori_pos = s.occ_pos = s.c1 = s.c2 = 0;
        s.start = s.end = 0; s.n_cigar = s.cigar_off = s.nm = s.md_len = s.md_off = s.n_multi = 0;
        s.multi_off = P.slots[r];
        sam_select(P.aln + 9 * P.aln_off[r], P.n_aln[r], P.n_occ, rng, s, P.multi + s.multi_off);
        P.rec[r] = s;
    }
    *P.x_end = rng.x;
}

}   // namespace hsa

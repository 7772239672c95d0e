// hsa_core.cuh -- the inexact-search state machine, one logical worker per CUDA thread.
//
// The file is plain C++ with a handful of macros so that the SAME source is compiled
//   * by nvcc for sm_100a (the product: hsa_kernels.cu), and
//   * by g++ for tests/emu (a host emulation used ONLY by the CPU test-suite to check the logic
//     against the oracle where no GPU exists; it is never linked into the product library).
//
// What it computes (bit-exact with the reference):
//   occ lookups   BWTAllOccValue / BWTOccValue          BWT.c:793-837 / 682-719
//   width pass    bwt_cal_width (type 1)                bwtaln.c:73-116
//   search        bwt_match_gap                         bwtgap.c:118-331
//                 (gap_push/gap_pop :46-92, gap_shadow :94-105, bwt_match_exact 2BWT-Interface.c:365-388,
//                  BWTAllSARangesBackward_Bidirection 2BWT-Interface.c:235-271)
//   whole-read driver logic of bwa_cal_sa_reg_gap       bwtaln.c:303-360, 371-372
//   splice seed calls of bwt_splice_match               bwtgap.c:797-820
//
// How it differs from the reference in structure (not in results) -- see DESIGN.md:
//   * device index layout: one 32-byte sector per 64 BWT symbols = {occ[4] at block start, 4 packed words},
//     so one occ lookup touches exactly one sector instead of two (+ a major-table row);
//   * every step of every phase (width pass, node expansion, exact-match tail) is the same
//     "occ4 at k and at l+1" memory operation, so divergent workers still share the load/popcount code;
//   * the score-bucketed stack is a per-worker arena of 16-byte records threaded into per-bucket LIFO
//     lists, bucket heads in shared memory, non-empty buckets in a 128-bit register mask;
//   * children are pushed LAZILY: the (up to 4) deletion children of a node are one "family" record and
//     its (up to 4) mismatch children another, each holding the parent's interval and a 4-bit mask of the
//     children that exist.  A child is materialised -- one more occ4 pair at the parent's interval -- only
//     when it is actually popped AND survives the reference's pop-time pruning (bwtgap.c:161-173).  About
//     nine of ten pushed entries of the reference are never popped, so this removes most stack traffic;
//   * the lowest-score child (the exact-match extension) is never pushed: it is the next node popped by
//     construction (pushed last into the currently-lowest bucket), so it is carried in registers;
//   * children whose score already exceeds best_score + s_mm after the first hit are only counted
//     (they can never be popped, bwtgap.c:158-159), keeping the max_entries test exact.
#pragma once
#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define HSA_WARP_SYNC() __syncwarp()
#else
#define HSA_WARP_SYNC() ((void)0)
#endif

#if defined(__CUDACC__)
#define HSA_HD __host__ __device__ __forceinline__
#define HSA_D  __device__ __forceinline__
#else
#define HSA_HD inline
#define HSA_D  inline
#endif

namespace hsa {

struct alignas(16) u32x4 { uint32_t x, y, z, w; };
struct alignas(8)  u32x2 { uint32_t x, y; };

// ---- memory access wrappers ---------------------------------------------------------------------
HSA_HD u32x4 ld_ro4(const u32x4 *p)
{
#if defined(__CUDA_ARCH__)
    uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));   // read-only path, 128-bit
    u32x4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r;
#else
    return *p;
#endif
}
HSA_HD uint32_t ld_ro1(const uint32_t *p)
{
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
HSA_HD uint32_t ld_ro_u8(const uint8_t *p)
{
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
// bring one 32-byte sector towards L1 ahead of the load that needs it (no register cost)
HSA_HD void prefetch_sector(const void *p)
{
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}
HSA_HD int popc64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
HSA_HD int popc32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
HSA_HD int ffs64(uint64_t x)     // index of lowest set bit, x != 0
{
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}

HSA_HD int clz32(uint32_t x)     // x != 0
{
#if defined(__CUDA_ARCH__)
    return __clz((int)x);
#else
    return __builtin_clz(x);
#endif
}

// a[c] for a run-time c without forcing the array into local memory
HSA_HD uint32_t sel4(const uint32_t a[4], uint32_t c)
{
    return c == 0 ? a[0] : c == 1 ? a[1] : c == 2 ? a[2] : a[3];
}

// ---- index --------------------------------------------------------------------------------------
// Device layout of one BWT direction: block b = symbols [64b, 64b+64) of the '$'-less BWT string.
//   blocks[2b]   = occ of A,C,G,T in [0, 64b)
//   blocks[2b+1] = the four packed words of the block, first symbol in the two MSBs (BWT.c:954)
struct DevBwt {
    const u32x4 *blocks;
    uint32_t n_blocks;
    uint32_t text_length;
    uint32_t inverse_sa0;
    uint32_t cum[5];          // cumulativeFreq
};

// Reference layout (as loaded by BWTLoad): used by the re-pack kernel and by the rank parity test.
struct RefBwt {
    const uint32_t *bwt_code;
    const uint32_t *occ_value;
    const uint32_t *occ_major;
    uint32_t text_length;
    uint32_t inverse_sa0;
};

struct DevIndex { DevBwt fwd, rev; };

// counts of C,G,T (A by subtraction) among the first `t` (0..32) symbols of a 64-bit MSB-first group
HSA_HD void count_prefix64(uint64_t g, uint32_t t, uint32_t &c, uint32_t &gg, uint32_t &tt)
{
    uint64_t x = t ? (g >> (64u - 2u * t)) : 0ull;      // keep the first t symbols; zero (=A) padding
    uint64_t lo = x & 0x5555555555555555ull;
    uint64_t hi = (x >> 1) & 0x5555555555555555ull;
    uint32_t both = (uint32_t)popc64(lo & hi);
    tt += both;
    gg += (uint32_t)popc64(hi) - both;
    c  += (uint32_t)popc64(lo) - both;
}

// occ of all four symbols at SA-coordinate `index` on the device layout == BWTAllOccValue (BWT.c:793)
HSA_HD void occ4_from_sector(const u32x4 &cnt, const u32x4 &w, uint32_t off, uint32_t occ[4])
{
    uint32_t c = 0, g = 0, t = 0;
    uint32_t t0 = off < 32u ? off : 32u, t1 = off - t0;
    count_prefix64(((uint64_t)w.x << 32) | w.y, t0, c, g, t);
    count_prefix64(((uint64_t)w.z << 32) | w.w, t1, c, g, t);
    occ[0] = cnt.x + (off - c - g - t);
    occ[1] = cnt.y + c;
    occ[2] = cnt.z + g;
    occ[3] = cnt.w + t;
}

HSA_HD void occ4_dev(const DevBwt &b, uint32_t index, uint32_t occ[4])
{
    index -= (index > b.inverse_sa0);                   // BWT.c:804
    uint32_t blk = index >> 6;
    u32x4 cnt = ld_ro4(b.blocks + 2 * (size_t)blk);
    u32x4 w = ld_ro4(b.blocks + 2 * (size_t)blk + 1);
    occ4_from_sector(cnt, w, index & 63u, occ);
}

// ---- rank on the REFERENCE layout (BWT.c:793-837, 1018-1059, 532-679) -------------------------------
HSA_HD void count_pairs_ref(uint32_t w, uint32_t a, uint32_t b, uint32_t cnt[4])
{
    if (a >= b) return;
    uint32_t m = 0x55555555u;
    if (a > 0)  m &= 0xFFFFFFFFu >> (2 * a);
    if (b < 16) m &= ~(0xFFFFFFFFu >> (2 * b));
    uint32_t lo = w & m, hi = (w >> 1) & m;
    cnt[3] += (uint32_t)popc32(lo & hi);
    cnt[2] += (uint32_t)popc32(hi & ~lo);
    cnt[1] += (uint32_t)popc32(lo & ~hi);
    cnt[0] += (uint32_t)popc32(m & ~(lo | hi));
}

// occ at raw '$'-less stream position p (0..n) on the reference layout
HSA_HD void occ4_ref_raw(const RefBwt &b, uint32_t p, uint32_t occ[4])
{
    uint32_t e = (p + 127u) >> 8;                        // nearest 256-sample, may lie above p (BWT.c:813)
    uint32_t base = e << 8;
    uint32_t sh = (e & 1u) ? 0u : 16u;                   // even sample -> high half-word (BWT.c:1045-1047)
    for (int c = 0; c < 4; ++c) {
        uint32_t major = ld_ro1(b.occ_major + (size_t)(e >> 8) * 4 + c);
        uint32_t minor = (ld_ro1(b.occ_value + (size_t)(e >> 1) * 4 + c) >> sh) & 0xFFFFu;
        occ[c] = major + minor;
    }
    if (p == base) return;
    uint32_t lo = p < base ? p : base, hi = p < base ? base : p;
    uint32_t cnt[4] = {0, 0, 0, 0};
    for (uint32_t q = lo; q < hi;) {
        uint32_t wi = q >> 4, wend = (wi + 1) << 4;
        uint32_t bb = hi < wend ? (hi & 15u) : 16u;
        count_pairs_ref(ld_ro1(b.bwt_code + wi), q & 15u, bb, cnt);
        q = hi < wend ? hi : wend;
    }
    for (int c = 0; c < 4; ++c) occ[c] = p > base ? occ[c] + cnt[c] : occ[c] - cnt[c];
}

HSA_HD void occ4_ref(const RefBwt &b, uint32_t index, uint32_t occ[4])
{
    index -= (index > b.inverse_sa0);
    occ4_ref_raw(b, index, occ);
}

// ---- options / tasks ------------------------------------------------------------------------------
struct DevOpt {                 // the fields of gap_opt_t (bwtaln.h:133-143) that the path reads
    int s_mm, s_gapo, s_gape;
    int mode;
    int indel_end_skip, max_del_occ, max_entries;
    int max_diff, max_gapo, max_gape;
    int max_seed_diff, seed_len;
    int max_top2;
    int pad[3];
};

enum : uint32_t { MODE_GAPE = 0x01, MODE_LOGGAP = 0x04, MODE_NONSTOP = 0x10 };
enum : uint32_t { SEED_NONE = 0, SEED_TAIL = 1, SEED_ALIAS = 2 };
enum : uint32_t { KIND_TASKS = 0, KIND_WHOLE = 1, KIND_SEEDS = 2, KIND_WIDTH = 3 };
enum : uint32_t { ST_M = 0, ST_I = 1, ST_D = 2 };

// item status written next to n_aln
enum : uint8_t { STATUS_OK = 0, STATUS_NEED_STRICT = 1, STATUS_BAD_SCORE = 2, STATUS_OUT_FULL = 3, STATUS_PENDING = 0xFF };

struct Task {                   // == hsa_task_t (include/hsa_b200.h)
    uint64_t read_off;
    uint32_t read_len;
    uint32_t strand;
    uint32_t sub_off;
    uint32_t len;
    uint32_t wsrc_off;
    uint32_t seed_mode;
    uint32_t opt_idx;
    uint32_t reserved;
};

struct Hit {                    // worker-private hit record (2 x 16 bytes)
    uint32_t k, l, rev_k, rev_l;
    uint32_t counts;            // n_mm | n_gapo << 16 | n_gape << 24   (== word 0 of bwt_aln1_t)
    int32_t  score;
    uint32_t pad0, pad1;
};

// counters block (uint64 each)
enum { CNT_WORK = 0, CNT_ALN = 1, CNT_LOOKUPS = 2, CNT_STRICT = 3, CNT_BAD = 4, CNT_POPS = 5, CNT_STEPS = 6, CNT_EXTRA = 7, CNT_N = 8 };
// internal block behind the public statistics: 8 work-queue cursors, 8 pass-2 list counters, diagnostics
enum { CNT_CURSOR0 = CNT_N, CNT_NEXT0 = CNT_N + 8, CNT_DIAG_WARP_ITERS = CNT_N + 16, CNT_DIAG_MAX_ITEM_STEPS = CNT_N + 17, CNT_TOTAL = CNT_N + 24 };

struct Params {
    DevIndex ix;
    const uint8_t *codes;           // base codes, 0..3, N = 4
    uint32_t kind;
    uint32_t n_groups;              // work items: tasks (KIND_TASKS) or reads (KIND_WHOLE / KIND_SEEDS)
    const uint32_t *group_list;     // optional indirection (strict re-runs): work index -> group id
    const Task *tasks;              // KIND_TASKS
    const uint64_t *read_off;       // KIND_WHOLE / KIND_SEEDS
    const uint32_t *read_len;
    const DevOpt *opts;             // option table (device memory; staged to shared memory by the kernel)
    uint32_t n_opts;
    const uint16_t *len2opt;        // KIND_WHOLE: read length -> opts[] index ; KIND_SEEDS: unused (opts[0])
    uint32_t max_len;               // longest read in the batch
    int32_t  filter_max_n;          // KIND_WHOLE: local_opt.max_diff of bwtaln.c:273-274 (N filter :314-317)
    // worker-private scratch, indexed by worker slot
    u32x4 *arena;  void *links;  uint32_t arena_cap;         // stack records + per-bucket / free-list links (LinkT)
    u32x2 *width;  uint32_t width_stride;                     // [2*(max_len+1)] per worker: back, seed
    Hit *hits;     uint32_t hit_cap;
    uint32_t n_buckets;             // size of the score-indexed head table (<= 128)
    // outputs
    int32_t  *n_aln;                // [n_items]
    uint64_t *aln_off;              // [n_items]
    uint8_t  *status;               // [n_items]
    uint32_t *aln;                  // n_aln_cap x 9 words (== hsa_aln1_t)
    uint64_t  aln_cap;
    unsigned long long *counters;   // CNT_N
    uint32_t *strict_list;          // groups that must be re-run with larger capacities
    u32x2 *width_out;               // KIND_WIDTH: read r's len+1 entries at read_off[r] + r
    int32_t *bid_out;               // KIND_WIDTH: bwt_cal_width's return value per read
    // split pipeline (width kernel -> search kernel, one strand per pass); unused by the fused worker
    uint32_t pass;                  // KIND_WHOLE: 1 = reverse-complement strand pass, 2 = forward strand pass
    uint32_t group_base;            // first group of this chunk (work index 0)
    const uint32_t *n_groups_dev;   // if set, the number of work items is read from device memory (pass 2)
    u32x2 *item_width;              // [work item][item_width_stride]: width_back, then width_seed
    uint32_t item_width_stride;
    uint32_t *next_list;            // pass 1: reads without a hit are appended here ...
    uint32_t *next_count;           // ... and counted here (device memory)
    unsigned long long *cursor;     // atomic work-queue cursor of this launch
};

// ---- tasks ------------------------------------------------------------------------------------------
struct TaskDesc {                   // one bwt_match_gap call, resolved
    const uint8_t *rd;              // the read (forward strand, as given)
    uint32_t rd_len, strand, sub_off, len, wsrc_off, seed_mode, opt_idx, out_idx;
    int32_t aln_start, aln_end;     // start/end stamped on hits (seeds); -1 = leave zero
};

// base p of the strand-resolved read (seq_reverse(len, seq, 1): bwaseqio.c:73-90)
HSA_HD uint32_t task_base(const TaskDesc &t, uint32_t p)
{
    if (t.strand) { uint32_t c = ld_ro_u8(t.rd + (t.rd_len - 1 - p)); return c < 4 ? 3 - c : c; }
    return ld_ro_u8(t.rd + p);
}

// sub-task `sub` of group `gid`:  KIND_TASKS: the task itself;  KIND_WHOLE: sub 0 = reverse-complement strand,
// sub 1 = forward strand (bwtaln.c:343);  KIND_SEEDS: sub = strand*3 + segment (bwtgap.c:797-812);  KIND_WIDTH.
HSA_HD TaskDesc make_task(const Params &P, const DevOpt *opts, uint32_t gid, uint32_t sub)
{
    TaskDesc t;
    t.aln_start = t.aln_end = -1;
    if (P.kind == KIND_TASKS) {
        const Task &k = P.tasks[gid];
        t.rd = P.codes + k.read_off; t.rd_len = k.read_len; t.strand = k.strand; t.sub_off = k.sub_off; t.len = k.len;
        t.wsrc_off = k.wsrc_off; t.seed_mode = k.seed_mode; t.opt_idx = k.opt_idx; t.out_idx = gid;
    } else if (P.kind == KIND_WHOLE) {
        t.rd = P.codes + P.read_off[gid]; t.rd_len = P.read_len[gid];
        t.strand = 1 - sub;                                 // revcomp first (bwtaln.c:343)
        t.sub_off = 0; t.len = t.rd_len; t.wsrc_off = 0;
        t.opt_idx = P.len2opt[t.rd_len];
        t.seed_mode = (int32_t)t.rd_len > opts[t.opt_idx].seed_len ? SEED_TAIL : SEED_NONE;   // bwtaln.c:332,344
        t.out_idx = gid;
    } else if (P.kind == KIND_WIDTH) {                      // bwt_cal_width alone (type 1), read as given
        t.rd = P.codes + P.read_off[gid]; t.rd_len = P.read_len[gid];
        t.strand = 0; t.sub_off = 0; t.len = t.rd_len; t.wsrc_off = 0; t.seed_mode = SEED_NONE; t.opt_idx = 0; t.out_idx = gid;
    } else {                                                // KIND_SEEDS, bwtgap.c:797-812
        t.rd = P.codes + P.read_off[gid]; t.rd_len = P.read_len[gid];
        uint32_t seg = sub % 3, sl = t.rd_len / 3;
        t.strand = sub / 3;
        t.len = sl + (seg == 2 ? t.rd_len % 3 : 0);
        t.sub_off = seg * sl;
        t.wsrc_off = 0;                                     // width on the read PREFIX (bwtgap.c:807-808)
        t.seed_mode = SEED_ALIAS;                           // bwtgap.c:809
        t.opt_idx = 0;
        t.out_idx = gid * 6 + sub;
        t.aln_start = (int32_t)t.sub_off; t.aln_end = (int32_t)(t.sub_off + t.len - 1);
    }
    return t;
}

// per-read filters of bwa_cal_sa_reg_gap (bwtaln.c:314-317 too many N, 324-325 poly-A / poly-T prefix)
HSA_HD bool read_filtered(const uint8_t *r, uint32_t L, int32_t max_n)
{
    uint32_t nn = 0;
    bool pa = L >= 15, pt = L >= 15;
    for (uint32_t j = 0; j < L; ++j) {
        uint32_t c = ld_ro_u8(r + j);
        nn += c > 3;
        if (j < 15) { pa = pa && c == 0; pt = pt && c == 3; }
    }
    return (int32_t)nn > max_n || pa || pt;
}

// occ of ONE symbol at SA coordinate `index` (BWTOccValue, BWT.c:682-719) on the device layout
HSA_HD uint32_t occ1_dev(const DevBwt &b, uint32_t index, uint32_t c)
{
    index -= (index > b.inverse_sa0);
    const u32x4 cnt = ld_ro4(b.blocks + 2 * (size_t)(index >> 6));
    const u32x4 w = ld_ro4(b.blocks + 2 * (size_t)(index >> 6) + 1);
    const uint32_t off = index & 63u, t0 = off < 32u ? off : 32u, t1 = off - t0;
    // symbols equal to c among the first t of a 64-bit group: xor with c replicated, then both bits zero
    const uint64_t pat = 0x5555555555555555ull * c;
    const uint64_t g0 = (((uint64_t)w.x << 32) | w.y) ^ pat, g1 = (((uint64_t)w.z << 32) | w.w) ^ pat;
    const uint64_t m0 = ~(g0 | (g0 >> 1)) & 0x5555555555555555ull, m1 = ~(g1 | (g1 >> 1)) & 0x5555555555555555ull;
    const uint64_t k0 = t0 ? (~0ull << (64u - 2u * t0)) : 0ull, k1 = t1 ? (~0ull << (64u - 2u * t1)) : 0ull;
    const uint32_t base = c == 0 ? cnt.x : c == 1 ? cnt.y : c == 2 ? cnt.z : cnt.w;
    return base + (uint32_t)popc64(m0 & k0) + (uint32_t)popc64(m1 & k1);
}

// bwt_cal_width, type 1 (bwtaln.c:73-97, 113-114): `n` bases starting at strand-resolved position `src`.
// Returns bid; *lookups += the BWTOccValue calls the reference issues.
HSA_HD int32_t cal_width_dev(const DevIndex &ix, const TaskDesc &t, uint32_t src, uint32_t n, u32x2 *dst, uint64_t &lookups)
{
    uint32_t k = 0, l = ix.fwd.text_length;
    int32_t bid = 0;
    for (uint32_t j = 0; j < n; ++j) {
        const uint32_t c = task_base(t, src + j);
        if (c < 4) {                                        // BWTSARangeForeward, 2BWT-Interface.c:121-132
            const uint32_t a = occ1_dev(ix.rev, k, c), b = occ1_dev(ix.rev, l + 1, c);
            k = ix.fwd.cum[c] + a + 1;
            l = ix.fwd.cum[c] + b;
            lookups += 2;
        }
        if (k > l || c > 3) { k = 0; l = ix.fwd.text_length; ++bid; }
        u32x2 v; v.x = l - k + 1; v.y = (uint32_t)bid;
        dst[j] = v;
    }
    u32x2 last; last.x = 0; last.y = (uint32_t)(++bid);
    dst[n] = last;
    return bid;
}

// Split pipeline, width kernel body for work item `w`: resolves the task, applies the whole-read filters in
// pass 1, writes width_back (+ width_seed) of the item.  A filtered read is marked by width[0].bid = ~0.
HSA_HD void width_item(const Params &P, const DevOpt *opts, uint32_t w, uint64_t &lookups)
{
    uint32_t gid, sub;
    if (P.kind == KIND_SEEDS) { gid = P.group_base + w / 6; sub = w % 6; }
    else if (P.kind == KIND_WHOLE) { gid = P.pass == 2 ? P.group_list[w] : P.group_base + w; sub = P.pass == 2 ? 1 : 0; }
    else { gid = P.group_base + w; sub = 0; }
    const TaskDesc t = make_task(P, opts, gid, sub);
    if (P.kind == KIND_WIDTH) {                              // hsa_cal_width_batch: straight to the caller's layout
        P.bid_out[gid] = cal_width_dev(P.ix, t, 0, t.len, P.width_out + P.read_off[gid] + gid, lookups);
        return;
    }
    u32x2 *wb = P.item_width + (size_t)w * P.item_width_stride;
    if (P.kind == KIND_WHOLE && P.pass == 1 && read_filtered(t.rd, t.rd_len, P.filter_max_n)) {
        u32x2 v; v.x = 0; v.y = 0xFFFFFFFFu;
        wb[0] = v;
        P.n_aln[t.out_idx] = 0; P.aln_off[t.out_idx] = 0; P.status[t.out_idx] = STATUS_OK;
        return;
    }
    uint64_t lk = 0;
    cal_width_dev(P.ix, t, t.wsrc_off, t.len, wb, lk);
    if (t.seed_mode == SEED_TAIL) {
        const uint32_t sl = (uint32_t)opts[t.opt_idx].seed_len;
        cal_width_dev(P.ix, t, t.sub_off + (t.len - sl), sl, wb + (P.max_len + 1), lk);
    }
    // the item's last slot carries its lookup count; the search worker adds it when the item completes, so
    // items that are re-run with the large-capacity kernel are not counted twice
    u32x2 c; c.x = (uint32_t)lk; c.y = 0;
    wb[P.item_width_stride - 1] = c;
    (void)lookups;
}

// ---- the worker -------------------------------------------------------------------------------------
// Stack record (16 bytes) = {k, l, rev_l, meta}; meta = i:12 | diff:1 | state:2 | mm:5 | go:4 | ge:5 | kind:2.
//   kind PLAIN : one search node (gap_entry_t, bwtaln.h:52-58), rev_k == rev_l - (l - k) always holds.
//   kind FAM_D : the deletion children (bwtgap.c:276-282 / 292-298) of the node stored in the record:
//                k,l,rev_l,mm,go,ge,state are the PARENT's, i is the children's i (parent's i before --i).
//   kind FAM_MM: the mismatch children (bwtgap.c:303-313) of the parent; i is the parent's i before --i.
// A family's 4-bit child mask lives in the top bits of the record's link word (LinkT: 16 bit = 12-bit
// link + mask for the fast kernel, 32 bit = 28-bit link + mask for the large-capacity kernel).
enum : uint32_t { KIND_PLAIN = 0, KIND_FAM_D = 1, KIND_FAM_MM = 2 };

// FUSED = true : one worker runs width passes and both strands of its group itself (large-capacity re-runs,
//                 the host emulation's reference flow);
// FUSED = false: split pipeline -- widths come from the width kernel's per-item buffer, one task per work item.
template <typename LinkT, bool FUSED>
struct Worker {
    static constexpr uint32_t LINK_BITS = sizeof(LinkT) * 8 - 4;
    static constexpr uint32_t NIL = (1u << LINK_BITS) - 1u;

    // environment
    const Params &P;
    uint32_t slot;                  // worker slot -> scratch
    LinkT *heads;                   // bucket heads of this worker: heads[b * head_stride]
    uint32_t head_stride;
    const DevOpt *opts;             // option table as seen by this worker (shared memory on the device)

    // group / task bookkeeping
    uint32_t gid, n_sub, sub;       // current group, number of sub-tasks in it, current sub-task
    uint32_t work;                  // work-queue index of the current item (split pipeline: width buffer slot)
    bool short_circuit;
    // current task
    const uint8_t *rd;              // the read
    uint32_t rd_len, strand, sub_off, len, wsrc_off, seed_mode, opt_idx, out_idx;
    int32_t aln_start, aln_end;     // start/end stamped on hits (seeds); -1 = leave zero
    // phase
    enum Phase : uint32_t { IDLE = 0, WSEED = 1, WBACK = 2, SEARCH = 3, RETIRED = 4 };
    uint32_t phase;
    // width pass state
    uint32_t wk, wl, wj, wn; int32_t wbid; uint32_t wsrc; u32x2 *wdst;
    // search state
    uint64_t mask0, mask1;          // non-empty buckets
    uint32_t n_live, n_phantom;     // entries the reference would hold: stored (families count their children); counted-only
    uint32_t top, free_head;        // arena bump pointer and free list
    int32_t best_score, max_diff, best_cnt;
    uint32_t n_hits;
    bool failed;                    // task ran out of capacity -> group must be re-run strict
    uint8_t fail_code;
    // candidate node
    bool have, direct, exact;
    uint32_t pend;                  // 0: interval known; else KIND_FAM_*: ck/cl/crl are the PARENT's, child pend_j
    uint32_t pend_j;
    uint32_t ck, cl, crl;           // k, l, rev_l  (rev_k == rev_l - (l - k) for every node ever created)
    uint32_t ci, c_mm, c_gapo, c_gape, c_state, c_diff;
    uint32_t ci_at_pop;             // e.info & 0xffff of the entry being processed (for last_diff_pos)
    uint32_t zflags;                // which of k,l,rev_k,rev_l were zero when bwt_match_exact was entered
    int32_t m_cur, m_seed_cur;      // remaining diffs of the candidate (bwtgap.c:161-171), set by acquire_vet()
    bool look;                      // this lane issues an occ4 pair in the current iteration
    bool ending;                    // the current task is over (handled once, at the end of the iteration)
    // statistics
    uint64_t lookups, lookups_group, pops, steps, extra;
    uint64_t steps_item0; uint32_t max_item_steps;  // diagnostics: iterations of the longest single item

    HSA_HD Worker(const Params &p, uint32_t slot_, LinkT *heads_, uint32_t stride_, const DevOpt *opts_)
        : P(p), slot(slot_), heads(heads_), head_stride(stride_), opts(opts_), phase(IDLE),
          lookups(0), lookups_group(0), pops(0), steps(0), extra(0), steps_item0(0), max_item_steps(0) {}

    HSA_HD bool idle() const { return phase == IDLE; }
    HSA_HD bool retired() const { return phase == RETIRED; }
    HSA_HD void retire() { phase = RETIRED; }

    // base p of the strand-resolved read (seq_reverse(len, seq, 1): bwaseqio.c:73-90)
    HSA_HD uint32_t base_at(uint32_t p) const
    {
        if (strand) { uint32_t c = ld_ro_u8(rd + (rd_len - 1 - p)); return c < 4 ? 3 - c : c; }
        return ld_ro_u8(rd + p);
    }
    HSA_HD u32x2 *wback() const
    {
        return FUSED ? P.width + (size_t)slot * P.width_stride : P.item_width + (size_t)work * P.item_width_stride;
    }
    HSA_HD u32x2 *wseed() const
    {
        if (seed_mode == SEED_ALIAS) return wback();
        return FUSED ? P.width + (size_t)slot * P.width_stride + (P.max_len + 1)
                     : P.item_width + (size_t)work * P.item_width_stride + (P.max_len + 1);
    }
    HSA_HD u32x4 *arena() const { return P.arena + (size_t)slot * P.arena_cap; }
    HSA_HD LinkT *links() const { return reinterpret_cast<LinkT *>(P.links) + (size_t)slot * P.arena_cap; }

    // ---------------------------------------------------------------- group / task setup
    HSA_HD void start_group(uint32_t work_idx)
    {
        work = work_idx;
        lookups_group = 0;
        steps_item0 = steps;
        failed = false; fail_code = STATUS_OK;
        if (FUSED) {
            gid = P.group_list ? P.group_list[work_idx] : P.group_base + work_idx;
            sub = 0;
            if (P.kind == KIND_WHOLE) {
                n_sub = 2; short_circuit = true;
                if (read_filtered(P.codes + P.read_off[gid], P.read_len[gid], P.filter_max_n)) {
                    finish_item(gid, 0, 0); phase = IDLE; return;
                }
            } else if (P.kind == KIND_SEEDS) { n_sub = 6; short_circuit = false; }
            else { n_sub = 1; short_circuit = false; }
            setup_task();
            return;
        }
        // split pipeline: exactly one task per work item; the width kernel has already run for it
        short_circuit = false;
        if (P.kind == KIND_SEEDS) { gid = P.group_base + work_idx / 6; sub = work_idx % 6; }
        else if (P.kind == KIND_WHOLE) { gid = P.pass == 2 ? P.group_list[work_idx] : P.group_base + work_idx; sub = P.pass == 2 ? 1 : 0; }
        else { gid = P.group_base + work_idx; sub = 0; }
        n_sub = sub + 1;
        load_task();
        if (wback()[0].y == 0xFFFFFFFFu) { phase = IDLE; return; }      // filtered by the width kernel (pass 1)
        begin_search();
    }

    HSA_HD void load_task()
    {
        const TaskDesc t = make_task(P, opts, gid, sub);
        rd = t.rd; rd_len = t.rd_len; strand = t.strand; sub_off = t.sub_off; len = t.len; wsrc_off = t.wsrc_off;
        seed_mode = t.seed_mode; opt_idx = t.opt_idx; out_idx = t.out_idx; aln_start = t.aln_start; aln_end = t.aln_end;
    }

    HSA_HD void setup_task()            // FUSED only
    {
        load_task();
        if (seed_mode == SEED_TAIL) {
            uint32_t sl = (uint32_t)opts[opt_idx].seed_len;
            begin_width(WSEED, sub_off + (len - sl), sl, wseed());
        } else begin_width(WBACK, wsrc_off, len, wback());
    }

    HSA_HD void begin_width(uint32_t ph, uint32_t src, uint32_t n, u32x2 *dst)
    {
        // n >= 1 always: the host rejects empty reads / tasks (no recursion with end_width, so that the
        // whole worker stays in registers)
        phase = ph; wk = 0; wl = P.ix.fwd.text_length; wj = 0; wn = n; wbid = 0; wsrc = src; wdst = dst;
    }

    HSA_HD void end_width()
    {
        u32x2 last; last.x = 0; last.y = (uint32_t)(++wbid);    // bwtaln.c:113-114
        wdst[wn] = last;
        if (phase == WSEED) begin_width(WBACK, wsrc_off, len, wback());
        else if (P.kind == KIND_WIDTH) {
            u32x2 *dst = P.width_out + P.read_off[gid] + gid;
            for (uint32_t j = 0; j <= wn; ++j) dst[j] = wdst[j];
            P.bid_out[gid] = wbid;
            lookups += lookups_group;
            phase = IDLE;
        } else begin_search();
    }

    HSA_HD void begin_search()
    {
        const DevOpt &o = opts[opt_idx];
        phase = SEARCH;
        mask0 = mask1 = 0; n_live = 0; n_phantom = 0; top = 0; free_head = NIL;
        best_score = (o.max_diff + 1) * o.s_mm + (o.max_gapo + 1) * o.s_gapo + (o.max_gape + 1) * o.s_gape; // :128
        max_diff = o.max_diff; best_cnt = 0; n_hits = 0;
        // root (bwtgap.c:142) is the first node popped; carry it directly
        have = true; direct = true; exact = false; pend = 0; pend_j = 0;
        ck = 0; cl = P.ix.fwd.text_length; crl = P.ix.fwd.text_length;
        ci = len; c_mm = c_gapo = c_gape = 0; c_state = ST_M; c_diff = 0; zflags = 0; ci_at_pop = len;
    }

    // ---------------------------------------------------------------- stack
    HSA_HD bool bucket_nonempty(uint32_t b) const { return b < 64 ? (mask0 >> b) & 1ull : (mask1 >> (b - 64)) & 1ull; }

    // one record of `kind` standing for `n_children` reference entries (1 for PLAIN), all of score `sc`
    HSA_HD void push_record(int32_t sc, uint32_t kind, uint32_t cmask, uint32_t n_children,
                            uint32_t i, uint32_t k, uint32_t l, uint32_t rl,
                            uint32_t mm, uint32_t go, uint32_t ge, uint32_t state, uint32_t is_diff)
    {
        const DevOpt &o = opts[opt_idx];
        // bwtgap.c:158-159: once a hit exists, anything above best_score + s_mm can never be popped -> count only
        if (n_hits && !(o.mode & MODE_NONSTOP) && sc > best_score + o.s_mm) { n_phantom += n_children; return; }
        if ((uint32_t)sc >= P.n_buckets) { failed = true; fail_code = STATUS_BAD_SCORE; return; }
        uint32_t s;
        LinkT *lk = links();
        if (free_head != NIL) { s = free_head; free_head = (uint32_t)lk[s] & NIL; }
        else if (top < P.arena_cap) s = top++;
        else { failed = true; fail_code = STATUS_NEED_STRICT; return; }
        u32x4 e;
        e.x = k; e.y = l; e.z = rl;
        e.w = i | is_diff << 12 | state << 13 | mm << 15 | go << 20 | ge << 24 | kind << 29;
        arena()[s] = e;
        uint32_t b = (uint32_t)sc;
        uint32_t prev = bucket_nonempty(b) ? (uint32_t)heads[b * head_stride] : NIL;
        lk[s] = (LinkT)(prev | cmask << LINK_BITS);
        heads[b * head_stride] = (LinkT)s;
        if (b < 64) mask0 |= 1ull << b; else mask1 |= 1ull << (b - 64);
        n_live += n_children;
    }

    // gap_pop (bwtgap.c:80-92): last entry of the lowest non-empty bucket -> candidate registers.
    // A family record yields its last-pushed remaining child and stays in place until it is empty.
    HSA_HD int32_t pop()
    {
        uint32_t b = mask0 ? (uint32_t)ffs64(mask0) : 64u + (uint32_t)ffs64(mask1);
        LinkT *lk = links();
        uint32_t s = heads[b * head_stride];
        u32x4 e = arena()[s];
        uint32_t lw = (uint32_t)lk[s];
        uint32_t nx = lw & NIL, cmask = lw >> LINK_BITS;
        uint32_t kind = e.w >> 29;
        uint32_t j = 0;
        if (kind != KIND_PLAIN) {
            j = 31u - (uint32_t)clz32(cmask);           // last pushed child first
            cmask &= ~(1u << j);
        }
        if (kind == KIND_PLAIN || cmask == 0) {
            if (nx == NIL) { if (b < 64) mask0 &= ~(1ull << b); else mask1 &= ~(1ull << (b - 64)); }
            else heads[b * head_stride] = (LinkT)nx;
            lk[s] = (LinkT)free_head; free_head = s;
        } else lk[s] = (LinkT)(nx | cmask << LINK_BITS);
        --n_live;
        ck = e.x; cl = e.y; crl = e.z;
        uint32_t ei = e.w & 0xFFFu, est = (e.w >> 13) & 3u;
        c_mm = (e.w >> 15) & 31u; c_gapo = (e.w >> 20) & 15u; c_gape = (e.w >> 24) & 31u;
        if (kind == KIND_PLAIN) { ci = ei; c_diff = (e.w >> 12) & 1u; c_state = est; pend = 0; }
        else if (kind == KIND_FAM_D) {
            ci = ei; c_diff = 1; c_state = ST_D; pend = KIND_FAM_D; pend_j = j;
            if (est == ST_M) ++c_gapo; else ++c_gape;           // gap open (bwtgap.c:281) / extension (:297)
        } else {
            ci = ei - 1; c_diff = 1; c_state = ST_M; pend = KIND_FAM_MM; pend_j = j;
            ++c_mm;                                             // bwtgap.c:312
        }
        have = true; direct = false; exact = false;
        ++pops;
        return (int32_t)b;
    }

    // ---------------------------------------------------------------- hits
    // action for found hits, bwtgap.c:188-241.  returns false when the search must stop (top2b break).
    HSA_HD bool record_hit(uint32_t k, uint32_t l, uint32_t rk, uint32_t rl)
    {
        const DevOpt &o = opts[opt_idx];
        int32_t score = (int32_t)c_mm * o.s_mm + (int32_t)c_gapo * o.s_gapo + (int32_t)c_gape * o.s_gape;
        bool do_add = true;
        Hit *hs = P.hits + (size_t)slot * P.hit_cap;
        if (n_hits == 0) {
            best_score = score;
            int32_t best_diff = (int32_t)(c_mm + c_gapo);
            if (o.mode & MODE_GAPE) best_diff += (int32_t)c_gape;
            if (!(o.mode & MODE_NONSTOP)) max_diff = (best_diff + 1 > o.max_diff) ? o.max_diff : best_diff + 1;
        }
        if (score == best_score) best_cnt = (int32_t)((uint32_t)best_cnt + (l - k + 1));
        else if (best_cnt > o.max_top2) return false;
        if (c_gapo) {
            for (uint32_t j = 0; j < n_hits; ++j)
                if (hs[j].k == k && hs[j].l == l) { do_add = false; break; }
        }
        if (do_add) {
            // gap_shadow (bwtgap.c:94-105) on width_back, in place
            uint32_t x = l - k + 1, ldp = c_diff ? ci_at_pop : 0u, jj = 0;
            u32x2 *w = wback();
            for (uint32_t i = 0; i < ldp; ++i) {
                u32x2 v = w[i];
                if (v.x > x) { v.x -= x; w[i] = v; }
                else if (v.x == x) { v.y = 1; v.x = P.ix.fwd.text_length - (++jj); w[i] = v; }
            }
            if (n_hits >= P.hit_cap) { failed = true; fail_code = STATUS_NEED_STRICT; return false; }
            Hit h;
            h.k = k; h.l = l; h.rev_k = rk; h.rev_l = rl;
            h.counts = c_mm | c_gapo << 16 | c_gape << 24;
            h.score = score; h.pad0 = h.pad1 = 0;
            hs[n_hits++] = h;
        }
        return true;
    }

    // write the current task's hits to the output arena
    HSA_HD void finish_item(uint32_t item, uint32_t n, uint32_t strand_stamp)
    {
        uint64_t off = 0;
        uint8_t st = STATUS_OK;
        if (n) {
#if defined(__CUDA_ARCH__)
            off = atomicAdd(&P.counters[CNT_ALN], (unsigned long long)n);
#else
            off = P.counters[CNT_ALN]; P.counters[CNT_ALN] += n;
#endif
            if (off + n > P.aln_cap) { st = STATUS_OUT_FULL; n = 0; }
            const Hit *hs = P.hits + (size_t)slot * P.hit_cap;
            for (uint32_t j = 0; j < n; ++j) {
                uint32_t *w = P.aln + (off + j) * 9;
                Hit h = hs[j];
                w[0] = h.counts; w[1] = h.k; w[2] = h.l; w[3] = h.rev_k; w[4] = h.rev_l;
                w[5] = strand_stamp << 30;                       // type:30 = 0, strand:2
                int32_t s0 = 0, e0 = 0;
                if (P.kind == KIND_SEEDS) { s0 = aln_start; e0 = aln_end; }                 // bwtgap.c:816-819
                else if (P.kind == KIND_WHOLE && j == 0) { s0 = 0; e0 = (int32_t)rd_len - 1; } // bwtaln.c:371-372
                w[6] = (uint32_t)s0; w[7] = (uint32_t)e0; w[8] = (uint32_t)h.score;
            }
        }
        P.n_aln[item] = (int32_t)n;
        P.aln_off[item] = off;
        P.status[item] = st;
    }

    HSA_HD void fail_group()
    {
        // discard everything this group produced; the host re-runs it with the large-capacity (fused) kernel
        uint32_t first = out_idx, cnt = 1;
        if (FUSED && P.kind == KIND_SEEDS) { first = gid * 6; cnt = 6; }
        for (uint32_t j = 0; j < cnt; ++j) { P.n_aln[first + j] = 0; P.aln_off[first + j] = 0; P.status[first + j] = fail_code; }
#if defined(__CUDA_ARCH__)
        unsigned long long idx = atomicAdd(&P.counters[fail_code == STATUS_NEED_STRICT ? CNT_STRICT : CNT_BAD], 1ull);
#else
        unsigned long long idx = P.counters[fail_code == STATUS_NEED_STRICT ? CNT_STRICT : CNT_BAD]++;
#endif
        if (fail_code == STATUS_NEED_STRICT && P.strict_list) P.strict_list[idx] = gid;
        phase = IDLE;
    }

    HSA_HD void end_task()
    {
        if ((uint32_t)(steps - steps_item0) > max_item_steps) max_item_steps = (uint32_t)(steps - steps_item0);
        if (failed) { fail_group(); return; }
        if (!FUSED) {
            const uint64_t mine = lookups_group + wback()[P.item_width_stride - 1].x;   // + the width kernel's share
            phase = IDLE;
            if (P.kind == KIND_WHOLE && P.pass == 1 && n_hits == 0) {
                // bwtaln.c:351-358: nothing on the reverse-complement strand -> the forward strand is searched (pass 2).
                // The read's pass-1 lookup count rides in aln_off[] until pass 2 completes, so that a read which
                // is later re-run with the large-capacity kernel is not counted twice.
                P.aln_off[out_idx] = mine;
#if defined(__CUDA_ARCH__)
                uint32_t idx = atomicAdd(P.next_count, 1u);
#else
                uint32_t idx = (*P.next_count)++;
#endif
                P.next_list[idx] = gid;
                return;
            }
            lookups += mine + ((P.kind == KIND_WHOLE && P.pass == 2) ? P.aln_off[out_idx] : 0ull);
            finish_item(out_idx, n_hits, strand);
            return;
        }
        bool last = (sub + 1 == n_sub) || (short_circuit && n_hits);
        if (P.kind != KIND_WHOLE || n_hits || last) finish_item(out_idx, n_hits, strand);
        if (last) { lookups += lookups_group; phase = IDLE; return; }
        ++sub;
        setup_task();
    }

    // ---------------------------------------------------------------- one iteration
    // MAX_POPS bounds the number of stack pops tried per iteration when candidates keep dying.
    // prefetch the two index sectors the candidate (ck, cl) will need and its width entries
    HSA_HD void prefetch_candidate()
    {
        const DevBwt &B = P.ix.fwd;
        uint32_t pk = ck, pl = cl + 1;
        pk -= (pk > B.inverse_sa0); pl -= (pl > B.inverse_sa0);
        prefetch_sector(B.blocks + 2 * (size_t)(pk >> 6));
        prefetch_sector(B.blocks + 2 * (size_t)(pl >> 6));
    }

    // One trip of "get a candidate and apply the reference's pop-time tests" (bwtgap.c:144-186).
    // Sets `look` when the candidate needs its occ4 pair (expansion, exact-match step, or family child
    // materialisation); sets `ending` when the task is over; otherwise the candidate died or was a hit and
    // another trip may follow.  (end_task / record_hit have exactly one call site each to keep the loop small.)
    HSA_HD void acquire_vet()
    {
        const DevOpt &o = opts[opt_idx];
        if (!have) {
            // loop top of bwtgap.c:144-159
            if (n_live == 0 || (int64_t)n_live + n_phantom > (int64_t)o.max_entries) { ending = true; return; }
            int32_t sc = pop();
            if (!(o.mode & MODE_NONSTOP) && sc > best_score + o.s_mm) { ending = true; return; }
            prefetch_candidate();                           // overlaps with the width load of the bound test below
        } else if (direct && !exact) {
            // the carried child would have been pushed and popped: same loop-top test, entry included
            if ((int64_t)n_live + n_phantom + 1 > (int64_t)o.max_entries) { ending = true; return; }
            direct = false;
        }
        bool hit = false;
        uint32_t hk = ck, hl = cl, hrk = crl - (cl - ck), hrl = crl;
        if (exact) {
            // still inside bwt_match_exact (2BWT-Interface.c:365-388)
            if (ci != 0) { look = true; return; }
            // matched down to the first base: write-back quirk of 2BWT-Interface.c:383-386
            if (zflags & 1u) hk = 0;
            if (zflags & 2u) hl = 0;
            if (zflags & 4u) hrk = 0;
            if (zflags & 8u) hrl = 0;
            hit = true;
        } else {
            m_cur = max_diff - (int32_t)(c_mm + c_gapo);                               // :161-164
            if (o.mode & MODE_GAPE) m_cur -= (int32_t)c_gape;
            if (m_cur < 0) { have = false; return; }
            if (seed_mode != SEED_NONE) {
                m_seed_cur = o.max_seed_diff - (int32_t)(c_mm + c_gapo);
                if (o.mode & MODE_GAPE) m_seed_cur -= (int32_t)c_gape;
            }
            if (ci > 0 && m_cur < (int32_t)wback()[ci - 1].y) { have = false; return; } // :172-173
            if (pend) { look = true; return; }              // survived pop-time pruning: materialise it
            ci_at_pop = ci;
            if (ci == 0) hit = true;                                                   // :177-179
            else {
                if (m_cur == 0 && (c_state == ST_M || (o.mode & MODE_GAPE) || (int32_t)c_gape == o.max_gape)) { // :180
                    exact = true;
                    zflags = (ck == 0) | (cl == 0) << 1 | (hrk == 0) << 2 | (crl == 0) << 3;
                }
                look = true;
                return;
            }
        }
        if (hit) {
            have = false;
            if (!record_hit(hk, hl, hrk, hrl) || failed) ending = true;
        }
    }

    // ---------------------------------------------------------------- one iteration
    // Every lane of a warp calls iterate() every loop trip (idle and retired lanes too) and passes the same
    // warp barriers: the divergent candidate handling above is followed by a re-convergence point, so the
    // occ4 loads and popcounts below are issued once for all lanes that need them.  MAX_POPS bounds the
    // number of acquire/vet trips per iteration when candidates keep dying.
    template <int MAX_POPS>
    HSA_HD void iterate()
    {
        const bool active = !(phase == IDLE || phase == RETIRED);
        if (active) ++steps;
        look = false; ending = false;
#if defined(__CUDACC__)
#pragma unroll 1
#endif
        for (int t = 0; t < MAX_POPS; ++t) {
            HSA_WARP_SYNC();
            if (active && !look && !ending && phase == SEARCH) acquire_vet();
        }
        if (FUSED && active && (phase == WSEED || phase == WBACK)) look = true;
        HSA_WARP_SYNC();
        if (look) lookup_and_step();
        HSA_WARP_SYNC();
        if (ending) end_task();
    }

    HSA_HD void lookup_and_step()
    {
        const DevOpt &o = opts[opt_idx];
        const int32_t m = m_cur, m_seed = m_seed_cur;
        // ---- the one memory operation every phase shares: occ4 at k and at l + 1 -------------------
        const bool searching = !FUSED || phase == SEARCH;
        const DevBwt &B = searching ? P.ix.fwd : P.ix.rev;
        uint32_t pk = searching ? ck : wk, pl = (searching ? cl : wl) + 1;
        pk -= (pk > B.inverse_sa0); pl -= (pl > B.inverse_sa0);
        u32x4 kc = ld_ro4(B.blocks + 2 * (size_t)(pk >> 6)), kw = ld_ro4(B.blocks + 2 * (size_t)(pk >> 6) + 1);
        u32x4 lc = ld_ro4(B.blocks + 2 * (size_t)(pl >> 6)), lw = ld_ro4(B.blocks + 2 * (size_t)(pl >> 6) + 1);

        if (FUSED && phase != SEARCH) {
            // ---- bwt_cal_width step (bwtaln.c:86-97) -------------------------------------------------
            uint32_t c = base_at(wsrc + wj);
            uint32_t oL[4], oR[4];
            occ4_from_sector(kc, kw, pk & 63u, oL);
            occ4_from_sector(lc, lw, pl & 63u, oR);
            if (c < 4) {
                uint32_t a = sel4(oL, c), b = sel4(oR, c);
                wk = P.ix.fwd.cum[c] + a + 1;                  // BWTSARangeForeward, 2BWT-Interface.c:121-132
                wl = P.ix.fwd.cum[c] + b;
                lookups_group += 2;
            }
            if (wk > wl || c > 3) { wk = 0; wl = P.ix.fwd.text_length; ++wbid; }
            u32x2 v; v.x = wl - wk + 1; v.y = (uint32_t)wbid;
            wdst[wj] = v;
            if (++wj == wn) end_width();
            return;
        }

        // ---- materialise a family child / expand a node (bwtgap.c:244-325) / one bwt_match_exact step ---
        // i: index of the base the lookup extends by.  For a pending family child the lookup is the PARENT's:
        // FAM_D children were produced by the parent's step at i = ci - 1 ... ci is the child's i (== parent's
        // pre-decrement i), FAM_MM children have ci == parent's post-decrement i.
        uint32_t i = pend == KIND_FAM_MM ? ci : ci - 1;
        uint32_t sc_ = base_at(sub_off + i);
        u32x2 w0, w1, s0, s1;                                   // width[i-1], width[i], width_seed[ii-1], width_seed[ii]
        bool use_seed = false;
        w0.x = w0.y = w1.x = w1.y = s0.x = s0.y = s1.x = s1.y = 0;
        if (!exact && !pend && i > 0) {
            const u32x2 *w = wback();
            w0 = w[i - 1]; w1 = w[i];
            if (seed_mode != SEED_NONE) {
                int32_t ii = seed_mode == SEED_ALIAS ? (int32_t)i : (int32_t)i - ((int32_t)len - o.seed_len);   // :253
                if (ii > 0) { const u32x2 *ws = wseed(); s0 = ws[ii - 1]; s1 = ws[ii]; use_seed = true; }
            }
        }
        uint32_t oL[4], oR[4], sk[4], sl[4], rsl[4];
        occ4_from_sector(kc, kw, pk & 63u, oL);
        occ4_from_sector(lc, lw, pl & 63u, oR);
        {
            // BWTAllSARangesBackward_Bidirection, 2BWT-Interface.c:235-271
            uint32_t oc = 0;
            for (int c = 3; c >= 0; --c) {
                sk[c] = P.ix.fwd.cum[c] + oL[c] + 1;
                sl[c] = P.ix.fwd.cum[c] + oR[c];
                rsl[c] = crl - oc;
                oc += oR[c] - oL[c];
            }
        }

        if (pend) {
            // the child the reference pushed at bwtgap.c:281/297 (deletion j) or :312 (mismatch j+1)
            uint32_t c = pend == KIND_FAM_D ? pend_j : (sc_ + pend_j + 1u) & 3u;
            ck = sel4(sk, c); cl = sel4(sl, c); crl = sel4(rsl, c);
            pend = 0;
            ++extra;
            prefetch_candidate();
            return;                                             // next iteration: hit test / exact entry / expansion
        }

        if (exact) {
            if (sc_ > 3) { have = false; return; }             // 2BWT-Interface.c:376-377 (no lookup issued there)
            lookups_group += 2;
            uint32_t nk = sel4(sk, sc_), nl = sel4(sl, sc_), nr = sel4(rsl, sc_);
            if (nk > nl) { have = false; return; }
            ck = nk; cl = nl; crl = nr; ci = i;
            prefetch_candidate();
            return;
        }

        lookups_group += 2;
        uint32_t occ = cl - ck + 1;
        bool allow_diff = true, allow_M = true;
        if (i > 0) {                                                                        // :252-265
            if ((int32_t)w0.y > m - 1) allow_diff = false;
            else if ((int32_t)w0.y == m - 1 && (int32_t)w1.y == m - 1 && w0.x == w1.x) allow_M = false;
            if (use_seed) {
                if ((int32_t)s0.y > m_seed - 1) allow_diff = false;
                else if ((int32_t)s0.y == m_seed - 1 && (int32_t)s1.y == m_seed - 1 && s0.x == s1.x) allow_M = false;
            }
        }
        const uint32_t e_mm = c_mm, e_go = c_gapo, e_ge = c_gape, e_state = c_state;
        const uint32_t pk0 = ck, pl0 = cl, prl0 = crl;
        const int32_t e_score = (int32_t)e_mm * o.s_mm + (int32_t)e_go * o.s_gapo + (int32_t)e_ge * o.s_gape;
        const uint32_t vmask = (uint32_t)(sk[0] <= sl[0]) | (uint32_t)(sk[1] <= sl[1]) << 1 |
                               (uint32_t)(sk[2] <= sl[2]) << 2 | (uint32_t)(sk[3] <= sl[3]) << 3;
        int32_t tmp;
        if (o.mode & MODE_LOGGAP) {                                                         // :267 + int_log2 :107-116
            uint32_t v = e_ge + e_go; int32_t lg = 0;
            while (v > 1) { v >>= 1; ++lg; }
            tmp = lg / 2 + 1;
        } else tmp = (int32_t)(e_go + e_ge);
        if (allow_diff && (int32_t)i >= o.indel_end_skip + tmp && (int32_t)len - (int32_t)i >= o.indel_end_skip + tmp) {
            bool ins = false, del = false;
            if (e_state == ST_M) ins = del = (int32_t)e_go < o.max_gapo;                   // :269-282
            else if (e_state == ST_I) ins = (int32_t)e_ge < o.max_gape;                    // :283-285
            else del = (int32_t)e_ge < o.max_gape &&                                       // :286-299
                       ((int32_t)(e_ge + e_go) < max_diff || occ < (uint32_t)o.max_del_occ);
            const int32_t gsc = e_score + (e_state == ST_M ? o.s_gapo : o.s_gape);
            if (ins)
                push_record(gsc, KIND_PLAIN, 0, 1, i, pk0, pl0, prl0, e_mm, e_go + (e_state == ST_M), e_ge + (e_state != ST_M), ST_I, 1);
            if (del && vmask)
                push_record(gsc, KIND_FAM_D, vmask, (uint32_t)popc32(vmask), i + 1, pk0, pl0, prl0, e_mm, e_go, e_ge, e_state, 1);
        }
        have = false;
        if (allow_diff && allow_M) {                                                        // :302-314
            // children j = 1..3 (and j = 4 when seq[i] is N) are mismatches: one family record, bit j-1
            uint32_t mmask = 0;
            for (uint32_t j = 1; j <= 3; ++j) mmask |= ((vmask >> ((sc_ + j) & 3u)) & 1u) << (j - 1);
            if (sc_ > 3) mmask |= ((vmask >> (sc_ & 3u)) & 1u) << 3;
            if (mmask)
                push_record(e_score + o.s_mm, KIND_FAM_MM, mmask, (uint32_t)popc32(mmask), i + 1, pk0, pl0, prl0, e_mm, e_go, e_ge, e_state, 1);
            if (sc_ < 4 && ((vmask >> sc_) & 1u)) carry(i, sel4(sk, sc_), sel4(sl, sc_), sel4(rsl, sc_), e_mm, e_go, e_ge);
        } else if (sc_ < 4) {                                                               // :315-325
            if ((vmask >> sc_) & 1u) carry(i, sel4(sk, sc_), sel4(sl, sc_), sel4(rsl, sc_), e_mm, e_go, e_ge);
        }
        if (failed) ending = true;
    }

    // the exact-match child: would be pushed last into the lowest bucket and popped next -> keep in registers
    HSA_HD void carry(uint32_t i, uint32_t k, uint32_t l, uint32_t rl, uint32_t mm, uint32_t go, uint32_t ge)
    {
        have = true; direct = true; exact = false; pend = 0;
        ck = k; cl = l; crl = rl; ci = i; c_mm = mm; c_gapo = go; c_gape = ge; c_state = ST_M; c_diff = 0;
        prefetch_candidate();
    }
};

} // namespace hsa

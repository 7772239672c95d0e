// hsa_core.cuh -- the inexact-search state machine, one search per CUDA thread, phase-voted per warp.
//
// The file is plain C++ with a handful of macros so that the SAME source is compiled
//   * by nvcc for sm_100a (the product: hsa_b200.cu), and
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
//   * device index layout: one 32-byte sector per 64 BWT symbols = {occ[4] at block start, two 64-bit planes},
//     so one occ lookup touches exactly one sector instead of two (+ a major-table row);
//   * a search is a per-lane state machine with four kinds of step -- POP (take the next stack entry and
//     apply the reference's pop-time pruning), LOOKUP (one "occ4 at k and at l+1" pair: node expansion,
//     bwt_match_exact step, or materialisation of a lazily stored child), HIT (record a hit, gap_shadow)
//     and END/START (results out, next work item in).  A warp executes ONE kind of step per trip, chosen by
//     vote (phase_vote), so that all participating lanes run the same instruction stream;
//   * ONE 16-byte stack record per node expansion: the expanded node's interval and counts.  Its children
//     (bwtgap.c:267-325) are two "memberships" of that record -- the gap children (insertion + up to four
//     deletions) in the bucket of score+gap, the mismatch children in the bucket of score+s_mm -- each a
//     5-bit mask in the record's link word.  A child is materialised (one more occ4 pair at the parent's
//     interval) only when it is popped AND survives the pop-time pruning (bwtgap.c:161-173).  About nine of
//     ten entries the reference pushes are never popped, so this removes most stack traffic;
//   * the lowest-score child (the exact-match extension) is never pushed: it is the next node popped by
//     construction (pushed last into the currently-lowest bucket), so it is carried in registers;
//   * children whose score already exceeds best_score + s_mm after the first hit are only counted
//     (they can never be popped, bwtgap.c:158-159), keeping the max_entries test exact;
//   * of bwt_width_t only what the hot path compares is kept close: one byte per position
//     (min(bid,63) | (w[i-1]==w[i]) << 7) in shared memory; the 32-bit w values stay in the item's row in
//     global memory for gap_shadow.
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
#define HSA_HD_CALL __host__ __device__ __noinline__      // large bodies with many call sites (hsa_splice.cuh): real calls
#else
#define HSA_HD inline
#define HSA_D  inline
#define HSA_HD_CALL
#endif

// Per-block shared memory (device) / per-run scratch (host emulation): option table, then per lane the bucket
// heads and the bound bytes.
#if defined(__CUDACC__)
extern __shared__ __align__(16) unsigned char hsa_smem[];
#endif
#if defined(__CUDA_ARCH__)
#define HSA_SMEM hsa_smem
#else
static thread_local unsigned char *hsa_smem_host = nullptr;
#define HSA_SMEM hsa_smem_host
static unsigned long long hsa_host_pair_count = 0, hsa_host_pair_same_sector = 0;   // emulation statistics
static unsigned long long hsa_host_pop_out[8];           // emulation statistics: what a POP step led to (by next state)
static unsigned long long hsa_host_hist[3][2][64];     // emulation statistics: lookup pairs by kind {expand, exact, materialise} x {two sectors, one} x depth
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
// one 32-byte index sector {counts, packed words} in a single 256-bit read-only load (sm_100: LDG.E.256)
HSA_HD void ld_sector(const u32x4 *p, u32x4 &a, u32x4 &b)
{
#if defined(__CUDA_ARCH__)
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p));
#else
    a = p[0]; b = p[1];
#endif
}
HSA_HD void st4(void *p, uint32_t x, uint32_t y, uint32_t z, uint32_t w)      // 16-byte store
{
    u32x4 v; v.x = x; v.y = y; v.z = z; v.w = w;
    *reinterpret_cast<u32x4 *>(p) = v;
}
// one 32-byte stack slot {record, link word(s) + padding} in a single 256-bit access (the slot is private to its lane)
HSA_HD void ld_slot(const u32x4 *p, u32x4 &a, u32x4 &b)
{
#if defined(__CUDA_ARCH__)
    asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p) : "memory");
#else
    a = p[0]; b = p[1];
#endif
}
HSA_HD void st_slot(u32x4 *p, const u32x4 &a, const u32x4 &b)
{
#if defined(__CUDA_ARCH__)
    asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
#else
    p[0] = a; p[1] = b;
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
//   blocks[2b+1] = the block's symbols as two 64-bit planes (planes_of): low code bits, high code bits
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

// The 64 symbols of a block as two bit planes: symbol j (0 = first) sits at bit j & 31 of word j >> 5; words x, y hold
// the low bits of the 2-bit codes, words z, w the high bits.  `v` = the block's four packed words as the reference
// stores them (first symbol in the two MSBs, BWT.c:954).
HSA_HD u32x4 planes_of(const u32x4 &v)
{
    const uint32_t in[4] = {v.x, v.y, v.z, v.w};
    uint32_t lo[2] = {0, 0}, hi[2] = {0, 0};
    for (uint32_t j = 0; j < 64; ++j) {
        const uint32_t code = (in[j >> 4] >> (30u - 2u * (j & 15u))) & 3u;
        lo[j >> 5] |= (code & 1u) << (j & 31u);
        hi[j >> 5] |= (code >> 1) << (j & 31u);
    }
    u32x4 r; r.x = lo[0]; r.y = lo[1]; r.z = hi[0]; r.w = hi[1];
    return r;
}

// masks of the first `off` (0..63) symbols of a block in its two plane words
HSA_HD void prefix_masks(uint32_t off, uint32_t &m0, uint32_t &m1)
{
    m0 = off >= 32u ? 0xFFFFFFFFu : (1u << off) - 1u;
    m1 = off > 32u ? (1u << (off - 32u)) - 1u : 0u;
}

// occ of all four symbols at SA-coordinate `index` on the device layout == BWTAllOccValue (BWT.c:793)
HSA_HD void occ4_from_sector(const u32x4 &cnt, const u32x4 &w, uint32_t off, uint32_t occ[4])
{
    uint32_t m0, m1;
    prefix_masks(off, m0, m1);
    const uint32_t l0 = w.x & m0, l1 = w.y & m1, h0 = w.z & m0, h1 = w.w & m1;
    const uint32_t t = (uint32_t)(popc32(l0 & h0) + popc32(l1 & h1));          // code 3
    const uint32_t h = (uint32_t)(popc32(h0) + popc32(h1));                    // codes 2, 3
    const uint32_t l = (uint32_t)(popc32(l0) + popc32(l1));                    // codes 1, 3
    occ[0] = cnt.x + (off + t - h - l);
    occ[1] = cnt.y + (l - t);
    occ[2] = cnt.z + (h - t);
    occ[3] = cnt.w + t;
}

HSA_HD void occ4_dev(const DevBwt &b, uint32_t index, uint32_t occ[4])
{
    index -= (index > b.inverse_sa0);                   // BWT.c:804
    u32x4 cnt, w;
    ld_sector(b.blocks + 2 * (size_t)(index >> 6), cnt, w);
    occ4_from_sector(cnt, w, index & 63u, occ);
}

// ---- SA index -> text position ----------------------------------------------------
// BWTPsiMinusValue (BWT.c:1142-1165) on the device layout: the SA index of the suffix one text position to the left =
// C[c] + occ(c, index + 1), c being the BWT symbol at `index` (BWTOccValueOnSpot, BWT.c:924-965).  The symbol and its
// rank come out of the SAME 32-byte sector: the symbol sits at in-block offset `off`, the rank counts the block's first
// off + 1 symbols.  index != inverseSa0 (that case is the '$' itself: BWT.c:1161-1163 returns 0).
HSA_HD uint32_t psi_minus_dev(const DevBwt &b, uint32_t index)
{
    uint32_t p = index + 1;
    p -= (p > b.inverse_sa0);                           // BWT.c:948; p in 1..n, the symbol is raw position p - 1
    const uint32_t q = p - 1, off = q & 63u;
    u32x4 cnt, w;
    ld_sector(b.blocks + 2 * (size_t)(q >> 6), cnt, w);
    const uint32_t lo = off < 32u ? w.x : w.y, hi = off < 32u ? w.z : w.w;
    const uint32_t c = ((lo >> (off & 31u)) & 1u) | ((hi >> (off & 31u)) & 1u) << 1;
    // the first off + 1 (1..64) symbols of the block
    const uint32_t t = off + 1u;
    const uint32_t m0 = t >= 32u ? 0xFFFFFFFFu : (1u << t) - 1u;
    const uint32_t m1 = t > 32u ? (t == 64u ? 0xFFFFFFFFu : (1u << (t - 32u)) - 1u) : 0u;
    const uint32_t fl = (c & 1u) - 1u, fh = (c >> 1) - 1u;                     // all ones where c's bit is 0
    const uint32_t base = c == 0 ? cnt.x : c == 1 ? cnt.y : c == 2 ? cnt.z : cnt.w;
    return b.cum[c] + base + (uint32_t)popc32((w.x ^ fl) & (w.z ^ fh) & m0) + (uint32_t)popc32((w.y ^ fl) & (w.w ^ fh) & m1);
}

// BWTSaValue, BWT.c:1195-1225: walk left until a sampled SA index; sa_value as loaded (entry 0 = -1, BWT.c:222)
HSA_HD uint32_t sa_value_dev(const DevBwt &b, const uint32_t *sa_value, uint32_t sa_interval, uint32_t sa_index, uint32_t &steps)
{
    steps = 0;
    while (sa_index % sa_interval != 0) {
        ++steps;
        sa_index = sa_index == b.inverse_sa0 ? 0u : psi_minus_dev(b, sa_index);
    }
    return ld_ro1(sa_value + sa_index / sa_interval) + steps;
}

// the block search of BWTRetrievePositionFromSAIndex (2BWT-Interface.c:339-361) over HSP::blockList rows
// {chrID, blockStart, blockEnd, ori}: the block that holds packed-text position occ_pos -> chromosome id and 1-based
// position in it; positions outside every block (SA[0] = -1) leave the outputs as they are, like the reference
HSA_HD bool locate_dev(const uint32_t *b4, uint32_t n_blocks, uint32_t occ_pos, uint32_t &seq_id, uint32_t &ori_pos)
{
    uint32_t l = 0, h = n_blocks;
    while (l < h) {
        const uint32_t m = (l + h) >> 1;
        const uint32_t start = ld_ro1(b4 + 4 * m + 1);
        if (start > occ_pos) h = m;
        else if (ld_ro1(b4 + 4 * m + 2) < occ_pos) l = m + 1;
        else { seq_id = ld_ro1(b4 + 4 * m); ori_pos = occ_pos - start + ld_ro1(b4 + 4 * m + 3) + 1; return true; }
    }
    return false;
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
enum { CNT_CURSOR0 = CNT_N, CNT_NEXT0 = CNT_N + 8, CNT_DIAG_WARP_ITERS = CNT_N + 16, CNT_DIAG_MAX_ITEM_STEPS = CNT_N + 17,
       CNT_STRICT2 = CNT_N + 18, CNT_DIAG_COOP_WAVES = CNT_N + 19 /* +1 wave_steps, +2 lane steps */,
       CNT_DIAG_FAST_LOOKUPS = CNT_N + 22 /* occ lookups of the searches the per-lane kernels completed (no width pass) */, CNT_TOTAL = CNT_N + 24 };

struct Params {
    DevIndex ix;
    const uint8_t *codes;           // base codes, 0..3, N = 4
    uint32_t kind;
    uint32_t pass;                  // KIND_WHOLE: 1 = reverse-complement strand pass, 2 = forward strand pass
    uint32_t n_work;                // work items of this launch ...
    const uint32_t *n_work_dev;     // ... or, if set, read from device memory (pass 2): items [n_work_skip, *n_work_dev) capped at n_work
    uint32_t n_work_skip;
    const uint32_t *work_list;      // optional: work index -> item id (pass-2 list, large-capacity re-runs)
    uint32_t work_base;             // item id of work index 0 when there is no list
    const Task *tasks;              // KIND_TASKS
    const uint64_t *read_off;       // KIND_WHOLE / KIND_SEEDS / KIND_WIDTH
    const uint32_t *read_len;
    const DevOpt *opts;             // option table (device memory; staged to shared memory by the kernels)
    uint32_t n_opts;
    const uint16_t *len2opt;        // KIND_WHOLE: read length -> opts[] index ; otherwise unused
    uint32_t max_len;               // longest searched sequence in the batch
    int32_t  filter_max_n;          // KIND_WHOLE: local_opt.max_diff of bwtaln.c:273-274 (N filter :314-317)
    // per-work-item rows written by the width kernel (row of work index w at rows + w * row_stride):
    //   [0, 4*(max_len+1))            w of width_back, uint32 each
    //   [row_bid_off, +max_len+1)     bound bytes of width_back: min(bid,63) | (w[i-1]==w[i]) << 7
    //   [row_seed_off, +seed_cap)     bound bytes of width_seed (HSA_SEED_TAIL only)
    //   [row_tail_off, +8)            {occ lookups of the width pass, flags (bit 0: read filtered)}
    uint8_t *rows;
    uint32_t row_stride, row_bid_off, row_seed_off, row_tail_off;
    // worker-private scratch, indexed by worker slot
    u32x4 *arena;  uint32_t arena_cap;                    // arena: one 32-byte slot per record = 2 x u32x4 (record, link word + padding)
    Hit *hits;     uint32_t hit_cap;
    uint32_t n_buckets;             // size of the score-indexed head table (<= 128)
    uint32_t smem_opts_bytes;       // shared memory: option table first ...
    uint32_t smem_lane_stride;      // ... then per lane: heads[n_buckets], bound bytes back, bound bytes seed
    uint32_t smem_bid_off, smem_seed_off;   // offsets of the bound bytes inside a lane's region
    uint32_t smem_stats_off;        // per-lane search kernels: the block's five 64-bit statistics words (8-byte aligned)
    // outputs
    int32_t  *n_aln;                // [n_items]
    uint64_t *aln_off;              // [n_items]
    uint8_t  *status;               // [n_items]
    uint32_t *aln;                  // aln_cap x 9 words (== hsa_aln1_t)
    uint64_t  aln_cap;
    unsigned long long *counters;   // CNT_TOTAL
    uint32_t *strict_list;          // items that must be re-run with larger capacities
    u32x2 *width_out;               // KIND_WIDTH: read r's len+1 entries at read_off[r] + r
    int32_t *bid_out;               // KIND_WIDTH: bwt_cal_width's return value per read
    uint32_t *next_list;            // pass 1: reads without a hit are appended here ...
    uint32_t *next_count;           // ... and counted here (device memory)
    unsigned long long *cursor;     // atomic work-queue cursor of this launch
    // warp-cooperative kernel (hsa_coop.cuh): per-warp global scratch (warp w at index w of each array group)
    u32x4 *coop_payload; uint32_t *coop_info; uint32_t *coop_prev;    // cap_chunks chunks of 32 records per warp
    u32x4 *coop_out_payload; uint32_t *coop_out_info;                 // 32 x COOP_OUT_CAP per warp
    Hit *coop_hits;                                                    // COOP_HIT_CAP per warp
    uint32_t coop_cap_chunks, coop_warp_smem;
    unsigned long long *strict_count;   // where items this launch cannot hold are counted (indexes strict_list)
    uint32_t step_budget;           // fast kernel: searches still running after this many steps are handed on (0 = off)
    uint32_t drain_budget;          // ... the same once the work queue has run dry (bounds the launch's serial tail)
    uint32_t vote_slow_min;         // phase_vote(): lanes that must wait for a SLOW step before one is run
    int32_t  vote_pop_bias;         // phase_vote(): LOOKUP runs if n_lookup + bias >= n_pop
};

enum : uint32_t { ROW_FLAG_FILTERED = 1u, ROW_FLAG_HAS_N = 2u /* the bound bytes' base bits do not hold every base */ };

// Row / shared-memory geometry of one launch configuration (shared by the host code and the emulation).
// seed_cap = longest width_seed in the batch + 1 (0 if none); head_bytes = 2 (fast kernel) or 4 (large capacity).
HSA_HD void set_layout(Params &P, uint32_t max_len, uint32_t seed_cap, uint32_t n_buckets, uint32_t n_opts,
                       uint32_t head_bytes, bool bids_smem)
{
    // every row region holds a multiple of 16 entries so that the width kernel can store 16 bytes at a time
    const uint32_t nb4 = (max_len + 1 + 15u) & ~15u, ns4 = (seed_cap + 15u) & ~15u;
    P.max_len = max_len; P.n_buckets = n_buckets; P.n_opts = n_opts;
    P.row_bid_off = 4u * nb4;
    P.row_seed_off = P.row_bid_off + nb4;
    P.row_tail_off = P.row_seed_off + ns4;
    P.row_stride = (P.row_tail_off + 8u + 31u) & ~31u;      // rows start on a 32-byte sector
    P.smem_opts_bytes = n_opts * (uint32_t)sizeof(DevOpt);
    const uint32_t heads = (n_buckets * head_bytes + 3u) & ~3u;
    P.smem_bid_off = heads;
    P.smem_seed_off = heads + (bids_smem ? nb4 : 0u);
    uint32_t stride = P.smem_seed_off + (bids_smem ? ns4 : 0u);
    if (((stride >> 2) & 1u) == 0) stride += 4;         // odd number of words per lane: lanes spread over the banks
    P.smem_lane_stride = stride;
}

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

// work items of a launch: P.n_work, or -- with a device-side count -- what is left of it beyond n_work_skip, capped at P.n_work
HSA_HD uint32_t work_count(const Params &P)
{
    if (!P.n_work_dev) return P.n_work;
    const uint32_t n = *P.n_work_dev, left = n > P.n_work_skip ? n - P.n_work_skip : 0u;
    return left < P.n_work ? left : P.n_work;
}

// Item ids:  KIND_TASKS: task index;  KIND_WHOLE: read index (the strand comes from P.pass: pass 1 = reverse
// complement, pass 2 = forward, bwtaln.c:343);  KIND_SEEDS: read * 6 + strand * 3 + segment (bwtgap.c:797-812);
// KIND_WIDTH: read index.
HSA_HD uint32_t work_item(const Params &P, uint32_t w) { return P.work_list ? P.work_list[w] : P.work_base + w; }

HSA_HD TaskDesc make_task(const Params &P, const DevOpt *opts, uint32_t item)
{
    TaskDesc t;
    t.aln_start = t.aln_end = -1;
    t.out_idx = item;
    if (P.kind == KIND_TASKS) {
        const Task &k = P.tasks[item];
        t.rd = P.codes + k.read_off; t.rd_len = k.read_len; t.strand = k.strand; t.sub_off = k.sub_off; t.len = k.len;
        t.wsrc_off = k.wsrc_off; t.seed_mode = k.seed_mode; t.opt_idx = k.opt_idx;
    } else if (P.kind == KIND_WHOLE) {
        t.rd = P.codes + P.read_off[item]; t.rd_len = P.read_len[item];
        t.strand = P.pass == 2 ? 0u : 1u;                   // revcomp first (bwtaln.c:343)
        t.sub_off = 0; t.len = t.rd_len; t.wsrc_off = 0;
        t.opt_idx = P.len2opt[t.rd_len];
        t.seed_mode = (int32_t)t.rd_len > opts[t.opt_idx].seed_len ? SEED_TAIL : SEED_NONE;   // bwtaln.c:332,344
    } else if (P.kind == KIND_WIDTH) {                      // bwt_cal_width alone (type 1), read as given
        t.rd = P.codes + P.read_off[item]; t.rd_len = P.read_len[item];
        t.strand = 0; t.sub_off = 0; t.len = t.rd_len; t.wsrc_off = 0; t.seed_mode = SEED_NONE; t.opt_idx = 0;
    } else {                                                // KIND_SEEDS, bwtgap.c:797-812
        const uint32_t gid = item / 6, sub = item % 6;
        t.rd = P.codes + P.read_off[gid]; t.rd_len = P.read_len[gid];
        const uint32_t seg = sub % 3, sl = t.rd_len / 3;
        t.strand = sub / 3;
        t.len = sl + (seg == 2 ? t.rd_len % 3 : 0);
        t.sub_off = seg * sl;
        t.wsrc_off = 0;                                     // width on the read PREFIX (bwtgap.c:807-808)
        t.seed_mode = SEED_ALIAS;                           // bwtgap.c:809
        t.opt_idx = 0;
        t.aln_start = (int32_t)t.sub_off; t.aln_end = (int32_t)(t.sub_off + t.len - 1);   // bwtgap.c:816-819
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
    u32x4 cnt, w;
    ld_sector(b.blocks + 2 * (size_t)(index >> 6), cnt, w);
    uint32_t m0, m1;
    prefix_masks(index & 63u, m0, m1);
    // symbols equal to c: both plane bits agree with c's
    const uint32_t fl = (c & 1u) - 1u, fh = ((c >> 1) & 1u) - 1u;            // all ones where c's bit is 0
    const uint32_t base = c == 0 ? cnt.x : c == 1 ? cnt.y : c == 2 ? cnt.z : cnt.w;
    return base + (uint32_t)popc32((w.x ^ fl) & (w.z ^ fh) & m0) + (uint32_t)popc32((w.y ^ fl) & (w.w ^ fh) & m1);
}

// the bound byte of one bwt_width_t entry given the previous entry's w (0xFFFFFFFF for entry 0: never equal,
// because w <= textLength + 1 < 2^32 - 1 for every index the uint32 SA coordinates can hold)
// Layout: bits 0..4 = min(bid, 31) (every comparison is against m <= max_diff <= 30), bits 5..6 = the base the search
// reads at this step (BB_BASE_SHIFT; set by the width pass for the width_back array, see base_step), bit 7 = (w[i-1] == w[i]).
enum : uint32_t { BB_BID = 31u, BB_BASE_SHIFT = 5u, BB_BASE = 3u << 5, BB_EQ = 0x80u };
HSA_HD uint8_t bound_byte(uint32_t bid, uint32_t w, uint32_t w_prev)
{
    return (uint8_t)((bid < BB_BID ? bid : BB_BID) | (w == w_prev ? BB_EQ : 0u));
}

// bwt_cal_width, type 1 (bwtaln.c:73-97, 113-114): `n` bases starting at strand-resolved position `src`.
// Writes any of: w_out (uint32 per entry), b_out (bound byte per entry), pair_out (bwt_width_t per entry), n + 1
// entries each.  w_out / b_out are 16-byte aligned and padded to a multiple of 16 entries: they are written 16
// bytes at a time (entries past n are zero).  Returns bid; lookups += the BWTOccValue calls the reference issues.
// base_src != NO_BASE: byte j of b_out also carries the base at strand-resolved position base_src + j (what the search
// reads at step j; for the splice seeds that is not the position the width is computed on), and has_n reports whether any
// of those n bases is not A/C/G/T (such items read their bases from the codes array instead).
enum : uint32_t { NO_BASE = 0xFFFFFFFFu };
HSA_HD int32_t cal_width_dev(const DevIndex &ix, const TaskDesc &t, uint32_t src, uint32_t n,
                             uint32_t *w_out, uint8_t *b_out, u32x2 *pair_out, uint32_t &lookups,
                             uint32_t base_src = NO_BASE, bool *has_n = nullptr)
{
    uint32_t k = 0, l = ix.fwd.text_length, w_prev = 0xFFFFFFFFu;
    int32_t bid = 0;
    for (uint32_t j16 = 0; j16 <= n; j16 += 16) {
        uint32_t bq[4];
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (uint32_t q = 0; q < 4; ++q) {
            uint32_t wq[4];
            uint32_t bw = 0;
#if defined(__CUDACC__)
#pragma unroll
#endif
            for (uint32_t r = 0; r < 4; ++r) {
                const uint32_t j = j16 + 4 * q + r;
                uint32_t w = 0, byte = 0;
                if (j < n) {
                    const uint32_t c = task_base(t, src + j);
                    if (c < 4) {                            // BWTSARangeForeward, 2BWT-Interface.c:121-132
                        const uint32_t a = occ1_dev(ix.rev, k, c), b = occ1_dev(ix.rev, l + 1, c);
                        k = ix.fwd.cum[c] + a + 1;
                        l = ix.fwd.cum[c] + b;
                        lookups += 2;
                    }
                    if (k > l || c > 3) { k = 0; l = ix.fwd.text_length; ++bid; }
                    w = l - k + 1;
                    byte = bound_byte((uint32_t)bid, w, w_prev);
                    if (base_src != NO_BASE) {
                        const uint32_t cb = base_src == src ? c : task_base(t, base_src + j);
                        if (cb > 3) *has_n = true; else byte |= cb << BB_BASE_SHIFT;
                    }
                    if (pair_out) { u32x2 v; v.x = w; v.y = (uint32_t)bid; pair_out[j] = v; }
                    w_prev = w;
                } else if (j == n) {                        // bwtaln.c:113-114
                    ++bid;
                    byte = bound_byte((uint32_t)bid, 0, w_prev);
                    if (pair_out) { u32x2 v; v.x = 0; v.y = (uint32_t)bid; pair_out[j] = v; }
                }
                wq[r] = w;
                bw |= byte << (8 * r);
            }
            if (w_out && j16 + 4 * q <= n) st4(w_out + j16 + 4 * q, wq[0], wq[1], wq[2], wq[3]);
            bq[q] = bw;
        }
        if (b_out) st4(b_out + j16, bq[0], bq[1], bq[2], bq[3]);
    }
    return bid;
}

// Width kernel body for work index `w`: resolves the task, applies the whole-read filters in pass 1 and writes
// the item's row (w of width_back, bound bytes of width_back and width_seed, lookup count).
HSA_HD void width_item(const Params &P, const DevOpt *opts, uint32_t w)
{
    const uint32_t item = work_item(P, w);
    const TaskDesc t = make_task(P, opts, item);
    uint32_t lk = 0;
    if (P.kind == KIND_WIDTH) {                              // hsa_cal_width_batch: straight to the caller's layout
        P.bid_out[item] = cal_width_dev(P.ix, t, 0, t.len, nullptr, nullptr, P.width_out + P.read_off[item] + item, lk);
#if defined(__CUDA_ARCH__)
        atomicAdd(&P.counters[CNT_LOOKUPS], (unsigned long long)lk);
#else
        P.counters[CNT_LOOKUPS] += lk;
#endif
        return;
    }
    uint8_t *row = P.rows + (size_t)w * P.row_stride;
    uint32_t *tail = reinterpret_cast<uint32_t *>(row + P.row_tail_off);
    if (P.kind == KIND_WHOLE && P.pass == 1 && read_filtered(t.rd, t.rd_len, P.filter_max_n)) {
        tail[0] = 0; tail[1] = ROW_FLAG_FILTERED;
        P.n_aln[t.out_idx] = 0; P.aln_off[t.out_idx] = 0; P.status[t.out_idx] = STATUS_OK;
        return;
    }
    bool has_n = false;
    cal_width_dev(P.ix, t, t.wsrc_off, t.len, reinterpret_cast<uint32_t *>(row), row + P.row_bid_off, nullptr, lk, t.sub_off, &has_n);
    if (t.seed_mode == SEED_TAIL) {
        const uint32_t sl = (uint32_t)opts[t.opt_idx].seed_len;
        cal_width_dev(P.ix, t, t.sub_off + (t.len - sl), sl, nullptr, row + P.row_seed_off, nullptr, lk);
    }
    // the row's tail carries the width pass's lookup count; the search worker adds it when the item completes,
    // so items that are re-run with the large-capacity kernel are not counted twice
    tail[0] = lk; tail[1] = has_n ? ROW_FLAG_HAS_N : 0u;
}

// ---- the worker -------------------------------------------------------------------------------------
// Stack record (16 bytes) = the EXPANDED node: {k, l, rev_l, meta}; meta = i:12 | state:2 @13 | mm:5 @15 |
// go:4 @20 | ge:5 @24, i being the node's i before bwtgap.c:244's --i.  rev_k == rev_l - (l - k) holds for
// every node ever created, so it is not stored.
// Link word of a record = two halves, one per membership:
//   half A (low)  : gap children, in the bucket of score + (state == M ? s_gapo : s_gape);
//                   mask bit 0 = the insertion (bwtgap.c:274 / :284), bits 1..4 = deletion of symbol 0..3
//                   (bwtgap.c:276-282 / :292-298)
//   half B (high) : mismatch children, in the bucket of score + s_mm; mask bit j-1 = the reference's loop
//                   index j = 1..4 (bwtgap.c:303-313; j = 4 exists only when the base is N)
//   each half = ref:NEXT_BITS+1 | mask:5, ref = slot | membership << NEXT_BITS naming the NEXT list element: a
//   bucket list threads through halves, and the same refs sit in the bucket heads, so a pop knows which
//   membership of the head record it is looking at (and the bucket index is the child's score).
//   Children are taken highest bit first = the reverse of the reference's push order, as its LIFO buckets do.
//   A half leaves its bucket list when its mask empties; the record is freed when both masks are empty.
enum : uint32_t { PEND_NONE = 0, PEND_DEL = 4, PEND_MM = 8 };        // | child index in the low two bits
// (the two low bits of a state are its phase: LOOKUP 1, POP 2, retired 3, everything else waits for a SLOW step)
enum : uint32_t { LS_IDLE = 0, LS_LOOKUP = 1, LS_POP = 2, LS_RETIRED = 3, LS_HIT = 4, LS_END = 8 };
enum : uint32_t { PHASE_SLOW = 0, PHASE_LOOKUP = 1, PHASE_POP = 2, PHASE_NONE = 3 };
enum : uint32_t { META_STATE_SHIFT = 13, META_MM_SHIFT = 15, META_GO_SHIFT = 20, META_GE_SHIFT = 24 };

// Which kind of step a warp executes next, from the number of lanes waiting for each kind.  SLOW steps (hit
// recording, task end, work fetch + task start) are long and rare: they run once enough lanes have queued up
// for them, or when nothing else can run.
enum { VOTE_SLOW_MIN_DEFAULT = 4, VOTE_POP_BIAS_DEFAULT = -12 };
HSA_HD uint32_t phase_vote(const Params &P, uint32_t n_lookup, uint32_t n_pop, uint32_t n_slow)
{
    if (n_slow >= P.vote_slow_min || (n_lookup == 0 && n_pop == 0)) return PHASE_SLOW;
    if (n_pop == 0) return PHASE_LOOKUP;
    if (n_lookup == 0) return PHASE_POP;
    return (int32_t)n_lookup + P.vote_pop_bias >= (int32_t)n_pop ? PHASE_LOOKUP : PHASE_POP;
}

template <typename LinkT, bool BIDS_SMEM>
struct Worker {
    static constexpr bool WIDE = sizeof(LinkT) == 8;                 // large-capacity configuration
    static constexpr uint32_t HALF_BITS = WIDE ? 32 : 16;
    static constexpr uint32_t NEXT_BITS = HALF_BITS - 6;             // 10 (fast: <= 1022 records) or 26
    static constexpr uint32_t NIL = (1u << NEXT_BITS) - 1u;
    static constexpr uint32_t REF_MASK = (1u << (NEXT_BITS + 1)) - 1u;
    static constexpr uint32_t MASK_SHIFT = NEXT_BITS + 1;
    static constexpr uint32_t HALF_MASK = WIDE ? 0xFFFFFFFFu : 0xFFFFu;
    static constexpr uint32_t MAX_BUCKETS = WIDE ? 128 : 64;         // the fast kernel keeps one 64-bit bucket mask

    // environment
    const Params &P;
    uint32_t slot;                  // worker slot -> scratch
    uint32_t sm_heads, sm_bid, sm_seed;   // byte offsets of this lane's regions in shared memory

    // current task
    uint32_t st;                    // LS_*
    uint32_t work;                  // work-queue index of the current item == its row
    const uint8_t *rd;              // the read
    uint32_t rd_len, strand, sub_off, len, seed_mode, seed_shift, opt_idx, out_idx;
    uint32_t rd_word, rd_widx;      // the four read bytes last loaded (base_at) and their word index; rd_widx == RD_IN_BYTES: the
                                    // item has no N and its bases ride in the bound bytes (base_step)
    uint8_t *row;                   // the item's row (global)
    // search state
    uint64_t mask0, mask1;          // non-empty buckets (mask1: large-capacity configuration only)
    uint32_t n_live, n_phantom;     // entries the reference would hold: stored children; counted-only children
    uint32_t top, free_head;        // arena bump pointer and free list
    int32_t best_score, max_diff, best_cnt;
    int32_t pop_cut;                // entries above this score end the search when popped (bwtgap.c:158-159)
    uint32_t n_hits;
    uint32_t fail_code;             // != STATUS_OK: the task ran out of capacity / hit a sizing error
    // candidate node
    bool direct, exact, c_diff;
    uint32_t pend;                  // PEND_*: the interval registers hold the PARENT's, child pend & 3
    uint32_t ck, cl, crl;           // k, l, rev_l  (rev_k == rev_l - (l - k))
    uint32_t ci;
    uint32_t c_meta;                // state / n_mm / n_gapo / n_gape in the stack record's bit layout
    int32_t c_score;                // aln_score(n_mm, n_gapo, n_gape)
    int32_t c_nd;                   // n_mm + n_gapo (+ n_gape with MODE_GAPE): the diffs bwtgap.c:161-164 counts
    uint32_t ci_at_pop;             // e.info & 0xffff of the entry being processed (for last_diff_pos)
    uint32_t zflags;                // which of k,l,rev_k,rev_l were zero when bwt_match_exact was entered
    int32_t m_cur;                  // remaining diffs of the candidate (bwtgap.c:161-164), set by vet()
    // statistics
    uint32_t lookups_item;          // occ lookups of the current item's search
    uint32_t steps32, pops32;
    uint32_t budget;                // step budget in force (Params::step_budget, or drain_budget once the queue is dry)
    uint64_t lookups, pops, steps, search_lookups;
    uint32_t max_item_steps;

    HSA_HD Worker(const Params &p, uint32_t slot_, uint32_t lane_in_block)
        : P(p), slot(slot_), st(LS_IDLE), steps32(0), pops32(0), budget(p.step_budget), lookups(0), pops(0), steps(0), search_lookups(0), max_item_steps(0)
    {
        sm_heads = P.smem_opts_bytes + lane_in_block * P.smem_lane_stride;
        sm_bid = sm_heads + P.smem_bid_off;
        sm_seed = sm_heads + P.smem_seed_off;
    }

    HSA_HD const DevOpt &opt() const { return reinterpret_cast<const DevOpt *>(HSA_SMEM)[opt_idx]; }

    // ---------------------------------------------------------------- small accessors
    HSA_HD uint32_t head_get(uint32_t b) const
    {
        if (!WIDE) return reinterpret_cast<const uint16_t *>(HSA_SMEM + sm_heads)[b];
        return reinterpret_cast<const uint32_t *>(HSA_SMEM + sm_heads)[b];
    }
    HSA_HD void head_set(uint32_t b, uint32_t ref)
    {
        if (!WIDE) reinterpret_cast<uint16_t *>(HSA_SMEM + sm_heads)[b] = (uint16_t)ref;
        else reinterpret_cast<uint32_t *>(HSA_SMEM + sm_heads)[b] = ref;
    }
    // bound bytes: width_back entry i / width_seed entry i
    HSA_HD uint32_t bb(uint32_t i) const { return BIDS_SMEM ? HSA_SMEM[sm_bid + i] : row[P.row_bid_off + i]; }
    HSA_HD void bb_set(uint32_t i, uint32_t v)
    {
        if (BIDS_SMEM) HSA_SMEM[sm_bid + i] = (uint8_t)v; else row[P.row_bid_off + i] = (uint8_t)v;
    }
    HSA_HD uint32_t bs(uint32_t i) const
    {
        if (seed_mode == SEED_ALIAS) return bb(i);
        return BIDS_SMEM ? HSA_SMEM[sm_seed + i] : row[P.row_seed_off + i];
    }
    // stack slot s of this lane: 32 bytes = {the 16-byte record, the link word, padding} -- one sector per push / pop
    HSA_HD u32x4 *slot_at(uint32_t s) const { return P.arena + ((size_t)slot * P.arena_cap + s) * 2; }
    HSA_HD LinkT *link_at(uint32_t s) const { return reinterpret_cast<LinkT *>(slot_at(s) + 1); }
    static HSA_HD LinkT link_of(const u32x4 &aux) { return WIDE ? (LinkT)((uint64_t)aux.x | (uint64_t)aux.y << 32) : (LinkT)aux.x; }
    static HSA_HD u32x4 aux_of(LinkT v) { u32x4 a; a.x = (uint32_t)v; a.y = WIDE ? (uint32_t)((uint64_t)v >> 32) : 0u; a.z = 0; a.w = 0; return a; }
    HSA_HD bool bucket_nonempty(uint32_t b) const
    {
        if (!WIDE) return (mask0 >> b) & 1ull;
        return b < 64 ? (mask0 >> b) & 1ull : (mask1 >> (b - 64)) & 1ull;
    }
    HSA_HD void bucket_set(uint32_t b)
    {
        if (!WIDE) mask0 |= 1ull << b;
        else if (b < 64) mask0 |= 1ull << b; else mask1 |= 1ull << (b - 64);
    }
    HSA_HD void bucket_clear(uint32_t b)
    {
        if (!WIDE) mask0 &= ~(1ull << b);
        else if (b < 64) mask0 &= ~(1ull << b); else mask1 &= ~(1ull << (b - 64));
    }
    HSA_HD uint32_t bucket_lowest() const
    {
        if (!WIDE) return (uint32_t)ffs64(mask0);
        return mask0 ? (uint32_t)ffs64(mask0) : 64u + (uint32_t)ffs64(mask1);
    }
    HSA_HD bool idle() const { return st == LS_IDLE; }
    HSA_HD bool retired() const { return st == LS_RETIRED; }
    HSA_HD void retire() { st = LS_RETIRED; }
    HSA_HD uint32_t cls() const          // which phase this lane waits for
    {
        return st & 3u;
    }
    HSA_HD uint32_t c_state() const { return (c_meta >> META_STATE_SHIFT) & 3u; }
    HSA_HD uint32_t c_mm() const { return (c_meta >> META_MM_SHIFT) & 31u; }
    HSA_HD uint32_t c_gapo() const { return (c_meta >> META_GO_SHIFT) & 15u; }
    HSA_HD uint32_t c_gape() const { return (c_meta >> META_GE_SHIFT) & 31u; }

    // base p of the strand-resolved read (seq_reverse(len, seq, 1): bwaseqio.c:73-90)
    HSA_HD uint32_t base_at(uint32_t p)
    {
        // four bases per 32-bit load: consecutive steps read consecutive bases, so three of four steps need no load at
        // all (-7 % on the 3.1 Gb genome, where these bytes competed with the index sectors for L2).  Aligned words: the
        // codes buffer must be readable up to the next 4-byte boundary (include/hsa_b200.h).
        const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(rd) & 3u);
        const uint32_t a = (strand ? rd_len - 1 - p : p) + mis, wi = a >> 2;
        if (wi != rd_widx) { rd_word = ld_ro1(reinterpret_cast<const uint32_t *>(rd - mis) + wi); rd_widx = wi; }
        const uint32_t c = (rd_word >> (8u * (a & 3u))) & 0xFFu;
        return (strand && c < 4) ? 3 - c : c;
    }
    // base the search reads at step i (== base_at(sub_off + i)): bits 5..6 of bound byte i, which the width pass filled and
    // which the step loads anyway -- unless the item has an N somewhere (ROW_FLAG_HAS_N), then from the codes array.
    // ncu on the 3.1 Gb genome: the 4-byte read-word loads were 15 % of the sectors the SMs asked L2 for.
    enum : uint32_t { RD_IN_BYTES = 0xFFFFFFFEu };
    HSA_HD uint32_t base_step(uint32_t i, uint32_t byte_i)
    {
        if (rd_widx == RD_IN_BYTES) return (byte_i >> BB_BASE_SHIFT) & 3u;
        return base_at(sub_off + i);
    }
    HSA_HD void fail(uint32_t code) { if (fail_code == STATUS_OK) fail_code = code; }
    // statistics: per-block words in shared memory on the device (they would cost nine registers per lane otherwise),
    // plain members in the host emulation
    enum : uint32_t { STAT_LOOKUPS = 0, STAT_POPS = 1, STAT_STEPS = 2, STAT_SEARCH_LOOKUPS = 3, STAT_MAX_ITEM_STEPS = 4, STAT_N = 5 };
    HSA_HD void stat_add(uint32_t which, uint64_t v)
    {
#if defined(__CUDA_ARCH__)
        atomicAdd(reinterpret_cast<unsigned long long *>(HSA_SMEM + P.smem_stats_off) + which, (unsigned long long)v);
#else
        (which == STAT_LOOKUPS ? lookups : which == STAT_POPS ? pops : which == STAT_STEPS ? steps : search_lookups) += v;
#endif
    }
    HSA_HD void stat_max_steps(uint32_t v)
    {
#if defined(__CUDA_ARCH__)
        atomicMax(reinterpret_cast<unsigned long long *>(HSA_SMEM + P.smem_stats_off) + STAT_MAX_ITEM_STEPS, (unsigned long long)v);
#else
        if (v > max_item_steps) max_item_steps = v;
#endif
    }

    // ---------------------------------------------------------------- START: next work item in
    HSA_HD void start(uint32_t work_idx)
    {
        work = work_idx;
        const DevOpt *opts = reinterpret_cast<const DevOpt *>(HSA_SMEM);
        const TaskDesc t = make_task(P, opts, work_item(P, work_idx));
        rd = t.rd; rd_len = t.rd_len; strand = t.strand; sub_off = t.sub_off; len = t.len;
        seed_mode = t.seed_mode; opt_idx = t.opt_idx; out_idx = t.out_idx;
        row = P.rows + (size_t)work * P.row_stride;
        const uint32_t *tail = reinterpret_cast<const uint32_t *>(row + P.row_tail_off);
        if (tail[1] & ROW_FLAG_FILTERED) { st = LS_IDLE; return; }       // filtered by the width kernel (pass 1)
        const DevOpt &o = opt();
        seed_shift = seed_mode == SEED_TAIL ? len - (uint32_t)o.seed_len : 0u;     // ii = i - (len - seed_len), :253
        if (BIDS_SMEM) {                                   // bound bytes -> this lane's shared-memory region
            const uint32_t *src = reinterpret_cast<const uint32_t *>(row + P.row_bid_off);
            uint32_t *dst = reinterpret_cast<uint32_t *>(HSA_SMEM + sm_bid);
            for (uint32_t j = 0; j < (len + 4) / 4; ++j) dst[j] = src[j];
            if (seed_mode == SEED_TAIL) {
                const uint32_t *s2 = reinterpret_cast<const uint32_t *>(row + P.row_seed_off);
                uint32_t *d2 = reinterpret_cast<uint32_t *>(HSA_SMEM + sm_seed);
                for (uint32_t j = 0; j < ((uint32_t)o.seed_len + 4) / 4; ++j) d2[j] = s2[j];
            }
        }
        rd_word = 0; rd_widx = (tail[1] & ROW_FLAG_HAS_N) ? 0xFFFFFFFFu : (uint32_t)RD_IN_BYTES;
        lookups_item = 0; steps32 = 0; pops32 = 0; fail_code = STATUS_OK;
        mask0 = mask1 = 0; n_live = 0; n_phantom = 0; top = 0; free_head = NIL;
        best_score = (o.max_diff + 1) * o.s_mm + (o.max_gapo + 1) * o.s_gapo + (o.max_gape + 1) * o.s_gape; // :128
        pop_cut = (o.mode & MODE_NONSTOP) ? 0x7FFFFFFF : best_score + o.s_mm;
        max_diff = o.max_diff; best_cnt = 0; n_hits = 0;
        // root (bwtgap.c:142) is the first node popped; carry it directly
        direct = true; exact = false; pend = PEND_NONE; c_diff = false;
        ck = 0; cl = P.ix.fwd.text_length; crl = P.ix.fwd.text_length;
        ci = len; c_meta = 0; c_score = 0; c_nd = 0; zflags = 0; ci_at_pop = len;
        vet();
        if (st == LS_POP) st = LS_END;                     // root pruned (wrong strand): the stack is empty
    }

    // ---------------------------------------------------------------- pop-time tests (bwtgap.c:150-186)
    // The candidate is in the c* registers.  Sets st: LS_POP (candidate dropped), LS_LOOKUP (needs its occ4
    // pair: expansion, bwt_match_exact step, or materialisation), LS_HIT, LS_END.
    HSA_HD void vet()
    {
        const DevOpt &o = opt();
        if (direct) {
            // the carried child would have been pushed and popped: same loop-top test, entry included (:150-151)
            if (n_live + n_phantom + 1 > (uint32_t)o.max_entries) { st = LS_END; return; }
            direct = false;
        }
        m_cur = max_diff - c_nd;                                                     // :161-164
        // (single-exit form: every early return costs a set of register moves on its way to the loop's back edge)
        const bool drop = m_cur < 0 || (ci > 0 && m_cur < (int32_t)(bb(ci ? ci - 1 : 0) & BB_BID));      // :172-173
        if (drop) st = LS_POP;
        else if (pend) st = LS_LOOKUP;                     // survived pop-time pruning: materialise it first
        else classify();
    }

    // hit test / bwt_match_exact entry (bwtgap.c:176-186) of a candidate whose interval is known
    HSA_HD void classify()
    {
        const DevOpt &o = opt();
        ci_at_pop = ci;
        const bool ex = ci != 0 && m_cur == 0 &&                                      // :177-180
                        (c_state() == ST_M || (o.mode & MODE_GAPE) || (int32_t)c_gape() == o.max_gape);
        if (ex) {
            exact = true;
            // only the root has k == 0 (k' = C[c] + occ + 1 >= 1), and with it rev_k == 0; l, rev_l and every other
            // node's rev_k (= the parent's rev_k + the interval's smaller symbols) are >= 1
            zflags = 0;
            if (ck == 0) zflags = 1u | (cl == 0) << 1 | (crl - (cl - ck) == 0) << 2 | (crl == 0) << 3;
        }
        st = ci == 0 ? LS_HIT : LS_LOOKUP;
    }

    // ---------------------------------------------------------------- POP: gap_pop (bwtgap.c:80-92) + vet
    HSA_HD void do_pop()
    {
        const DevOpt &o = opt();
        ++steps32;
        // loop top of bwtgap.c:144-159
        if (n_live == 0 || n_live + n_phantom > (uint32_t)o.max_entries) { st = LS_END; return; }
        if (budget && steps32 > budget) { fail(STATUS_NEED_STRICT); st = LS_END; return; }   // heavy: hand it on
        const uint32_t b = bucket_lowest();
        const uint32_t ref = head_get(b);
        const uint32_t s = ref & NIL;
        const bool isB = (ref >> NEXT_BITS) != 0;
        u32x4 e, aux;
        ld_slot(slot_at(s), e, aux);
        const LinkT lw = link_of(aux);
        ++pops32;
        --n_live;
        if ((int32_t)b > pop_cut) { st = LS_END; return; }                            // :158-159
        uint32_t halfA = (uint32_t)(lw & (LinkT)HALF_MASK), halfB = (uint32_t)(lw >> HALF_BITS);
        uint32_t half = isB ? halfB : halfA;
        uint32_t cmask = half >> MASK_SHIFT;
        const uint32_t nref = half & REF_MASK;
        const uint32_t j = 31u - (uint32_t)clz32(cmask);                // last pushed child first
        cmask &= ~(1u << j);
        half = nref | cmask << MASK_SHIFT;
        if (isB) halfB = half; else halfA = half;
        LinkT nw = (LinkT)halfA | (LinkT)halfB << HALF_BITS;
        if (cmask == 0) {
            // this membership is exhausted: it leaves its bucket list
            if ((nref & NIL) == NIL) bucket_clear(b); else head_set(b, nref);
            const uint32_t other = isB ? halfA >> MASK_SHIFT : halfB >> MASK_SHIFT;
            if (other == 0) { nw = (LinkT)free_head; free_head = s; }
        }
        *link_at(s) = nw;
        // the child (bwtgap.c:267-314) as a candidate; its score is the bucket it was filed under
        const uint32_t pm = e.w;
        const uint32_t pi = pm & 0xFFFu, pst = (pm >> META_STATE_SHIFT) & 3u;
        const bool gape_counts = (o.mode & MODE_GAPE) != 0;
        ck = e.x; cl = e.y; crl = e.z;
        c_score = (int32_t)b;
        c_nd = (int32_t)(((pm >> META_MM_SHIFT) & 31u) + ((pm >> META_GO_SHIFT) & 15u) +
                         (gape_counts ? (pm >> META_GE_SHIFT) & 31u : 0u));
        c_diff = true; direct = false; exact = false;
        uint32_t m = pm & ~(0xFFFu | 3u << META_STATE_SHIFT);          // counts only, state M
        if (isB) {                                          // mismatch j+1 (bwtgap.c:303-313)
            ci = pi - 1; m += 1u << META_MM_SHIFT; ++c_nd; pend = PEND_MM | j;
        } else {
            if (pst == ST_M) { m += 1u << META_GO_SHIFT; ++c_nd; }          // gap open (:274, :281)
            else { m += 1u << META_GE_SHIFT; c_nd += gape_counts ? 1 : 0; } // gap extension (:284, :297)
            if (j == 0) { ci = pi - 1; m |= ST_I << META_STATE_SHIFT; pend = PEND_NONE; }        // insertion: same interval
            else { ci = pi; m |= ST_D << META_STATE_SHIFT; pend = PEND_DEL | (j - 1); }       // deletion of symbol j-1
        }
        c_meta = m;
        vet();
#if !defined(__CUDA_ARCH__)
        ++hsa_host_pop_out[st == LS_POP ? 0 : st == LS_LOOKUP ? (pend ? 1 : 2) : st == LS_HIT ? 3 : 4];
#endif
    }

    // ---------------------------------------------------------------- LOOKUP: one occ4 pair + what follows
    // A LOOKUP step runs in three stages so that the kernel can re-converge the warp between them (the lanes reach
    // the long stages by different routes, and without a barrier each route would run them on its own):
    //   lookup_a      the occ4 pair; complete for materialisations and bwt_match_exact steps (-> LK_DONE); for a
    //                 node expansion decides whether children that differ from the read may exist (-> LK_PUSH) or
    //                 only the exact-match child (-> LK_CHILD)
    //   lookup_push   files those children as one stack record (-> LK_CHILD, or LK_DONE if the item failed)
    //   lookup_child  continues with the exact-match child in registers
    enum : uint32_t { LK_DONE = 0, LK_PUSH = 1, LK_CHILD = 2 };
    struct LookupCarry { uint32_t nk, nl, nr, vmask, sc, i; bool allow_M; };

    HSA_HD void do_lookup()
    {
        LookupCarry c;
        uint32_t stage = lookup_a(c);
        if (stage == LK_PUSH) stage = lookup_push(c);
        if (stage == LK_CHILD) lookup_child(c);
    }

    HSA_HD uint32_t lookup_a(LookupCarry &out)
    {
        const DevOpt &o = opt();
        ++steps32;
        // ---- the one memory operation every kind of step shares: occ4 at k and at l + 1 -----------------
        const DevBwt &B = P.ix.fwd;
        uint32_t pk = ck, pl = cl + 1;
        pk -= (pk > B.inverse_sa0); pl -= (pl > B.inverse_sa0);
        u32x4 kc, kw, lc, lw;
        ld_sector(B.blocks + 2 * (size_t)(pk >> 6), kc, kw);
        ld_sector(B.blocks + 2 * (size_t)(pl >> 6), lc, lw);
#if !defined(__CUDA_ARCH__)
        ++hsa_host_pair_count; hsa_host_pair_same_sector += (pk >> 6) == (pl >> 6);    // emulation only: SURVEY 8d's deduplicated figure
        { const uint32_t dp = len - ci > 63u ? 63u : len - ci; ++hsa_host_hist[pend ? 2 : exact ? 1 : 0][(pk >> 6) == (pl >> 6)][dp]; }
#endif
        // i: index of the base the lookup extends by.  For a pending child the lookup is the PARENT's: deletion
        // children have ci == the parent's pre-decrement i, mismatch children ci == its post-decrement i.
        const uint32_t i = (pend & PEND_MM) ? ci : ci - 1;
        const uint32_t byte_i = bb(i);                       // bound byte of step i: lower bound, equal-width flag, base
        const uint32_t sc_ = base_step(i, byte_i);
        uint32_t oL[4], oR[4], sk[4], sl[4], rsl[4];
        occ4_from_sector(kc, kw, pk & 63u, oL);
        occ4_from_sector(lc, lw, pl & 63u, oR);
        {
            // BWTAllSARangesBackward_Bidirection, 2BWT-Interface.c:235-271
            uint32_t oc = 0;
            for (int c = 3; c >= 0; --c) {
                sk[c] = B.cum[c] + oL[c] + 1;
                sl[c] = B.cum[c] + oR[c];
                rsl[c] = crl - oc;
                oc += oR[c] - oL[c];
            }
        }
        const uint32_t vmask = (uint32_t)(sk[0] <= sl[0]) | (uint32_t)(sk[1] <= sl[1]) << 1 |
                               (uint32_t)(sk[2] <= sl[2]) << 2 | (uint32_t)(sk[3] <= sl[3]) << 3;
        // the one child every kind of step continues with: the pending child, else the read's base
        const uint32_t csel = ((pend & PEND_DEL) ? pend : (pend & PEND_MM) ? sc_ + pend + 1u : sc_) & 3u;
        const uint32_t nk = sel4(sk, csel), nl = sel4(sl, csel), nr = sel4(rsl, csel);
        const bool alive = (vmask >> csel) & 1u;

        out.nk = nk; out.nl = nl; out.nr = nr; out.vmask = vmask; out.sc = sc_; out.i = i; out.allow_M = true;
        if (pend) {
            // the child the reference pushed at bwtgap.c:281/297 (deletion) or :312 (mismatch): its interval
            ck = nk; cl = nl; crl = nr; pend = PEND_NONE;
            classify();
            return LK_DONE;
        }
        if (exact) {
            // one step of bwt_match_exact (2BWT-Interface.c:365-388)
            if (sc_ > 3) { st = LS_POP; return LK_DONE; }           // :376-377 (no lookup issued there)
            lookups_item += 2;
            if (!alive) { st = LS_POP; return LK_DONE; }
            ck = nk; cl = nl; crl = nr; ci = i;
            if (ci == 0) st = LS_HIT;
            return LK_DONE;
        }

        // ---- node expansion (bwtgap.c:244-325) --------------------------------------------------------
        lookups_item += 2;
        const int32_t m = m_cur;
        bool allow_diff = true, allow_M = true;
        if (i > 0) {                                                                        // :252-265
            const uint32_t b0 = bb(i - 1), b1 = byte_i;
            if ((int32_t)(b0 & BB_BID) > m - 1) allow_diff = false;
            else if ((int32_t)(b0 & BB_BID) == m - 1 && (int32_t)(b1 & BB_BID) == m - 1 && (b1 & BB_EQ)) allow_M = false;
            const int32_t ii = (int32_t)i - (int32_t)seed_shift;                            // :253
            if (seed_mode != SEED_NONE && ii > 0) {
                const int32_t m_seed = o.max_seed_diff - c_nd;                              // :167-171
                const uint32_t s0 = bs((uint32_t)ii - 1), s1 = bs((uint32_t)ii);
                if ((int32_t)(s0 & BB_BID) > m_seed - 1) allow_diff = false;
                else if ((int32_t)(s0 & BB_BID) == m_seed - 1 && (int32_t)(s1 & BB_BID) == m_seed - 1 && (s1 & BB_EQ)) allow_M = false;
            }
        }
        out.allow_M = allow_M;
        return allow_diff ? LK_PUSH : LK_CHILD;
    }

    // node expansion, bwtgap.c:267-314: the children that differ from the read, as ONE stack record
    HSA_HD uint32_t lookup_push(const LookupCarry &in)
    {
        const DevOpt &o = opt();
        const uint32_t vmask = in.vmask, sc_ = in.sc, i = in.i;
        const bool allow_M = in.allow_M;
        {
            const uint32_t e_go = c_gapo(), e_ge = c_gape(), e_state = c_state();
            uint32_t maskA = 0, maskB = 0;
            int32_t tmp;
            if (o.mode & MODE_LOGGAP) {                                                     // :267 + int_log2 :107-116
                uint32_t v = e_ge + e_go; int32_t lg = 0;
                while (v > 1) { v >>= 1; ++lg; }
                tmp = lg / 2 + 1;
            } else tmp = (int32_t)(e_go + e_ge);
            if ((int32_t)i >= o.indel_end_skip + tmp && (int32_t)len - (int32_t)i >= o.indel_end_skip + tmp) {
                bool ins = false, del = false;
                if (e_state == ST_M) ins = del = (int32_t)e_go < o.max_gapo;               // :269-282
                else if (e_state == ST_I) ins = (int32_t)e_ge < o.max_gape;                // :283-285
                else del = (int32_t)e_ge < o.max_gape &&                                   // :286-299
                           ((int32_t)(e_ge + e_go) < max_diff || cl - ck + 1 < (uint32_t)o.max_del_occ);
                maskA = (ins ? 1u : 0u) | (del ? vmask << 1 : 0u);
            }
            if (allow_M) {                                                                  // :302-314
                // children j = 1..3 (and j = 4 when seq[i] is N) are mismatches: bit j-1
                // = vmask rotated right by sc_ + 1 (a 4-bit rotation: two copies side by side, shifted)
                maskB = ((vmask * 0x11u) >> ((sc_ + 1u) & 3u)) & (sc_ > 3 ? 15u : 7u);
            }
            const int32_t gsc = c_score + (e_state == ST_M ? o.s_gapo : o.s_gape), msc = c_score + o.s_mm;
            const int32_t cut = n_hits ? pop_cut : 0x7FFFFFFF;      // bwtgap.c:158-159: can never be popped -> count only
            uint32_t nA = (uint32_t)popc32(maskA), nB = (uint32_t)popc32(maskB);
            if (gsc > cut) { n_phantom += nA; maskA = 0; nA = 0; }
            if (msc > cut) { n_phantom += nB; maskB = 0; nB = 0; }
            if (maskA | maskB) {
                const uint32_t hi_sc = (uint32_t)((maskA ? gsc : 0) > (maskB ? msc : 0) ? (maskA ? gsc : 0) : (maskB ? msc : 0));
                if (hi_sc >= P.n_buckets)
                    fail(WIDE ? STATUS_BAD_SCORE : STATUS_NEED_STRICT);      // the fast kernel has 64 buckets
                else {
                    uint32_t s = NIL;
                    if (free_head != NIL) { s = free_head; free_head = (uint32_t)(*link_at(s) & (LinkT)NIL); }
                    else if (top < P.arena_cap) s = top++;
                    else fail(STATUS_NEED_STRICT);
                    if (s != NIL) {
                        u32x4 e;
                        e.x = ck; e.y = cl; e.z = crl; e.w = c_meta | (i + 1);
                        uint32_t halfA = NIL, halfB = NIL;
                        if (maskA) {
                            halfA = (bucket_nonempty((uint32_t)gsc) ? head_get((uint32_t)gsc) : NIL) | maskA << MASK_SHIFT;
                            head_set((uint32_t)gsc, s); bucket_set((uint32_t)gsc);
                        }
                        if (maskB) {    // after A: if both share a bucket, B links to the record's own half A
                            halfB = (bucket_nonempty((uint32_t)msc) ? head_get((uint32_t)msc) : NIL) | maskB << MASK_SHIFT;
                            head_set((uint32_t)msc, s | 1u << NEXT_BITS); bucket_set((uint32_t)msc);
                        }
                        st_slot(slot_at(s), e, aux_of((LinkT)halfA | (LinkT)halfB << HALF_BITS));
                        n_live += nA + nB;
                    }
                }
                if (fail_code != STATUS_OK) { st = LS_END; return LK_DONE; }
            }
        }
        return LK_CHILD;
    }

    // the exact-match child (bwtgap.c:303-313 with j = 4, or :315-325): would be pushed last into the lowest
    // bucket and popped next -> it stays in registers.  Same counts as its parent, so m_cur stands.
    HSA_HD void lookup_child(const LookupCarry &in)
    {
        const DevOpt &o = opt();
        const bool exists = in.sc < 4 && ((in.vmask >> (in.sc & 3u)) & 1u);
        // (values without meaning when the child does not exist or is dropped: the next POP overwrites all of them)
        ck = in.nk; cl = in.nl; crl = in.nr; ci = in.i; c_diff = false;
        c_meta &= ~(3u << META_STATE_SHIFT);            // STATE_M
        const bool over = n_live + n_phantom + 1 > (uint32_t)o.max_entries;                           // :150-151
        const bool pruned = ci > 0 && m_cur < (int32_t)(bb(ci ? ci - 1 : 0) & BB_BID);                  // :172-173
        if (!exists) st = LS_POP;
        else if (over) st = LS_END;
        else if (pruned) st = LS_POP;
        else classify();
    }

    // ---------------------------------------------------------------- HIT: action for found hits, bwtgap.c:188-241
    HSA_HD void do_hit()
    {
        const DevOpt &o = opt();
        uint32_t k = ck, l = cl, rk = crl - (cl - ck), rl = crl;
        if (exact) {
            // matched down to the first base: write-back quirk of 2BWT-Interface.c:383-386
            if (zflags & 1u) k = 0;
            if (zflags & 2u) l = 0;
            if (zflags & 4u) rk = 0;
            if (zflags & 8u) rl = 0;
        }
        st = LS_POP;
        const int32_t score = c_score;
        Hit *hs = P.hits + (size_t)slot * P.hit_cap;
        if (n_hits == 0) {
            best_score = score;
            const int32_t best_diff = c_nd;                 // n_mm + n_gapo (+ n_gape with MODE_GAPE), :192-196
            if (!(o.mode & MODE_NONSTOP)) {
                max_diff = (best_diff + 1 > o.max_diff) ? o.max_diff : best_diff + 1;
                pop_cut = best_score + o.s_mm;
            }
        }
        if (score == best_score) best_cnt = (int32_t)((uint32_t)best_cnt + (l - k + 1));
        else if (best_cnt > o.max_top2) { st = LS_END; return; }                      // top2b break
        if (c_gapo())
            for (uint32_t j = 0; j < n_hits; ++j)
                if (hs[j].k == k && hs[j].l == l) return;                              // :205-213 already found
        // gap_shadow (bwtgap.c:94-105) on width_back, in place; the bound bytes follow
        const uint32_t x = l - k + 1, ldp = c_diff ? ci_at_pop : 0u;
        uint32_t *w = reinterpret_cast<uint32_t *>(row);
        uint32_t jj = 0, w_prev = 0xFFFFFFFFu;
        for (uint32_t i = 0; i < ldp; ++i) {
            const uint32_t old = bb(i);
            uint32_t v = w[i], bid = old & BB_BID;
            if (v > x) { v -= x; w[i] = v; }
            else if (v == x) { bid = 1; v = P.ix.fwd.text_length - (++jj); w[i] = v; }
            bb_set(i, bound_byte(bid, v, w_prev) | (old & BB_BASE));
            w_prev = v;
        }
        if (ldp > 0 && ldp <= len) { const uint32_t old = bb(ldp); bb_set(ldp, bound_byte(old & BB_BID, w[ldp], w_prev) | (old & BB_BASE)); }
        if (n_hits >= P.hit_cap) { fail(STATUS_NEED_STRICT); st = LS_END; return; }
        Hit h;
        h.k = k; h.l = l; h.rev_k = rk; h.rev_l = rl;
        h.counts = c_mm() | c_gapo() << 16 | c_gape() << 24;
        h.score = score; h.pad0 = h.pad1 = 0;
        hs[n_hits++] = h;
    }

    // ---------------------------------------------------------------- END: results out
    HSA_HD void finish_item(uint32_t item, uint32_t n, uint32_t strand_stamp)
    {
        uint64_t off = 0;
        uint8_t stt = STATUS_OK;
        if (n) {
#if defined(__CUDA_ARCH__)
            off = atomicAdd(&P.counters[CNT_ALN], (unsigned long long)n);
#else
            off = P.counters[CNT_ALN]; P.counters[CNT_ALN] += n;
#endif
            if (off + n > P.aln_cap) { stt = STATUS_OUT_FULL; n = 0; }
            const Hit *hs = P.hits + (size_t)slot * P.hit_cap;
            for (uint32_t j = 0; j < n; ++j) {
                uint32_t *w = P.aln + (off + j) * 9;
                const Hit h = hs[j];
                w[0] = h.counts; w[1] = h.k; w[2] = h.l; w[3] = h.rev_k; w[4] = h.rev_l;
                w[5] = strand_stamp << 30;                       // type:30 = 0, strand:2
                int32_t s0 = 0, e0 = 0;
                if (P.kind == KIND_SEEDS) { s0 = (int32_t)sub_off; e0 = (int32_t)(sub_off + len - 1); }   // bwtgap.c:816-819
                else if (P.kind == KIND_WHOLE && j == 0) { s0 = 0; e0 = (int32_t)rd_len - 1; } // bwtaln.c:371-372
                w[6] = (uint32_t)s0; w[7] = (uint32_t)e0; w[8] = (uint32_t)h.score;
            }
        }
        P.n_aln[item] = (int32_t)n;
        P.aln_off[item] = off;
        P.status[item] = stt;
    }

    HSA_HD void do_end()
    {
        stat_max_steps(steps32);
        stat_add(STAT_STEPS, steps32); stat_add(STAT_POPS, pops32);
        st = LS_IDLE;
        if (fail_code != STATUS_OK) {
            // discard what this item produced; the host re-runs it with the large-capacity kernel
            P.n_aln[out_idx] = 0; P.aln_off[out_idx] = 0; P.status[out_idx] = (uint8_t)fail_code;
            unsigned long long *cnt = fail_code == STATUS_NEED_STRICT ? P.strict_count : &P.counters[CNT_BAD];
#if defined(__CUDA_ARCH__)
            const unsigned long long idx = atomicAdd(cnt, 1ull);
#else
            const unsigned long long idx = (*cnt)++;
#endif
            if (fail_code == STATUS_NEED_STRICT && P.strict_list) P.strict_list[idx] = out_idx;
            return;
        }
        stat_add(STAT_SEARCH_LOOKUPS, lookups_item);
        const uint64_t mine = (uint64_t)lookups_item + reinterpret_cast<const uint32_t *>(row + P.row_tail_off)[0];
        if (P.kind == KIND_WHOLE && P.pass == 1 && n_hits == 0) {
            // bwtaln.c:351-358: nothing on the reverse-complement strand -> the forward strand is searched (pass 2).
            // The read's pass-1 lookup count rides in aln_off[] until pass 2 completes, so that a read which
            // is later re-run with the large-capacity kernel is not counted twice.
            P.aln_off[out_idx] = mine;
#if defined(__CUDA_ARCH__)
            const uint32_t idx = atomicAdd(P.next_count, 1u);
#else
            const uint32_t idx = (*P.next_count)++;
#endif
            P.next_list[idx] = out_idx;
            return;
        }
        stat_add(STAT_LOOKUPS, mine + ((P.kind == KIND_WHOLE && P.pass == 2) ? P.aln_off[out_idx] : 0ull));
        finish_item(out_idx, n_hits, strand);
        if (P.kind == KIND_TASKS && P.width_out) {
            // per-call form (hsa_match_gap_call): width_back is an in/out argument of bwt_match_gap -- gap_shadow rewrites
            // it in place (bwtgap.c:94-105, :217) and the splice path reads it again afterwards (bwtgap.c:1176-1192)
            const uint32_t *w = reinterpret_cast<const uint32_t *>(row);
            for (uint32_t i = 0; i <= len; ++i) { u32x2 v; v.x = w[i]; v.y = bb(i) & BB_BID; P.width_out[i] = v; }
        }
    }
};

} // namespace hsa

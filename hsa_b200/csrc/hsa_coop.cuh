// hsa_coop.cuh -- warp-cooperative bwt_match_gap for HEAVY searches: one warp per search.
//
// The fast kernel (hsa_core.cuh: Worker) runs one search per lane; a search that needs > 10^4 steps then
// holds a whole warp (and, at the end of a launch, the whole GPU) hostage at lone-lane latency.  Such
// searches are handed to this kernel, which finds parallelism INSIDE one search without changing its result:
//
//   * the reference pops the entries of its lowest-score bucket one after the other (LIFO, bwtgap.c:80-92);
//     the node popped and everything expanded from it in registers ("chain": its exact-match descendants,
//     bwtgap.c:303-325) only PUSHES into higher buckets (scores are positive), so until a hit is found the
//     chains of one bucket are independent of each other;
//   * a WAVE takes the next (up to) 32 children of the lowest bucket in exactly the reference's pop order,
//     runs their chains in the 32 lanes speculatively, and then COMMITS them in that order: the pushed
//     records are appended to their buckets as the sequential run would have left them (prefix sums over
//     per-lane, per-bucket counts give every record its position);
//   * a chain that ends in a hit stops the commit after itself: the hit is recorded (first-hit max_diff
//     shrink, best_cnt / max_top2, dedupe, gap_shadow -- bwtgap.c:188-241), the later chains of the wave are
//     thrown away and re-run in the next wave under the new state;
//   * the max_entries test (bwtgap.c:150-151) is checked against an upper bound for the whole wave; if that
//     bound fails the wave is re-run one chain at a time with the exact test.
//
// Buckets here are the reference's own representation -- per-score LIFO arrays (in chunks of 32 records) --
// because a wave must read the top 32 records of a bucket at once.  Records are the lazily expanded
// "families" of the fast kernel, one record per membership.
//
// The orchestration is written as a sequence of lane-parallel phases that communicate through the warp's
// shared-memory block only, so the SAME source runs on the GPU (one thread per lane, __syncwarp between
// phases) and in the host emulation (a loop over 32 lanes per phase).
#pragma once
#include "hsa_core.cuh"

namespace hsa {

enum : uint32_t { COOP_CHUNK = 32, COOP_OUT_CAP = 192, COOP_NB = 64, COOP_HIT_CAP = 4096 };
enum : uint32_t { CH_IDLE = 0, CH_RUN = 1, CH_DEAD = 2, CH_HIT = 3, CH_SUSP = 4, CH_ENTRIES = 5, CH_FAIL = 6 };
enum : uint32_t { COOP_NONE = 0xFFFFFFFFu };

// shared-memory read-modify-writes the lanes of one phase may issue together (plain in the host emulation)
HSA_HD uint32_t coop_add32(uint32_t *p, uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, v);
#else
    const uint32_t o = *p; *p = o + v; return o;
#endif
}
HSA_HD void coop_or32(uint32_t *p, uint32_t v)
{
#if defined(__CUDA_ARCH__)
    atomicOr(p, v);
#else
    *p |= v;
#endif
}
HSA_HD void coop_max32(uint32_t *p, uint32_t v)
{
#if defined(__CUDA_ARCH__)
    atomicMax(p, v);
#else
    if (v > *p) *p = v;
#endif
}
HSA_HD void coop_or64(unsigned long long *p, unsigned long long v)
{
#if defined(__CUDA_ARCH__)
    atomicOr(p, v);
#else
    *p |= v;
#endif
}
HSA_HD void coop_and64(unsigned long long *p, unsigned long long v)
{
#if defined(__CUDA_ARCH__)
    atomicAnd(p, v);
#else
    *p &= v;
#endif
}

struct CoopCand {                   // one search node in flight (the c* registers of the fast kernel's Worker)
    uint32_t ck, cl, crl, ci, c_meta, pend, ci_at_pop, zflags;
    int32_t c_score, c_nd, m_cur;
    uint32_t flags;                 // bit 0 exact, bit 1 c_diff, bit 2 direct (carried root), bit 3 resume (already vetted)
};

struct CoopWarp {                   // one per warp, in shared memory
    // task
    const uint8_t *rd; uint8_t *row;
    uint32_t rd_len, strand, sub_off, len, seed_mode, seed_shift, opt_idx, out_idx, work, have_item;
    // search state
    unsigned long long mask;        // non-empty buckets
    uint32_t n_live, n_phantom, n_chunks, n_hits, fail_code, done;
    int32_t best_score, max_diff, best_cnt, pop_cut;
    uint32_t top_chunk[COOP_NB]; uint8_t top_cnt[COOP_NB];
    // wave
    uint32_t b, T, W, carried, width, exact_mode, sticky_exact, n_valid, ev_lane, ev_kind, redo, base_entries;
    uint32_t S[COOP_CHUNK + 1]; uint8_t slot_owner[COOP_CHUNK]; uint8_t cc[COOP_CHUNK];      // cc: children per loaded record
    uint32_t evmask, acc_pushed, acc_lookups, acc_steps, longest, n_full;                        // wave accumulators
    unsigned long long touched;                                                                  // buckets the committing lanes pushed to
    u32x4 ent_payload[COOP_CHUNK]; uint32_t ent_info[COOP_CHUNK];
    uint16_t cnt[32][COOP_NB];      // pushes per (lane, bucket) of the wave; exclusive prefix over lanes after the scan
    uint32_t newbase[COOP_NB]; uint16_t total[COOP_NB];
    uint8_t lane_status[32]; uint16_t lane_nout[32];
    CoopCand carry;
    // statistics of the item / the warp
    uint32_t lookups_item, steps_item, pops_item, waves_item;
    unsigned long long lookups, pops, steps, waves, wave_steps;   // wave_steps: sum over waves of the longest chain
    unsigned long long prof_chain, prof_commit, prof_total, prof_t0, prof_t1;   // -DHSA_PHASE_PROF: cycles in the chain phase / rest
};

struct CoopScratch {                // per-warp global memory (Params::coop_* point at warp 0's)
    u32x4 *payload; uint32_t *info; uint32_t *prev;     // bucket records: chunk c = entries [32c, 32c+32)
    u32x4 *out_payload; uint32_t *out_info;             // per-lane push buffers of the current wave
    Hit *hits;
};

// ---- per-lane chain execution ---------------------------------------------------------------------------
struct CoopLane {
    CoopCand c;
    uint32_t status, n_out, live, ph, lookups, steps;
    uint64_t touched;               // buckets this lane's chain pushed to (its non-zero entries of CoopWarp::cnt)
};

struct Coop {
    const Params &P;
    CoopWarp &sh;
    uint8_t *bids;                  // bound bytes of width_back, then of width_seed (shared memory, after sh)
    CoopScratch g;
    uint32_t cap_chunks;

    HSA_HD const DevOpt &opt() const { return reinterpret_cast<const DevOpt *>(HSA_SMEM)[sh.opt_idx]; }
    HSA_HD uint32_t bb(uint32_t i) const { return bids[i]; }
    HSA_HD uint32_t bs(uint32_t i) const { return sh.seed_mode == SEED_ALIAS ? bids[i] : bids[P.row_seed_off - P.row_bid_off + i]; }
    HSA_HD uint32_t base_at(uint32_t p) const
    {
        const uint32_t c = ld_ro_u8(sh.strand ? sh.rd + (sh.rd_len - 1 - p) : sh.rd + p);
        return (sh.strand && c < 4) ? 3 - c : c;
    }

    // pop-time tests (bwtgap.c:150-186), as Worker::vet / classify
    HSA_HD void classify(CoopLane &me) const
    {
        const DevOpt &o = opt();
        me.c.ci_at_pop = me.c.ci;
        if (me.c.ci == 0) { me.status = CH_HIT; return; }
        const uint32_t st = (me.c.c_meta >> META_STATE_SHIFT) & 3u, ge = (me.c.c_meta >> META_GE_SHIFT) & 31u;
        if (me.c.m_cur == 0 && (st == ST_M || (o.mode & MODE_GAPE) || (int32_t)ge == o.max_gape)) {
            me.c.flags |= 1u;
            me.c.zflags = (me.c.ck == 0) | (me.c.cl == 0) << 1 | (me.c.crl - (me.c.cl - me.c.ck) == 0) << 2 | (me.c.crl == 0) << 3;
        }
        me.status = CH_RUN;
    }
    HSA_HD bool entries_exceeded(const CoopLane &me) const      // exact mode only: the loop-top test for a carried child
    {
        return sh.exact_mode && (int64_t)sh.base_entries + me.live + me.ph + 1 > (int64_t)opt().max_entries;
    }
    HSA_HD void vet(CoopLane &me) const
    {
        if (me.c.flags & 4u) {                              // carried root: loop-top test, entry included
            if (entries_exceeded(me)) { me.status = CH_ENTRIES; return; }
            me.c.flags &= ~4u;
        }
        me.c.m_cur = sh.max_diff - me.c.c_nd;
        if (me.c.m_cur < 0) { me.status = CH_DEAD; return; }
        if (me.c.ci > 0 && me.c.m_cur < (int32_t)(bb(me.c.ci - 1) & BB_BID)) { me.status = CH_DEAD; return; }
        if (me.c.pend) { me.status = CH_RUN; return; }
        classify(me);
    }

    // record one push of the current chain (a family of children of the node just expanded)
    HSA_HD void emit(CoopLane &me, uint32_t lane, uint32_t bucket, uint32_t cmask, uint32_t kind, const u32x4 &payload)
    {
        const uint32_t rank = sh.cnt[lane][bucket]++;
        me.touched |= 1ull << bucket;
        g.out_payload[(size_t)lane * COOP_OUT_CAP + me.n_out] = payload;
        g.out_info[(size_t)lane * COOP_OUT_CAP + me.n_out] = bucket | cmask << 8 | kind << 15 | rank << 16;
        ++me.n_out;
    }

    // one occ4 pair + what follows (Worker::do_lookup with the pushes redirected to emit)
    HSA_HD void lookup_step(CoopLane &me, uint32_t lane)
    {
        const DevOpt &o = opt();
        CoopCand &c = me.c;
        const bool is_exact = c.flags & 1u;
        if (!c.pend && !is_exact && me.n_out + 2 > COOP_OUT_CAP) { me.status = CH_SUSP; return; }   // before touching anything
        ++me.steps;
        const DevBwt &B = P.ix.fwd;
        uint32_t pk = c.ck, pl = c.cl + 1;
        pk -= (pk > B.inverse_sa0); pl -= (pl > B.inverse_sa0);
        u32x4 kc, kw, lc, lw;
        ld_sector(B.blocks + 2 * (size_t)(pk >> 6), kc, kw);
        ld_sector(B.blocks + 2 * (size_t)(pl >> 6), lc, lw);
        const uint32_t i = (c.pend & PEND_MM) ? c.ci : c.ci - 1;
        const uint32_t sc_ = base_at(sh.sub_off + i);
        uint32_t oL[4], oR[4], sk[4], sl[4], rsl[4];
        occ4_from_sector(kc, kw, pk & 63u, oL);
        occ4_from_sector(lc, lw, pl & 63u, oR);
        {
            uint32_t oc = 0;
            for (int x = 3; x >= 0; --x) {
                sk[x] = B.cum[x] + oL[x] + 1;
                sl[x] = B.cum[x] + oR[x];
                rsl[x] = c.crl - oc;
                oc += oR[x] - oL[x];
            }
        }
        const uint32_t vmask = (uint32_t)(sk[0] <= sl[0]) | (uint32_t)(sk[1] <= sl[1]) << 1 |
                               (uint32_t)(sk[2] <= sl[2]) << 2 | (uint32_t)(sk[3] <= sl[3]) << 3;
        const uint32_t csel = ((c.pend & PEND_DEL) ? c.pend : (c.pend & PEND_MM) ? sc_ + c.pend + 1u : sc_) & 3u;
        const uint32_t nk = sel4(sk, csel), nl = sel4(sl, csel), nr = sel4(rsl, csel);
        const bool alive = (vmask >> csel) & 1u;
        if (c.pend) {
            c.ck = nk; c.cl = nl; c.crl = nr; c.pend = PEND_NONE;
            classify(me);
            return;
        }
        if (is_exact) {
            if (sc_ > 3) { me.status = CH_DEAD; return; }
            me.lookups += 2;
            if (!alive) { me.status = CH_DEAD; return; }
            c.ck = nk; c.cl = nl; c.crl = nr; c.ci = i;
            if (c.ci == 0) me.status = CH_HIT;
            return;
        }
        // ---- node expansion (bwtgap.c:244-325) ----
        me.lookups += 2;
        const int32_t m = c.m_cur;
        bool allow_diff = true, allow_M = true;
        if (i > 0) {
            const uint32_t b0 = bb(i - 1), b1 = bb(i);
            if ((int32_t)(b0 & BB_BID) > m - 1) allow_diff = false;
            else if ((int32_t)(b0 & BB_BID) == m - 1 && (int32_t)(b1 & BB_BID) == m - 1 && (b1 & BB_EQ)) allow_M = false;
            const int32_t ii = (int32_t)i - (int32_t)sh.seed_shift;
            if (sh.seed_mode != SEED_NONE && ii > 0) {
                const int32_t m_seed = o.max_seed_diff - c.c_nd;
                const uint32_t s0 = bs((uint32_t)ii - 1), s1 = bs((uint32_t)ii);
                if ((int32_t)(s0 & BB_BID) > m_seed - 1) allow_diff = false;
                else if ((int32_t)(s0 & BB_BID) == m_seed - 1 && (int32_t)(s1 & BB_BID) == m_seed - 1 && (s1 & BB_EQ)) allow_M = false;
            }
        }
        if (allow_diff) {
            const uint32_t e_go = (c.c_meta >> META_GO_SHIFT) & 15u, e_ge = (c.c_meta >> META_GE_SHIFT) & 31u;
            const uint32_t e_state = (c.c_meta >> META_STATE_SHIFT) & 3u;
            uint32_t maskA = 0, maskB = 0;
            int32_t tmp;
            if (o.mode & MODE_LOGGAP) {
                uint32_t v = e_ge + e_go; int32_t lg = 0;
                while (v > 1) { v >>= 1; ++lg; }
                tmp = lg / 2 + 1;
            } else tmp = (int32_t)(e_go + e_ge);
            if ((int32_t)i >= o.indel_end_skip + tmp && (int32_t)sh.len - (int32_t)i >= o.indel_end_skip + tmp) {
                bool ins = false, del = false;
                if (e_state == ST_M) ins = del = (int32_t)e_go < o.max_gapo;
                else if (e_state == ST_I) ins = (int32_t)e_ge < o.max_gape;
                else del = (int32_t)e_ge < o.max_gape &&
                           ((int32_t)(e_ge + e_go) < sh.max_diff || c.cl - c.ck + 1 < (uint32_t)o.max_del_occ);
                maskA = (ins ? 1u : 0u) | (del ? vmask << 1 : 0u);
            }
            if (allow_M)
                maskB = ((vmask >> ((sc_ + 1u) & 3u)) & 1u) | ((vmask >> ((sc_ + 2u) & 3u)) & 1u) << 1 |
                        ((vmask >> ((sc_ + 3u) & 3u)) & 1u) << 2 | (sc_ > 3 ? (vmask & 1u) << 3 : 0u);
            const int32_t gsc = c.c_score + (e_state == ST_M ? o.s_gapo : o.s_gape), msc = c.c_score + o.s_mm;
            const int32_t cut = sh.n_hits ? sh.pop_cut : 0x7FFFFFFF;
            uint32_t nA = (uint32_t)popc32(maskA), nB = (uint32_t)popc32(maskB);
            if (gsc > cut) { me.ph += nA; maskA = 0; nA = 0; }
            if (msc > cut) { me.ph += nB; maskB = 0; nB = 0; }
            if (maskA | maskB) {
                if ((maskA && (uint32_t)gsc >= COOP_NB) || (maskB && (uint32_t)msc >= COOP_NB)) { me.status = CH_FAIL; return; }
                u32x4 e;
                e.x = c.ck; e.y = c.cl; e.z = c.crl; e.w = c.c_meta | (i + 1);
                if (maskA) emit(me, lane, (uint32_t)gsc, maskA, 0, e);      // pushed first (bwtgap.c:274-299) ...
                if (maskB) emit(me, lane, (uint32_t)msc, maskB, 1, e);      // ... then the mismatches (:303-313)
                me.live += nA + nB;
            }
        }
        if (sc_ < 4 && alive) {
            c.ck = nk; c.cl = nl; c.crl = nr; c.ci = i; c.flags &= ~2u;
            c.c_meta &= ~(3u << META_STATE_SHIFT);
            if (entries_exceeded(me)) { me.status = CH_ENTRIES; return; }
            if (c.ci > 0 && m < (int32_t)(bb(c.ci - 1) & BB_BID)) { me.status = CH_DEAD; return; }
            classify(me);
        } else me.status = CH_DEAD;
    }

    HSA_HD void run_chain(CoopLane &me, uint32_t lane)
    {
        if (me.status != CH_RUN) return;
        if (me.c.flags & 8u) me.c.flags &= ~8u;             // resumed chain: already vetted
        else vet(me);
        while (me.status == CH_RUN) lookup_step(me, lane);
    }

    // candidate = child `j` of record (payload, kind) popped from bucket b (Worker::do_pop's derivation)
    HSA_HD void child_of(CoopLane &me, const u32x4 &e, uint32_t kind, uint32_t j, uint32_t b) const
    {
        const DevOpt &o = opt();
        const uint32_t pm = e.w, pi = pm & 0xFFFu, pst = (pm >> META_STATE_SHIFT) & 3u;
        const bool gape_counts = (o.mode & MODE_GAPE) != 0;
        CoopCand &c = me.c;
        c.ck = e.x; c.cl = e.y; c.crl = e.z;
        c.c_score = (int32_t)b;
        c.c_nd = (int32_t)(((pm >> META_MM_SHIFT) & 31u) + ((pm >> META_GO_SHIFT) & 15u) + (gape_counts ? (pm >> META_GE_SHIFT) & 31u : 0u));
        c.flags = 2u;                                       // c_diff
        c.zflags = 0; c.m_cur = 0;
        uint32_t m = pm & ~(0xFFFu | 3u << META_STATE_SHIFT);
        if (kind) { c.ci = pi - 1; m += 1u << META_MM_SHIFT; ++c.c_nd; c.pend = PEND_MM | j; }
        else {
            if (pst == ST_M) { m += 1u << META_GO_SHIFT; ++c.c_nd; }
            else { m += 1u << META_GE_SHIFT; c.c_nd += gape_counts ? 1 : 0; }
            if (j == 0) { c.ci = pi - 1; m |= ST_I << META_STATE_SHIFT; c.pend = PEND_NONE; }
            else { c.ci = pi; m |= ST_D << META_STATE_SHIFT; c.pend = PEND_DEL | (j - 1); }
        }
        c.c_meta = m;
        c.ci_at_pop = c.ci;
    }
};

// ---- orchestration: lane-parallel phases over the warp's shared block ------------------------------------
#if defined(__CUDA_ARCH__)
#define COOP_FOR_LANES(lane) { const uint32_t lane = threadIdx.x & 31u;
#define COOP_END_LANES } __syncwarp();
#define COOP_ME(lane) me_
#else
#define COOP_FOR_LANES(lane) for (uint32_t lane = 0; lane < 32; ++lane) {
#define COOP_END_LANES }
#define COOP_ME(lane) me_[lane]
#endif

// the entry `t` positions below the top of bucket b (t < entries available in its top two chunks)
HSA_HD uint32_t coop_entry_index(const CoopWarp &sh, const uint32_t *prev, uint32_t b, uint32_t t)
{
    const uint32_t cnt = sh.top_cnt[b], chunk = sh.top_chunk[b];
    if (t < cnt) return chunk * COOP_CHUNK + (cnt - 1 - t);
    return prev[chunk] * COOP_CHUNK + (COOP_CHUNK - 1 - (t - cnt));
}

// Runs the searches of this warp's queue share.  `me_` is the lane-private state: one CoopLane on the device, an
// array of 32 in the emulation.
template <typename MeT>
HSA_HD void coop_run(const Params &P, CoopWarp &sh, uint8_t *bids, const CoopScratch &g, uint32_t cap_chunks,
                     uint32_t n_work, MeT &me_)
{
    Coop C{P, sh, bids, g, cap_chunks};
    const DevOpt *opts = reinterpret_cast<const DevOpt *>(HSA_SMEM);
    COOP_FOR_LANES(lane)
    { uint32_t *row = reinterpret_cast<uint32_t *>(&sh.cnt[lane][0]); for (uint32_t q = 0; q < COOP_NB / 2; ++q) row[q] = 0; COOP_ME(lane).touched = 0; }
    COOP_END_LANES
    COOP_FOR_LANES(lane) if (lane == 0) { sh.have_item = 0; sh.lookups = sh.pops = sh.steps = sh.waves = sh.wave_steps = 0; sh.prof_chain = sh.prof_commit = sh.prof_total = 0; } COOP_END_LANES
    for (;;) {
        // ---------------------------------------------------------------- next item
        COOP_FOR_LANES(lane)
        if (lane == 0) {
#if defined(__CUDA_ARCH__)
            const unsigned long long idx = atomicAdd(P.cursor, 1ull);
#else
            const unsigned long long idx = (*P.cursor)++;
#endif
            sh.have_item = idx < n_work;
            if (sh.have_item) {
                sh.work = (uint32_t)idx;
                const TaskDesc t = make_task(P, opts, work_item(P, sh.work));
                sh.rd = t.rd; sh.rd_len = t.rd_len; sh.strand = t.strand; sh.sub_off = t.sub_off; sh.len = t.len;
                sh.seed_mode = t.seed_mode; sh.opt_idx = t.opt_idx; sh.out_idx = t.out_idx;
                sh.row = P.rows + (size_t)sh.work * P.row_stride;
                const DevOpt &o = opts[t.opt_idx];
                sh.seed_shift = t.seed_mode == SEED_TAIL ? t.len - (uint32_t)o.seed_len : 0u;
                sh.mask = 0; sh.n_live = 0; sh.n_phantom = 0; sh.n_chunks = 0; sh.n_hits = 0; sh.fail_code = STATUS_OK;
                sh.best_score = (o.max_diff + 1) * o.s_mm + (o.max_gapo + 1) * o.s_gapo + (o.max_gape + 1) * o.s_gape;
                sh.pop_cut = (o.mode & MODE_NONSTOP) ? 0x7FFFFFFF : sh.best_score + o.s_mm;
                sh.max_diff = o.max_diff; sh.best_cnt = 0;
                sh.lookups_item = sh.steps_item = sh.pops_item = sh.waves_item = 0;
                const uint32_t *tail = reinterpret_cast<const uint32_t *>(sh.row + P.row_tail_off);
                sh.done = (tail[1] & ROW_FLAG_FILTERED) ? 2u : 0u;          // 2: nothing to do, nothing to write
                // the root (bwtgap.c:142) is the first node popped: it is carried into wave 0
                CoopCand &c = sh.carry;
                c.ck = 0; c.cl = P.ix.fwd.text_length; c.crl = P.ix.fwd.text_length; c.ci = t.len; c.c_meta = 0;
                c.pend = PEND_NONE; c.ci_at_pop = t.len; c.zflags = 0; c.c_score = 0; c.c_nd = 0; c.m_cur = 0; c.flags = 4u;
                sh.carried = 1; sh.width = 32; sh.redo = 0; sh.sticky_exact = 0;
            }
        }
        COOP_END_LANES
        if (!sh.have_item) break;
        COOP_FOR_LANES(lane)                                  // bound bytes -> shared memory, all lanes
        if (sh.done == 0) {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(sh.row + P.row_bid_off);
            uint32_t *dst = reinterpret_cast<uint32_t *>(bids);
            const uint32_t words = (P.row_tail_off - P.row_bid_off) / 4;
            for (uint32_t j = lane; j < words; j += 32) dst[j] = src[j];
        }
        COOP_END_LANES

        // ---------------------------------------------------------------- waves
        while (sh.done == 0) {
            // P0: pick the bucket and the wave's shape (lane 0)
            COOP_FOR_LANES(lane)
            if (lane == 0) {
                const DevOpt &o = opts[sh.opt_idx];
                if (sh.redo) sh.sticky_exact = 1;           // max_entries is within reach: one chain at a time from here on
                sh.exact_mode = sh.sticky_exact;
                sh.width = sh.exact_mode ? 1u : 32u;
                sh.redo = 0; sh.T = 0; sh.W = 0; sh.ev_lane = COOP_NONE;
                sh.evmask = 0; sh.touched = 0; sh.acc_pushed = 0; sh.acc_lookups = 0; sh.acc_steps = 0; sh.longest = 0; sh.n_full = 0;
                ++sh.waves_item;
                if (!sh.carried) {
                    // loop top of bwtgap.c:144-159 for the first pop of the wave
                    if (sh.n_live == 0 || (sh.exact_mode && (int64_t)sh.n_live + sh.n_phantom > (int64_t)o.max_entries)) sh.done = 1;
                    else {
                        sh.b = (uint32_t)ffs64(sh.mask);
                        if ((int32_t)sh.b > sh.pop_cut) sh.done = 1;
                    }
                }
                if (sh.done == 0 && sh.width > sh.carried && sh.mask && !(sh.carried && (sh.carry.flags & 4u))) {
                    // records to look at: the top of the bucket's array, at most one chunk's worth
                    if (sh.carried) sh.b = (uint32_t)ffs64(sh.mask);
                    const bool usable = !sh.carried || (int32_t)sh.b <= sh.pop_cut;
                    // a carried (resumed) chain came from bucket carry.c_score; new children join only from that same bucket
                    if (usable && (!sh.carried || (int32_t)sh.b == sh.carry.c_score)) {
                        const uint32_t chunk = sh.top_chunk[sh.b];
                        const uint32_t avail = sh.top_cnt[sh.b] + (g.prev[chunk] != COOP_NONE ? COOP_CHUNK : 0u);
                        sh.T = avail < 32u ? avail : 32u;
                    }
                }
                sh.base_entries = sh.n_live + sh.n_phantom - (sh.carried ? 0u : 1u);
            }
            COOP_END_LANES
            if (sh.done) break;
            // P1: load the top T records of the bucket (lane t: the record t below the top)
            COOP_FOR_LANES(lane)
            if (lane < sh.T) {
                const uint32_t idx = coop_entry_index(sh, g.prev, sh.b, lane);
                sh.ent_payload[lane] = g.payload[idx];
                const uint32_t inf = g.info[idx];
                sh.ent_info[lane] = inf;
                sh.cc[lane] = (uint8_t)popc32(inf & 31u);
            }
            COOP_END_LANES
            // P2: children in pop order -> lanes: every record's lane sums the child counts above it and claims its slots
            COOP_FOR_LANES(lane)
            if (lane < sh.T) {
                uint32_t s0 = 0;
                for (uint32_t t = 0; t < lane; ++t) s0 += sh.cc[t];
                sh.S[lane] = s0;
                const uint32_t c = sh.cc[lane], room = sh.width - sh.carried;
                for (uint32_t q = s0; q < s0 + c && q < room; ++q) sh.slot_owner[q] = (uint8_t)lane;
                if (lane + 1 == sh.T) { sh.S[sh.T] = s0 + c; sh.W = (s0 + c < room ? s0 + c : room) + sh.carried; }
            } else if (lane == 0 && sh.T == 0) { sh.S[0] = 0; sh.W = sh.carried; }
            COOP_END_LANES
            // P3: every lane of the wave gets its candidate; P4: and runs its chain
            COOP_FOR_LANES(lane)
            {
                CoopLane &me = COOP_ME(lane);
                me.status = CH_IDLE; me.n_out = 0; me.live = 0; me.ph = 0; me.lookups = 0; me.steps = 0; me.touched = 0;
                if (lane < sh.W) {
                    if (sh.carried && lane == 0) { me.c = sh.carry; }
                    else {
                        const uint32_t q = lane - sh.carried, t = sh.slot_owner[q], r = q - sh.S[t];
                        uint32_t cm = sh.ent_info[t] & 31u, j = 0;
                        for (uint32_t x = 0; x <= r; ++x) { j = 31u - (uint32_t)clz32(cm); cm &= ~(1u << j); }   // r-th highest bit
                        C.child_of(me, sh.ent_payload[t], (sh.ent_info[t] >> 7) & 1u, j, sh.b);
                    }
                    me.status = CH_RUN;
                    C.run_chain(me, lane);
                    sh.lane_status[lane] = (uint8_t)me.status; sh.lane_nout[lane] = (uint16_t)me.n_out;
                    if (me.status != CH_DEAD) coop_or32(&sh.evmask, 1u << lane);
                    coop_max32(&sh.longest, me.steps);
                }
            }
            COOP_END_LANES
            // P5: how much of the wave commits: everything up to and including the first chain that did not just die
            COOP_FOR_LANES(lane)
            if (lane == 0) {
                const uint32_t ev = sh.evmask ? (uint32_t)ffs64(sh.evmask) : COOP_NONE;
                sh.ev_lane = ev; sh.n_valid = ev == COOP_NONE ? sh.W : ev + 1;
                sh.ev_kind = ev == COOP_NONE ? (uint32_t)CH_DEAD : sh.lane_status[ev];
                sh.wave_steps += sh.longest;
            }
            COOP_END_LANES
            COOP_FOR_LANES(lane)
            if (lane < sh.n_valid) {
                const CoopLane &me = COOP_ME(lane);
                if (me.live + me.ph) coop_add32(&sh.acc_pushed, me.live + me.ph);
                if (me.lookups) coop_add32(&sh.acc_lookups, me.lookups);
                if (me.steps) coop_add32(&sh.acc_steps, me.steps);
                if (me.touched) coop_or64(&sh.touched, me.touched);
            }
            COOP_END_LANES
            COOP_FOR_LANES(lane)
            if (lane == 0) {
                const DevOpt &o = opts[sh.opt_idx];
                if (!sh.exact_mode && (int64_t)sh.n_live + sh.n_phantom + (int64_t)sh.acc_pushed + 1 > (int64_t)o.max_entries)
                    sh.redo = 1;                            // the max_entries test might fire inside this wave: one chain at a time
                else { sh.lookups_item += sh.acc_lookups; sh.steps_item += sh.acc_steps; }
            }
            COOP_END_LANES
            if (sh.redo) {
                COOP_FOR_LANES(lane)
                {
                    CoopLane &me = COOP_ME(lane);
                    for (uint64_t m = me.touched; m; m &= m - 1) sh.cnt[lane][ffs64(m)] = 0;
                    me.touched = 0;
                }
                COOP_END_LANES
                continue;
            }
            // P6: per touched bucket (lane j owns the j-th of them): exclusive prefix of the counts over the committing
            // lanes, the total, and the chunks the bucket grows by; in parallel, the records of bucket b the wave used up
            COOP_FOR_LANES(lane)
            {
                for (uint32_t skip = lane; skip < COOP_NB; skip += 32) {    // (a second round only if > 32 buckets were touched)
                    uint64_t m = sh.touched;
                    for (uint32_t j = 0; j < skip && m; ++j) m &= m - 1;
                    if (!m) break;
                    const uint32_t bb_ = (uint32_t)ffs64(m);
                    uint32_t run = 0;
                    // (only the lanes that pushed there get their prefix: a lane cleans up exactly the entries it touched)
                    for (uint32_t l = 0; l < sh.n_valid; ++l) { const uint32_t c = sh.cnt[l][bb_]; if (c) { sh.cnt[l][bb_] = (uint16_t)run; run += c; } }
                    sh.total[bb_] = (uint16_t)run;
                    const bool empty = !((sh.mask >> bb_) & 1ull);
                    const uint32_t cnt = empty ? COOP_CHUNK : sh.top_cnt[bb_];          // an empty bucket starts a fresh chunk
                    const uint32_t need = (cnt + run - 1) / COOP_CHUNK;                // new chunks (ordinal q >= 1)
                    if (need) {
                        const uint32_t base = coop_add32(&sh.n_chunks, need);
                        if (base + need > C.cap_chunks) sh.fail_code = STATUS_NEED_STRICT;
                        else {
                            sh.newbase[bb_] = base;
                            for (uint32_t q = 0; q < need; ++q)
                                g.prev[base + q] = q ? base + q - 1 : (empty ? COOP_NONE : sh.top_chunk[bb_]);
                        }
                    }
                }
                // consumption of bucket b: K children were taken from its top records
                const uint32_t K = sh.n_valid - (sh.n_valid ? sh.carried : 0u);
                if (lane < sh.T && K) {
                    const uint32_t c = sh.cc[lane], s0 = sh.S[lane];
                    if (s0 + c <= K) coop_add32(&sh.n_full, 1u);
                    else if (s0 < K) {                      // partially consumed: drop its (K - s0) highest children
                        uint32_t cm = sh.ent_info[lane] & 31u;
                        for (uint32_t x = 0; x < K - s0; ++x) cm &= ~(1u << (31u - (uint32_t)clz32(cm)));
                        g.info[coop_entry_index(sh, g.prev, sh.b, lane)] = (sh.ent_info[lane] & ~31u) | cm;
                    }
                }
            }
            COOP_END_LANES
            if (sh.fail_code != STATUS_OK) {                  // out of record chunks: the item goes to the large-capacity kernel
                COOP_FOR_LANES(lane)
                {
                    CoopLane &me = COOP_ME(lane);
                    for (uint64_t m = me.touched; m; m &= m - 1) sh.cnt[lane][ffs64(m)] = 0;
                    me.touched = 0;
                    if (lane == 0) sh.done = 1;
                }
                COOP_END_LANES
                break;
            }
            // P7: every committing lane writes its pushes to their places (pre-append geometry of the target buckets)
            COOP_FOR_LANES(lane)
            if (lane < sh.n_valid) {
                const uint32_t n = sh.lane_nout[lane];
                for (uint32_t p = 0; p < n; ++p) {
                    const uint32_t inf = g.out_info[(size_t)lane * COOP_OUT_CAP + p];
                    const uint32_t bb_ = inf & 0xFFu, rank = (inf >> 16) + sh.cnt[lane][bb_];
                    const bool empty = !((sh.mask >> bb_) & 1ull);
                    const uint32_t cnt = empty ? COOP_CHUNK : sh.top_cnt[bb_];
                    const uint32_t pos = cnt + rank, q = pos / COOP_CHUNK, off = pos % COOP_CHUNK;
                    const uint32_t chunk = q == 0 ? sh.top_chunk[bb_] : sh.newbase[bb_] + q - 1;
                    g.payload[(size_t)chunk * COOP_CHUNK + off] = g.out_payload[(size_t)lane * COOP_OUT_CAP + p];
                    g.info[(size_t)chunk * COOP_CHUNK + off] = ((inf >> 8) & 31u) | ((inf >> 15) & 1u) << 7;
                }
            }
            COOP_END_LANES
            // P8: bucket b loses the records that were used up (lane 0); the grown buckets get their new tops (owners)
            COOP_FOR_LANES(lane)
            {
                if (lane == 0) {
                    const uint32_t K = sh.n_valid - (sh.n_valid ? sh.carried : 0u);
                    if (K) {
                        uint32_t full = sh.n_full, cnt = sh.top_cnt[sh.b], chunk = sh.top_chunk[sh.b];
                        while (full) {
                            const uint32_t take = full < cnt ? full : cnt;
                            cnt -= take; full -= take;
                            if (cnt == 0) { chunk = g.prev[chunk]; cnt = chunk == COOP_NONE ? 0u : COOP_CHUNK; if (chunk == COOP_NONE) break; }
                        }
                        if (chunk == COOP_NONE) { coop_and64(&sh.mask, ~(1ull << sh.b)); sh.top_chunk[sh.b] = COOP_NONE; sh.top_cnt[sh.b] = 0; }
                        else { sh.top_chunk[sh.b] = chunk; sh.top_cnt[sh.b] = (uint8_t)cnt; }
                        sh.n_live -= K; sh.pops_item += K;
                    }
                    sh.n_live += sh.acc_pushed;             // corrected for the phantoms just below
                    sh.carried = 0;
                }
                for (uint32_t skip = lane; skip < COOP_NB; skip += 32) {
                    uint64_t m = sh.touched;
                    for (uint32_t j = 0; j < skip && m; ++j) m &= m - 1;
                    if (!m) break;
                    const uint32_t bb_ = (uint32_t)ffs64(m), tot = sh.total[bb_];
                    if (tot) {
                        const bool empty = !((sh.mask >> bb_) & 1ull);
                        const uint32_t cnt = empty ? COOP_CHUNK : sh.top_cnt[bb_];
                        const uint32_t pos = cnt + tot - 1, q = pos / COOP_CHUNK;
                        sh.top_chunk[bb_] = q == 0 ? sh.top_chunk[bb_] : sh.newbase[bb_] + q - 1;
                        sh.top_cnt[bb_] = (uint8_t)(pos % COOP_CHUNK + 1);
                        coop_or64(&sh.mask, 1ull << bb_);
                    }
                }
            }
            COOP_END_LANES
            // P9: phantom bookkeeping, per-lane clean-up of the count table, and the event of the wave
            COOP_FOR_LANES(lane)
            {
                CoopLane &me = COOP_ME(lane);
                if (lane < sh.n_valid && me.ph) { coop_add32(&sh.n_phantom, me.ph); coop_add32(&sh.n_live, 0u - me.ph); }
                for (uint64_t m = me.touched; m; m &= m - 1) sh.cnt[lane][ffs64(m)] = 0;
                me.touched = 0;
                if (lane == sh.ev_lane) {
                    const DevOpt &o = opts[sh.opt_idx];
                    if (sh.ev_kind == CH_SUSP) { sh.carry = me.c; sh.carry.flags |= 8u; sh.carried = 1; }
                    else if (sh.ev_kind == CH_ENTRIES) sh.done = 1;
                    else if (sh.ev_kind == CH_FAIL) { sh.fail_code = STATUS_NEED_STRICT; sh.done = 1; }
                    else if (sh.ev_kind == CH_HIT) {
                        // action for found hits, bwtgap.c:188-241 (Worker::do_hit)
                        const CoopCand &c = me.c;
                        uint32_t k = c.ck, l = c.cl, rk = c.crl - (c.cl - c.ck), rl = c.crl;
                        if (c.flags & 1u) {
                            if (c.zflags & 1u) k = 0;
                            if (c.zflags & 2u) l = 0;
                            if (c.zflags & 4u) rk = 0;
                            if (c.zflags & 8u) rl = 0;
                        }
                        const int32_t score = c.c_score;
                        bool add = true;
                        if (sh.n_hits == 0) {
                            sh.best_score = score;
                            if (!(o.mode & MODE_NONSTOP)) {
                                sh.max_diff = (c.c_nd + 1 > o.max_diff) ? o.max_diff : c.c_nd + 1;
                                sh.pop_cut = sh.best_score + o.s_mm;
                            }
                        }
                        if (score == sh.best_score) sh.best_cnt = (int32_t)((uint32_t)sh.best_cnt + (l - k + 1));
                        else if (sh.best_cnt > o.max_top2) { sh.done = 1; add = false; }
                        if (add && ((c.c_meta >> META_GO_SHIFT) & 15u))
                            for (uint32_t j = 0; j < sh.n_hits; ++j)
                                if (g.hits[j].k == k && g.hits[j].l == l) { add = false; break; }
                        if (add) {
                            const uint32_t x = l - k + 1, ldp = (c.flags & 2u) ? c.ci_at_pop : 0u;
                            uint32_t *w = reinterpret_cast<uint32_t *>(sh.row);
                            uint32_t jj = 0, w_prev = 0xFFFFFFFFu;
                            for (uint32_t i = 0; i < ldp; ++i) {
                                uint32_t v = w[i], bid = bids[i] & BB_BID;
                                if (v > x) { v -= x; w[i] = v; }
                                else if (v == x) { bid = 1; v = P.ix.fwd.text_length - (++jj); w[i] = v; }
                                bids[i] = bound_byte(bid, v, w_prev);
                                w_prev = v;
                            }
                            if (ldp > 0 && ldp <= sh.len) bids[ldp] = bound_byte(bids[ldp] & BB_BID, w[ldp], w_prev);
                            if (sh.n_hits >= COOP_HIT_CAP) { sh.fail_code = STATUS_NEED_STRICT; sh.done = 1; }
                            else {
                                Hit h;
                                h.k = k; h.l = l; h.rev_k = rk; h.rev_l = rl;
                                h.counts = ((c.c_meta >> META_MM_SHIFT) & 31u) | ((c.c_meta >> META_GO_SHIFT) & 15u) << 16 |
                                           ((c.c_meta >> META_GE_SHIFT) & 31u) << 24;
                                h.score = score; h.pad0 = h.pad1 = 0;
                                g.hits[sh.n_hits++] = h;
                            }
                        }
                    }
                }
            }
            COOP_END_LANES
        }

        // ---------------------------------------------------------------- results out (lane 0)
        COOP_FOR_LANES(lane)
        if (lane == 0 && sh.done != 2) {
            sh.steps += sh.steps_item; sh.pops += sh.pops_item; sh.waves += sh.waves_item;
            if (sh.fail_code != STATUS_OK) {
                P.n_aln[sh.out_idx] = 0; P.aln_off[sh.out_idx] = 0; P.status[sh.out_idx] = (uint8_t)sh.fail_code;
#if defined(__CUDA_ARCH__)
                const unsigned long long idx = atomicAdd(P.strict_count, 1ull);
#else
                const unsigned long long idx = (*P.strict_count)++;
#endif
                if (P.strict_list) P.strict_list[idx] = sh.out_idx;
            } else {
                const uint64_t mine = (uint64_t)sh.lookups_item + reinterpret_cast<const uint32_t *>(sh.row + P.row_tail_off)[0];
                if (P.kind == KIND_WHOLE && P.pass == 1 && sh.n_hits == 0) {
                    P.aln_off[sh.out_idx] = mine;           // rides along until pass 2 completes (Worker::do_end)
#if defined(__CUDA_ARCH__)
                    const uint32_t idx = atomicAdd(P.next_count, 1u);
#else
                    const uint32_t idx = (*P.next_count)++;
#endif
                    P.next_list[idx] = sh.out_idx;
                } else {
                    sh.lookups += mine + ((P.kind == KIND_WHOLE && P.pass == 2) ? P.aln_off[sh.out_idx] : 0ull);
                    uint64_t off = 0;
                    uint32_t n = sh.n_hits;
                    uint8_t stt = STATUS_OK;
                    if (n) {
#if defined(__CUDA_ARCH__)
                        off = atomicAdd(&P.counters[CNT_ALN], (unsigned long long)n);
#else
                        off = P.counters[CNT_ALN]; P.counters[CNT_ALN] += n;
#endif
                        if (off + n > P.aln_cap) { stt = STATUS_OUT_FULL; n = 0; }
                        for (uint32_t j = 0; j < n; ++j) {
                            uint32_t *w = P.aln + (off + j) * 9;
                            const Hit h = g.hits[j];
                            w[0] = h.counts; w[1] = h.k; w[2] = h.l; w[3] = h.rev_k; w[4] = h.rev_l;
                            w[5] = sh.strand << 30;
                            int32_t s0 = 0, e0 = 0;
                            if (P.kind == KIND_SEEDS) { s0 = (int32_t)sh.sub_off; e0 = (int32_t)(sh.sub_off + sh.len - 1); }
                            else if (P.kind == KIND_WHOLE && j == 0) { s0 = 0; e0 = (int32_t)sh.rd_len - 1; }
                            w[6] = (uint32_t)s0; w[7] = (uint32_t)e0; w[8] = (uint32_t)h.score;
                        }
                    }
                    P.n_aln[sh.out_idx] = (int32_t)n; P.aln_off[sh.out_idx] = off; P.status[sh.out_idx] = stt;
                    if (P.kind == KIND_TASKS && P.width_out) {            // per-call form: see Worker::do_end
                        const uint32_t *w = reinterpret_cast<const uint32_t *>(sh.row);
                        for (uint32_t i = 0; i <= sh.len; ++i) { u32x2 v; v.x = w[i]; v.y = bids[i] & BB_BID; P.width_out[i] = v; }
                    }
                }
            }
        }
        COOP_END_LANES
    }
}

} // namespace hsa

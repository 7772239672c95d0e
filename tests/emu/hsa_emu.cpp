// hsa_emu.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles hsa_b200/csrc/hsa_core.cuh (the exact source the CUDA kernels are built from) with g++ and
// runs the worker state machine serially on the host, so the CPU test-suite (`-m "not gpu"`) can check
// the device algorithm against the oracle in a container without a GPU.  It is NOT a fallback: nothing
// in hsa_b200/ loads this library, and the product's C ABI fails loudly without CUDA.
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include "../../hsa_b200/csrc/hsa_core.cuh"
#include "../../hsa_b200/csrc/hsa_coop.cuh"
#include "../../hsa_b200/csrc/hsa_splice.cuh"
#include "../../hsa_b200/csrc/hsa_sam.cuh"
#include "../../include/hsa_b200.h"

using namespace hsa;

struct EmuIndex {
    std::vector<u32x4> blocks[2];
    DevIndex ix;
};

static void repack(const hsa_bwt_view_t *v, std::vector<u32x4> &out, DevBwt &d)
{
    RefBwt r;
    r.bwt_code = v->bwtCode; r.occ_value = v->occValue; r.occ_major = v->occValueMajor;
    r.text_length = v->textLength; r.inverse_sa0 = v->inverseSa0;
    uint32_t nb = v->textLength / 64 + 1;
    out.resize((size_t)nb * 2);
    for (uint32_t b = 0; b < nb; ++b) {
        uint32_t occ[4];
        occ4_ref_raw(r, b * 64u, occ);
        u32x4 c, w;
        c.x = occ[0]; c.y = occ[1]; c.z = occ[2]; c.w = occ[3];
        w.x = v->bwtCode[4 * (size_t)b]; w.y = v->bwtCode[4 * (size_t)b + 1];
        w.z = v->bwtCode[4 * (size_t)b + 2]; w.w = v->bwtCode[4 * (size_t)b + 3];
        out[2 * (size_t)b] = c; out[2 * (size_t)b + 1] = planes_of(w);
    }
    d.blocks = out.data(); d.n_blocks = nb; d.text_length = v->textLength; d.inverse_sa0 = v->inverseSa0;
    for (int i = 0; i < 5; ++i) d.cum[i] = v->cumulativeFreq[i];
}

static uint64_t g_last_steps = 0;
static uint32_t *g_item_steps = nullptr;      // optional diagnostic: worker steps per work item, launches concatenated
static size_t g_item_steps_pos = 0;
// SIMT simulation (diagnostic): 32 workers stepped in lockstep under phase_vote(); counts how often each phase
// runs and how many lanes take part, so scheduling policies can be compared on the CPU.
static int g_simt = 0;
static uint32_t g_rerun_cap = 0;
static int g_use_coop = 0;                       // re-runs go through the warp-cooperative kernel first
static uint32_t g_step_budget = 0;               // fast configuration: hand searches on after this many steps                 // 0 = report flagged items instead of re-running them
static uint64_t g_flagged_first = 0;
static const uint8_t *g_rows_host = nullptr; static size_t g_rows_host_bytes = 0;   // per-call form: the caller's widths as item 0's row
static uint32_t g_vote_slow_min = VOTE_SLOW_MIN_DEFAULT; static int32_t g_vote_pop_bias = VOTE_POP_BIAS_DEFAULT;
static uint64_t g_phase_runs[3], g_phase_lanes[3];

struct Smem {                                   // stands in for one block's shared memory
    std::vector<unsigned char> buf;
};

template <typename LinkT, bool BIDS_SMEM>
static void run_worker(const Params &P, uint32_t n_work, uint64_t st[4])
{
    if (!g_simt) {
        Worker<LinkT, BIDS_SMEM> w(P, 0, 0);
        uint32_t next = 0;
        uint64_t steps0 = 0;
        for (;;) {
            switch (w.st) {
            case LS_IDLE:
                if (g_item_steps && next > 0) { g_item_steps[g_item_steps_pos++] = (uint32_t)(w.steps - steps0); steps0 = w.steps; }
                if (next < n_work) w.start(next++); else w.retire();
                break;
            case LS_POP: w.do_pop(); break;
            case LS_LOOKUP: w.do_lookup(); break;
            case LS_HIT: w.do_hit(); break;
            case LS_END: w.do_end(); break;
            default: break;
            }
            if (w.retired()) break;
        }
        st[0] += w.lookups; st[1] += w.pops; st[3] += w.steps;
        return;
    }
    // lockstep warp: needs per-lane scratch (arena / links / hits are indexed by slot, shared memory by lane)
    const int W = 32;
    std::vector<Worker<LinkT, BIDS_SMEM>> ws;
    for (int l = 0; l < W; ++l) ws.emplace_back(P, (uint32_t)l, (uint32_t)l);
    uint32_t next = 0;
    for (;;) {
        uint32_t n[3] = {0, 0, 0};
        int alive = 0;
        for (auto &w : ws) if (!w.retired()) { ++n[w.cls()]; ++alive; }
        if (!alive) break;
        const uint32_t ph = phase_vote(P, n[PHASE_LOOKUP], n[PHASE_POP], n[PHASE_SLOW]);
        ++g_phase_runs[ph]; g_phase_lanes[ph] += n[ph];
        for (auto &w : ws) {
            if (w.retired() || w.cls() != ph) continue;
            if (ph == PHASE_LOOKUP) w.do_lookup();
            else if (ph == PHASE_POP) w.do_pop();
            else {
                if (w.st == LS_HIT) w.do_hit();
                if (w.st == LS_END) w.do_end();
                if (w.st == LS_IDLE) { if (next < n_work) w.start(next++); else w.retire(); }
            }
        }
    }
    for (auto &w : ws) { st[0] += w.lookups; st[1] += w.pops; st[3] += w.steps; }
}

// the warp-cooperative kernel (hsa_coop.cuh) for one "warp": the same source, lanes looped per phase
static uint64_t g_coop_waves = 0, g_coop_wave_steps = 0, g_coop_steps = 0;
static void run_coop(Params P, uint32_t n_work, uint64_t st[4])
{
    const uint32_t cap_chunks = 4096;
    const size_t bid_bytes = P.row_tail_off - P.row_bid_off;
    std::vector<u32x4> wsh((sizeof(CoopWarp) + bid_bytes + 64) / 16 + 1);
    CoopWarp *sh = reinterpret_cast<CoopWarp *>(wsh.data());
    uint8_t *bids = reinterpret_cast<uint8_t *>(sh) + ((sizeof(CoopWarp) + 15) & ~size_t(15));
    std::vector<u32x4> payload((size_t)cap_chunks * COOP_CHUNK), outp((size_t)32 * COOP_OUT_CAP);
    std::vector<uint32_t> info((size_t)cap_chunks * COOP_CHUNK), prev(cap_chunks), outi((size_t)32 * COOP_OUT_CAP);
    std::vector<Hit> hits(COOP_HIT_CAP);
    CoopScratch g{payload.data(), info.data(), prev.data(), outp.data(), outi.data(), hits.data()};
    static CoopLane me[32];
    coop_run(P, *sh, bids, g, cap_chunks, n_work, me);
    st[0] += sh->lookups; st[1] += sh->pops; st[3] += sh->steps; g_coop_waves += sh->waves; g_coop_wave_steps += sh->wave_steps; g_coop_steps += sh->steps;
}

extern "C" {

void *emu_index_new(const hsa_bwt_view_t *fwd, const hsa_bwt_view_t *rev)
{
    EmuIndex *e = new EmuIndex();
    repack(fwd, e->blocks[0], e->ix.fwd);
    repack(rev, e->blocks[1], e->ix.rev);
    return e;
}
void emu_index_free(void *p) { delete (EmuIndex *)p; }

void emu_occ(void *p, int which, int layout, const hsa_bwt_view_t *refview, const uint32_t *idx, size_t n, uint32_t *out4)
{
    EmuIndex *e = (EmuIndex *)p;
    for (size_t i = 0; i < n; ++i) {
        if (layout == 1) occ4_dev(which == 0 ? e->ix.fwd : e->ix.rev, idx[i], out4 + 4 * i);
        else {
            RefBwt r;
            r.bwt_code = refview->bwtCode; r.occ_value = refview->occValue; r.occ_major = refview->occValueMajor;
            r.text_length = refview->textLength; r.inverse_sa0 = refview->inverseSa0;
            occ4_ref(r, idx[i], out4 + 4 * i);
        }
    }
}

// BWTSaValue through the device code (hsa_core.cuh: sa_value_dev / psi_minus_dev) on the re-packed layout
void emu_sa_values(void *p, const uint32_t *sa_value, uint32_t sa_interval, const uint32_t *idx, size_t n, uint32_t *out,
                   uint32_t *steps_out)
{
    EmuIndex *e = (EmuIndex *)p;
    for (size_t i = 0; i < n; ++i) {
        uint32_t st = 0;
        out[i] = sa_value_dev(e->ix.fwd, sa_value, sa_interval, idx[i], st);
        steps_out[i] = st;
    }
}

// the block search of BWTRetrievePositionFromSAIndex through the device code (hsa_core.cuh: locate_dev)
void emu_locate(const uint32_t *blocks4, uint32_t n_blocks, const uint32_t *pos, size_t n, uint32_t *seq_id, uint32_t *ori_pos)
{
    for (size_t i = 0; i < n; ++i) {
        uint32_t a = 0xFFFFFFFFu, b = 0xFFFFFFFFu;
        locate_dev(blocks4, n_blocks, pos[i], a, b);
        seq_id[i] = a; ori_pos[i] = b;
    }
}

static void to_devopt(const hsa_gap_opt_t &o, DevOpt &d)
{
    memset(&d, 0, sizeof(d));
    d.s_mm = o.s_mm; d.s_gapo = o.s_gapo; d.s_gape = o.s_gape; d.mode = o.mode;
    d.indel_end_skip = o.indel_end_skip; d.max_del_occ = o.max_del_occ; d.max_entries = o.max_entries;
    d.max_diff = o.max_diff; d.max_gapo = o.max_gapo; d.max_gape = o.max_gape;
    d.max_seed_diff = o.max_seed_diff; d.seed_len = o.seed_len; d.max_top2 = o.max_top2;
}

// Run one batch through the worker.  kind: 0 tasks, 1 whole, 2 seeds, 3 width.
// opts: already-resolved per-call options (tasks: indexed by task.opt_idx; whole: indexed through len2opt;
// seeds: opts[0]).  Outputs: n_aln[n_items], aln_off[n_items], status[n_items], aln[aln_cap*9].
// Returns the number of hits written; *lookups gets the reference-equivalent occ lookup count;
// *n_strict the number of groups that ran out of capacity (arena_cap / hit_cap).
long emu_run(void *p, uint32_t kind, const uint8_t *codes, const hsa_task_t *tasks, const uint64_t *read_off,
             const uint32_t *read_len, uint32_t n_groups, const hsa_gap_opt_t *opts, uint32_t n_opts,
             const uint16_t *len2opt, uint32_t max_len, int32_t filter_max_n, uint32_t arena_cap,
             uint32_t hit_cap, int32_t *n_aln, uint64_t *aln_off, uint8_t *status, uint32_t *aln, uint64_t aln_cap,
             uint32_t *width_out, int32_t *bid_out, uint64_t *lookups, uint64_t *n_strict, uint64_t *pops)
{
    EmuIndex *e = (EmuIndex *)p;
    std::vector<DevOpt> dopts(n_opts ? n_opts : 1);
    uint32_t nb = 1;
    for (uint32_t i = 0; i < n_opts; ++i) {
        to_devopt(opts[i], dopts[i]);
        uint32_t b = (uint32_t)((opts[i].max_diff + 1) * opts[i].s_mm + (opts[i].max_gapo + 1) * opts[i].s_gapo +
                                (opts[i].max_gape + 1) * opts[i].s_gape + 1);
        if (b > nb) nb = b;
    }
    if (nb > 128) return -1;
    // arena_cap <= 1022: the fast configuration (16-bit link halves, 64 buckets, bound bytes in "shared memory");
    // larger: the large-capacity configuration (32-bit halves, bound bytes read from the rows).  As in
    // hsa_b200.cu's run_batch, items the fast configuration flags are re-run with the large one when
    // `rerun_cap` is set (emu_set_rerun).
    const int lanes = g_simt ? 32 : 1;
    const uint32_t per = kind == KIND_SEEDS ? 6u : 1u, n_work_all = n_groups * per;
    uint32_t max_seed = 0;
    for (uint32_t i = 0; i < n_opts; ++i)
        if (opts[i].seed_len > 0 && (uint32_t)opts[i].seed_len < max_len) max_seed = std::max(max_seed, (uint32_t)opts[i].seed_len);
    const uint32_t seed_cap = (kind == KIND_TASKS || kind == KIND_WHOLE) && max_seed ? max_seed + 1 : 0;
    std::vector<uint32_t> strict(n_work_all + 1), strict_in;
    unsigned long long counters[CNT_TOTAL];
    memset(counters, 0, sizeof(counters));
    uint64_t st[4] = {0, 0, 0, 0};
    uint64_t flagged_first = 0;

    // stages as hsa_b200.cu runs them: the given configuration; then (if re-runs are on) the warp-cooperative kernel
    // for what it flagged; then the large-capacity configuration for what that flagged
    for (int round = 0; round < 3; ++round) {
        const bool coop = round == 1;
        if (coop && !g_use_coop) continue;
        const uint32_t cap = round == 0 ? arena_cap : g_rerun_cap;
        const uint32_t hcap = round == 0 ? hit_cap : 4096u;
        const bool wide = cap > 1022;
        if (round > 0 && strict_in.empty()) break;
        const uint32_t n_work = round == 0 ? n_work_all : (uint32_t)strict_in.size();
        Params P;
        memset(&P, 0, sizeof(P));
        set_layout(P, max_len, seed_cap, wide ? nb : std::min(nb, 64u), n_opts ? n_opts : 1, wide ? 4 : 2, !wide && !coop);
        std::vector<unsigned char> smem((size_t)P.smem_opts_bytes + (size_t)lanes * P.smem_lane_stride + 16);
        memcpy(smem.data(), dopts.data(), dopts.size() * sizeof(DevOpt));
        hsa_smem_host = smem.data();
        std::vector<u32x4> arena((size_t)cap * lanes * 2);            // 32-byte slots: record + link
        std::vector<uint64_t> links((size_t)cap * lanes);          // large enough for either link width
        std::vector<Hit> hits((size_t)hcap * lanes);
        std::vector<u32x4> rows(((size_t)std::max(n_work, 1u) * P.row_stride + 15) / 16);
        std::vector<uint32_t> next_list(n_groups + 1);
        uint32_t next_count = 0;

        P.ix = e->ix; P.codes = codes; P.kind = kind;
        P.tasks = (const Task *)tasks; P.read_off = read_off; P.read_len = read_len;
        P.opts = dopts.data(); P.len2opt = len2opt; P.filter_max_n = filter_max_n;
        P.rows = reinterpret_cast<uint8_t *>(rows.data());
        P.arena = arena.data(); P.arena_cap = cap;
        P.hits = hits.data(); P.hit_cap = hcap;
        P.n_aln = n_aln; P.aln_off = aln_off; P.status = status; P.aln = aln; P.aln_cap = aln_cap;
        P.counters = counters; P.strict_list = strict.data(); P.strict_count = &counters[CNT_STRICT];
        P.step_budget = round == 0 ? g_step_budget : 0;
        P.width_out = (u32x2 *)width_out; P.bid_out = bid_out;
        P.vote_slow_min = g_vote_slow_min; P.vote_pop_bias = g_vote_pop_bias;
        unsigned long long cursor = 0;
        P.cursor = &cursor;
        // the split pipeline, launch for launch as hsa_b200.cu's issue_chunk enqueues it
        P.pass = 1; P.work_base = 0; P.work_list = round == 0 ? nullptr : strict_in.data(); P.n_work = n_work;
        P.next_list = next_list.data(); P.next_count = &next_count;
        for (uint32_t w = 0; w < n_work; ++w) width_item(P, dopts.data(), w);
        if (g_rows_host && n_work == 1) memcpy(P.rows, g_rows_host, std::min<size_t>(g_rows_host_bytes, P.row_stride));   // hsa_match_gap_call: Batch::rows_host
        if (kind != KIND_WIDTH) {
            if (coop) run_coop(P, n_work, st);
            else if (wide) run_worker<uint64_t, false>(P, n_work, st); else run_worker<uint32_t, true>(P, n_work, st);
            if (kind == KIND_WHOLE) {
                P.pass = 2; P.work_list = next_list.data(); P.n_work = next_count; P.next_list = nullptr; P.next_count = nullptr;
                cursor = 0;
                for (uint32_t w = 0; w < next_count; ++w) width_item(P, dopts.data(), w);
                if (coop) run_coop(P, next_count, st);
                else if (wide) run_worker<uint64_t, false>(P, next_count, st); else run_worker<uint32_t, true>(P, next_count, st);
            }
        }
        if (round == 0) flagged_first = counters[CNT_STRICT];
        if (!g_rerun_cap || counters[CNT_BAD]) break;
        strict_in.assign(strict.begin(), strict.begin() + counters[CNT_STRICT]);
        if (round < 2) counters[CNT_STRICT] = 0;
    }
    g_flagged_first = flagged_first;
    hsa_smem_host = nullptr;
    *lookups = st[0] + counters[CNT_LOOKUPS];
    *n_strict = counters[CNT_STRICT] + counters[CNT_BAD];
    *pops = st[1];
    g_last_steps = st[3];
    return (long)counters[CNT_ALN];
}

void emu_set_rerun(uint32_t cap) { g_rerun_cap = cap; }
void emu_set_rows_host(const uint8_t *row, size_t bytes) { g_rows_host = row; g_rows_host_bytes = bytes; }
void emu_set_coop(int on, uint32_t step_budget) { g_use_coop = on; g_step_budget = step_budget; }
uint64_t emu_coop_waves(void) { return g_coop_waves; }
void emu_coop_stats(uint64_t *o) { o[0] = g_coop_waves; o[1] = g_coop_wave_steps; o[2] = g_coop_steps; g_coop_waves = g_coop_wave_steps = g_coop_steps = 0; }
uint64_t emu_flagged_first(void) { return g_flagged_first; }
void emu_set_item_steps(uint32_t *buf) { g_item_steps = buf; g_item_steps_pos = 0; }
void emu_set_vote(uint32_t slow_min, int32_t pop_bias) { g_vote_slow_min = slow_min; g_vote_pop_bias = pop_bias; }
void emu_pair_stats(uint64_t *o) { o[0] = hsa_host_pair_count; o[1] = hsa_host_pair_same_sector; hsa_host_pair_count = hsa_host_pair_same_sector = 0; }
// bwt_splice_match for every read, one after the other, on the device code of hsa_splice.cuh.
// opts: n_opts gap_opt_t as the driver holds them in aux->opt; opt_idx per read (may be null: all reads use opts[0]).
// n_aln_out[n], aln_out[n * 18] (two hsa_aln1_t per read), status_out[n]; returns the occ lookups issued.
uint64_t emu_splice(void *p, const uint32_t *sa_value, uint32_t sa_interval, const uint32_t *blocks4, uint32_t n_blocks,
                    const uint32_t *packed_dna, uint32_t dna_length, const uint8_t *codes, const uint64_t *off, const uint32_t *len,
                    size_t n, const hsa_gap_opt_t *opts, size_t n_opts, const uint32_t *opt_idx, uint32_t arena_cap, uint32_t aln_cap,
                    int32_t *n_aln_out, uint32_t *aln_out, uint8_t *status_out)
{
    EmuIndex *e = (EmuIndex *)p;
    std::vector<DevOpt> dopts(n_opts);
    for (size_t i = 0; i < n_opts; ++i) {
        const hsa_gap_opt_t &o = opts[i]; DevOpt &d = dopts[i];
        memset(&d, 0, sizeof(d));
        d.s_mm = o.s_mm; d.s_gapo = o.s_gapo; d.s_gape = o.s_gape; d.mode = o.mode;
        d.indel_end_skip = o.indel_end_skip; d.max_del_occ = o.max_del_occ; d.max_entries = o.max_entries;
        d.max_diff = o.max_diff; d.max_gapo = o.max_gapo; d.max_gape = o.max_gape;
        d.max_seed_diff = o.max_seed_diff; d.seed_len = o.seed_len; d.max_top2 = o.max_top2;
    }
    uint32_t max_len = 0;
    for (size_t i = 0; i < n; ++i) max_len = std::max(max_len, len[i]);
    SpliceParams P;
    memset(&P, 0, sizeof(P));
    P.env.ix = e->ix; P.env.sa_value = sa_value; P.env.sa_interval = sa_interval; P.env.blocks4 = blocks4; P.env.n_blocks = n_blocks;
    P.env.packed_dna = packed_dna; P.env.dna_length = dna_length;
    P.codes = codes; P.read_off = off; P.read_len = len; P.opts = dopts.data(); P.opt_idx = opt_idx;
    P.n_work = (uint32_t)n; P.max_len = max_len;
    std::vector<SEntry> arena(arena_cap);
    std::vector<uint32_t> heads(SPL_BUCKETS), sites(4096);
    std::vector<SWidth> widths(3 * ((size_t)max_len + 1) + 16);
    std::vector<SAln> lists((size_t)6 * aln_cap);
    std::vector<SPos> pos(SPL_POS_CAP);
    P.arena = arena.data(); P.arena_cap = arena_cap; P.heads = heads.data(); P.widths = widths.data();
    P.lists = lists.data(); P.aln_cap = aln_cap; P.site_pos = sites.data(); P.site_cap = (uint32_t)sites.size(); P.pos_info = pos.data();
    P.n_aln = n_aln_out; P.aln = aln_out; P.status = status_out;
    std::vector<uint32_t> fail_list(n + 1);
    unsigned long long fail_count = 0, lookups = 0;
    P.fail_list = fail_list.data(); P.fail_count = &fail_count; P.lookups = &lookups;
    for (size_t i = 0; i < n; ++i) splice_item(P, (uint32_t)i, 0);
    return lookups;
}

void emu_pop_out(uint64_t *o) { memcpy(o, hsa_host_pop_out, sizeof(hsa_host_pop_out)); memset(hsa_host_pop_out, 0, sizeof(hsa_host_pop_out)); }
void emu_hist(uint64_t *o) { memcpy(o, hsa_host_hist, sizeof(hsa_host_hist)); memset(hsa_host_hist, 0, sizeof(hsa_host_hist)); }
void emu_set_simt(int on) { g_simt = on; for (int i = 0; i < 3; ++i) { g_phase_runs[i] = 0; g_phase_lanes[i] = 0; } }
void emu_phase_stats(uint64_t *runs, uint64_t *lanes) { for (int i = 0; i < 3; ++i) { runs[i] = g_phase_runs[i]; lanes[i] = g_phase_lanes[i]; } }
uint64_t emu_last_extra(void) { return 0; }
uint64_t emu_last_steps(void) { return g_last_steps; }

// dp_global (hsa_sam.cuh) on one (reference window, read) pair: score and the CIGAR in path order (op << 28 | len), as
// aln_global_core + bwa_aln_path2cigar return them; returns n_cigar, or -1 if the band does not fit W
int emu_dp(const uint8_t *ref, int32_t len1, const uint8_t *read, int32_t len2, uint32_t W, int32_t *score_out, uint32_t *cigar_out)
{
    std::vector<uint8_t> bytes((size_t)(len2 + 1) * W + len1 + 2);
    std::vector<int32_t> rows(std::max<size_t>(3 * ((size_t)len1 + 1), (size_t)len1 + len2 + 2) + 8);
    DpScratch S;
    S.T = S.RT = 1; S.W = W; S.len1_cap = (uint32_t)len1; S.len2_cap = (uint32_t)len2; S.run_cap = (uint32_t)rows.size();
    S.cells = bytes.data(); S.ref = bytes.data() + (size_t)(len2 + 1) * W; S.rows = rows.data();
    for (int32_t i = 0; i < len1; ++i) S.refb(i + 1) = ref[i];
    const SamRead q{read, (uint32_t)len2, 0u, 0u};
    int32_t n_runs = 0;
    *score_out = dp_global(S, len1, len2, q, n_runs);
    if (n_runs < 0) return -1;
    for (int32_t k = 0; k < n_runs; ++k) cigar_out[k] = (uint32_t)S.run(n_runs - 1 - k);
    return n_runs;
}

static int g_sel_sequential = 0; static uint32_t g_sel_dependent = 0;
void emu_sam_set_sequential(int on) { g_sel_sequential = on; }
uint32_t emu_sam_dependent_reads(void) { return g_sel_dependent; }
// The SAM-field stage (hsa_sam.cuh) for every read, one after the other: the host selection, sam_pos_item, sam_dp_item.
// counts_out: {multi slots, cigar words, md bytes, reads refined, status}; returns 0, or -1 when an output array is too small.
int emu_sam(void *p, const uint32_t *sa_value, uint32_t sa_interval, const uint32_t *blocks4, uint32_t n_blocks,
            const uint32_t *packed_dna, uint32_t dna_length, const uint8_t *codes, const uint64_t *off, const uint32_t *len, size_t n,
            const int32_t *n_aln, const uint64_t *aln_off, const uint32_t *aln, const int32_t *maxdiff_by_len, int32_t max_mm, int n_occ, uint64_t *rng_state,
            SamRec *rec_out, SamMulti *multi_out, size_t multi_cap, uint32_t *cigar_out, size_t cigar_cap, char *md_out, size_t md_cap,
            uint64_t *counts_out)
{
    EmuIndex *e = (EmuIndex *)p;
    size_t n_multi = 0; uint32_t max_len = 0, max_ext = 0;
    memset(rec_out, 0, n * sizeof(SamRec));
    {
        // the selection, cut into the pieces the kernels run (hsa_sam.cuh): classify, prefix sums, descriptors of the reads
        // with several best hits, the chain over them, every read's own jump -- or, with g_sel_sequential, the sequential statement
        std::vector<uint32_t> fixed(n + 1), dep(n + 1), slots(n + 1), vcum(n + 1);
        std::vector<SelDesc> desc(n + 1); std::vector<SelPick> pick(n + 1);
        uint64_t x_end = 0; uint32_t rare = 0;
        SelParams S;
        memset(&S, 0, sizeof(S));
        S.n_aln = n_aln; S.aln_off = aln_off; S.aln = aln; S.n_reads = (uint32_t)n; S.n_occ = n_occ;
        S.fixed = fixed.data(); S.dep = dep.data(); S.slots = slots.data(); S.desc = desc.data(); S.pick = pick.data(); S.vcum = vcum.data();
        S.rec = rec_out; S.multi = multi_out; S.x0 = *rng_state; S.x_end = &x_end; S.rare = &rare;
        for (size_t i = 0; i < n; ++i) sel_classify_item(S, (uint32_t)i);
        auto scan = [&](std::vector<uint32_t> &v) { uint32_t acc = 0; for (size_t i = 0; i <= n; ++i) { const uint32_t t = i < n ? v[i] : 0; v[i] = acc; acc += t; } };
        scan(fixed); scan(dep); scan(slots);
        n_multi = slots[n];
        if (n_multi + 16 > multi_cap) return -1;
        if (!g_sel_sequential) {
            for (size_t i = 0; i < n; ++i) sel_desc_item(S, (uint32_t)i);
            sel_chain(S);
            for (size_t i = 0; i < n; ++i) sel_finish_item(S, (uint32_t)i);
        }
        if (g_sel_sequential || rare) sel_sequential(S);
        g_sel_dependent = dep[n];
        *rng_state = x_end;
    }
    for (size_t i = 0; i < n; ++i) {
        max_len = std::max(max_len, len[i]); max_ext = std::max(max_ext, rec_out[i].n_gapo + rec_out[i].n_gape);
        for (uint32_t j = 0; j < rec_out[i].n_multi; ++j) max_ext = std::max(max_ext, multi_out[rec_out[i].multi_off + j].gap);
    }
    SamParams P;
    memset(&P, 0, sizeof(P));
    P.env.ix = e->ix; P.env.sa_value = sa_value; P.env.sa_interval = sa_interval; P.env.blocks4 = blocks4; P.env.n_blocks = n_blocks;
    P.env.packed_dna = packed_dna; P.env.dna_length = dna_length;
    P.codes = codes; P.read_off = off; P.read_len = len; P.n_reads = (uint32_t)n;
    P.n_aln = n_aln; P.aln_off = aln_off; P.aln = aln; P.rec = rec_out; P.multi = multi_out;
    P.maxdiff_by_len = maxdiff_by_len; P.max_mm = max_mm; P.max_len = max_len;
    unsigned long long cnt[6] = {0, 0, 0, 0, 0, 0}; uint32_t status = 0;
    std::vector<uint32_t> list(n + 1);
    P.cigar = cigar_out; P.cigar_cap = cigar_cap; P.cigar_used = cnt; P.md = md_out; P.md_cap = md_cap; P.md_used = cnt + 1;
    P.dp_list = list.data(); P.dp_count = cnt + 2; P.cursor = cnt + 3; P.status = &status; P.dp_tasks = cnt + 5;
    // (a reference window clipped at the end of the text needs full-width trace-back rows: second pass, as the library does)
    std::vector<SamRec> rec0(rec_out, rec_out + n); std::vector<SamMulti> multi0(multi_out, multi_out + n_multi);
    for (int wide = 0; wide < 2; ++wide) {
        if (wide) { memcpy(rec_out, rec0.data(), n * sizeof(SamRec)); memcpy(multi_out, multi0.data(), n_multi * sizeof(SamMulti)); memset(cnt, 0, sizeof(cnt)); status = 0; }
        for (size_t i = 0; i < n; ++i) sam_pos_item(P, (uint32_t)i);
        const uint32_t len1_cap = max_len + max_ext, len2_cap = max_len, W = wide ? len1_cap + 1u : std::min<uint32_t>(2u * DP_BAND + max_ext + 1u, len1_cap + 1u);
        std::vector<uint8_t> bytes((size_t)(len2_cap + 1u) * W + len1_cap + 1u);
        std::vector<int32_t> rows(std::max<size_t>(3 * ((size_t)len1_cap + 1), (size_t)len1_cap + len2_cap + 2));
        P.dp_bytes = bytes.data(); P.dp_rows = rows.data(); P.dp_workers = 1; P.dp_w = W; P.dp_len1_cap = len1_cap; P.dp_len2_cap = len2_cap;
        { const DpScratch S0 = dp_scratch_of(P, 0); for (unsigned long long w = 0; w < cnt[2]; ++w) sam_dp_item(P, list[w], S0); }
        if (status != SAM_SCRATCH) break;
    }
    counts_out[0] = n_multi; counts_out[1] = cnt[0]; counts_out[2] = cnt[1]; counts_out[3] = cnt[2]; counts_out[4] = status;
    return (cnt[0] > cigar_cap || cnt[1] > md_cap) ? -1 : 0;
}

} // extern "C"

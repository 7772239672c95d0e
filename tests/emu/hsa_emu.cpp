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
        out[2 * (size_t)b] = c; out[2 * (size_t)b + 1] = w;
    }
    d.blocks = out.data(); d.n_blocks = nb; d.text_length = v->textLength; d.inverse_sa0 = v->inverseSa0;
    for (int i = 0; i < 5; ++i) d.cum[i] = v->cumulativeFreq[i];
}

static uint64_t g_last_extra = 0, g_last_steps = 0;
static uint32_t *g_item_steps = nullptr;      // optional diagnostic: worker iterations per work item, launches concatenated
static size_t g_item_steps_pos = 0;

template <typename LinkT, bool FUSED>
static void run_worker(const Params &P, LinkT *heads, const DevOpt *dopts, uint32_t n_work, uint64_t st[4])
{
    Worker<LinkT, FUSED> w(P, 0, heads, 1, dopts);
    uint32_t next = 0;
    uint64_t steps0 = 0;
    for (;;) {
        if (w.idle()) {
            if (g_item_steps && next > 0) { g_item_steps[g_item_steps_pos++] = (uint32_t)(w.steps - steps0); steps0 = w.steps; }
            if (next < n_work) w.start_group(next++);
            else break;
            continue;
        }
        w.template iterate<2>();
    }
    st[0] += w.lookups; st[1] += w.pops; st[2] += w.extra; st[3] += w.steps;
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
    std::vector<u32x4> arena(arena_cap);
    std::vector<uint32_t> links(arena_cap);          // large enough for either link width
    std::vector<u32x2> width(2 * (size_t)(max_len + 1));
    std::vector<Hit> hits(hit_cap);
    std::vector<uint32_t> heads(nb);
    std::vector<uint32_t> strict(n_groups + 1);
    unsigned long long counters[CNT_N];
    memset(counters, 0, sizeof(counters));

    Params P;
    memset(&P, 0, sizeof(P));
    P.ix = e->ix; P.codes = codes; P.kind = kind; P.n_groups = n_groups; P.group_list = nullptr;
    P.tasks = (const Task *)tasks; P.read_off = read_off; P.read_len = read_len;
    P.opts = dopts.data(); P.n_opts = n_opts; P.len2opt = len2opt; P.max_len = max_len; P.filter_max_n = filter_max_n;
    P.arena = arena.data(); P.links = links.data(); P.arena_cap = arena_cap;
    P.width = width.data(); P.width_stride = 2 * (max_len + 1);
    P.hits = hits.data(); P.hit_cap = hit_cap; P.n_buckets = nb;
    P.n_aln = n_aln; P.aln_off = aln_off; P.status = status; P.aln = aln; P.aln_cap = aln_cap;
    P.counters = counters; P.strict_list = strict.data();
    P.width_out = (u32x2 *)width_out; P.bid_out = bid_out;

    uint64_t st[4] = {0, 0, 0, 0};
    unsigned long long cursor = 0;
    P.cursor = &cursor;
    if (arena_cap > 4094) {
        // the large-capacity configuration: fused flow (width passes + both strands inside the worker), 32-bit links
        run_worker<uint32_t, true>(P, heads.data(), dopts.data(), n_groups, st);
    } else {
        // the split pipeline, launch for launch as hsa_b200.cu's run_batch enqueues it
        const uint32_t per = kind == KIND_SEEDS ? 6u : 1u, n_work = n_groups * per;
        uint32_t max_seed = 0;
        for (uint32_t i = 0; i < n_opts; ++i)
            if (opts[i].seed_len > 0 && (uint32_t)opts[i].seed_len < max_len) max_seed = std::max(max_seed, (uint32_t)opts[i].seed_len);
        const uint32_t seed_cap = (kind == KIND_TASKS || kind == KIND_WHOLE) ? max_seed + 1 : 0;
        const uint32_t wstride = (max_len + 1) + seed_cap + 1;
        std::vector<u32x2> item_width((size_t)std::max(n_work, 1u) * wstride);
        std::vector<uint32_t> next_list(n_groups + 1);
        uint32_t next_count = 0;
        P.item_width = item_width.data(); P.item_width_stride = wstride;
        P.pass = 1; P.group_base = 0; P.n_groups = n_work; P.next_list = next_list.data(); P.next_count = &next_count;
        for (uint32_t w = 0; w < n_work; ++w) width_item(P, dopts.data(), w, st[0]);
        if (kind != KIND_WIDTH) {
            run_worker<uint16_t, false>(P, (uint16_t *)heads.data(), dopts.data(), n_work, st);
            if (kind == KIND_WHOLE) {
                P.pass = 2; P.group_list = next_list.data(); P.n_groups = next_count; P.next_list = nullptr; P.next_count = nullptr;
                for (uint32_t w = 0; w < next_count; ++w) width_item(P, dopts.data(), w, st[0]);
                run_worker<uint16_t, false>(P, (uint16_t *)heads.data(), dopts.data(), next_count, st);
            }
        }
    }
    *lookups = st[0];
    *n_strict = counters[CNT_STRICT] + counters[CNT_BAD];
    *pops = st[1];
    g_last_extra = st[2]; g_last_steps = st[3];
    return (long)counters[CNT_ALN];
}

void emu_set_item_steps(uint32_t *buf) { g_item_steps = buf; g_item_steps_pos = 0; }
uint64_t emu_last_extra(void) { return g_last_extra; }
uint64_t emu_last_steps(void) { return g_last_steps; }

} // extern "C"

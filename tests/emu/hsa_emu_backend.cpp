// hsa_emu_backend.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The part of the C ABI of include/hsa_b200.h that shim/hsa_gpu_shim.c calls, implemented on the host emulation of the device
// sources (hsa_emu.cpp: the exact .cuh files the CUDA kernels are built from, run serially by g++-compiled code).  oracle/Makefile
// links it with the shim, the harness and the UNMODIFIED reference objects into oracle/_ref/hsa_ref_emu, so that the CPU suite
// (tests/test_shim_emu.py, build container only) can run the drop-in end to end -- bwa_cal_sa_reg_gap_gpu with its option switch,
// helper threads and splice batch, generate_sam_se_core_gpu with its formatter -- against the stock program without a GPU.
// It is NOT a fallback: nothing in hsa_b200/ knows about it, libhsa_b200.so does not contain it, and the product's C ABI fails
// loudly without CUDA.
#include "hsa_emu.cpp"
#include <string>

extern "C" int bwa_cal_maxdiff(int l, double err, double thres);       // the reference's own (bwtaln.c:46-58), linked in

struct hsa_index {
    void *emu = nullptr;
    const uint32_t *sa_value = nullptr; uint32_t sa_interval = 0;
    std::vector<uint32_t> blocks4; uint32_t n_blocks = 0;
    const uint32_t *packed_dna = nullptr; uint32_t dna_length = 0;
};

static std::string g_err;
static int fail(int code, const char *msg) { g_err = msg; return code; }

static std::vector<uint8_t> padded(const uint8_t *codes, const uint64_t *off, const uint32_t *len, size_t n)
{
    size_t bytes = 0;
    for (size_t i = 0; i < n; ++i) bytes = std::max(bytes, (size_t)off[i] + len[i]);
    std::vector<uint8_t> c(bytes + 16, 0);                              // as the product's device copy (upload_reads)
    if (bytes) memcpy(c.data(), codes, bytes);
    return c;
}

extern "C" {

const char *hsa_last_error(void) { return g_err.c_str(); }

int hsa_index_upload(int, const hsa_bwt_view_t *fwd, const hsa_bwt_view_t *rev, hsa_index_t **out)
{
    if (!fwd || !rev || !out) return fail(HSA_E_ARG, "null argument");
    hsa_index *ix = new hsa_index();
    ix->emu = emu_index_new(fwd, rev);
    *out = ix;
    return HSA_OK;
}
void hsa_index_free(hsa_index_t *ix) { if (ix) { emu_index_free(ix->emu); delete ix; } }
int hsa_index_attach_sa(hsa_index_t *ix, const uint32_t *sa_value, size_t, uint32_t sa_interval)
{
    ix->sa_value = sa_value; ix->sa_interval = sa_interval; return HSA_OK;
}
int hsa_index_attach_blocks(hsa_index_t *ix, const uint32_t *blocks4, uint32_t n_blocks)
{
    ix->blocks4.assign(blocks4, blocks4 + 4 * (size_t)n_blocks); ix->n_blocks = n_blocks; return HSA_OK;
}
int hsa_index_attach_packed_dna(hsa_index_t *ix, const uint32_t *packed_dna, uint32_t dna_length)
{
    ix->packed_dna = packed_dna; ix->dna_length = dna_length; return HSA_OK;
}
int hsa_sa_values(const hsa_index_t *ix, const uint32_t *sa_index, size_t n, uint32_t *sa_value_out, uint64_t *steps_total)
{
    if (!ix->sa_value) return fail(HSA_E_ARG, "SA samples not attached");
    std::vector<uint32_t> steps(n ? n : 1);
    emu_sa_values(ix->emu, ix->sa_value, ix->sa_interval, sa_index, n, sa_value_out, steps.data());
    if (steps_total) { *steps_total = 0; for (size_t i = 0; i < n; ++i) *steps_total += steps[i]; }
    return HSA_OK;
}
// bwt_match_gap itself, one call with the frame's own arguments: as hsa_b200.cu's hsa_match_gap_call -- the caller's widths become
// the item's row (bound bytes + bases), the search runs as a one-task batch, width_back comes back as gap_shadow left it
int hsa_match_gap_call(const hsa_index_t *ix, const uint8_t *seq, uint32_t len, int strand, hsa_width_t *width_back, hsa_width_t *width_seed,
                       const hsa_gap_opt_t *opt, int *n_aln_out, hsa_aln1_t **aln_out)
{
    if (!ix || !seq || !len || !width_back || !opt || !n_aln_out || !aln_out) return fail(HSA_E_ARG, "null / empty argument");
    uint32_t seed_mode = HSA_SEED_NONE;
    if (width_seed == width_back) {
        if (opt->seed_len != (int)len) return fail(HSA_E_ARG, "width_seed aliasing width_back needs opt->seed_len == len");
        seed_mode = HSA_SEED_ALIAS;
    } else if (width_seed && opt->seed_len > 0 && (uint32_t)opt->seed_len < len) seed_mode = HSA_SEED_TAIL;
    const uint32_t seed_cap = (opt->seed_len > 0 && (uint32_t)opt->seed_len < len) ? (uint32_t)opt->seed_len + 1 : 0;
    Params L;
    memset(&L, 0, sizeof(L));
    set_layout(L, len, seed_cap, 64, 1, 2, true);                     // the row's geometry depends on max_len and seed_cap only
    std::vector<uint8_t> row(L.row_stride, 0);
    uint32_t *w = reinterpret_cast<uint32_t *>(row.data());
    uint32_t w_prev = 0xFFFFFFFFu;
    bool has_n = false;
    for (uint32_t i = 0; i <= len; ++i) {
        w[i] = width_back[i].w;
        uint8_t byte = bound_byte((uint32_t)width_back[i].bid, width_back[i].w, w_prev);
        if (i < len) { if (seq[i] > 3) has_n = true; else byte |= (uint8_t)(seq[i] << BB_BASE_SHIFT); }
        row[L.row_bid_off + i] = byte;
        w_prev = width_back[i].w;
    }
    reinterpret_cast<uint32_t *>(row.data() + L.row_tail_off)[1] = has_n ? ROW_FLAG_HAS_N : 0u;
    if (seed_mode == HSA_SEED_TAIL) {
        w_prev = 0xFFFFFFFFu;
        for (uint32_t i = 0; i <= (uint32_t)opt->seed_len; ++i) {
            row[L.row_seed_off + i] = bound_byte((uint32_t)width_seed[i].bid, width_seed[i].w, w_prev);
            w_prev = width_seed[i].w;
        }
    }
    hsa_task_t t; memset(&t, 0, sizeof(t));
    t.read_off = 0; t.read_len = len; t.strand = 0; t.sub_off = 0; t.len = len; t.wsrc_off = 0; t.seed_mode = seed_mode; t.opt_idx = 0;
    std::vector<uint8_t> c((size_t)len + 16, 0);
    memcpy(c.data(), seq, len);
    const uint64_t off0 = 0; const uint32_t len0 = len;
    int32_t n_aln = 0; uint64_t aln_off = 0; uint8_t status = 0xFF;
    std::vector<uint32_t> aln(9 * 8192);
    std::vector<u32x2> wout((size_t)len + 1);
    uint64_t lk = 0, ns = 0, pops = 0;
    emu_set_rerun(1u << 18); emu_set_coop(0, 0);
    emu_set_rows_host(row.data(), row.size());
    const long total = emu_run(ix->emu, KIND_TASKS, c.data(), &t, &off0, &len0, 1, opt, 1, nullptr, len, 0, 1022, 32, &n_aln, &aln_off, &status,
                               aln.data(), 8192, reinterpret_cast<uint32_t *>(wout.data()), nullptr, &lk, &ns, &pops);
    emu_set_rows_host(nullptr, 0);
    if (total < 0 || status != STATUS_OK) return fail(HSA_E_CAPACITY, "emulation: the call was left unprocessed");
    hsa_aln1_t *out = (hsa_aln1_t *)calloc((size_t)(n_aln < 10 ? 10 : n_aln), sizeof(hsa_aln1_t));
    if (n_aln) memcpy(out, aln.data() + aln_off * 9, (size_t)n_aln * sizeof(hsa_aln1_t));
    for (int q = 0; q < n_aln; ++q) out[q].strand = (uint32_t)strand & 3u;
    for (uint32_t i = 0; i <= len; ++i) {
        width_back[i].w = wout[i].x;
        if (wout[i].y < BB_BID) width_back[i].bid = (int)wout[i].y;
    }
    *n_aln_out = n_aln; *aln_out = out;
    return HSA_OK;
}

void hsa_result_free(hsa_result_t *r)
{
    if (!r) return;
    free(r->n_aln); free(r->aln_off); free(r->aln);
    memset(r, 0, sizeof(*r));
}

// the whole-read part of bwa_cal_sa_reg_gap for one caller opt: per-length option resolution as hsa_b200.cu's
// resolve_whole_opts (bwtaln.c:260-261, 273-276, 330-332), then the emulated kernels with the large-capacity re-run
int hsa_whole_reads(const hsa_index_t *ix, const uint8_t *codes, const uint64_t *off, const uint32_t *len, size_t n,
                    const hsa_gap_opt_t *opt, int keep_gape, hsa_result_t *res)
{
    if (!ix || !opt || !res) return fail(HSA_E_ARG, "null argument");
    res->n_items = 0; res->n_aln_total = 0;
    if (n == 0) return HSA_OK;
    uint32_t max_len = 0;
    for (size_t i = 0; i < n; ++i) max_len = std::max(max_len, len[i]);
    std::vector<uint8_t> seen((size_t)max_len + 1, 0);
    for (size_t i = 0; i < n; ++i) seen[len[i]] = 1;
    std::vector<hsa_gap_opt_t> opts; std::vector<uint16_t> l2o((size_t)max_len + 1, 0);
    for (uint32_t L = 0; L <= max_len; ++L) {
        if (!seen[L]) continue;
        hsa_gap_opt_t o = *opt;
        if (!keep_gape) o.mode &= ~HSA_MODE_GAPE;
        if (opt->fnr > 0.0) o.max_diff = bwa_cal_maxdiff((int)L, 0.02, opt->fnr);
        o.seed_len = opt->seed_len < (int)L ? opt->seed_len : 0x7fffffff;
        size_t j = 0;
        for (; j < opts.size(); ++j) if (opts[j].max_diff == o.max_diff && opts[j].seed_len == o.seed_len) break;
        if (j == opts.size()) opts.push_back(o);
        l2o[L] = (uint16_t)j;
    }
    const int32_t fmax = opt->fnr > 0.0 ? bwa_cal_maxdiff((int)max_len, 0.02, opt->fnr) : opt->max_diff;
    std::vector<uint8_t> c = padded(codes, off, len, n);
    std::vector<int32_t> n_aln(n); std::vector<uint64_t> aln_off(n); std::vector<uint8_t> status(n, 0xFF);
    uint64_t cap = std::max<uint64_t>(8 * n, 1024);
    for (int attempt = 0;; ++attempt) {
        std::vector<uint32_t> aln(cap * 9);
        uint64_t lk = 0, ns = 0, pops = 0;
        emu_set_rerun(1u << 18); emu_set_coop(0, 0);
        const long total = emu_run(ix->emu, KIND_WHOLE, c.data(), nullptr, off, len, (uint32_t)n, opts.data(), (uint32_t)opts.size(), l2o.data(),
                                   max_len, fmax, 1022, 32, n_aln.data(), aln_off.data(), status.data(), aln.data(), cap, nullptr, nullptr,
                                   &lk, &ns, &pops);
        if (total < 0) return fail(HSA_E_CAPACITY, "emulation: capacity");
        if ((uint64_t)total > cap && attempt < 3) { cap = (uint64_t)total + 1024; continue; }
        for (size_t i = 0; i < n; ++i) if (status[i] != STATUS_OK) return fail(HSA_E_CAPACITY, "emulation: an item was left unprocessed");
        res->n_aln = (int32_t *)realloc(res->n_aln, n * sizeof(int32_t));
        res->aln_off = (uint64_t *)realloc(res->aln_off, n * sizeof(uint64_t));
        res->aln = (hsa_aln1_t *)realloc(res->aln, std::max<size_t>((size_t)total, 1) * sizeof(hsa_aln1_t));
        memcpy(res->n_aln, n_aln.data(), n * sizeof(int32_t));
        memcpy(res->aln_off, aln_off.data(), n * sizeof(uint64_t));
        memcpy(res->aln, aln.data(), (size_t)total * 36);
        res->n_items = n; res->n_aln_total = (size_t)total; res->occ_lookups = lk; res->n_strict = ns; res->pops = pops;
        return HSA_OK;
    }
}

int hsa_splice_match_batch(const hsa_index_t *ix, const uint8_t *codes, const uint64_t *off, const uint32_t *len, size_t n,
                           const hsa_gap_opt_t *opts, size_t n_opts, const uint32_t *opt_idx, int32_t *n_aln_out, hsa_aln1_t *aln_out,
                           uint64_t *occ_lookups)
{
    if (occ_lookups) *occ_lookups = 0;
    if (n == 0) return HSA_OK;
    if (!ix->sa_value || ix->blocks4.empty() || !ix->packed_dna) return fail(HSA_E_ARG, "the splice path needs the full index");
    std::vector<uint8_t> c = padded(codes, off, len, n), status(n, 0);
    const uint64_t lk = emu_splice(ix->emu, ix->sa_value, ix->sa_interval, ix->blocks4.data(), ix->n_blocks, ix->packed_dna, ix->dna_length,
                                   c.data(), off, len, n, opts, n_opts, opt_idx, 1u << 16, 4096, n_aln_out, reinterpret_cast<uint32_t *>(aln_out),
                                   status.data());
    for (size_t i = 0; i < n; ++i) if (status[i]) return fail(HSA_E_CAPACITY, "emulation: a read outgrew the splice scratch");
    if (occ_lookups) *occ_lookups = lk;
    return HSA_OK;
}

void hsa_sam_result_free(hsa_sam_result_t *r)
{
    if (!r) return;
    free(r->rec); free(r->multi); free(r->cigar); free(r->md);
    memset(r, 0, sizeof(*r));
}

int hsa_sam_se_batch(const hsa_index_t *ix, const uint8_t *codes, const uint64_t *off, const uint32_t *len, size_t n,
                     const int32_t *n_aln, const uint64_t *aln_off, const hsa_aln1_t *aln, const hsa_gap_opt_t *opt, int n_occ,
                     uint64_t *rng48_state, hsa_sam_result_t *res)
{
    static_assert(sizeof(SamRec) == sizeof(hsa_sam1_t) && sizeof(SamMulti) == sizeof(hsa_multi1_t), "record mirrors");
    if (!ix || !opt || !rng48_state || !res) return fail(HSA_E_ARG, "null argument");
    res->n_reads = n; res->n_multi = res->n_cigar = res->md_bytes = 0; res->n_refined = 0; res->kernel_ms = 0;
    if (n == 0) return HSA_OK;
    if (!ix->sa_value || ix->blocks4.empty() || !ix->packed_dna) return fail(HSA_E_ARG, "the SAM stage needs the full index");
    uint32_t max_len = 0;
    for (size_t i = 0; i < n; ++i) max_len = std::max(max_len, len[i]);
    std::vector<int32_t> md_tab;
    if (opt->fnr > 0.0) { md_tab.resize((size_t)max_len + 1); for (uint32_t l = 0; l <= max_len; ++l) md_tab[l] = bwa_cal_maxdiff((int)l, 0.02, opt->fnr); }
    std::vector<uint8_t> c = padded(codes, off, len, n);
    size_t multi_cap = 128 * n + 1024, cigar_cap = 64 * n + 1024, md_cap = 512 * n + 1024;
    for (int attempt = 0; attempt < 4; ++attempt) {
        res->rec = (hsa_sam1_t *)realloc(res->rec, n * sizeof(hsa_sam1_t));
        res->multi = (hsa_multi1_t *)realloc(res->multi, multi_cap * sizeof(hsa_multi1_t));
        res->cigar = (uint32_t *)realloc(res->cigar, cigar_cap * 4);
        res->md = (char *)realloc(res->md, md_cap);
        uint64_t counts[5] = {0, 0, 0, 0, 0}, st = *rng48_state;
        const int rc = emu_sam(ix->emu, ix->sa_value, ix->sa_interval, ix->blocks4.data(), ix->n_blocks, ix->packed_dna, ix->dna_length, c.data(), off,
                               len, n, n_aln, aln_off, reinterpret_cast<const uint32_t *>(aln), md_tab.empty() ? nullptr : md_tab.data(), opt->max_diff, n_occ,
                               &st, reinterpret_cast<SamRec *>(res->rec), reinterpret_cast<SamMulti *>(res->multi), multi_cap, res->cigar, cigar_cap,
                               res->md, md_cap, counts);
        if (rc == 0 && counts[4] == 0) {
            *rng48_state = st;
            res->n_multi = (size_t)counts[0]; res->n_cigar = (size_t)counts[1]; res->md_bytes = (size_t)counts[2]; res->n_refined = counts[3];
            return HSA_OK;
        }
        multi_cap *= 4; cigar_cap *= 4; md_cap *= 4;               // an output array was too small: again with more room, same rng state
    }
    return fail(HSA_E_CAPACITY, "emulation: SAM arenas");
}

}   // extern "C"

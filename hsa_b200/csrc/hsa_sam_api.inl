// hsa_sam_api.inl -- kernels and C ABI of the SAM-field stage (hsa_sam.cuh); included at the end of hsa_b200.cu.

namespace hsa {

// every read: positions, mapQ, pairing of spliced parts, MD where nothing has to be refined (hsa_sam.cuh: sam_pos_item)
__global__ void __launch_bounds__(256) sam_pos_kernel(const __grid_constant__ SamParams P)
{
    for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < P.n_reads; r += (size_t)gridDim.x * blockDim.x)
        sam_pos_item(P, (uint32_t)r);
}

// the listed reads: banded global alignment(s), CIGAR, MD.  Persistent threads, each with its interleaved scratch slice.
// The trace-back cells (one byte each, written once, read once) always live in global memory; the score row and the
// reference bases, which every cell reads and rewrites, sit in shared memory when a block's share fits (dp_rows_smem).
__global__ void __launch_bounds__(128) sam_dp_kernel(const __grid_constant__ SamParams P, uint32_t n_work)
{
    extern __shared__ int32_t dp_smem[];
    const uint32_t worker = blockIdx.x * blockDim.x + threadIdx.x;
    if (worker >= P.dp_workers) return;
    DpScratch S = dp_scratch_of(P, worker);
    if (P.dp_rows_smem) {
        S.RT = blockDim.x;
        S.rows = dp_smem + threadIdx.x;
        S.ref = reinterpret_cast<uint8_t *>(dp_smem + (size_t)3u * (P.dp_len1_cap + 1u) * blockDim.x) + threadIdx.x;
        S.run_cap = 3u * (P.dp_len1_cap + 1u);
    }
    for (;;) {
        const unsigned long long w = atomicAdd(P.cursor, 1ull);
        if (w >= n_work) break;
        sam_dp_item(P, P.dp_list[w], S);
    }
}

// the hit selection, cut where the drand48 stream allows it (hsa_sam.cuh)
__global__ void __launch_bounds__(256) sel_classify_kernel(const __grid_constant__ SelParams P, uint32_t *max_gaps)
{
    uint32_t g = 0;
    for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < P.n_reads; r += (size_t)gridDim.x * blockDim.x) {
        sel_classify_item(P, (uint32_t)r);
        const uint32_t *a = P.aln + 9 * P.aln_off[r];
        for (int32_t j = 0; j < P.n_aln[r]; ++j) g = max(g, ((a[9 * j] >> 16) & 0xFFu) + (a[9 * j] >> 24));     // n_gapo + n_gape: sizes the DP scratch
    }
    for (int o = 16; o > 0; o >>= 1) g = max(g, __shfl_down_sync(0xffffffffu, g, o));
    if ((threadIdx.x & 31u) == 0 && g) atomicMax(max_gaps, g);
}
__global__ void __launch_bounds__(256) sel_desc_kernel(const __grid_constant__ SelParams P)
{
    for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < P.n_reads; r += (size_t)gridDim.x * blockDim.x) sel_desc_item(P, (uint32_t)r);
}
__global__ void sel_chain_kernel(const __grid_constant__ SelParams P) { sel_chain(P); }
__global__ void __launch_bounds__(256) sel_finish_kernel(const __grid_constant__ SelParams P)
{
    for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < P.n_reads; r += (size_t)gridDim.x * blockDim.x) sel_finish_item(P, (uint32_t)r);
}
__global__ void sel_sequential_kernel(const __grid_constant__ SelParams P) { sel_sequential(P); }

}   // namespace hsa

struct SamCache {
    DevBuf codes, off, len, n_aln, aln_off, aln, rec, multi, cigar, md, list, cnt, maxdiff, dp_bytes, dp_rows;
    DevBuf fixed, dep, slots, desc, pick, vcum, scan_tmp, sel_out;
};
static void sam_cache_free(SamCache *c) { delete c; }

template <typename T> static int host_grow(T *&p, size_t &cap, size_t need)
{
    if (need <= cap && p) return 0;
    const size_t want = std::max<size_t>({need, cap + cap / 2, (size_t)1024});
    T *q = (T *)realloc(p, want * sizeof(T));
    if (!q) return -1;
    p = q; cap = want;
    return 0;
}

extern "C" void hsa_sam_result_free(hsa_sam_result_t *res)
{
    if (!res) return;
    free(res->rec); free(res->multi); free(res->cigar); free(res->md);
    memset(res, 0, sizeof(*res));
}

// What the stage leaves on the device (valid until the next SAM call on the index)
struct SamDeviceOut { size_t n_multi = 0, n_cigar = 0, md_bytes = 0; uint64_t n_refined = 0, n_dependent = 0; float kernel_ms = 0; bool sequential = false; };

// The whole stage on device-resident inputs: selection (classify, prefix sums, chain, finish), positions, refinement.
static int sam_run_device(const hsa_index_t *ix, SamCache &C, const uint8_t *codes_dev, const uint64_t *off_dev, const uint32_t *len_dev,
                          size_t n_reads, uint32_t max_len, const int32_t *n_aln_dev, const uint64_t *aln_off_dev, const uint32_t *aln_dev,
                          const hsa_gap_opt_t *opt, int n_occ, uint64_t *rng48_state, cudaStream_t s, SamDeviceOut &out)
{
    const uint32_t n = (uint32_t)n_reads;
    const size_t n1 = n_reads + 1;
    if (C.fixed.alloc(n1 * 4) || C.dep.alloc(n1 * 4) || C.slots.alloc(n1 * 4) || C.rec.alloc(n_reads * sizeof(SamRec)) ||
        C.list.alloc(n_reads * 4) || C.cnt.alloc(8 * sizeof(unsigned long long)) || C.maxdiff.alloc(((size_t)max_len + 1) * 4) ||
        C.sel_out.alloc(4 * sizeof(unsigned long long)))
        return fail(HSA_E_CUDA, "out of device memory for the SAM batch");
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    CU(cudaEventRecord(e0, s));
    // ---- selection ----
    SelParams S;
    memset(&S, 0, sizeof(S));
    S.n_aln = n_aln_dev; S.aln_off = aln_off_dev; S.aln = aln_dev; S.n_reads = n; S.n_occ = n_occ;
    S.fixed = C.fixed.as<uint32_t>(); S.dep = C.dep.as<uint32_t>(); S.slots = C.slots.as<uint32_t>(); S.rec = C.rec.as<SamRec>();
    S.x0 = *rng48_state & 0xFFFFFFFFFFFFull;
    unsigned long long *so = C.sel_out.as<unsigned long long>();     // {stream state at the end, rare flag, max gaps of any hit}
    S.x_end = reinterpret_cast<uint64_t *>(so); S.rare = reinterpret_cast<uint32_t *>(so + 1);
    uint32_t *max_gaps = reinterpret_cast<uint32_t *>(so + 2);
    CU(cudaMemsetAsync(so, 0, 4 * sizeof(unsigned long long), s));
    CU(cudaMemsetAsync(S.fixed + n, 0, 4, s)); CU(cudaMemsetAsync(S.dep + n, 0, 4, s)); CU(cudaMemsetAsync(S.slots + n, 0, 4, s));
    const uint32_t grid = (uint32_t)std::min<size_t>((n_reads + 255) / 256, (size_t)ix->sm_count * 8);
    sel_classify_kernel<<<grid, 256, 0, s>>>(S, max_gaps);
    CU(cudaGetLastError());
    size_t tmp_bytes = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, S.fixed, S.fixed, (int)n1, s));
    if (C.scan_tmp.alloc(tmp_bytes + 256)) return fail(HSA_E_CUDA, "out of device memory for the SAM batch");
    for (uint32_t *v : {S.fixed, S.dep, S.slots}) {
        size_t tb = C.scan_tmp.cap;
        CU(cub::DeviceScan::ExclusiveSum(C.scan_tmp.p, tb, v, v, (int)n1, s));
    }
    uint32_t totals[3] = {0, 0, 0}, h_gaps = 0;                       // fixed draws, reads with several best hits, multi slots
    CU(cudaMemcpyAsync(&totals[0], S.fixed + n, 4, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(&totals[1], S.dep + n, 4, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(&totals[2], S.slots + n, 4, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(&h_gaps, max_gaps, 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    const size_t n_dep = totals[1], n_multi = totals[2];
    if (C.multi.alloc(std::max<size_t>(n_multi, 1) * sizeof(SamMulti)) || C.desc.alloc(std::max<size_t>(n_dep, 1) * sizeof(SelDesc)) ||
        C.pick.alloc(std::max<size_t>(n_dep, 1) * sizeof(SelPick)) || C.vcum.alloc(std::max<size_t>(n_dep, 1) * 4))
        return fail(HSA_E_CUDA, "out of device memory for the SAM batch");
    S.multi = C.multi.as<SamMulti>(); S.desc = C.desc.as<SelDesc>(); S.pick = C.pick.as<SelPick>(); S.vcum = C.vcum.as<uint32_t>();
    const bool force_seq = env_long("HSA_B200_SAM_SEQUENTIAL", 0) != 0;      // tests: the sequential statement on one device thread
    bool sequential = force_seq;
    if (!sequential) {
        sel_desc_kernel<<<grid, 256, 0, s>>>(S);
        sel_chain_kernel<<<1, 1, 0, s>>>(S);
        sel_finish_kernel<<<grid, 256, 0, s>>>(S);
        CU(cudaGetLastError());
        uint32_t rare = 0;
        CU(cudaMemcpyAsync(&rare, S.rare, 4, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        sequential = rare != 0;
    }
    if (sequential) { sel_sequential_kernel<<<1, 1, 0, s>>>(S); CU(cudaGetLastError()); }
    uint64_t x_end = 0;
    CU(cudaMemcpyAsync(&x_end, S.x_end, 8, cudaMemcpyDeviceToHost, s));
    // ---- positions, refinement ----
    std::vector<int32_t> maxdiff;
    if (opt->fnr > 0.0f) { maxdiff.resize((size_t)max_len + 1); for (uint32_t L = 0; L <= max_len; ++L) maxdiff[L] = hsa_cal_maxdiff((int)L, 0.02, opt->fnr); }
    if (!maxdiff.empty()) CU(cudaMemcpyAsync(C.maxdiff.p, maxdiff.data(), maxdiff.size() * 4, cudaMemcpyHostToDevice, s));
    SamParams P;
    memset(&P, 0, sizeof(P));
    P.env.ix = ix->ix; P.env.sa_value = ix->sa_value; P.env.sa_interval = ix->sa_interval;
    P.env.blocks4 = ix->blocks4; P.env.n_blocks = ix->n_blocks; P.env.packed_dna = ix->packed_dna; P.env.dna_length = ix->dna_length;
    P.codes = codes_dev; P.read_off = off_dev; P.read_len = len_dev; P.n_reads = n;
    P.n_aln = n_aln_dev; P.aln_off = aln_off_dev; P.aln = aln_dev;
    P.rec = C.rec.as<SamRec>(); P.multi = C.multi.as<SamMulti>();
    P.maxdiff_by_len = maxdiff.empty() ? nullptr : C.maxdiff.as<int32_t>(); P.max_mm = opt->max_diff; P.max_len = max_len;
    P.dp_list = C.list.as<uint32_t>();
    unsigned long long *cnt = C.cnt.as<unsigned long long>();        // {cigar words, md bytes, dp reads, cursor, status, dp tasks}
    P.cigar_used = cnt; P.md_used = cnt + 1; P.dp_count = cnt + 2; P.cursor = cnt + 3; P.status = reinterpret_cast<uint32_t *>(cnt + 4);
    P.dp_tasks = cnt + 5;
    // arena sizes: MD from the read count, CIGAR from the refinements the first kernel lists (12 words each); a batch that
    // outgrows them is run again with four times the room (a read stops at its first failed claim, so the counters of a
    // failed attempt are lower bounds, not sizes)
    const bool tiny = env_long("HSA_B200_SAM_TINY", 0) != 0;          // tests: start with arenas that are certainly too small
    size_t cigar_cap = tiny ? 16 : 4096, md_cap = tiny ? 64 : n_reads * 16 + 4096;
    const uint32_t max_ext = h_gaps;
    unsigned long long h[6] = {0, 0, 0, 0, 0, 0};
    bool wide = false;                                               // full-width trace-back rows (a window clipped at the text's end)
    for (int attempt = 0; attempt < 7; ++attempt) {
        if (C.md.alloc(md_cap)) return fail(HSA_E_CUDA, "out of device memory for the SAM batch");
        P.cigar = nullptr; P.cigar_cap = 0; P.md = C.md.as<char>(); P.md_cap = md_cap;
        CU(cudaMemsetAsync(cnt, 0, 8 * sizeof(unsigned long long), s));
        if (attempt) {                                               // the first kernel rewrites records in place: select again
            if (sequential) sel_sequential_kernel<<<1, 1, 0, s>>>(S); else sel_finish_kernel<<<grid, 256, 0, s>>>(S);
        }
        sam_pos_kernel<<<grid, 256, 0, s>>>(P);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(h, cnt, sizeof(h), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        const uint32_t n_dp = (uint32_t)h[2];
        if (!tiny) cigar_cap = std::max<size_t>(cigar_cap, (size_t)h[5] * 12 + 4096);
        if (C.cigar.alloc(cigar_cap * 4)) return fail(HSA_E_CUDA, "out of device memory for the SAM batch");
        P.cigar = C.cigar.as<uint32_t>(); P.cigar_cap = cigar_cap;
        if (n_dp) {
            // scratch: (len2 + 1) rows of W trace-back bytes + len1 reference bases + 3 x (len1 + 1) score words per worker
            const uint32_t len1_cap = max_len + max_ext, len2_cap = max_len;
            const uint32_t W = wide ? len1_cap + 1u : std::min<uint32_t>(2u * DP_BAND + max_ext + 1u, len1_cap + 1u);
            const size_t per_bytes = (size_t)(len2_cap + 1u) * W + len1_cap + 1u, per_rows = std::max<size_t>(3 * ((size_t)len1_cap + 1), (size_t)len1_cap + len2_cap + 2);
            // score row + reference bases in shared memory when 64 workers' share leaves room for two blocks per SM
            // (13 bytes per reference position and worker: reads up to ~125 bases), or one (~250 bases); else everything in
            // global memory with 128-thread blocks, six per SM (80 registers).  Scratch capped at 2 GB.
            const size_t smem64 = (size_t)64 * 13 * ((size_t)len1_cap + 1) + 64;
            const bool rows_smem = !wide && smem64 <= 200 * 1024 && env_long("HSA_B200_SAM_DP_GLOBAL", 0) == 0;
            const uint32_t block = rows_smem ? 64u : 128u;
            const size_t per_sm = rows_smem ? (smem64 <= 100 * 1024 ? 128 : 64) : 768;
            uint32_t workers = (uint32_t)std::min<size_t>({(size_t)n_dp, (size_t)ix->sm_count * per_sm, std::max<size_t>(128, ((size_t)2 << 30) / (per_bytes + 4 * per_rows))});
            workers = (workers + block - 1u) / block * block;
            if (C.dp_bytes.alloc(per_bytes * workers) || C.dp_rows.alloc(per_rows * 4 * workers)) return fail(HSA_E_CUDA, "out of device memory for the DP scratch");
            P.dp_bytes = C.dp_bytes.as<uint8_t>(); P.dp_rows = C.dp_rows.as<int32_t>(); P.dp_workers = workers; P.dp_w = W;
            P.dp_len1_cap = len1_cap; P.dp_len2_cap = len2_cap; P.dp_rows_smem = rows_smem ? 1u : 0u;
            if (rows_smem) CU(cudaFuncSetAttribute(sam_dp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem64));
            sam_dp_kernel<<<workers / block, block, rows_smem ? smem64 : 0, s>>>(P, n_dp);
            CU(cudaGetLastError());
        }
        CU(cudaEventRecord(e1, s));
        CU(cudaMemcpyAsync(h, cnt, sizeof(h), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        const uint32_t status = (uint32_t)(h[4] & 0xFFFFFFFFull);
        if (status == SAM_SCRATCH) {
            if (wide) return fail(HSA_E_CAPACITY, "an alignment exceeded the DP scratch (gap count beyond the batch's maximum)");
            wide = true;
            continue;
        }
        if (h[0] <= cigar_cap && h[1] <= md_cap && status == SAM_OK) break;
        if (attempt == 6) return fail(HSA_E_CAPACITY, "CIGAR / MD arenas overflowed repeatedly");
        if (h[0] > cigar_cap || status == SAM_CIGAR_FULL) cigar_cap = std::max<size_t>(4 * cigar_cap, 2 * (size_t)h[0]);
        if (h[1] > md_cap || status == SAM_MD_FULL) md_cap = std::max<size_t>(4 * md_cap, 2 * (size_t)h[1]);
    }
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *rng48_state = x_end;
    out.n_multi = n_multi; out.n_cigar = (size_t)h[0]; out.md_bytes = (size_t)h[1]; out.n_refined = h[2]; out.n_dependent = n_dep;
    out.kernel_ms = ms; out.sequential = sequential;
    return HSA_OK;
}

extern "C" int hsa_sam_se_batch(const hsa_index_t *ix, const uint8_t *codes, const uint64_t *off, const uint32_t *len, size_t n_reads,
                                const int32_t *n_aln, const uint64_t *aln_off, const hsa_aln1_t *aln, const hsa_gap_opt_t *opt,
                                int n_occ, uint64_t *rng48_state, hsa_sam_result_t *res)
{
    static_assert(sizeof(SamRec) == sizeof(hsa_sam1_t) && sizeof(SamMulti) == sizeof(hsa_multi1_t), "record mirrors");
    if (!ix || !opt || !rng48_state || !res) return fail(HSA_E_ARG, "null argument");
    res->n_reads = n_reads; res->n_multi = res->n_cigar = res->md_bytes = 0; res->n_refined = 0; res->kernel_ms = 0;
    if (n_reads == 0) return HSA_OK;
    if (!codes || !off || !len || !n_aln || !aln_off) return fail(HSA_E_ARG, "null argument");
    if (!ix->sa_value || !ix->blocks4 || !ix->packed_dna)
        return fail(HSA_E_ARG, "the SAM stage needs the SA samples, the block list and the packed text "
                               "(hsa_index_attach_sa / _blocks / _packed_dna)");
    if (n_reads > 0x3FFFFFF0ull) return fail(HSA_E_ARG, "too many reads in one batch");
    if (n_occ < 0 || n_occ > 1000) return fail(HSA_E_ARG, "n_occ out of range");
    const bool trace = env_long("HSA_B200_TRACE", 0) != 0;
    const auto t_begin = std::chrono::steady_clock::now();
    auto since = [&](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
    size_t bytes = 0, n_hits = 0; uint32_t max_len = 0;
    for (size_t i = 0; i < n_reads; ++i) {
        if (len[i] > 4095) return fail(HSA_E_ARG, "reads longer than 4095 bases are not supported");
        if (n_aln[i] < 0) return fail(HSA_E_ARG, "negative n_aln");
        if (n_aln[i] && !aln) return fail(HSA_E_ARG, "hits without a hit array");
        bytes = std::max(bytes, (size_t)off[i] + len[i]); max_len = std::max(max_len, len[i]);
        if (n_aln[i]) n_hits = std::max(n_hits, (size_t)aln_off[i] + (size_t)n_aln[i]);
    }
    CU(cudaSetDevice(ix->device));
    cudaStream_t s = ix->stream;
    hsa_index *mix = const_cast<hsa_index *>(ix);
    if (!mix->sam_cache) mix->sam_cache = new SamCache();
    SamCache &C = *mix->sam_cache;
    if (C.codes.alloc(bytes + 16) || C.off.alloc(n_reads * 8) || C.len.alloc(n_reads * 4) || C.n_aln.alloc(n_reads * 4) ||
        C.aln_off.alloc(n_reads * 8) || C.aln.alloc(std::max<size_t>(n_hits, 1) * 36))
        return fail(HSA_E_CUDA, "out of device memory for the SAM batch");
    CU(cudaMemcpyAsync(C.codes.p, codes, bytes, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(C.off.p, off, n_reads * 8, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(C.len.p, len, n_reads * 4, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(C.n_aln.p, n_aln, n_reads * 4, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(C.aln_off.p, aln_off, n_reads * 8, cudaMemcpyHostToDevice, s));
    if (n_hits) CU(cudaMemcpyAsync(C.aln.p, aln, n_hits * 36, cudaMemcpyHostToDevice, s));
    const double ms_h2d = since(t_begin);
    SamDeviceOut o;
    int rc = sam_run_device(ix, C, C.codes.as<uint8_t>(), C.off.as<uint64_t>(), C.len.as<uint32_t>(), n_reads, max_len, C.n_aln.as<int32_t>(),
                            C.aln_off.as<uint64_t>(), C.aln.as<uint32_t>(), opt, n_occ, rng48_state, s, o);
    if (rc) return rc;
    const double ms_dev = since(t_begin) - ms_h2d;
    res->kernel_ms = o.kernel_ms; res->n_refined = o.n_refined; res->n_multi = o.n_multi; res->n_cigar = o.n_cigar; res->md_bytes = o.md_bytes;
    if (host_grow(res->rec, res->cap_rec, n_reads) || host_grow(res->multi, res->cap_multi, o.n_multi) ||
        host_grow(res->cigar, res->cap_cigar, o.n_cigar) || host_grow(res->md, res->cap_md, o.md_bytes)) return fail(HSA_E_NOMEM, "out of host memory");
    CU(cudaMemcpyAsync(res->rec, C.rec.p, n_reads * sizeof(SamRec), cudaMemcpyDeviceToHost, s));
    if (o.n_multi) CU(cudaMemcpyAsync(res->multi, C.multi.p, o.n_multi * sizeof(SamMulti), cudaMemcpyDeviceToHost, s));
    if (o.n_cigar) CU(cudaMemcpyAsync(res->cigar, C.cigar.p, o.n_cigar * 4, cudaMemcpyDeviceToHost, s));
    if (o.md_bytes) CU(cudaMemcpyAsync(res->md, C.md.p, o.md_bytes, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (trace)
        fprintf(stderr, "[hsa_b200 trace] sam_se: %zu reads | checks + H2D %.1f ms | device stage %.1f ms (%.2f ms between its first and last kernel; "
                        "%llu reads with several best hits on the chain%s) | D2H %.1f ms | refined %llu, cigar words %zu, md bytes %zu\n",
                n_reads, ms_h2d, ms_dev, o.kernel_ms, (unsigned long long)o.n_dependent, o.sequential ? "; SEQUENTIAL selection" : "",
                since(t_begin) - ms_h2d - ms_dev, (unsigned long long)o.n_refined, o.n_cigar, o.md_bytes);
    return HSA_OK;
}

extern "C" int hsa_copy_from_device(const hsa_index_t *ix, void *dst_host, const void *src_dev, size_t bytes)
{
    if (!ix || (bytes && (!dst_host || !src_dev))) return fail(HSA_E_ARG, "null argument");
    CU(cudaSetDevice(ix->device));
    if (bytes) CU(cudaMemcpy(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost));
    return HSA_OK;
}

extern "C" int hsa_sam_se_device(const hsa_index_t *ix, const uint8_t *codes_dev, const uint64_t *off_dev, const uint32_t *len_dev, size_t n_reads,
                                 uint32_t max_len, const int32_t *n_aln_dev, const uint64_t *aln_off_dev, const hsa_aln1_t *aln_dev,
                                 const hsa_gap_opt_t *opt, int n_occ, uint64_t *rng48_state, void *stream, hsa_sam_device_t *out)
{
    if (!ix || !opt || !rng48_state || !out) return fail(HSA_E_ARG, "null argument");
    memset(out, 0, sizeof(*out));
    if (n_reads == 0) return HSA_OK;
    if (!codes_dev || !off_dev || !len_dev || !n_aln_dev || !aln_off_dev || !aln_dev) return fail(HSA_E_ARG, "null argument");
    if (!ix->sa_value || !ix->blocks4 || !ix->packed_dna)
        return fail(HSA_E_ARG, "the SAM stage needs the SA samples, the block list and the packed text "
                               "(hsa_index_attach_sa / _blocks / _packed_dna)");
    if (n_reads > 0x3FFFFFF0ull || max_len > 4095 || n_occ < 0 || n_occ > 1000) return fail(HSA_E_ARG, "argument out of range");
    CU(cudaSetDevice(ix->device));
    hsa_index *mix = const_cast<hsa_index *>(ix);
    if (!mix->sam_cache) mix->sam_cache = new SamCache();
    SamCache &C = *mix->sam_cache;
    SamDeviceOut o;
    int rc = sam_run_device(ix, C, codes_dev, off_dev, len_dev, n_reads, max_len, n_aln_dev, aln_off_dev, reinterpret_cast<const uint32_t *>(aln_dev),
                            opt, n_occ, rng48_state, (cudaStream_t)stream, o);
    if (rc) return rc;
    out->rec_dev = C.rec.as<hsa_sam1_t>(); out->multi_dev = C.multi.as<hsa_multi1_t>(); out->cigar_dev = C.cigar.as<uint32_t>();
    out->md_dev = C.md.as<char>();
    out->n_multi = o.n_multi; out->n_cigar = o.n_cigar; out->md_bytes = o.md_bytes; out->n_refined = o.n_refined; out->n_several_best = o.n_dependent;
    out->kernel_ms = o.kernel_ms;
    return HSA_OK;
}

// ---- bwa_print_sam1 (bwtse.c:677-835), the single-end / no-quality / no-read-group case -------------------------------
namespace {
struct TextBuf {
    char *p = nullptr; size_t n = 0, cap = 0; bool bad = false;
    void need(size_t k) { if (n + k > cap) { size_t c = std::max(cap * 2, n + k + 4096); char *q = (char *)realloc(p, c); if (!q) { bad = true; return; } p = q; cap = c; } }
    void put(char c) { need(1); if (!bad) p[n++] = c; }
    void str(const char *s) { const size_t k = strlen(s); need(k); if (!bad) { memcpy(p + n, s, k); n += k; } }
    void num(long long v) { char t[32]; snprintf(t, sizeof(t), "%lld", v); str(t); }
};
}

extern "C" int hsa_sam_format(const hsa_sam_result_t *res, size_t first, size_t count, const uint8_t *codes, const uint64_t *off,
                              const uint32_t *len, const char *const *names, const char *const *chr_names, size_t n_chr,
                              const hsa_gap_opt_t *opt, char **text_out, size_t *bytes_out)
{
    if (!res || !codes || !off || !len || !chr_names || !opt || !text_out || !bytes_out) return fail(HSA_E_ARG, "null argument");
    if (first + count > res->n_reads) return fail(HSA_E_ARG, "read range outside the result");
    TextBuf T;
    for (size_t r = first; r < first + count; ++r) {
        const hsa_sam1_t &p = res->rec[r];
        if (p.type == TYPE_NO_MATCH) continue;                                      // bwtse.c:923-924
        if (p.seq_id >= n_chr) { free(T.p); return fail(HSA_E_ARG, "seq_id outside chr_names"); }
        const uint8_t *seq = codes + off[r]; const uint32_t L = len[r];
        char nm[32];
        if (names && names[r]) T.str(names[r]); else { snprintf(nm, sizeof(nm), "r%zu", r); T.str(nm); }
        T.put('\t'); T.num(p.strand ? 16 : 0); T.put('\t'); T.str(chr_names[p.seq_id]); T.put('\t');
        T.num((int)p.ori_pos); T.put('\t'); T.num(p.mapQ); T.put('\t');
        if (p.n_cigar) for (uint32_t j = 0; j < p.n_cigar; ++j) {
            const uint32_t c = res->cigar[p.cigar_off + j], op = c >> 28;
            T.num((int)(c & 0x0FFFFFFFu)); T.put(op < 9 ? "MIDNSHP=X"[op] : '?');      // (the reference indexes past its table there)
        } else { T.num(L); T.put('M'); }
        T.str("\t*\t0\t0\t");
        if (!p.strand) for (uint32_t j = 0; j < L; ++j) T.put("ACGTN"[seq[j] > 4 ? 4 : seq[j]]);
        else for (uint32_t j = 0; j < L; ++j) T.put("TGCAN"[seq[L - 1 - j] > 4 ? 4 : seq[L - 1 - j]]);
        T.str("\t*");
        T.str("\tXT:A:"); T.put("NURMS"[p.type > 4 ? 0 : p.type]);
        T.str((opt->mode & HSA_MODE_COMPREAD) ? "\tNM:i:" : "\tCM:i:"); T.num(p.nm);
        if (p.type != TYPE_MATESW) {
            T.str("\tX0:i:"); T.num(p.c1);
            if ((long long)p.c1 <= (long long)opt->max_top2) { T.str("\tX1:i:"); T.num(p.c2); }
        }
        T.str("\tXM:i:"); T.num(p.n_mm); T.str("\tXO:i:"); T.num(p.n_gapo); T.str("\tXG:i:"); T.num(p.n_gapo + p.n_gape);
        if (p.md_len) { T.str("\tMD:Z:"); T.need(p.md_len); if (!T.bad) { memcpy(T.p + T.n, res->md + p.md_off, p.md_len); T.n += p.md_len; } }
        if (p.n_multi) {
            T.str("\tXA:Z:");
            for (uint32_t i = 0; i < p.n_multi; ++i) {
                const hsa_multi1_t &q = res->multi[p.multi_off + i];
                if (q.seq_id >= n_chr) { free(T.p); return fail(HSA_E_ARG, "seq_id outside chr_names"); }
                T.str(chr_names[q.seq_id]); T.put(','); T.put(q.strand ? '-' : '+'); T.num((int)q.ori_pos);
                if (q.n_cigar) for (uint32_t k = 0; k < q.n_cigar; ++k) {
                    const uint32_t c = res->cigar[q.cigar_off + k], op = c >> 28;
                    T.num((int)(c & 0x0FFFFFFFu)); T.put(op < 5 ? "MIDNS"[op] : '?');
                } else { T.put(','); T.num(q.gap + q.mm); T.put(';'); }
            }
        }
        T.put('\n');
        if (T.bad) { free(T.p); return fail(HSA_E_NOMEM, "out of host memory"); }
    }
    *text_out = T.p; *bytes_out = T.n;
    return HSA_OK;
}

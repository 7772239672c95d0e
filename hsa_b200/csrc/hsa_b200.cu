// hsa_b200.cu -- CUDA kernels (sm_100a) and the C ABI of include/hsa_b200.h.
//
// Kernels:
//   repack_kernel        reference-layout BWT + occ tables  ->  one 32-byte sector per 64 symbols
//   occ_kernel           BWTAllOccValue / BWTOccValue on either layout (rank parity)
//   width_kernel         bwt_cal_width for every work item: one thread per item, writes the item's row
//   search_kernel        persistent, atomic work queue; one search (hsa_core.cuh Worker) per thread, one kind
//                        of step per warp trip chosen by vote: the bounded backtracking search bwt_match_gap
//   probe_kernel         random 32-byte-sector loads: the roofline denominator of SURVEY.md 8d
//
// There is no CPU fallback anywhere in this file: every entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <chrono>
#include <string>
#include <vector>
#include <algorithm>
#include "hsa_core.cuh"
#include "hsa_coop.cuh"
#include "hsa_splice.cuh"
#include "hsa_sam.cuh"
#include "../../include/hsa_b200.h"

using namespace hsa;

static_assert(sizeof(hsa_task_t) == sizeof(Task), "hsa_task_t must mirror hsa::Task");
static_assert(sizeof(hsa_aln1_t) == 36, "hsa_aln1_t must be 36 bytes like bwt_aln1_t");
static_assert(sizeof(hsa_gap_opt_t) == 64, "hsa_gap_opt_t must be 64 bytes like gap_opt_t");
static_assert(sizeof(hsa_width_t) == 8, "hsa_width_t must be 8 bytes like bwt_width_t");
static_assert(sizeof(DevOpt) == 64, "DevOpt is staged as 16 ints");

enum { MAX_PIPES = 4 };
// behind the core's counters: per-pipe slots {cursor pass 1, cursor pass 2, pass-2 list count, pad}, then the
// per-phase cycle counters of the -DHSA_PHASE_PROF build {cycles, runs, lanes} x {SLOW, LOOKUP, POP}
enum { CNT_PIPE0 = CNT_TOTAL, CNT_PROF0 = CNT_PIPE0 + 4 * (MAX_PIPES + 2), CNT_ALLOC = CNT_PROF0 + 12 };

// =====================================================================================================
// kernels
// =====================================================================================================

__global__ void repack_kernel(RefBwt ref, u32x4 *blocks, uint32_t n_blocks)
{
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    uint32_t occ[4];
    occ4_ref_raw(ref, b * 64u, occ);
    u32x4 c, w;
    c.x = occ[0]; c.y = occ[1]; c.z = occ[2]; c.w = occ[3];
    const uint4 v = *reinterpret_cast<const uint4 *>(ref.bwt_code + 4 * (size_t)b);
    w.x = v.x; w.y = v.y; w.z = v.z; w.w = v.w;
    blocks[2 * (size_t)b] = c;
    blocks[2 * (size_t)b + 1] = planes_of(w);
}

__global__ void occ_kernel(DevBwt dev, RefBwt ref, int layout, const uint32_t *idx, size_t n,
                           uint32_t *occ4_out, uint32_t *occ1_out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t occ[4];
    if (layout == 1) occ4_dev(dev, idx[i], occ);
    else occ4_ref(ref, idx[i], occ);
    for (int c = 0; c < 4; ++c) { occ4_out[4 * i + c] = occ[c]; occ1_out[4 * i + c] = occ[c]; }
}

// SA index -> text position (BWTSaValue, BWT.c:1195-1225): a chain of dependent sector loads per query whose length
// is geometric (one SA index in sa_interval is sampled), so lanes are refilled from a work queue the moment their chain
// ends instead of idling until the warp's longest chain is done.  A warp reserves SA_CHUNK queries at a time (one
// atomic), hands them to its free lanes in order, and every lane makes one PsiMinus step per trip.
enum : uint32_t { SA_CHUNK = 1024 };
__global__ void __launch_bounds__(256) sa_kernel(DevBwt fwd, const uint32_t *sa_value, uint32_t sa_interval,
                                                 const uint32_t *sa_index, size_t n, uint32_t *out,
                                                 unsigned long long *cursor, unsigned long long *steps_total,
                                                 const uint32_t *blocks4, uint32_t n_blocks, uint32_t *seq_id_out, uint32_t *ori_pos_out)
{
    const uint32_t lane = threadIdx.x & 31u;
    unsigned long long base = 0, my = 0, steps = 0;
    uint32_t remain = 0, cur = 0, walked = 0;
    bool busy = false, dry = false;
    for (;;) {
        const unsigned idle = __ballot_sync(0xffffffffu, !busy);
        if (idle && !dry) {
            if (remain == 0) {                          // reserve the warp's next chunk
                if (lane == 0) base = atomicAdd(cursor, (unsigned long long)SA_CHUNK);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base >= n) dry = true;
                else remain = (uint32_t)min((unsigned long long)SA_CHUNK, (unsigned long long)n - base);
            }
            if (!dry) {
                const uint32_t rank = __popc(idle & ((1u << lane) - 1u)), take = min(remain, (uint32_t)__popc(idle));
                if (!busy && rank < take) { my = base + rank; cur = sa_index[my]; walked = 0; busy = true; }
                base += take; remain -= take;
            }
        }
        if (__ballot_sync(0xffffffffu, busy) == 0) { if (dry) break; else continue; }
        if (busy) {
            if (cur % sa_interval == 0) {               // a sampled SA index: done (BWT.c:1223)
                const uint32_t pos = __ldg(sa_value + cur / sa_interval) + walked;
                out[my] = pos;
                if (blocks4) {                          // the rest of BWTRetrievePositionFromSAIndex (2BWT-Interface.c:339-361)
                    uint32_t sid = 0xFFFFFFFFu, op = 0xFFFFFFFFu;
                    locate_dev(blocks4, n_blocks, pos, sid, op);
                    seq_id_out[my] = sid; ori_pos_out[my] = op;
                }
                steps += walked;
                busy = false;
            } else {
                ++walked;
                cur = cur == fwd.inverse_sa0 ? 0u : psi_minus_dev(fwd, cur);
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) steps += __shfl_down_sync(0xffffffffu, steps, o);
    if (lane == 0 && steps_total) atomicAdd(steps_total, steps);
}

// Pre-binning of ragged batches (north star: "reads pre-binned by length and max_diff"): a counting sort of the batch's
// work order by read length, longest first.  max_diff is a function of the length (bwtaln.c:330-331), so one bin = one
// (length, max_diff) class: the lanes of a warp then walk reads of one shape through the width pass and start searches of
// similar depth, and the longest searches start first.  Only the ORDER of work changes: results are written per item.
enum : uint32_t { BIN_LENS = 4096 };
__global__ void __launch_bounds__(256) bin_count_kernel(const uint32_t *len, uint32_t n, uint32_t *hist)
{
    __shared__ uint32_t sh[BIN_LENS];
    for (uint32_t i = threadIdx.x; i < BIN_LENS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) atomicAdd(&sh[min(len[i], BIN_LENS - 1u)], 1u);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < BIN_LENS; i += blockDim.x) if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
// hist[L] -> first work index of length L, lengths in descending order (one block of BIN_LENS / 4 threads)
__global__ void __launch_bounds__(1024) bin_scan_kernel(uint32_t *hist)
{
    __shared__ uint32_t part[1024];
    const uint32_t t = threadIdx.x;
    uint32_t v[4], sum = 0;
    for (int q = 0; q < 4; ++q) { v[q] = hist[BIN_LENS - 1u - (4u * t + q)]; sum += v[q]; }
    part[t] = sum;
    __syncthreads();
    for (uint32_t d = 1; d < 1024; d <<= 1) {
        const uint32_t add = t >= d ? part[t - d] : 0u;
        __syncthreads();
        part[t] += add;
        __syncthreads();
    }
    uint32_t run = part[t] - sum;
    for (int q = 0; q < 4; ++q) { hist[BIN_LENS - 1u - (4u * t + q)] = run; run += v[q]; }
}
__global__ void __launch_bounds__(256) bin_scatter_kernel(const uint32_t *len, uint32_t n, uint32_t *next, uint32_t *work_list)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        work_list[atomicAdd(&next[min(len[i], BIN_LENS - 1u)], 1u)] = i;
}

// Width kernel of the split pipeline: one thread per work item, every thread of a warp walks reads of the
// same shape, so the loop is divergence-free (bwt_cal_width is a strictly sequential chain per read).
template <int MINB>
__global__ void __launch_bounds__(256, MINB) width_kernel(const __grid_constant__ Params P)
{
    DevOpt *sopt = reinterpret_cast<DevOpt *>(hsa_smem);
    for (uint32_t i = threadIdx.x; i < P.n_opts * (sizeof(DevOpt) / 4); i += blockDim.x)
        reinterpret_cast<int *>(sopt)[i] = reinterpret_cast<const int *>(P.opts)[i];
    __syncthreads();
    const uint32_t n = work_count(P);                                              // n_work caps a device-side count
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < n; w += gridDim.x * blockDim.x)
        width_item(P, sopt, w);
}

// Search kernel: persistent grid, one search per thread, atomic work queue.  Every trip of the loop the warp
// votes for ONE kind of step (hsa_core.cuh: phase_vote) and only the lanes waiting for that kind take it, so
// the lanes that execute share one instruction stream.
// MINB = minimum resident blocks per SM the register allocation must allow; picked at run time
// (HSA_B200_MINB) so occupancy can be tuned on the GPU.
template <int BLOCK, int MINB, typename LinkT, bool BIDS_SMEM>
__global__ void __launch_bounds__(BLOCK, MINB) search_kernel(const __grid_constant__ Params P)
{
    for (uint32_t i = threadIdx.x; i < P.n_opts * (sizeof(DevOpt) / 4); i += blockDim.x)
        reinterpret_cast<int *>(hsa_smem)[i] = reinterpret_cast<const int *>(P.opts)[i];
    if (threadIdx.x < 5) reinterpret_cast<unsigned long long *>(hsa_smem + P.smem_stats_off)[threadIdx.x] = 0ull;
    __syncthreads();

    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_work = work_count(P);
    Worker<LinkT, BIDS_SMEM> w(P, slot, threadIdx.x);
    unsigned long long warp_iters = 0;
    uint32_t pop_trips = 0;
#ifdef HSA_PHASE_PROF
    unsigned long long prof_cyc[3] = {0, 0, 0}, prof_runs[3] = {0, 0, 0}, prof_lanes[3] = {0, 0, 0};
#endif

    for (;;) {
        // classes are two-bit codes (SLOW 0, LOOKUP 1, POP 2, retired 3): two ballots give all four counts
        const uint32_t c = w.cls();
        const unsigned b0 = __ballot_sync(0xffffffffu, c & 1u), b1 = __ballot_sync(0xffffffffu, c & 2u);
        if ((b0 & b1) == 0xffffffffu) break;
        ++warp_iters;
        const uint32_t ph = phase_vote(P, __popc(b0 & ~b1), __popc(b1 & ~b0), __popc(~(b0 | b1)));
#ifdef HSA_PHASE_PROF
        const long long t_ph0 = clock64();
        const uint32_t n_ph = ph == PHASE_LOOKUP ? __popc(b0 & ~b1) : ph == PHASE_POP ? __popc(b1 & ~b0) : __popc(~(b0 | b1));
#endif
        if (ph == PHASE_LOOKUP) {
            // three stages with the warp re-converged in between (hsa_core.cuh: lookup_a / lookup_push / lookup_child)
            typename Worker<LinkT, BIDS_SMEM>::LookupCarry lc;
            uint32_t stage = Worker<LinkT, BIDS_SMEM>::LK_DONE;
            if (c == PHASE_LOOKUP) stage = w.lookup_a(lc);
            __syncwarp();
            if (stage == Worker<LinkT, BIDS_SMEM>::LK_PUSH) stage = w.lookup_push(lc);
            __syncwarp();
            if (stage == Worker<LinkT, BIDS_SMEM>::LK_CHILD) w.lookup_child(lc);
        } else if (ph == PHASE_POP) {
            if (P.drain_budget && (++pop_trips & 31u) == 0 && w.budget != P.drain_budget) {
                // once the queue is dry, the searches still running may only use drain_budget steps in total: what is
                // heavier goes to the warp-cooperative kernel instead of holding the launch at lone-lane speed
                // (the budget is looked at when an entry is popped, so POP trips are where it is refreshed)
                const unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(P.cursor);
                if (cur >= n_work) w.budget = P.drain_budget;
            }
            if (c == PHASE_POP) w.do_pop();
        } else {
            if (w.st == LS_HIT) w.do_hit();
            __syncwarp();
            if (w.st == LS_END) w.do_end();
            __syncwarp();
            const bool need = w.idle();
            const unsigned bal = __ballot_sync(0xffffffffu, need);
            if (bal) {
                // warp-aggregated fetch from the atomic work queue
                unsigned long long base = 0;
                const int leader = __ffs(bal) - 1;
                if ((int)lane == leader) base = atomicAdd(P.cursor, (unsigned long long)__popc(bal));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (need) {
                    const unsigned long long idx = base + __popc(bal & ((1u << lane) - 1u));
                    if (idx < n_work) w.start((uint32_t)idx);
                    else w.retire();
                }
            }
        }
        __syncwarp();
#ifdef HSA_PHASE_PROF
        if (lane == 0) { prof_cyc[ph] += (unsigned long long)(clock64() - t_ph0); prof_runs[ph] += 1; prof_lanes[ph] += n_ph; }
#endif
    }
#ifdef HSA_PHASE_PROF
    if (lane == 0)
        for (int i = 0; i < 3; ++i) {
            atomicAdd(&P.counters[CNT_PROF0 + 3 * i], prof_cyc[i]);
            atomicAdd(&P.counters[CNT_PROF0 + 3 * i + 1], prof_runs[i]);
            atomicAdd(&P.counters[CNT_PROF0 + 3 * i + 2], prof_lanes[i]);
        }
#endif

    // statistics: the block's shared-memory words, one set of atomics per block
    if (lane == 0) atomicAdd(&P.counters[CNT_DIAG_WARP_ITERS], warp_iters);
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long *sv = reinterpret_cast<const unsigned long long *>(hsa_smem + P.smem_stats_off);
        typedef Worker<LinkT, BIDS_SMEM> W;
        atomicMax(&P.counters[CNT_DIAG_MAX_ITEM_STEPS], sv[W::STAT_MAX_ITEM_STEPS]);
        atomicAdd(&P.counters[CNT_LOOKUPS], sv[W::STAT_LOOKUPS]);
        atomicAdd(&P.counters[CNT_DIAG_FAST_LOOKUPS], sv[W::STAT_SEARCH_LOOKUPS]);
        atomicAdd(&P.counters[CNT_POPS], sv[W::STAT_POPS]);
        atomicAdd(&P.counters[CNT_STEPS], sv[W::STAT_STEPS]);
    }
}

// Warp-cooperative kernel for the heavy searches (hsa_coop.cuh): persistent grid, one search per WARP.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) coop_kernel(const __grid_constant__ Params P)
{
    for (uint32_t i = threadIdx.x; i < P.n_opts * (sizeof(DevOpt) / 4); i += blockDim.x)
        reinterpret_cast<int *>(hsa_smem)[i] = reinterpret_cast<const int *>(P.opts)[i];
    __syncthreads();
    const uint32_t wib = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const size_t wg = (size_t)blockIdx.x * (BLOCK / 32) + wib;
    unsigned char *base = hsa_smem + P.smem_opts_bytes + (size_t)wib * P.coop_warp_smem;
    CoopWarp &sh = *reinterpret_cast<CoopWarp *>(base);
    uint8_t *bids = base + ((sizeof(CoopWarp) + 15) & ~size_t(15));
    CoopScratch g;
    g.payload = P.coop_payload + wg * P.coop_cap_chunks * COOP_CHUNK;
    g.info = P.coop_info + wg * P.coop_cap_chunks * COOP_CHUNK;
    g.prev = P.coop_prev + wg * P.coop_cap_chunks;
    g.out_payload = P.coop_out_payload + wg * 32 * COOP_OUT_CAP;
    g.out_info = P.coop_out_info + wg * 32 * COOP_OUT_CAP;
    g.hits = P.coop_hits + wg * COOP_HIT_CAP;
    const uint32_t n_work = work_count(P);
    CoopLane me;
    coop_run(P, sh, bids, g, P.coop_cap_chunks, n_work, me);
    if (lane == 0) {
        atomicAdd(&P.counters[CNT_LOOKUPS], sh.lookups);
        atomicAdd(&P.counters[CNT_POPS], sh.pops);
        atomicAdd(&P.counters[CNT_STEPS], sh.steps);
        atomicAdd(&P.counters[CNT_DIAG_COOP_WAVES], sh.waves);
        atomicAdd(&P.counters[CNT_DIAG_COOP_WAVES + 1], sh.wave_steps);
        atomicAdd(&P.counters[CNT_DIAG_COOP_WAVES + 2], sh.steps);
#ifdef HSA_PHASE_PROF
        atomicAdd(&P.counters[CNT_PROF0 + 9], sh.prof_chain);
        atomicAdd(&P.counters[CNT_PROF0 + 10], sh.prof_commit);
        atomicAdd(&P.counters[CNT_PROF0 + 11], sh.prof_total);
#endif
    }
}

// Splice fallback (hsa_splice.cuh): persistent grid, one read per thread at a time, atomic work queue.  Every thread owns
// a slice of the scratch arrays (stack arena, hit lists, width arrays); reads that outgrow it are listed for a re-run
// with larger slices.
__global__ void __launch_bounds__(128) splice_kernel(const __grid_constant__ SpliceParams P, uint32_t lane_stride)
{
    // Small batches are latency-bound: 32 lanes on 32 different control flows run one after the other, so a read takes
    // 32 x longer than on a lane that has its warp to itself.  lane_stride > 1 leaves only every lane_stride-th lane of a
    // warp working (the host picks it so that the batch still fills the machine once).
    const uint32_t lane = threadIdx.x & 31u;
    if (lane % lane_stride) return;
    const size_t worker = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / lane_stride;
    for (;;) {
        const unsigned long long w = atomicAdd(P.cursor, 1ull);
        if (w >= P.n_work) break;
        splice_item(P, (uint32_t)w, worker);
    }
}

// Random-sector probe: every thread walks `iters` pseudo-random 32-byte sectors of `buf` (n_sectors),
// two independent 16-byte loads per sector like an occ lookup, `CHAINS` independent chains per thread.
template <int CHAINS>
__global__ void probe_kernel(const uint4 *buf, uint64_t n_sectors, int iters, unsigned long long *sink)
{
    uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t s[CHAINS];
    uint32_t acc = 0;
    for (int c = 0; c < CHAINS; ++c) s[c] = (tid * CHAINS + c) * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            s[c] = s[c] * 6364136223846793005ull + 1442695040888963407ull;
            uint64_t sec = (s[c] >> 20) % n_sectors;
            uint4 a = __ldg(buf + 2 * sec), b = __ldg(buf + 2 * sec + 1);
            acc += a.x ^ b.w;
            s[c] ^= (uint64_t)(a.y & 1u);            // make the next address depend on the load
        }
    }
    if (acc == 0xDEADBEEFu) atomicAdd(sink, 1ull);
}

// Second probe variant: one 256-bit load per sector (as the kernels issue them), sector picked by multiply-shift
// instead of a 64-bit modulo, and UNROLL loads in flight per thread whose addresses do not depend on each other's data
// (the upper bound a kernel with more memory-level parallelism could reach).
template <int UNROLL>
__global__ void probe_kernel_mlp(const uint4 *buf, uint64_t n_sectors, int iters, unsigned long long *sink)
{
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t s = tid * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        uint32_t v[UNROLL][8];
#pragma unroll
        for (int c = 0; c < UNROLL; ++c) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            const uint64_t sec = __umul64hi(s, n_sectors);
            asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v[c][0]), "=r"(v[c][1]), "=r"(v[c][2]), "=r"(v[c][3]), "=r"(v[c][4]), "=r"(v[c][5]), "=r"(v[c][6]), "=r"(v[c][7])
                         : "l"(buf + 2 * sec));
        }
#pragma unroll
        for (int c = 0; c < UNROLL; ++c) acc += v[c][0] ^ v[c][7];
    }
    if (acc == 0xDEADBEEFu) atomicAdd(sink, 1ull);
}

// Third probe variant, shaped like sa_kernel's traffic: per iteration one 256-bit sector load and one 4-byte load from
// another random sector (an SA sample), UNROLL of each in flight per thread.  Counts two sectors per pair.
template <int UNROLL>
__global__ void probe_kernel_mix(const uint4 *buf, uint64_t n_sectors, int iters, unsigned long long *sink)
{
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t s = tid * 0x9E3779B97F4A7C15ull + 0x7654321ull;
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        uint32_t v[UNROLL][8], u[UNROLL];
#pragma unroll
        for (int c = 0; c < UNROLL; ++c) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            const uint64_t sec = __umul64hi(s, n_sectors);
            asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v[c][0]), "=r"(v[c][1]), "=r"(v[c][2]), "=r"(v[c][3]), "=r"(v[c][4]), "=r"(v[c][5]), "=r"(v[c][6]), "=r"(v[c][7])
                         : "l"(buf + 2 * sec));
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            const uint64_t sec2 = __umul64hi(s, n_sectors);
            u[c] = __ldg(reinterpret_cast<const uint32_t *>(buf + 2 * sec2) + (c & 7));
        }
#pragma unroll
        for (int c = 0; c < UNROLL; ++c) acc += v[c][0] ^ v[c][7] ^ u[c];
    }
    if (acc == 0xDEADBEEFu) atomicAdd(sink, 1ull);
}

// Fourth probe variant, shaped like the LF walk of sa_kernel (the one shipped kernel that out-ran the other variants,
// VERDICT round 1): ONE dependent chain per thread, the next sector chosen from the data just loaded, chains of geometric
// length (p = 1/8) that end with a 4-byte load from a second array; 2048 threads per SM, no shared memory.  Counts 1 + 1/8
// sectors per step.  The first half of `buf` plays the index, the second half the SA samples.
__global__ void __launch_bounds__(256) probe_kernel_walk(const uint4 *buf, uint64_t n_sectors, int iters, unsigned long long *sink)
{
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, n_units = n_sectors / 2;
    uint64_t s = (tid + 4242) * 0x9E3779B97F4A7C15ull;
    uint32_t acc = 0;
    const uint32_t *tail = reinterpret_cast<const uint32_t *>(buf + 2 * n_units);
    for (int it = 0; it < iters; ++it) {
        const uint64_t u = __umul64hi(s, n_units);
        uint32_t v[8];
        asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(buf + 2 * u));
        const uint32_t x = v[0] ^ v[7];
        s = s * 6364136223846793005ull + 1442695040888963407ull + ((uint64_t)x << 32);
        if (((s >> 40) & 7u) == 0u) {
            acc += __ldg(tail + __umul64hi(s * 0x9E3779B97F4A7C15ull, n_units) * 8);
            s ^= s >> 29;
        }
        acc += x;
    }
    if (acc == 0xDEADBEEFu) atomicAdd(sink, 1ull);
}

// =====================================================================================================
// host side
// =====================================================================================================

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }

#define CU(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail(HSA_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));               \
    } while (0)

static long env_long(const char *name, long dflt)
{
    const char *v = getenv(name);
    return (v && *v) ? atol(v) : dflt;
}

// HSA_B200_HOST_TIMING=1: wall seconds of the host-buffer entry points by phase, printed when the index is freed
static double g_host_t[12];
static int g_host_timing = -1;
static inline bool host_timing() { if (g_host_timing < 0) g_host_timing = env_long("HSA_B200_HOST_TIMING", 0) ? 1 : 0; return g_host_timing > 0; }
static inline double wall_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
struct HostTick {
    double t; bool on;
    HostTick() : t(0), on(host_timing()) { if (on) t = wall_s(); }
    void mark(int k) { if (on) { const double n = wall_s(); g_host_t[k] += n - t; t = n; } }
};

struct hsa_index {
    int device = 0;
    cudaStream_t stream = nullptr;
    uint32_t *ref_code[2] = {nullptr, nullptr}, *ref_occ[2] = {nullptr, nullptr}, *ref_major[2] = {nullptr, nullptr};
    bool own_ref = false;
    u32x4 *blocks[2] = {nullptr, nullptr};
    bool own_blocks = false;
    DevIndex ix;
    RefBwt ref[2];
    bool have_ref = false;
    struct hsa_workspace *ws = nullptr;      // == pool[0]; lazily created for the host-buffer entry points
    struct hsa_workspace *pool[3] = {nullptr, nullptr, nullptr};   // one per in-flight job
    bool pool_busy[3] = {false, false, false};
    cudaStream_t h2d = nullptr, d2h = nullptr;                     // copy streams of the job pipeline
    int sm_count = 0;
    uint32_t *sa_value = nullptr; size_t sa_words = 0; uint32_t sa_interval = 0;    // forward text's SA samples (optional)
    enum { SA_SLOTS = 8 };
    unsigned long long *sa_counters = nullptr;                                        // SA_SLOTS x {work cursor, PsiMinus steps}: one slot per call, round-robin
    mutable uint32_t sa_seq = 0;
    uint32_t *blocks4 = nullptr; uint32_t n_blocks = 0;                               // HSP::blockList rows (optional)
    uint32_t *packed_dna = nullptr; uint32_t dna_length = 0;                          // HSP::packedDNA (optional; splice path)
    struct SpliceCache *splice_cache = nullptr;                                       // per-worker scratch of the splice path, kept between calls
    struct SamCache *sam_cache = nullptr;                                             // device buffers of the SAM-field stage, kept between calls
    bool splice_stack_set = false;                                                    // device stack limit raised for splice_kernel
};

struct Scratch {                             // worker-private device memory for one launch configuration
    uint32_t n_workers = 0, arena_cap = 0, hit_cap = 0, link_bytes = 0;
    u32x4 *arena = nullptr; Hit *hits = nullptr;
    void release()
    {
        cudaFree(arena); cudaFree(hits);
        arena = nullptr; hits = nullptr; n_workers = 0;
    }
};

struct Pipe {                                // one in-flight chunk: stream, worker scratch, item rows, pass-2 list
    cudaStream_t stream = nullptr;           // internal stream (used when a batch runs on more than one pipe)
    cudaEvent_t done = nullptr;
    Scratch sc;
    uint8_t *rows = nullptr; size_t rows_cap = 0;
    uint32_t *next_list = nullptr; size_t next_cap = 0;
    // warp-cooperative stage: per-warp record chunks, per-lane push buffers, hit buffers
    uint32_t coop_warps = 0, coop_cap_chunks = 0;
    u32x4 *coop_payload = nullptr, *coop_out_payload = nullptr;
    uint32_t *coop_info = nullptr, *coop_prev = nullptr, *coop_out_info = nullptr;
    Hit *coop_hits = nullptr;
    void release()
    {
        sc.release(); cudaFree(rows); cudaFree(next_list);
        cudaFree(coop_payload); cudaFree(coop_out_payload); cudaFree(coop_info); cudaFree(coop_prev); cudaFree(coop_out_info);
        cudaFree(coop_hits);
        coop_payload = coop_out_payload = nullptr; coop_info = coop_prev = coop_out_info = nullptr; coop_hits = nullptr; coop_warps = 0;
        if (stream) cudaStreamDestroy(stream);
        if (done) cudaEventDestroy(done);
        rows = nullptr; next_list = nullptr; stream = nullptr; done = nullptr; rows_cap = next_cap = 0;
    }
};


struct hsa_workspace {
    const hsa_index *idx = nullptr;
    Pipe pipes[MAX_PIPES];
    Pipe heavy;                                      // warp-cooperative stage (slot MAX_PIPES)
    Pipe strict;                                     // large-capacity re-runs (slot MAX_PIPES + 1)
    uint32_t *strict2_list = nullptr; size_t strict2_list_cap = 0;   // what the cooperative stage hands on
    bool heavy_enqueued = false; uint32_t heavy_cap = 0;             // the cooperative stage was queued behind the fast one
    unsigned long long *counters = nullptr;          // CNT_ALLOC
    uint32_t *strict_list = nullptr; size_t strict_list_cap = 0;
    DevOpt *opts_dev = nullptr; size_t opts_cap = 0;
    uint16_t *len2opt_dev = nullptr; size_t len2opt_cap = 0;
    // pinned staging of the option table + length map: copies need no host sync.  A ring of OPT_STAGES buffers, each
    // guarded by an event recorded behind its H2D copy, so that an asynchronous caller (hsa_whole_reads_device) may
    // change the options from one call to the next without overwriting a copy that has not run yet.
    enum { OPT_STAGES = 4 };
    DevOpt *opts_stage[OPT_STAGES] = {nullptr, nullptr, nullptr, nullptr}; size_t opts_stage_cap[OPT_STAGES] = {0, 0, 0, 0};
    uint16_t *l2o_stage[OPT_STAGES] = {nullptr, nullptr, nullptr, nullptr}; size_t l2o_stage_cap[OPT_STAGES] = {0, 0, 0, 0};
    cudaEvent_t stage_ev[OPT_STAGES] = {nullptr, nullptr, nullptr, nullptr}; bool stage_used[OPT_STAGES] = {false, false, false, false};
    uint32_t stage_next = 0;
    DevOpt *opts_host = nullptr;                                       // the stage of the current batch (host-side reads)
    uint8_t *status_dev = nullptr; size_t status_cap = 0;
    uint32_t *bin_list = nullptr; size_t bin_list_cap = 0; uint32_t *bin_hist = nullptr;    // pre-binned work order of a ragged batch
    // staging for the host-buffer entry points
    uint8_t *codes_dev = nullptr; size_t codes_cap = 0;
    uint64_t *off_dev = nullptr; uint32_t *len_dev = nullptr; size_t reads_cap = 0;
    Task *tasks_dev = nullptr; size_t tasks_cap = 0;
    int32_t *n_aln_dev = nullptr; size_t items_cap = 0; uint64_t *aln_off_dev = nullptr; size_t items2_cap = 0;
    uint32_t *aln_dev = nullptr; size_t aln_cap = 0;
    u32x2 *width_out_dev = nullptr; size_t width_out_cap = 0; int32_t *bid_dev = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaEvent_t> trace_ev;             // HSA_B200_TRACE=1: one event after every launch
    std::vector<const char *> trace_name;
    int trace = -1;
    uint32_t last_launches = 0;
    // what hsa_workspace_check needs to know about the last device-resident call
    cudaStream_t last_stream = nullptr; bool last_valid = false; uint64_t last_n = 0, last_aln_cap = 0;
    // launch configuration (environment overrides are read once)
    bool configured = false;
    uint32_t block = 128; int minb = 5; int blocks_per_sm_cap = 0;
    bool minb_auto = true; long nb_fast_env = -1;   // HSA_B200_MINB / HSA_B200_NB_FAST not given: dense_fast() decides per batch
    bool dense_now = false;                          // ... and this is its verdict on the batch batch_params() saw last
    uint32_t last_config[4] = {0, 0, 0, 0};          // hsa_workspace_last_config
    uint32_t n_pipes = 1; uint64_t chunk_items = 12u << 20;
    uint32_t arena_cap = 1022, hit_cap = 32;
    uint32_t vote_slow_min = VOTE_SLOW_MIN_DEFAULT; int32_t vote_pop_bias = VOTE_POP_BIAS_DEFAULT;
    bool use_coop = true; uint32_t step_budget = 0, drain_budget = 1000;
    bool prebin = true;
};

template <typename T>
static int ensure(T *&p, size_t &cap, size_t need)
{
    if (need <= cap && p) return HSA_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t n = need + need / 8 + 64;
    CU(cudaMalloc((void **)&p, n * sizeof(T)));
    cap = n;
    return HSA_OK;
}

static void to_devopt(const hsa_gap_opt_t &o, DevOpt &d)
{
    memset(&d, 0, sizeof(d));
    d.s_mm = o.s_mm; d.s_gapo = o.s_gapo; d.s_gape = o.s_gape; d.mode = o.mode;
    d.indel_end_skip = o.indel_end_skip; d.max_del_occ = o.max_del_occ; d.max_entries = o.max_entries;
    d.max_diff = o.max_diff; d.max_gapo = o.max_gapo; d.max_gape = o.max_gape;
    d.max_seed_diff = o.max_seed_diff; d.seed_len = o.seed_len; d.max_top2 = o.max_top2;
}

static int check_opt(const hsa_gap_opt_t &o, uint32_t max_len, uint32_t *n_buckets)
{
    if (o.s_mm < 0 || o.s_gapo < 0 || o.s_gape < 0) return fail(HSA_E_ARG, "negative scores are not supported");
    if (o.max_diff < 0) return fail(HSA_E_ARG, "max_diff < 0 after resolution (fnr <= 0 and max_diff unset?)");
    if (o.max_entries < 0) return fail(HSA_E_ARG, "max_entries < 0");
    if (o.max_diff > 30 || o.max_gapo > 15 || o.max_gape > 31 || o.max_gapo < 0 || o.max_gape < 0)
        return fail(HSA_E_ARG, "max_diff/max_gapo/max_gape exceed the packed stack-record fields (30/15/31)");
    if (max_len > 4095) return fail(HSA_E_ARG, "reads longer than 4095 bases are not supported");
    long nb = (long)(o.max_diff + 1) * o.s_mm + (long)(o.max_gapo + 1) * o.s_gapo + (long)(o.max_gape + 1) * o.s_gape + 1;
    if (nb > 128) return fail(HSA_E_ARG, "score range exceeds 128 buckets: (max_diff+1)*s_mm+(max_gapo+1)*s_gapo+(max_gape+1)*s_gape+1 > 128");
    if ((uint32_t)nb > *n_buckets) *n_buckets = (uint32_t)nb;
    return HSA_OK;
}

// ---------------------------------------------------------------------------------------------- basics
extern "C" int hsa_b200_abi_version(void) { return HSA_B200_ABI_VERSION; }
extern "C" const char *hsa_last_error(void) { return g_err.c_str(); }

extern "C" void hsa_gap_opt_default(hsa_gap_opt_t *o)          // gap_init_opt, bwtaln.c:21-44
{
    memset(o, 0, sizeof(*o));
    o->s_mm = 3; o->s_gapo = 11; o->s_gape = 4;
    o->max_diff = -1; o->max_gapo = 1; o->max_gape = 6;
    o->indel_end_skip = 5; o->max_del_occ = 10; o->max_entries = 2000000;
    o->mode = HSA_MODE_GAPE | HSA_MODE_COMPREAD;
    o->seed_len = 32; o->max_seed_diff = 2;
    o->fnr = 0.04f;
    o->n_threads = 1; o->max_top2 = 30; o->trim_qual = 0;
}

extern "C" int hsa_cal_maxdiff(int l, double err, double thres) // bwa_cal_maxdiff, bwtaln.c:46-58
{
    double elambda = exp(-l * err), sum, y = 1.0;
    int k, x = 1;
    for (k = 1, sum = elambda; k < 1000; ++k) {
        y *= l * err;
        x = (int)((unsigned)x * (unsigned)k);       // the reference's int product wraps for k > 12; keep that
        sum += elambda * y / x;
        if (1.0 - sum < thres) return k;
    }
    return 2;
}

// ---------------------------------------------------------------------------------------------- index
// The path's memory traffic is random 32-byte sectors, and an L2 miss brings in more than the sector asked for (ncu,
// 3.1 Gb genome: 831 GB read from DRAM for 396 GB delivered to the SMs).  cudaLimitMaxL2FetchGranularity is the knob
// CUDA offers for that; on B200 it changes nothing measurable (HSA_B200_L2_FETCH = 32 / 64 / 128 / default: 646.3 /
// 645.3 / 646.2 / 647.4 ms per 12.5 M-read batch, probe 1220 GB/s each time), so the driver default is left alone
// unless the variable is set.
static void apply_l2_fetch_granularity()
{
    static bool done = false;
    if (done) return;
    done = true;
    const long g = env_long("HSA_B200_L2_FETCH", 0);
    if (g > 0) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)g);
}

static int init_index_common(hsa_index *ix, int device)
{
    ix->device = device;
    CU(cudaSetDevice(device));
    CU(cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&ix->h2d, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&ix->d2h, cudaStreamNonBlocking));
    CU(cudaDeviceGetAttribute(&ix->sm_count, cudaDevAttrMultiProcessorCount, device));
    apply_l2_fetch_granularity();
    return HSA_OK;
}

static void fill_dev_bwt(DevBwt &d, const hsa_bwt_view_t *v, const u32x4 *blocks)
{
    d.blocks = blocks; d.n_blocks = v->textLength / 64 + 1; d.text_length = v->textLength; d.inverse_sa0 = v->inverseSa0;
    for (int i = 0; i < 5; ++i) d.cum[i] = v->cumulativeFreq[i];
}

static int repack_direction(hsa_index *ix, int which, const hsa_bwt_view_t *v)
{
    uint32_t nb = v->textLength / 64 + 1;
    if ((size_t)4 * nb > v->bwtSizeInWord)
        return fail(HSA_E_ARG, "bwtCode array shorter than the resident size BWTLoad allocates (BWT.c:176)");
    CU(cudaMalloc((void **)&ix->blocks[which], (size_t)nb * 2 * sizeof(u32x4)));
    ix->own_blocks = true;
    repack_kernel<<<(nb + 255) / 256, 256, 0, ix->stream>>>(ix->ref[which], ix->blocks[which], nb);
    CU(cudaGetLastError());
    fill_dev_bwt(which == 0 ? ix->ix.fwd : ix->ix.rev, v, ix->blocks[which]);
    return HSA_OK;
}

static int validate_view(const hsa_bwt_view_t *v)
{
    if (!v || !v->bwtCode || !v->occValue || !v->occValueMajor) return fail(HSA_E_ARG, "null BWT view");
    if (v->cumulativeFreq[4] != v->textLength) return fail(HSA_E_ARG, "cumulativeFreq[4] != textLength");
    uint32_t n = v->textLength, num = (n + 255) / 256 + 1;
    if (v->occSizeInWord < (num + 1) / 2 * 4 || v->occMajorSizeInWord < (num + 255) / 256 * 4)
        return fail(HSA_E_ARG, "occ tables shorter than BWTOccValue{Minor,Major}SizeInWord (BWT.c:1097-1116)");
    return HSA_OK;
}

extern "C" void hsa_index_free(hsa_index_t *ix);

extern "C" int hsa_index_upload(int device, const hsa_bwt_view_t *fwd, const hsa_bwt_view_t *rev, hsa_index_t **out)
{
    if (!out) return fail(HSA_E_ARG, "out is null");
    int rc;
    if ((rc = validate_view(fwd)) || (rc = validate_view(rev))) return rc;
    hsa_index *ix = new hsa_index();
    if ((rc = init_index_common(ix, device))) { delete ix; return rc; }
    const hsa_bwt_view_t *vs[2] = {fwd, rev};
    ix->own_ref = true; ix->have_ref = true;
    auto upload_all = [&]() -> int {
        for (int d = 0; d < 2; ++d) {
            const hsa_bwt_view_t *v = vs[d];
            CU(cudaMalloc((void **)&ix->ref_code[d], (size_t)v->bwtSizeInWord * 4));
            CU(cudaMalloc((void **)&ix->ref_occ[d], (size_t)v->occSizeInWord * 4));
            CU(cudaMalloc((void **)&ix->ref_major[d], (size_t)v->occMajorSizeInWord * 4));
            CU(cudaMemcpyAsync(ix->ref_code[d], v->bwtCode, (size_t)v->bwtSizeInWord * 4, cudaMemcpyHostToDevice, ix->stream));
            CU(cudaMemcpyAsync(ix->ref_occ[d], v->occValue, (size_t)v->occSizeInWord * 4, cudaMemcpyHostToDevice, ix->stream));
            CU(cudaMemcpyAsync(ix->ref_major[d], v->occValueMajor, (size_t)v->occMajorSizeInWord * 4, cudaMemcpyHostToDevice, ix->stream));
            ix->ref[d].bwt_code = ix->ref_code[d]; ix->ref[d].occ_value = ix->ref_occ[d]; ix->ref[d].occ_major = ix->ref_major[d];
            ix->ref[d].text_length = v->textLength; ix->ref[d].inverse_sa0 = v->inverseSa0;
            int rc2;
            if ((rc2 = repack_direction(ix, d, v))) return rc2;
        }
        CU(cudaStreamSynchronize(ix->stream));
        return HSA_OK;
    };
    if ((rc = upload_all())) { const std::string keep = g_err; hsa_index_free(ix); g_err = keep; return rc; }   // nothing leaks on failure
    *out = ix;
    return HSA_OK;
}

extern "C" int hsa_index_from_device(int device, const hsa_bwt_view_t *fwd, const hsa_bwt_view_t *rev, hsa_index_t **out)
{
    if (!out) return fail(HSA_E_ARG, "out is null");
    int rc;
    if ((rc = validate_view(fwd)) || (rc = validate_view(rev))) return rc;
    hsa_index *ix = new hsa_index();
    if ((rc = init_index_common(ix, device))) { delete ix; return rc; }
    const hsa_bwt_view_t *vs[2] = {fwd, rev};
    ix->own_ref = false; ix->have_ref = true;
    for (int d = 0; d < 2 && !rc; ++d) {
        const hsa_bwt_view_t *v = vs[d];
        ix->ref[d].bwt_code = v->bwtCode; ix->ref[d].occ_value = v->occValue; ix->ref[d].occ_major = v->occValueMajor;
        ix->ref[d].text_length = v->textLength; ix->ref[d].inverse_sa0 = v->inverseSa0;
        rc = repack_direction(ix, d, v);
    }
    if (!rc && cudaStreamSynchronize(ix->stream) != cudaSuccess) rc = fail(HSA_E_CUDA, "re-pack kernel failed");
    if (rc) { const std::string keep = g_err; hsa_index_free(ix); g_err = keep; return rc; }
    // the caller keeps ownership of the reference-layout arrays and may free them now
    ix->have_ref = false;
    *out = ix;
    return HSA_OK;
}

extern "C" int hsa_index_blocks(const hsa_index_t *ix, int which, void **dev_ptr, size_t *bytes)
{
    if (!ix || which < 0 || which > 1) return fail(HSA_E_ARG, "bad index / direction");
    const DevBwt &b = which == 0 ? ix->ix.fwd : ix->ix.rev;
    if (dev_ptr) *dev_ptr = (void *)b.blocks;
    if (bytes) *bytes = (size_t)b.n_blocks * 2 * sizeof(u32x4);
    return HSA_OK;
}

extern "C" int hsa_index_meta(const hsa_index_t *ix, int which, uint32_t meta[7])
{
    if (!ix || which < 0 || which > 1) return fail(HSA_E_ARG, "bad index / direction");
    const DevBwt &b = which == 0 ? ix->ix.fwd : ix->ix.rev;
    meta[0] = b.text_length; meta[1] = b.inverse_sa0;
    for (int i = 0; i < 5; ++i) meta[2 + i] = b.cum[i];
    return HSA_OK;
}

extern "C" int hsa_index_from_blocks(int device, const uint32_t meta_fwd[7], const uint32_t meta_rev[7],
                                     void *blocks_fwd_dev, void *blocks_rev_dev, int take_ownership, hsa_index_t **out)
{
    if (!out || !blocks_fwd_dev || !blocks_rev_dev) return fail(HSA_E_ARG, "null argument");
    hsa_index *ix = new hsa_index();
    int rc;
    if ((rc = init_index_common(ix, device))) { delete ix; return rc; }
    const uint32_t *ms[2] = {meta_fwd, meta_rev};
    void *bs[2] = {blocks_fwd_dev, blocks_rev_dev};
    for (int d = 0; d < 2; ++d) {
        DevBwt &b = d == 0 ? ix->ix.fwd : ix->ix.rev;
        b.blocks = (const u32x4 *)bs[d]; b.text_length = ms[d][0]; b.inverse_sa0 = ms[d][1];
        b.n_blocks = ms[d][0] / 64 + 1;
        for (int i = 0; i < 5; ++i) b.cum[i] = ms[d][2 + i];
        ix->blocks[d] = (u32x4 *)bs[d];
    }
    ix->own_blocks = take_ownership != 0;
    *out = ix;
    return HSA_OK;
}

extern "C" void hsa_workspace_free(hsa_workspace_t *ws);
static void splice_cache_free(struct SpliceCache *c);
static void sam_cache_free(struct SamCache *c);

extern "C" void hsa_index_free(hsa_index_t *ix)
{
    if (!ix) return;
    if (host_timing())
        fprintf(stderr, "[hsa_b200 host] whole_reads: submit (checks + H2D + enqueue) %.4f | wait for the batch %.4f | results D2H %.4f || "
                        "splice: checks + H2D %.4f | pass 1 %.4f | large-slice pass %.4f | D2H %.4f\n",
                g_host_t[0], g_host_t[1], g_host_t[2], g_host_t[3], g_host_t[4], g_host_t[5], g_host_t[6]);
    cudaSetDevice(ix->device);
    for (hsa_workspace *w : ix->pool) if (w) hsa_workspace_free(w);
    if (ix->h2d) cudaStreamDestroy(ix->h2d);
    if (ix->d2h) cudaStreamDestroy(ix->d2h);
    if (ix->own_ref) for (int d = 0; d < 2; ++d) { cudaFree(ix->ref_code[d]); cudaFree(ix->ref_occ[d]); cudaFree(ix->ref_major[d]); }
    if (ix->own_blocks) for (int d = 0; d < 2; ++d) cudaFree(ix->blocks[d]);
    cudaFree(ix->sa_value); cudaFree(ix->sa_counters); cudaFree(ix->blocks4); cudaFree(ix->packed_dna);
    splice_cache_free(ix->splice_cache);
    sam_cache_free(ix->sam_cache);
    if (ix->stream) cudaStreamDestroy(ix->stream);
    delete ix;
}

// ---------------------------------------------------------------------------------------------- rank
extern "C" int hsa_occ_batch(const hsa_index_t *ix, int which, int layout, const uint32_t *indices, size_t n,
                             uint32_t *occ4_out, uint32_t *occ1_out)
{
    if (!ix || which < 0 || which > 1 || (layout != 0 && layout != 1)) return fail(HSA_E_ARG, "bad argument");
    if (layout == 0 && !ix->have_ref) return fail(HSA_E_ARG, "reference-layout arrays are not resident in this index");
    if (n == 0) return HSA_OK;
    CU(cudaSetDevice(ix->device));
    uint32_t *d_idx = nullptr, *d4 = nullptr, *d1 = nullptr;
    CU(cudaMalloc((void **)&d_idx, n * 4)); CU(cudaMalloc((void **)&d4, n * 16)); CU(cudaMalloc((void **)&d1, n * 16));
    CU(cudaMemcpyAsync(d_idx, indices, n * 4, cudaMemcpyHostToDevice, ix->stream));
    occ_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ix->stream>>>(which == 0 ? ix->ix.fwd : ix->ix.rev, ix->ref[which],
                                                                    layout, d_idx, n, d4, d1);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(occ4_out, d4, n * 16, cudaMemcpyDeviceToHost, ix->stream));
    if (occ1_out) CU(cudaMemcpyAsync(occ1_out, d1, n * 16, cudaMemcpyDeviceToHost, ix->stream));
    CU(cudaStreamSynchronize(ix->stream));
    cudaFree(d_idx); cudaFree(d4); cudaFree(d1);
    return HSA_OK;
}

// ---------------------------------------------------------------------------------------------- SA -> position
extern "C" int hsa_index_attach_sa(hsa_index_t *ix, const uint32_t *sa_value, size_t n_words, uint32_t sa_interval)
{
    if (!ix || !sa_value || sa_interval == 0) return fail(HSA_E_ARG, "bad argument");
    const uint32_t n = ix->ix.fwd.text_length;
    if (n_words != ((size_t)n + sa_interval) / sa_interval) return fail(HSA_E_ARG, "SA sample count does not match textLength / saInterval (BWT.c:219)");
    CU(cudaSetDevice(ix->device));
    cudaFree(ix->sa_value); ix->sa_value = nullptr;
    CU(cudaMalloc((void **)&ix->sa_value, n_words * sizeof(uint32_t)));
    if (!ix->sa_counters) CU(cudaMalloc((void **)&ix->sa_counters, 2 * hsa_index::SA_SLOTS * sizeof(unsigned long long)));
    CU(cudaMemcpyAsync(ix->sa_value, sa_value, n_words * sizeof(uint32_t), cudaMemcpyHostToDevice, ix->stream));
    const uint32_t minus1 = 0xFFFFFFFFu;                 // BWT.c:222: SA[0] is kept as -1 whatever the file holds
    CU(cudaMemcpyAsync(ix->sa_value, &minus1, sizeof(uint32_t), cudaMemcpyHostToDevice, ix->stream));
    CU(cudaStreamSynchronize(ix->stream));
    ix->sa_words = n_words; ix->sa_interval = sa_interval;
    return HSA_OK;
}

extern "C" int hsa_index_attach_blocks(hsa_index_t *ix, const uint32_t *blocks4, uint32_t n_blocks)
{
    if (!ix || !blocks4 || n_blocks == 0) return fail(HSA_E_ARG, "bad argument");
    for (uint32_t i = 0; i < n_blocks; ++i)
        if (blocks4[4 * i + 1] > blocks4[4 * i + 2] || (i && blocks4[4 * i + 1] <= blocks4[4 * i - 2]))
            return fail(HSA_E_ARG, "block list is not ascending and disjoint");
    CU(cudaSetDevice(ix->device));
    cudaFree(ix->blocks4); ix->blocks4 = nullptr;
    CU(cudaMalloc((void **)&ix->blocks4, (size_t)n_blocks * 16));
    CU(cudaMemcpy(ix->blocks4, blocks4, (size_t)n_blocks * 16, cudaMemcpyHostToDevice));
    ix->n_blocks = n_blocks;
    return HSA_OK;
}

// Every call takes its own {cursor, steps} slot, so calls on different streams never share a work cursor (up to SA_SLOTS
// calls may be in flight per index).  *slot_out = the slot's counters.
static int sa_launch(const hsa_index_t *ix, const uint32_t *idx_dev, size_t n, uint32_t *out_dev, cudaStream_t s,
                     unsigned long long **slot_out, uint32_t *seq_dev = nullptr, uint32_t *ori_dev = nullptr)
{
    unsigned long long *cnt = ix->sa_counters + 2 * (ix->sa_seq++ % hsa_index::SA_SLOTS);
    *slot_out = cnt;
    CU(cudaMemsetAsync(cnt, 0, 2 * sizeof(unsigned long long), s));
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sa_kernel, 256, 0));
    const size_t want = (n + SA_CHUNK - 1) / SA_CHUNK;   // one warp per chunk is the most that can find work
    const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>((size_t)ix->sm_count * std::max(occ, 1), (want + 7) / 8));
    sa_kernel<<<grid, 256, 0, s>>>(ix->ix.fwd, ix->sa_value, ix->sa_interval, idx_dev, n, out_dev, cnt, cnt + 1,
                                   seq_dev ? ix->blocks4 : nullptr, ix->n_blocks, seq_dev, ori_dev);
    CU(cudaGetLastError());
    return HSA_OK;
}

extern "C" int hsa_sa_values(const hsa_index_t *ix, const uint32_t *sa_index, size_t n, uint32_t *sa_value_out, uint64_t *steps_total)
{
    if (!ix || (n && (!sa_index || !sa_value_out))) return fail(HSA_E_ARG, "bad argument");
    if (!ix->sa_value) return fail(HSA_E_ARG, "no SA samples attached to this index (hsa_index_attach_sa)");
    if (steps_total) *steps_total = 0;
    if (n == 0) return HSA_OK;
    for (size_t i = 0; i < n; ++i)
        if (sa_index[i] > ix->ix.fwd.text_length) return fail(HSA_E_ARG, "SA index beyond textLength");
    CU(cudaSetDevice(ix->device));
    uint32_t *d_idx = nullptr, *d_out = nullptr;
    CU(cudaMalloc((void **)&d_idx, n * 4)); CU(cudaMalloc((void **)&d_out, n * 4));
    CU(cudaMemcpyAsync(d_idx, sa_index, n * 4, cudaMemcpyHostToDevice, ix->stream));
    unsigned long long *cnt = nullptr;
    int rc = sa_launch(ix, d_idx, n, d_out, ix->stream, &cnt);
    if (rc == HSA_OK) {
        unsigned long long st = 0;
        CU(cudaMemcpyAsync(sa_value_out, d_out, n * 4, cudaMemcpyDeviceToHost, ix->stream));
        CU(cudaMemcpyAsync(&st, cnt + 1, sizeof(st), cudaMemcpyDeviceToHost, ix->stream));
        CU(cudaStreamSynchronize(ix->stream));
        if (steps_total) *steps_total = st;
    }
    cudaFree(d_idx); cudaFree(d_out);
    return rc;
}

extern "C" int hsa_sa_locate(const hsa_index_t *ix, const uint32_t *sa_index, size_t n, uint32_t *occ_pos_out, uint32_t *seq_id_out,
                             uint32_t *ori_pos_out)
{
    if (!ix || (n && (!sa_index || !occ_pos_out || !seq_id_out || !ori_pos_out))) return fail(HSA_E_ARG, "bad argument");
    if (!ix->sa_value) return fail(HSA_E_ARG, "no SA samples attached to this index (hsa_index_attach_sa)");
    if (!ix->blocks4) return fail(HSA_E_ARG, "no block list attached to this index (hsa_index_attach_blocks)");
    if (n == 0) return HSA_OK;
    for (size_t i = 0; i < n; ++i)
        if (sa_index[i] > ix->ix.fwd.text_length) return fail(HSA_E_ARG, "SA index beyond textLength");
    CU(cudaSetDevice(ix->device));
    uint32_t *d = nullptr;                               // {indices, occ_pos, seq_id, ori_pos}
    CU(cudaMalloc((void **)&d, n * 16));
    CU(cudaMemcpyAsync(d, sa_index, n * 4, cudaMemcpyHostToDevice, ix->stream));
    unsigned long long *cnt = nullptr;
    int rc = sa_launch(ix, d, n, d + n, ix->stream, &cnt, d + 2 * n, d + 3 * n);
    if (rc == HSA_OK) {
        CU(cudaMemcpyAsync(occ_pos_out, d + n, n * 4, cudaMemcpyDeviceToHost, ix->stream));
        CU(cudaMemcpyAsync(seq_id_out, d + 2 * n, n * 4, cudaMemcpyDeviceToHost, ix->stream));
        CU(cudaMemcpyAsync(ori_pos_out, d + 3 * n, n * 4, cudaMemcpyDeviceToHost, ix->stream));
        CU(cudaStreamSynchronize(ix->stream));
    }
    cudaFree(d);
    return rc;
}

extern "C" int hsa_sa_values_device(const hsa_index_t *ix, const uint32_t *sa_index_dev, size_t n, uint32_t *sa_value_out_dev,
                                    uint64_t *steps_total_dev, void *stream)
{
    if (!ix || (n && (!sa_index_dev || !sa_value_out_dev))) return fail(HSA_E_ARG, "bad argument");
    if (!ix->sa_value) return fail(HSA_E_ARG, "no SA samples attached to this index (hsa_index_attach_sa)");
    if (n == 0) return HSA_OK;
    CU(cudaSetDevice(ix->device));
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long *cnt = nullptr;
    int rc = sa_launch(ix, sa_index_dev, n, sa_value_out_dev, s, &cnt);
    if (rc) return rc;
    if (steps_total_dev) CU(cudaMemcpyAsync(steps_total_dev, cnt + 1, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s));
    return HSA_OK;
}

// ---------------------------------------------------------------------------------------------- splice fallback
extern "C" int hsa_index_attach_packed_dna(hsa_index_t *ix, const uint32_t *packed_dna, uint32_t dna_length)
{
    if (!ix || !packed_dna || dna_length == 0) return fail(HSA_E_ARG, "bad argument");
    CU(cudaSetDevice(ix->device));
    const size_t words = ((size_t)dna_length + 15) / 16 + 1;       // DNALoadPacked allocates one spare word (TextConverter.c:705)
    cudaFree(ix->packed_dna); ix->packed_dna = nullptr;
    CU(cudaMalloc((void **)&ix->packed_dna, words * 4));
    CU(cudaMemset(ix->packed_dna, 0, words * 4));
    CU(cudaMemcpy(ix->packed_dna, packed_dna, (words - 1) * 4, cudaMemcpyHostToDevice));
    ix->dna_length = dna_length;
    return HSA_OK;
}

namespace {
struct DevBuf {                    // cudaMalloc'd array, grown on demand, released on scope exit
    void *p = nullptr; size_t cap = 0;
    ~DevBuf() { cudaFree(p); }
    int alloc(size_t bytes)
    {
        if (bytes <= cap && p) return 0;
        cudaFree(p); p = nullptr; cap = 0;
        if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) { p = nullptr; return -1; }
        cap = bytes;
        return 0;
    }
    void release() { cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T *as() const { return static_cast<T *>(p); }
};
}
struct SpliceCache { DevBuf arena, heads, widths, lists, sites, pos, codes, off, len, opts, oi, n, aln, status, fail, cnt, again; };
static void splice_cache_free(SpliceCache *c) { delete c; }

// one pass of splice_kernel over `n_work` reads with per-worker scratch of the given capacities
static int splice_pass(const hsa_index_t *ix, SpliceCache &C, SpliceParams P, uint32_t n_work, const uint32_t *work_list, uint32_t max_workers,
                       uint32_t arena_cap, uint32_t aln_cap, uint32_t site_cap, unsigned long long *counters, cudaStream_t s)
{
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, splice_kernel, 128, 0));
    const uint32_t grid_full = (uint32_t)ix->sm_count * (uint32_t)std::max(occ, 1);
    // lanes per warp that work: all 32 when the batch has enough reads for every lane of a full grid, fewer (down to one)
    // for small batches, so that a read's latency is not multiplied by the divergence of its warp
    uint32_t lane_stride = 32;
    while (lane_stride > 1 && (uint64_t)grid_full * 128 / lane_stride < n_work) lane_stride >>= 1;
    uint32_t workers = std::min<uint32_t>(std::min<uint32_t>(grid_full * 128 / lane_stride, max_workers), std::max<uint32_t>(n_work, 1));
    const uint32_t per_block = 128 / lane_stride;
    const uint32_t grid = std::max<uint32_t>(1, (workers + per_block - 1) / per_block);
    workers = grid * per_block;
    const size_t wl = (size_t)P.max_len + 1;
    if (C.arena.alloc((size_t)workers * arena_cap * sizeof(SEntry)) || C.heads.alloc((size_t)workers * SPL_BUCKETS * 4) ||
        C.widths.alloc((size_t)workers * (3 * wl + 16) * sizeof(SWidth)) || C.lists.alloc((size_t)workers * 6 * aln_cap * sizeof(SAln)) ||
        C.sites.alloc((size_t)workers * site_cap * 4) || C.pos.alloc((size_t)workers * SPL_POS_CAP * sizeof(SPos)))
        return fail(HSA_E_CUDA, "out of device memory for the splice scratch");
    CU(cudaMemsetAsync(C.widths.p, 0, (size_t)workers * (3 * wl + 16) * sizeof(SWidth), s));     // the driver's calloc (bwtaln.c:283-285)
    P.arena = C.arena.as<SEntry>(); P.arena_cap = arena_cap; P.heads = C.heads.as<uint32_t>(); P.widths = C.widths.as<SWidth>();
    P.lists = C.lists.as<SAln>(); P.aln_cap = aln_cap; P.site_pos = C.sites.as<uint32_t>(); P.site_cap = site_cap; P.pos_info = C.pos.as<SPos>();
    P.n_work = n_work; P.work_list = work_list;
    P.cursor = counters; P.fail_count = counters + 1; P.lookups = counters + 2;
    CU(cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned long long), s));                 // cursor + fail count; lookups accumulate
    splice_kernel<<<grid, 128, 0, s>>>(P, lane_stride);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s));
    return HSA_OK;
}

extern "C" int hsa_splice_match_batch(const hsa_index_t *ix, const uint8_t *codes, const uint64_t *off, const uint32_t *len,
                                      size_t n_reads, const hsa_gap_opt_t *opts, size_t n_opts, const uint32_t *opt_idx,
                                      int32_t *n_aln_out, hsa_aln1_t *aln_out, uint64_t *occ_lookups)
{
    if (occ_lookups) *occ_lookups = 0;
    if (n_reads == 0) return HSA_OK;
    if (!ix || !codes || !off || !len || !opts || !n_opts || !n_aln_out || !aln_out) return fail(HSA_E_ARG, "null / empty argument");
    if (!ix->sa_value || !ix->blocks4 || !ix->packed_dna)
        return fail(HSA_E_ARG, "the splice path needs the SA samples, the block list and the packed text "
                               "(hsa_index_attach_sa / _blocks / _packed_dna)");
    if (n_reads > 0x7FFFFFF0ull) return fail(HSA_E_ARG, "too many reads in one batch");
    HostTick tk;
    size_t bytes = 0; uint32_t max_len = 0;
    for (size_t i = 0; i < n_reads; ++i) {
        if (len[i] < 36) return fail(HSA_E_ARG, "bwt_splice_match needs reads of at least 36 bases (three seeds, 12-base anchors)");
        if (len[i] > 4095) return fail(HSA_E_ARG, "reads longer than 4095 bases are not supported");
        if (opt_idx && opt_idx[i] >= n_opts) return fail(HSA_E_ARG, "opt_idx out of range");
        bytes = std::max(bytes, (size_t)off[i] + len[i]); max_len = std::max(max_len, len[i]);
    }
    std::vector<DevOpt> dopts(n_opts);
    for (size_t i = 0; i < n_opts; ++i) {
        const hsa_gap_opt_t &o = opts[i];
        if (o.s_mm < 0 || o.s_gapo < 0 || o.s_gape < 0 || o.max_diff < 0 || o.max_gapo < 0 || o.max_gape < 0 || o.max_seed_diff < 0)
            return fail(HSA_E_ARG, "negative scores / limits are not supported");
        const long top = (long)(std::max(o.max_diff, o.max_seed_diff) + 2) * o.s_mm + (long)(o.max_gapo + 2) * o.s_gapo +
                         (long)(std::max(o.max_gape, 3) + 2) * o.s_gape;
        if (top >= (long)SPL_BUCKETS) return fail(HSA_E_ARG, "score range exceeds the splice stack's 256 buckets");
        to_devopt(o, dopts[i]);
    }
    CU(cudaSetDevice(ix->device));
    cudaStream_t s = ix->stream;
    hsa_index *mix = const_cast<hsa_index *>(ix);
    if (!mix->splice_cache) mix->splice_cache = new SpliceCache();
    SpliceCache &C = *mix->splice_cache;
    DevBuf &d_codes = C.codes, &d_off = C.off, &d_len = C.len, &d_opts = C.opts, &d_oi = C.oi, &d_n = C.n, &d_aln = C.aln,
           &d_status = C.status, &d_fail = C.fail, &d_cnt = C.cnt;
    if (d_codes.alloc(bytes + 16) || d_off.alloc(n_reads * 8) || d_len.alloc(n_reads * 4) || d_opts.alloc(n_opts * sizeof(DevOpt)) ||
        d_oi.alloc(n_reads * 4) || d_n.alloc(n_reads * 4) || d_aln.alloc(n_reads * 18 * 4) || d_status.alloc(n_reads) ||
        d_fail.alloc((n_reads + 1) * 4) || d_cnt.alloc(4 * sizeof(unsigned long long)))
        return fail(HSA_E_CUDA, "out of device memory for the splice batch");
    CU(cudaMemcpyAsync(d_codes.p, codes, bytes, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(d_off.p, off, n_reads * 8, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(d_len.p, len, n_reads * 4, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(d_opts.p, dopts.data(), n_opts * sizeof(DevOpt), cudaMemcpyHostToDevice, s));
    if (opt_idx) CU(cudaMemcpyAsync(d_oi.p, opt_idx, n_reads * 4, cudaMemcpyHostToDevice, s));
    CU(cudaMemsetAsync(d_cnt.p, 0, 4 * sizeof(unsigned long long), s));
    if (!mix->splice_stack_set) { CU(cudaDeviceSetLimit(cudaLimitStackSize, 8192)); mix->splice_stack_set = true; }   // real calls in splice_kernel
    SpliceParams P;
    memset(&P, 0, sizeof(P));
    P.env.ix = ix->ix; P.env.sa_value = ix->sa_value; P.env.sa_interval = ix->sa_interval;
    P.env.blocks4 = ix->blocks4; P.env.n_blocks = ix->n_blocks; P.env.packed_dna = ix->packed_dna; P.env.dna_length = ix->dna_length;
    P.codes = d_codes.as<uint8_t>(); P.read_off = d_off.as<uint64_t>(); P.read_len = d_len.as<uint32_t>();
    P.opts = d_opts.as<DevOpt>(); P.opt_idx = opt_idx ? d_oi.as<uint32_t>() : nullptr; P.max_len = max_len;
    P.n_aln = d_n.as<int32_t>(); P.aln = d_aln.as<uint32_t>(); P.status = d_status.as<uint8_t>(); P.fail_list = d_fail.as<uint32_t>();
    unsigned long long *cnt = d_cnt.as<unsigned long long>();
    int rc;
    // first pass: every read, small slices for many workers; then the reads that outgrew them, large slices for few
    const uint32_t arena_cap = (uint32_t)std::max<long>(64, env_long("HSA_B200_SPLICE_ARENA", 2048));
    const uint32_t aln_cap = (uint32_t)std::max<long>(16, env_long("HSA_B200_SPLICE_ALNS", 128));
    tk.mark(3);
    if ((rc = splice_pass(ix, C, P, (uint32_t)n_reads, nullptr, 1u << 20, arena_cap, aln_cap, 512, cnt, s))) return rc;
    unsigned long long h[3];
    CU(cudaMemcpy(h, cnt, sizeof(h), cudaMemcpyDeviceToHost));
    tk.mark(4);
    if (h[1]) {
        const uint32_t n_fail = (uint32_t)h[1];
        DevBuf &again = C.again;
        if (again.alloc((size_t)n_fail * 4)) return fail(HSA_E_CUDA, "out of device memory");
        CU(cudaMemcpy(again.p, d_fail.p, (size_t)n_fail * 4, cudaMemcpyDeviceToDevice));
        // large slices for few workers (2 MB of stack + 1.2 MB of hit lists each); the arrays are not kept afterwards
        C.arena.release(); C.lists.release(); C.sites.release();
        rc = splice_pass(ix, C, P, n_fail, again.as<uint32_t>(), 4096, 1u << 16, 1u << 12, 1u << 13, cnt, s);
        C.arena.release(); C.lists.release(); C.sites.release();
        if (rc) return rc;
        CU(cudaMemcpy(h, cnt, sizeof(h), cudaMemcpyDeviceToHost));
        if (h[1]) return fail(HSA_E_CAPACITY, "a read exceeded the splice path's large-capacity scratch (65536 stack entries, 4096 hits per seed)");
    }
    tk.mark(5);
    CU(cudaMemcpyAsync(n_aln_out, d_n.p, n_reads * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(aln_out, d_aln.p, n_reads * 18 * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    tk.mark(6);
    if (occ_lookups) *occ_lookups = h[2];
    return HSA_OK;
}

// ---------------------------------------------------------------------------------------------- workspace
static int scratch_alloc(Scratch &s, uint32_t n_workers, uint32_t arena_cap, uint32_t hit_cap, uint32_t link_bytes)
{
    if (s.n_workers >= n_workers && s.arena_cap == arena_cap && s.hit_cap == hit_cap && s.link_bytes == link_bytes) return HSA_OK;
    s.release();
    CU(cudaMalloc((void **)&s.arena, (size_t)n_workers * arena_cap * 2 * sizeof(u32x4)));     // 32-byte slots: record + link
    CU(cudaMalloc((void **)&s.hits, (size_t)n_workers * hit_cap * sizeof(Hit)));
    s.n_workers = n_workers; s.arena_cap = arena_cap; s.hit_cap = hit_cap; s.link_bytes = link_bytes;
    return HSA_OK;
}

extern "C" int hsa_workspace_create(const hsa_index_t *ix, size_t max_reads, uint32_t max_len, size_t aln_capacity,
                                    hsa_workspace_t **out)
{
    (void)max_reads; (void)aln_capacity; (void)max_len;
    if (!ix || !out) return fail(HSA_E_ARG, "null argument");
    CU(cudaSetDevice(ix->device));
    hsa_workspace *ws = new hsa_workspace();
    ws->idx = ix;
    CU(cudaMalloc((void **)&ws->counters, CNT_ALLOC * sizeof(unsigned long long)));
    CU(cudaEventCreateWithFlags(&ws->ev0, cudaEventDefault)); CU(cudaEventCreateWithFlags(&ws->ev1, cudaEventDefault));
    *out = ws;
    return HSA_OK;
}

extern "C" void hsa_workspace_free(hsa_workspace_t *ws)
{
    if (!ws) return;
    for (Pipe &p : ws->pipes) p.release();
    ws->strict.release(); ws->heavy.release();
    cudaFree(ws->counters); cudaFree(ws->strict_list); cudaFree(ws->strict2_list); cudaFree(ws->opts_dev); cudaFree(ws->len2opt_dev);
    cudaFree(ws->status_dev); cudaFree(ws->codes_dev); cudaFree(ws->off_dev); cudaFree(ws->len_dev); cudaFree(ws->bin_list); cudaFree(ws->bin_hist);
    cudaFree(ws->tasks_dev); cudaFree(ws->n_aln_dev); cudaFree(ws->aln_off_dev); cudaFree(ws->aln_dev);
    cudaFree(ws->width_out_dev); cudaFree(ws->bid_dev);
    for (int i = 0; i < hsa_workspace::OPT_STAGES; ++i) {
        cudaFreeHost(ws->opts_stage[i]); cudaFreeHost(ws->l2o_stage[i]);
        if (ws->stage_ev[i]) cudaEventDestroy(ws->stage_ev[i]);
    }
    for (cudaEvent_t e : ws->trace_ev) cudaEventDestroy(e);
    if (ws->ev0) cudaEventDestroy(ws->ev0);
    if (ws->ev1) cudaEventDestroy(ws->ev1);
    delete ws;
}

extern "C" uint32_t hsa_workspace_last_launches(const hsa_workspace_t *ws) { return ws ? ws->last_launches : 0; }
extern "C" void hsa_workspace_last_config(const hsa_workspace_t *ws, uint32_t out[4])
{
    if (!out) return;
    for (int i = 0; i < 4; ++i) out[i] = ws ? ws->last_config[i] : 0;
}

// Per-launch timing of the workspace's next calls (bench.py: the dominant kernel's own duration and algorithmic
// bytes).  enable: one CUDA event is recorded behind every kernel launch; read: waits for the device, returns for the
// last call the launches' names ("width1;search1;...") and durations, and the occ lookups of the searches the
// per-lane search kernel launches completed.
extern "C" int hsa_workspace_launch_timing(hsa_workspace_t *ws, int enable)
{
    if (!ws) return fail(HSA_E_ARG, "bad argument");
    for (cudaEvent_t e : ws->trace_ev) cudaEventDestroy(e);
    ws->trace_ev.clear(); ws->trace_name.clear();
    ws->trace = enable ? 2 : 0;
    return HSA_OK;
}

extern "C" int hsa_workspace_launch_times(hsa_workspace_t *ws, char *names, size_t names_cap, float *ms, size_t ms_cap,
                                          size_t *n_out, uint64_t *fast_search_lookups)
{
    if (!ws || !names || !ms || !n_out) return fail(HSA_E_ARG, "bad argument");
    CU(cudaSetDevice(ws->idx->device));
    CU(cudaDeviceSynchronize());
    std::string nm;
    size_t n = 0;
    cudaEvent_t prev = ws->ev0;
    for (size_t i = 0; i < ws->trace_ev.size() && n < ms_cap; ++i, ++n) {
        float t = 0;
        CU(cudaEventElapsedTime(&t, prev, ws->trace_ev[i]));
        ms[n] = t; prev = ws->trace_ev[i];
        if (i) nm += ";";
        nm += ws->trace_name[i];
    }
    if (nm.size() + 1 > names_cap) return fail(HSA_E_ARG, "names buffer too small");
    memcpy(names, nm.c_str(), nm.size() + 1);
    *n_out = n;
    if (fast_search_lookups) {
        unsigned long long v = 0;
        CU(cudaMemcpy(&v, ws->counters + CNT_DIAG_FAST_LOOKUPS, sizeof(v), cudaMemcpyDeviceToHost));
        *fast_search_lookups = v;
    }
    return HSA_OK;
}

// ---------------------------------------------------------------------------------------------- launch
// Kernel variants.  FAST: 16-bit link halves (<= 1022 records per worker, 64 score buckets), bound bytes in shared memory
// (FAST_ROWS: in the rows, for reads too long for shared memory).  LARGE: 32-bit halves, 64-thread blocks.
enum Variant { V_FAST = 0, V_FAST_ROWS = 1, V_LARGE = 2, V_COOP = 3 };

struct Batch;
static bool dense_fast(const hsa_workspace *ws, const Batch &b, uint32_t seed_cap);

static const void *search_fn(Variant v, int block, int minb)
{
    if (v == V_COOP) return (const void *)coop_kernel<128>;
    if (v == V_LARGE) return (const void *)search_kernel<64, 1, uint64_t, false>;
    if (v == V_FAST_ROWS) {
        if (block != 128) return (const void *)search_kernel<256, 2, uint32_t, false>;
        switch (minb) {
        case 5: return (const void *)search_kernel<128, 5, uint32_t, false>;
        case 6: return (const void *)search_kernel<128, 6, uint32_t, false>;
        default: return (const void *)search_kernel<128, 4, uint32_t, false>;
        }
    }
    if (block == 128) {
        switch (minb) {
        case 3: return (const void *)search_kernel<128, 3, uint32_t, true>;
        case 4: return (const void *)search_kernel<128, 4, uint32_t, true>;
        case 6: return (const void *)search_kernel<128, 6, uint32_t, true>;
        case 8: return (const void *)search_kernel<128, 8, uint32_t, true>;
        default: return (const void *)search_kernel<128, 5, uint32_t, true>;
        }
    }
    switch (minb) {
    case 3: return (const void *)search_kernel<256, 3, uint32_t, true>;
    case 4: return (const void *)search_kernel<256, 4, uint32_t, true>;
    default: return (const void *)search_kernel<256, 2, uint32_t, true>;
    }
}

struct Batch {                      // everything one batch needs, device pointers
    uint32_t kind = 0, n_groups = 0, n_items = 0, max_len = 0, n_opts = 0, n_buckets = 1, max_seed_len = 0;
    int32_t filter_max_n = 0;
    bool scores_positive = true;    // every option set has s_mm, s_gapo, s_gape >= 1 (the cooperative stage needs that)
    bool ragged = false;            // KIND_WHOLE: more than one read length in the batch -> pre-bin the work order
    const uint8_t *codes = nullptr; const Task *tasks = nullptr;
    const uint64_t *read_off = nullptr; const uint32_t *read_len = nullptr;
    int32_t *n_aln = nullptr; uint64_t *aln_off = nullptr; uint32_t *aln = nullptr; uint64_t aln_cap = 0;
    u32x2 *width_out = nullptr; int32_t *bid_out = nullptr;
    const uint8_t *rows_host = nullptr; size_t rows_host_bytes = 0;   // per-call form: the caller's widths as one ready-made row
};

// Six resident blocks per SM instead of five (24 warps instead of 20 waiting on their lookups) for the batches that allow it.
// The fast kernel is bound by the latency of its dependent lookups times the warps in flight (ncu: 50 % issue-active, 4 long-
// scoreboard stall cycles per issue at 3.1 Gb); a sixth block needs <= 85 registers (launch bound 6: no spills) and a shared-
// memory share that leaves the SM >= 60 KB of L1 -- below that the random-sector rate halves (tools/probe_sweep.cu, carve-out
// sweep).  40 score buckets per lane instead of 64 (80 instead of 128 bytes of bucket heads) give 31.3 KB per block: 188 KB for
// six.  Searches that push a record of score >= 40 are handed to the cooperative kernel like those that reached 64 before
// (+22 % of an 0.9 % share with the default options).  Measured (profiles/r02_occupancy_buckets.log): search kernels -7.3 % at
// 3.1 Gb, -8.5 % at 46 Mb, seed searches +7 %; 48 buckets fall off the L1 cliff (+30 %), eight blocks (64 registers, spills) +50 %.
// Not for small batches (one read per lane: nothing to overlap, 100 000 x 75 bp: +6 % time), nor for option sets whose scores go
// far beyond 64 anyway (the stress configuration hands 16 % more searches on and loses 3 %), nor for reads so long that six blocks'
// bound bytes no longer fit the 196 KB carve-out (> 111 bases with a seed region).
enum : long { DENSE_NB_FAST = 40 };
static bool dense_fast(const hsa_workspace *ws, const Batch &b, uint32_t seed_cap)
{
    if (!ws->minb_auto || ws->nb_fast_env >= 0 || ws->block != 128) return false;
    const uint64_t n_work = (uint64_t)b.n_groups * (b.kind == KIND_SEEDS ? 6u : 1u);
    // n_buckets = highest score an option set can reach + 1 (check_opt): 69 with the default options, 60 for -n 2 -o 1, 80 for the
    // stress configuration
    if (b.kind == KIND_WIDTH || b.n_buckets > 72 || n_work < 400000) return false;
    // six blocks (+ 1 KB each that the system reserves) must fit the 196 KB carve-out, the largest that leaves 60 KB of L1:
    // reads of up to 111 bases with a 32-base seed region
    Params T;
    memset(&T, 0, sizeof(T));
    set_layout(T, b.max_len, seed_cap, std::min<uint32_t>(b.n_buckets, (uint32_t)DENSE_NB_FAST), b.n_opts, 2, true);
    const size_t per_block = (size_t)T.smem_opts_bytes + 128u * (size_t)T.smem_lane_stride + 8 + 5 * sizeof(unsigned long long) + 1024;
    return 6 * per_block <= 196u * 1024u;
}

static void trace_mark(hsa_workspace *ws, const char *name, cudaStream_t stream)
{
    if (ws->trace <= 0) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, stream);
    ws->trace_ev.push_back(e); ws->trace_name.push_back(name);
}

static void trace_dump(hsa_workspace *ws, const unsigned long long *cnt_all)
{
    if (ws->trace <= 0 || ws->trace == 2) return;            // 2: kept for hsa_workspace_launch_times
    fprintf(stderr, "[hsa_b200 trace]");
    for (size_t i = 0; i < ws->trace_ev.size(); ++i) {
        float t = 0;
        cudaEventElapsedTime(&t, ws->ev0, ws->trace_ev[i]);     // completion time since the batch started
        fprintf(stderr, " %s@%.2f", ws->trace_name[i], t);
        cudaEventDestroy(ws->trace_ev[i]);
    }
    ws->trace_ev.clear(); ws->trace_name.clear();
    if (cnt_all)
        fprintf(stderr, " | steps=%llu warp_iters=%llu lanes_per_step=%.2f max_item_steps=%llu pops=%llu lookups=%llu",
                cnt_all[CNT_STEPS], cnt_all[CNT_DIAG_WARP_ITERS],
                cnt_all[CNT_DIAG_WARP_ITERS] ? (double)cnt_all[CNT_STEPS] / (double)cnt_all[CNT_DIAG_WARP_ITERS] : 0.0,
                cnt_all[CNT_DIAG_MAX_ITEM_STEPS], cnt_all[CNT_POPS], cnt_all[CNT_LOOKUPS]);
    if (cnt_all && cnt_all[CNT_DIAG_COOP_WAVES])
        fprintf(stderr, " | coop: heavy=%llu left=%llu waves=%llu wave_steps=%llu lane_steps=%llu lanes/wave_step=%.2f",
                cnt_all[CNT_STRICT], cnt_all[CNT_STRICT2], cnt_all[CNT_DIAG_COOP_WAVES], cnt_all[CNT_DIAG_COOP_WAVES + 1],
                cnt_all[CNT_DIAG_COOP_WAVES + 2],
                cnt_all[CNT_DIAG_COOP_WAVES + 1] ? (double)cnt_all[CNT_DIAG_COOP_WAVES + 2] / cnt_all[CNT_DIAG_COOP_WAVES + 1] : 0.0);
#ifdef HSA_PHASE_PROF
    if (cnt_all && cnt_all[CNT_PROF0 + 11])
        fprintf(stderr, " | coop cycles: chain=%.3g other=%.3g (chain share %.2f)", (double)cnt_all[CNT_PROF0 + 9],
                (double)cnt_all[CNT_PROF0 + 10], (double)cnt_all[CNT_PROF0 + 9] / (double)cnt_all[CNT_PROF0 + 11]);
    if (cnt_all) {
        static const char *nm[3] = {"SLOW", "LOOKUP", "POP"};
        for (int i = 0; i < 3; ++i) {
            const unsigned long long cyc = cnt_all[CNT_PROF0 + 3 * i], runs = cnt_all[CNT_PROF0 + 3 * i + 1], ln = cnt_all[CNT_PROF0 + 3 * i + 2];
            fprintf(stderr, " | %s runs=%llu cyc/run=%.0f lanes/run=%.2f", nm[i], runs, runs ? (double)cyc / runs : 0.0, runs ? (double)ln / runs : 0.0);
        }
    }
#endif
    fprintf(stderr, "\n");
}

static int configure(hsa_workspace *ws)
{
    if (ws->configured) return HSA_OK;
    ws->block = (uint32_t)env_long("HSA_B200_BLOCK", 128);
    if (ws->block != 128) ws->block = 256;
    ws->minb = (int)env_long("HSA_B200_MINB", ws->block == 128 ? 5 : 2);
    ws->minb_auto = getenv("HSA_B200_MINB") == nullptr;
    ws->nb_fast_env = env_long("HSA_B200_NB_FAST", -1);
    ws->blocks_per_sm_cap = (int)env_long("HSA_B200_BLOCKS_PER_SM", 0);
    ws->n_pipes = (uint32_t)std::min<long>(MAX_PIPES, std::max<long>(1, env_long("HSA_B200_PIPES", 1)));
    ws->chunk_items = (uint64_t)std::max<long>(6 * 1024, env_long("HSA_B200_CHUNK", 12 << 20));
    ws->chunk_items -= ws->chunk_items % 6;
    ws->arena_cap = (uint32_t)std::min<long>(1022, std::max<long>(16, env_long("HSA_B200_ARENA_CAP", 1022)));   // 10-bit slot ids
    ws->hit_cap = (uint32_t)std::max<long>(1, env_long("HSA_B200_HIT_CAP", 32));
    ws->vote_slow_min = (uint32_t)std::max<long>(1, env_long("HSA_B200_SLOW_MIN", VOTE_SLOW_MIN_DEFAULT));
    ws->vote_pop_bias = (int32_t)env_long("HSA_B200_POP_BIAS", VOTE_POP_BIAS_DEFAULT);
    ws->use_coop = env_long("HSA_B200_COOP", 1) != 0;
    ws->step_budget = (uint32_t)std::max<long>(0, env_long("HSA_B200_STEP_BUDGET", 0));
    ws->drain_budget = (uint32_t)std::max<long>(0, env_long("HSA_B200_DRAIN_BUDGET", 1000));
    ws->prebin = env_long("HSA_B200_PREBIN", 1) != 0;
    ws->configured = true;
    return HSA_OK;
}

// Where a stage of the pipeline takes its work from and where it sends what it cannot hold.
struct StageIO {
    const uint32_t *work_list = nullptr;     // work index -> item id (nullptr: work_base + index)
    uint32_t work_base = 0;
    uint32_t n_work = 0;                      // work items, or the cap of a device-side count
    const uint32_t *n_work_dev = nullptr;     // device-side count (low word of a counter)
    uint32_t n_work_skip = 0;                 // ... of which the first n_work_skip belong to earlier rounds
    uint32_t *flag_list = nullptr;            // items handed on to the next stage ...
    unsigned long long *flag_count = nullptr; // ... and their count
};

static int coop_scratch_alloc(Pipe &p, uint32_t warps, uint32_t cap_chunks)
{
    if (p.coop_warps >= warps && p.coop_cap_chunks == cap_chunks) return HSA_OK;
    cudaFree(p.coop_payload); cudaFree(p.coop_out_payload); cudaFree(p.coop_info); cudaFree(p.coop_prev);
    cudaFree(p.coop_out_info); cudaFree(p.coop_hits);
    p.coop_payload = p.coop_out_payload = nullptr; p.coop_info = p.coop_prev = p.coop_out_info = nullptr; p.coop_hits = nullptr;
    p.coop_warps = 0;
    const size_t ent = (size_t)warps * cap_chunks * COOP_CHUNK, out = (size_t)warps * 32 * COOP_OUT_CAP;
    CU(cudaMalloc((void **)&p.coop_payload, ent * sizeof(u32x4)));
    CU(cudaMalloc((void **)&p.coop_info, ent * 4));
    CU(cudaMalloc((void **)&p.coop_prev, (size_t)warps * cap_chunks * 4));
    CU(cudaMalloc((void **)&p.coop_out_payload, out * sizeof(u32x4)));
    CU(cudaMalloc((void **)&p.coop_out_info, out * 4));
    CU(cudaMalloc((void **)&p.coop_hits, (size_t)warps * COOP_HIT_CAP * sizeof(Hit)));
    p.coop_warps = warps; p.coop_cap_chunks = cap_chunks;
    return HSA_OK;
}

// One chunk of work items through the split pipeline on `stream`:
//   width(pass 1) -> search(pass 1) [-> width(pass 2) -> search(pass 2)]
// Pass 2 exists only for whole reads: the reads whose reverse-complement strand found nothing are appended to a
// device-side list by the pass-1 search kernel, and the pass-2 kernels read the list length from device memory,
// so no host synchronisation separates the four launches.  `v` picks the search kernel: the fast per-lane one,
// the warp-cooperative one (heavy searches) or the large-capacity per-lane one (last resort).
static int issue_chunk(hsa_workspace *ws, const Batch &b, Params P, Pipe &pipe, int pipe_slot, Variant v,
                       const StageIO &io, cudaStream_t stream)
{
    const hsa_index *ix = ws->idx;
    const bool large = v == V_LARGE, coop = v == V_COOP;
    const uint32_t block = large ? 64u : coop ? 128u : ws->block;
    if (!coop) P.smem_stats_off = (uint32_t)(((size_t)P.smem_opts_bytes + (size_t)block * P.smem_lane_stride + 7) & ~size_t(7));
    const size_t smem = coop ? (size_t)P.smem_opts_bytes + (size_t)(block / 32) * P.coop_warp_smem
                             : (size_t)P.smem_stats_off + 5 * sizeof(unsigned long long);
    const void *fn = search_fn(v, (int)block, (v == V_FAST && ws->dense_now) ? 6 : ws->minb);
    CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (v <= V_FAST_ROWS) {
        ws->last_config[0] = (uint32_t)((v == V_FAST && ws->dense_now) ? 6 : ws->minb); ws->last_config[1] = P.n_buckets;
        ws->last_config[2] = (uint32_t)smem; ws->last_config[3] = v == V_FAST ? 1u : 0u;
    }
    {   // experiment knob (north_star: "L2 persisting-access window"): HSA_B200_L2_PERSIST=1 marks the forward direction's blocks as
        // persisting in L2 for the kernels of this stream when they fit the device's persisting carve-out (the 46 Mb index: 23 MB)
        static const long persist = env_long("HSA_B200_L2_PERSIST", 0);
        if (persist) {
            cudaDeviceProp prop;
            CU(cudaGetDeviceProperties(&prop, ix->device));
            const size_t bytes = (size_t)ix->ix.fwd.n_blocks * 32;
            if (bytes <= (size_t)prop.persistingL2CacheMaxSize && bytes <= (size_t)prop.accessPolicyMaxWindowSize) {
                CU(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, bytes + (8u << 20))));
                cudaStreamAttrValue av;
                memset(&av, 0, sizeof(av));
                av.accessPolicyWindow.base_ptr = const_cast<u32x4 *>(ix->ix.fwd.blocks);
                av.accessPolicyWindow.num_bytes = bytes;
                av.accessPolicyWindow.hitRatio = 1.0f;
                av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                CU(cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &av));
            }
        }
    }
    {   // experiment knob: shared-memory carve-out in percent of the SM's unified L1/shared array (-1 = driver default)
        static const long carve = env_long("HSA_B200_CARVEOUT", -1);
        if (carve >= 0 && v <= V_FAST_ROWS) CU(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, (int)carve));
        // cooperative kernel: 8 blocks x 25 KB of shared memory take the largest carve-out and leave 28 KB of L1, the size at which
        // the random-sector rate halves.  With the index in HBM a 164 KB carve-out (six blocks, 92 KB of L1) makes the stage 10 %
        // faster (2 x 32.2 -> 2 x 28.8 ms per 12.5 M-read batch at 3.1 Gb); with the index in L2 the two extra blocks are worth more
        // than the L1 (stress configuration 13.05 vs 12.78 M reads/s), so the default stays there (profiles/r02_coop_carveout.log).
        static const long coop_carve_env = env_long("HSA_B200_COOP_CARVEOUT", -2);
        const long coop_carve = coop_carve_env >= -1 ? coop_carve_env : ((size_t)ix->ix.fwd.n_blocks * 32 > (96u << 20) ? 72 : -1);
        if (v == V_COOP) CU(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, (int)coop_carve));
    }
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, (int)block, smem));
    if (occ < 1) return fail(HSA_E_CUDA, "search kernel does not fit on an SM");
    if (v <= V_FAST_ROWS && ws->blocks_per_sm_cap > 0 && occ > ws->blocks_per_sm_cap) occ = ws->blocks_per_sm_cap;
    uint32_t grid_full = (uint32_t)ix->sm_count * (uint32_t)occ;
    if (large) grid_full = std::min<uint32_t>(grid_full, 32);           // 2048 workers x ~6.4 MB of stack each
    const uint32_t n_work = io.n_work;
    const uint32_t per_block = coop ? block / 32 : block;               // searches in flight per block
    const uint32_t grid = std::max<uint32_t>(1, std::min<uint32_t>(grid_full, (n_work + per_block - 1) / per_block));
    int rc;
    if (coop) {
        if ((rc = coop_scratch_alloc(pipe, grid_full * (block / 32), (uint32_t)env_long("HSA_B200_COOP_CHUNKS", 2048)))) return rc;
    } else if ((rc = scratch_alloc(pipe.sc, grid_full * block, large ? (1u << 18) : ws->arena_cap, large ? 4096u : ws->hit_cap,
                                   large ? 8u : 4u))) return rc;
    if ((rc = ensure(pipe.rows, pipe.rows_cap, std::max((size_t)n_work * P.row_stride, b.rows_host_bytes)))) return rc;
    if (b.kind == KIND_WHOLE && (rc = ensure(pipe.next_list, pipe.next_cap, (size_t)n_work + 1))) return rc;

    unsigned long long *slots = ws->counters + CNT_PIPE0 + 4 * pipe_slot;
    CU(cudaMemsetAsync(slots, 0, 4 * sizeof(unsigned long long), stream));
    P.rows = pipe.rows;
    P.arena = pipe.sc.arena; P.arena_cap = pipe.sc.arena_cap;
    P.hits = pipe.sc.hits; P.hit_cap = pipe.sc.hit_cap;
    P.coop_payload = pipe.coop_payload; P.coop_info = pipe.coop_info; P.coop_prev = pipe.coop_prev;
    P.coop_out_payload = pipe.coop_out_payload; P.coop_out_info = pipe.coop_out_info; P.coop_hits = pipe.coop_hits;
    P.coop_cap_chunks = pipe.coop_cap_chunks;
    P.strict_list = io.flag_list; P.strict_count = io.flag_count;
    if (v > V_FAST_ROWS) { P.step_budget = 0; P.drain_budget = 0; }
    P.pass = 1; P.work_list = io.work_list; P.work_base = io.work_base; P.n_work = n_work; P.n_work_dev = io.n_work_dev;
    P.n_work_skip = io.n_work_skip;
    P.next_list = pipe.next_list; P.next_count = reinterpret_cast<uint32_t *>(slots + 2);
    P.cursor = slots;
    // The width pass is a chain of dependent lookups per read and nothing else.  Measured (profiles/r02_width_ab.log): with the index
    // in L2 (46 Mb) a grid of exactly one wave of resident blocks is 14 % faster than the first version's 8 blocks per SM (the loop
    // is grid-strided with equal shares, so 1 184 blocks ran as a full wave of 740 and a second one of 444); with the index in HBM
    // (3.1 Gb) the pass sits at the address-translation limit, the second wave costs nothing and more resident warps only hurt
    // (launch bound 6 = 48 warps per SM: +30 % time), so the grid stays as it was there.
    // HSA_B200_WIDTH_MINB=6 / HSA_B200_WIDTH_GRID=<blocks per SM> select the other forms.
    static const long width_minb = env_long("HSA_B200_WIDTH_MINB", 5), width_grid = env_long("HSA_B200_WIDTH_GRID", 0);
    void (*wfn)(Params) = width_minb == 6 ? width_kernel<6> : width_kernel<5>;
    int wocc = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&wocc, wfn, 256, P.smem_opts_bytes));
    const bool index_in_l2 = (size_t)ix->ix.rev.n_blocks * 32 <= (96u << 20);
    const uint32_t wblocks = (uint32_t)ix->sm_count * (uint32_t)(width_grid > 0 ? width_grid : index_in_l2 ? std::max(wocc, 1) : 8);
    const uint32_t wgrid = std::max<uint32_t>(1, std::min<uint32_t>((n_work + 255) / 256, wblocks));
    void *args[] = {(void *)&P};
    const char *nm1 = large ? "search1L" : coop ? "search1C" : "search1", *nm2 = large ? "search2L" : coop ? "search2C" : "search2";
    if (b.rows_host) {
        // hsa_match_gap_call: the caller computed the widths (they are arguments of bwt_match_gap): no width pass
        CU(cudaMemcpyAsync(pipe.rows, b.rows_host, b.rows_host_bytes, cudaMemcpyHostToDevice, stream));
    } else {
        wfn<<<wgrid, 256, P.smem_opts_bytes, stream>>>(P);
        CU(cudaGetLastError());
        ++ws->last_launches; trace_mark(ws, "width1", stream);
    }
    if (b.kind == KIND_WIDTH) return HSA_OK;
    CU(cudaLaunchKernel(fn, dim3(grid), dim3(block), args, smem, stream));
    ++ws->last_launches; trace_mark(ws, nm1, stream);
    if (b.kind == KIND_WHOLE) {
        P.pass = 2; P.work_list = pipe.next_list; P.work_base = 0; P.n_work_dev = P.next_count; P.n_work_skip = 0;
        P.next_list = nullptr; P.next_count = nullptr; P.cursor = slots + 1;
        wfn<<<wgrid, 256, P.smem_opts_bytes, stream>>>(P);
        CU(cudaGetLastError());
        ++ws->last_launches; trace_mark(ws, "width2", stream);
        CU(cudaLaunchKernel(fn, dim3(grid), dim3(block), args, smem, stream));
        ++ws->last_launches; trace_mark(ws, nm2, stream);
    }
    return HSA_OK;
}

// The warp-cooperative stage needs positive scores (a chain must only push into higher buckets) and reads short
// enough for a warp's shared-memory block; otherwise heavy searches go straight to the large-capacity kernel.
static bool coop_usable(hsa_workspace *ws, const Batch &b)
{
    if (!ws->use_coop || b.kind == KIND_WIDTH || !b.scores_positive) return false;
    return b.max_len <= 2000;
}

// layout of the cooperative stage: rows as the large-capacity kernel reads them, one CoopWarp + bound bytes per warp
static void coop_layout(Params &S, const Batch &b, uint32_t seed_cap)
{
    set_layout(S, b.max_len, seed_cap, std::min<uint32_t>(b.n_buckets, COOP_NB), b.n_opts, 4, false);
    S.coop_warp_smem = (uint32_t)(((sizeof(CoopWarp) + 15) & ~size_t(15)) + (S.row_tail_off - S.row_bid_off) + 16) & ~15u;
}

// Launch parameters of a batch in the fast configuration (shared by the enqueue and the finish half).
static int batch_params(hsa_workspace *ws, const Batch &b, Params &P, Variant &v, uint32_t &seed_cap)
{
    const hsa_index *ix = ws->idx;
    seed_cap = (b.kind == KIND_TASKS || b.kind == KIND_WHOLE) && b.max_seed_len ? b.max_seed_len + 1 : 0;
    memset(&P, 0, sizeof(P));
    P.ix = ix->ix; P.codes = b.codes; P.kind = b.kind;
    P.tasks = b.tasks; P.read_off = b.read_off; P.read_len = b.read_len;
    P.opts = ws->opts_dev; P.len2opt = ws->len2opt_dev; P.filter_max_n = b.filter_max_n;
    P.n_aln = b.n_aln; P.aln_off = b.aln_off; P.status = ws->status_dev; P.aln = b.aln; P.aln_cap = b.aln_cap;
    P.counters = ws->counters; P.strict_list = ws->strict_list;
    P.width_out = b.width_out; P.bid_out = b.bid_out;
    P.vote_slow_min = ws->vote_slow_min; P.vote_pop_bias = ws->vote_pop_bias;
    P.step_budget = ws->step_budget; P.drain_budget = ws->drain_budget;
    // fast configuration: bound bytes in shared memory if a block's share leaves room for >= 4 blocks per SM
    ws->dense_now = dense_fast(ws, b, seed_cap);        // issue_chunk, called for this batch right after, picks the launch bound by it
    const long nb_want = ws->nb_fast_env >= 0 ? ws->nb_fast_env : ws->dense_now ? DENSE_NB_FAST : 64;
    const uint32_t nb_fast = std::min<uint32_t>(b.n_buckets, (uint32_t)std::min<long>(64, std::max<long>(8, nb_want)));   // scores >= nb_fast send the item on
    set_layout(P, b.max_len, seed_cap, nb_fast, b.n_opts, 2, true);
    v = V_FAST;
    if ((size_t)P.smem_opts_bytes + (size_t)ws->block * P.smem_lane_stride > 56 * 1024 || env_long("HSA_B200_FORCE_ROWS", 0)) {
        v = V_FAST_ROWS;
        set_layout(P, b.max_len, seed_cap, nb_fast, b.n_opts, 2, false);
    }
    if ((size_t)P.smem_opts_bytes + (size_t)ws->block * P.smem_lane_stride > 200 * 1024)
        return fail(HSA_E_ARG, "option table too large for shared memory");
    return HSA_OK;
}

// Enqueue half of a batch whose inputs/outputs are already on the device: nothing is synchronised.
// The batch is cut into chunks of work items that are issued round-robin on `n_pipes` internal streams
// (HSA_B200_PIPES / HSA_B200_CHUNK; default: one chunk on the caller's stream).
static int batch_enqueue(hsa_workspace *ws, const Batch &b, cudaStream_t stream, bool allow_trace)
{
    const hsa_index *ix = ws->idx;
    CU(cudaSetDevice(ix->device));
    int rc;
    if ((rc = configure(ws))) return rc;
    ws->last_launches = 0;
    if (ws->trace < 0) ws->trace = (int)env_long("HSA_B200_TRACE", 0);
    if (ws->trace == 2) { for (cudaEvent_t e : ws->trace_ev) cudaEventDestroy(e); ws->trace_ev.clear(); ws->trace_name.clear(); }
    const int trace_saved = ws->trace;
    if (!allow_trace && ws->trace != 2) ws->trace = 0;
    const uint32_t items_per_group = b.kind == KIND_SEEDS ? 6u : 1u;
    const uint64_t n_work_total = (uint64_t)b.n_groups * items_per_group;
    if ((rc = ensure(ws->status_dev, ws->status_cap, (size_t)b.n_items + 1))) return rc;
    if ((rc = ensure(ws->strict_list, ws->strict_list_cap, (size_t)n_work_total + 1))) return rc;
    Params P; Variant v; uint32_t seed_cap;
    if ((rc = batch_params(ws, b, P, v, seed_cap))) return rc;

    CU(cudaMemsetAsync(ws->counters, 0, CNT_ALLOC * sizeof(unsigned long long), stream));
    CU(cudaEventRecord(ws->ev0, stream));
    const uint32_t *binned = nullptr;
    if (b.kind == KIND_WHOLE && b.ragged && ws->prebin && b.n_items > 1) {
        if ((rc = ensure(ws->bin_list, ws->bin_list_cap, (size_t)b.n_items))) return rc;
        if (!ws->bin_hist) CU(cudaMalloc((void **)&ws->bin_hist, BIN_LENS * sizeof(uint32_t)));
        CU(cudaMemsetAsync(ws->bin_hist, 0, BIN_LENS * sizeof(uint32_t), stream));
        const unsigned g = (unsigned)std::min<uint64_t>((uint64_t)ix->sm_count * 4, ((uint64_t)b.n_items + 255) / 256);
        bin_count_kernel<<<g, 256, 0, stream>>>(b.read_len, b.n_items, ws->bin_hist);
        bin_scan_kernel<<<1, 1024, 0, stream>>>(ws->bin_hist);
        bin_scatter_kernel<<<g, 256, 0, stream>>>(b.read_len, b.n_items, ws->bin_hist, ws->bin_list);
        CU(cudaGetLastError());
        ws->last_launches += 3; trace_mark(ws, "prebin", stream);
        binned = ws->bin_list;
    }
    const uint64_t chunk = std::min<uint64_t>(ws->chunk_items, std::max<uint64_t>(n_work_total, 1));
    const uint32_t n_chunks = (uint32_t)((n_work_total + chunk - 1) / chunk);
    const uint32_t n_pipes = std::min<uint32_t>(ws->n_pipes, std::max<uint32_t>(n_chunks, 1));
    if (n_pipes > 1) {
        for (uint32_t p = 0; p < n_pipes; ++p) {
            Pipe &pp = ws->pipes[p];
            if (!pp.stream) CU(cudaStreamCreateWithFlags(&pp.stream, cudaStreamNonBlocking));
            if (!pp.done) CU(cudaEventCreateWithFlags(&pp.done, cudaEventDisableTiming));
            CU(cudaStreamWaitEvent(pp.stream, ws->ev0, 0));
        }
    }
    uint32_t c = 0;
    for (uint64_t w0 = 0; w0 < n_work_total; w0 += chunk, ++c) {
        const uint32_t nw = (uint32_t)std::min<uint64_t>(chunk, n_work_total - w0);
        Pipe &pp = ws->pipes[c % n_pipes];
        cudaStream_t s = n_pipes > 1 ? pp.stream : stream;
        StageIO io;
        io.work_base = (uint32_t)w0; io.n_work = nw;
        if (binned) io.work_list = binned + w0;               // work index -> read, longest reads first
        io.flag_list = ws->strict_list; io.flag_count = ws->counters + CNT_STRICT;
        if ((rc = issue_chunk(ws, b, P, pp, (int)(c % n_pipes), v, io, s))) return rc;
    }
    if (n_pipes > 1) {
        for (uint32_t p = 0; p < n_pipes; ++p) {
            CU(cudaEventRecord(ws->pipes[p].done, ws->pipes[p].stream));
            CU(cudaStreamWaitEvent(stream, ws->pipes[p].done, 0));
        }
    }
    // the heavy searches the fast kernel handed on (stack / hit capacity, scores >= 64, step budgets) go through the
    // warp-cooperative kernel right behind it: its work count is read from device memory, nothing is synchronised
    ws->heavy_enqueued = false;
    if (b.kind != KIND_WIDTH && coop_usable(ws, b)) {
        // in rounds of at most n / 4 searches (the stage's rows and record chunks are sized for one round); as many rounds
        // are queued as it takes to cover the whole batch, so that nothing is left over whatever share of it is heavy --
        // a round that finds its part of the list empty costs four empty launches
        const uint32_t round_cap = (uint32_t)std::min<uint64_t>(n_work_total, std::max<uint64_t>(65536, n_work_total / 4));
        const uint32_t rounds = (uint32_t)((n_work_total + round_cap - 1) / round_cap);
        ws->heavy_cap = (uint32_t)std::min<uint64_t>(n_work_total, (uint64_t)round_cap * rounds);
        if ((rc = ensure(ws->strict2_list, ws->strict2_list_cap, (size_t)n_work_total + 1))) return rc;
        Params S = P;
        coop_layout(S, b, seed_cap);
        for (uint32_t r = 0; r < rounds; ++r) {
            StageIO io;
            io.work_list = ws->strict_list + (size_t)r * round_cap; io.n_work = round_cap; io.n_work_skip = r * round_cap;
            io.n_work_dev = reinterpret_cast<const uint32_t *>(ws->counters + CNT_STRICT);
            io.flag_list = ws->strict2_list; io.flag_count = ws->counters + CNT_STRICT2;
            if ((rc = issue_chunk(ws, b, S, ws->heavy, MAX_PIPES, V_COOP, io, stream))) return rc;
        }
        ws->heavy_enqueued = true;
    }
    CU(cudaEventRecord(ws->ev1, stream));
    ws->trace = trace_saved;
    return HSA_OK;
}

// Finish half: waits for the batch, re-runs the items that ran out of stack / hit capacity through the same
// pipeline with the large-capacity kernel, returns the statistics block.
// One synchronous re-run stage over `n` listed items (device list `list`): returns the counters after it.
static int run_stage_sync(hsa_workspace *ws, const Batch &b, Params S, Pipe &pipe, int slot, Variant v, const uint32_t *list,
                          uint32_t n, uint32_t *flag_list, int flag_counter, cudaStream_t stream,
                          unsigned long long cnt[CNT_ALLOC], float *ms)
{
    int rc;
    CU(cudaMemsetAsync(ws->counters + flag_counter, 0, sizeof(unsigned long long), stream));
    StageIO io;
    io.work_list = list; io.n_work = n; io.flag_list = flag_list; io.flag_count = ws->counters + flag_counter;
    CU(cudaEventRecord(ws->ev0, stream));
    if ((rc = issue_chunk(ws, b, S, pipe, slot, v, io, stream))) return rc;
    CU(cudaEventRecord(ws->ev1, stream));
    CU(cudaMemcpyAsync(cnt, ws->counters, CNT_ALLOC * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    trace_dump(ws, cnt);
    float t = 0;
    CU(cudaEventElapsedTime(&t, ws->ev0, ws->ev1));
    *ms += t;
    return HSA_OK;
}

// Finish half: waits for the batch and completes what the queued stages could not: heavy searches beyond the
// cooperative stage's queue share, then whatever the cooperative stage handed on (or everything heavy, when that
// stage cannot be used) through the large-capacity kernel.  Returns the statistics block.
// `side` is the stream the statistics are read back on after the batch's end event; it differs from `stream`
// in the job pipeline, where later jobs may already be queued on the compute stream.
static int batch_finish(hsa_workspace *ws, const Batch &b, cudaStream_t stream, cudaStream_t side, uint64_t stats[CNT_N], float *ms)
{
    const hsa_index *ix = ws->idx;
    CU(cudaSetDevice(ix->device));
    int rc;
    Params P; Variant v; uint32_t seed_cap;
    if ((rc = batch_params(ws, b, P, v, seed_cap))) return rc;
    unsigned long long cnt[CNT_ALLOC];
    CU(cudaStreamWaitEvent(side, ws->ev1, 0));
    CU(cudaMemcpyAsync(cnt, ws->counters, sizeof(cnt), cudaMemcpyDeviceToHost, side));
    CU(cudaStreamSynchronize(side));
    trace_dump(ws, cnt);
    float t = 0;
    CU(cudaEventElapsedTime(&t, ws->ev0, ws->ev1));
    *ms = t;
    if (cnt[CNT_BAD]) return fail(HSA_E_ARG, "a score exceeded the bucket table (internal sizing error)");
    const uint64_t n_heavy = cnt[CNT_STRICT];                 // handed on by the fast kernel
    const bool coop_ok = coop_usable(ws, b);
    Params SC = P, SL = P;
    coop_layout(SC, b, seed_cap);
    set_layout(SL, b.max_len, seed_cap, b.n_buckets, b.n_opts, 4, false);
    uint64_t done_heavy = ws->heavy_enqueued ? std::min<uint64_t>(n_heavy, ws->heavy_cap) : 0;
    if (coop_ok) {
        if ((rc = ensure(ws->strict2_list, ws->strict2_list_cap, (size_t)std::max<uint64_t>(n_heavy, 1) + 1))) return rc;
        while (done_heavy < n_heavy) {                        // (rare) more heavy searches than the queued stage's share
            const uint32_t n = (uint32_t)std::min<uint64_t>(n_heavy - done_heavy, std::max<uint32_t>(ws->heavy_cap, 65536));
            // this round's leftovers are appended behind the earlier ones: keep CNT_STRICT2 running
            StageIO io;
            io.work_list = ws->strict_list + done_heavy; io.n_work = n;
            io.flag_list = ws->strict2_list; io.flag_count = ws->counters + CNT_STRICT2;
            CU(cudaEventRecord(ws->ev0, stream));
            if ((rc = issue_chunk(ws, b, SC, ws->heavy, MAX_PIPES, V_COOP, io, stream))) return rc;
            CU(cudaEventRecord(ws->ev1, stream));
            CU(cudaMemcpyAsync(cnt, ws->counters, sizeof(cnt), cudaMemcpyDeviceToHost, stream));
            CU(cudaStreamSynchronize(stream));
            trace_dump(ws, cnt);
            CU(cudaEventElapsedTime(&t, ws->ev0, ws->ev1));
            *ms += t;
            done_heavy += n;
        }
    }
    // last resort: the large-capacity per-lane kernel
    const uint32_t *last_list = coop_ok ? ws->strict2_list : ws->strict_list;
    const uint64_t n_last = coop_ok ? cnt[CNT_STRICT2] : n_heavy;
    if (n_last) {
        uint32_t *list_dev = nullptr;
        CU(cudaMalloc((void **)&list_dev, n_last * 4));
        CU(cudaMemcpyAsync(list_dev, last_list, n_last * 4, cudaMemcpyDeviceToDevice, stream));
        rc = run_stage_sync(ws, b, SL, ws->strict, MAX_PIPES + 1, V_LARGE, list_dev, (uint32_t)n_last, ws->strict_list, CNT_STRICT,
                            stream, cnt, ms);
        cudaFree(list_dev);
        if (rc) return rc;
        if (cnt[CNT_STRICT] || cnt[CNT_BAD])
            return fail(HSA_E_CAPACITY, "a search exceeded the large-capacity kernel's 262144-record stack or 4096-hit capacity");
    }
    for (int i = 0; i < CNT_N; ++i) stats[i] = cnt[i];
    stats[CNT_STRICT] = n_heavy;
    return HSA_OK;
}

static int run_batch(hsa_workspace *ws, const Batch &b, cudaStream_t stream, bool sync, uint64_t stats[CNT_N], float *ms)
{
    int rc = batch_enqueue(ws, b, stream, sync);
    if (rc || !sync) return rc;
    return batch_finish(ws, b, stream, stream, stats, ms);
}


// ---------------------------------------------------------------------------------------------- results
static int result_reserve(hsa_result_t *r, size_t n_items, size_t n_aln)
{
    if (n_items > r->cap_items || !r->n_aln) {
        if (r->n_aln) cudaFreeHost(r->n_aln);
        if (r->aln_off) cudaFreeHost(r->aln_off);
        r->n_aln = nullptr; r->aln_off = nullptr; r->cap_items = 0;
        size_t c = n_items + n_items / 8 + 16;
        CU(cudaHostAlloc((void **)&r->n_aln, c * sizeof(int32_t), cudaHostAllocDefault));
        CU(cudaHostAlloc((void **)&r->aln_off, c * sizeof(uint64_t), cudaHostAllocDefault));
        r->cap_items = c;
    }
    if (n_aln > r->cap_aln || !r->aln) {
        if (r->aln) cudaFreeHost(r->aln);
        r->aln = nullptr; r->cap_aln = 0;
        size_t c = n_aln + n_aln / 8 + 16;
        CU(cudaHostAlloc((void **)&r->aln, c * sizeof(hsa_aln1_t), cudaHostAllocDefault));
        r->cap_aln = c;
    }
    return HSA_OK;
}

extern "C" void hsa_result_free(hsa_result_t *r)
{
    if (!r) return;
    if (r->n_aln) cudaFreeHost(r->n_aln);
    if (r->aln_off) cudaFreeHost(r->aln_off);
    if (r->aln) cudaFreeHost(r->aln);
    memset(r, 0, sizeof(*r));
}

// ---------------------------------------------------------------------------------------------- jobs
// A job = one host-buffer batch in flight:  H2D (copy stream) -> kernels (compute stream) -> D2H (copy stream).
// hsa_*_submit enqueues the first two and returns; hsa_job_wait finishes the batch (large-capacity re-runs if
// any), copies the results out and releases the job.  With two jobs in flight the copies of one batch overlap
// the kernels of the other.  The blocking entry points are submit + wait.
struct hsa_job {
    hsa_index *ix = nullptr;
    hsa_workspace *ws = nullptr;
    int slot = -1;
    Batch b;
    size_t want = 0;                 // hit arena capacity of this attempt
    cudaEvent_t h2d_done = nullptr;
};

static int job_begin(const hsa_index_t *ix_c, hsa_job **out)
{
    hsa_index *ix = const_cast<hsa_index *>(ix_c);
    CU(cudaSetDevice(ix->device));
    int slot = -1;
    for (int i = 0; i < 3; ++i) if (!ix->pool_busy[i]) { slot = i; break; }
    if (slot < 0) return fail(HSA_E_ARG, "too many jobs in flight on this index (3): wait for one first");
    if (!ix->pool[slot]) {
        int rc = hsa_workspace_create(ix, 0, 0, 0, &ix->pool[slot]);
        if (rc) return rc;
        if (slot == 0) ix->ws = ix->pool[0];
    }
    hsa_job *j = new hsa_job();
    j->ix = ix; j->ws = ix->pool[slot]; j->slot = slot;
    if (cudaEventCreateWithFlags(&j->h2d_done, cudaEventDisableTiming) != cudaSuccess) { delete j; return fail(HSA_E_CUDA, "cudaEventCreate failed"); }
    ix->pool_busy[slot] = true;
    *out = j;
    return HSA_OK;
}

static void job_release(hsa_job *j)
{
    if (!j) return;
    if (j->slot >= 0) j->ix->pool_busy[j->slot] = false;
    if (j->h2d_done) cudaEventDestroy(j->h2d_done);
    delete j;
}

// inputs are queued on the copy stream: order the kernels behind them and enqueue the batch
static int job_launch(hsa_job *j)
{
    hsa_index *ix = j->ix; hsa_workspace *ws = j->ws; Batch &b = j->b;
    int rc;
    if ((rc = ensure(ws->n_aln_dev, ws->items_cap, (size_t)b.n_items + 1))) return rc;
    if ((rc = ensure(ws->aln_off_dev, ws->items2_cap, (size_t)b.n_items + 1))) return rc;
    j->want = std::max<size_t>((size_t)b.n_items * 2 + 1024, ws->aln_cap / 9);
    if ((rc = ensure(ws->aln_dev, ws->aln_cap, j->want * 9))) return rc;
    b.n_aln = ws->n_aln_dev; b.aln_off = ws->aln_off_dev; b.aln = ws->aln_dev; b.aln_cap = j->want;
    CU(cudaEventRecord(j->h2d_done, ix->h2d));
    CU(cudaStreamWaitEvent(ix->stream, j->h2d_done, 0));
    return batch_enqueue(ws, b, ix->stream, true);
}

static int job_finish(hsa_job *j, hsa_result_t *res)
{
    hsa_index *ix = j->ix; hsa_workspace *ws = j->ws; Batch &b = j->b;
    int rc;
    uint64_t stats[CNT_N];
    float ms = 0, ms_total = 0;
    uint32_t launches = ws->last_launches;
    HostTick tk;
    for (int attempt = 0;; ++attempt) {
        if ((rc = batch_finish(ws, b, ix->stream, ix->d2h, stats, &ms))) return rc;
        ms_total += ms;
        if (stats[CNT_ALN] <= j->want) break;
        if (attempt == 2) return fail(HSA_E_CAPACITY, "hit arena overflow persisted");
        j->want = stats[CNT_ALN] + 1024;             // the counter kept counting past the capacity: run again
        if ((rc = ensure(ws->aln_dev, ws->aln_cap, j->want * 9))) return rc;
        b.aln = ws->aln_dev; b.aln_cap = j->want;
        if ((rc = batch_enqueue(ws, b, ix->stream, true))) return rc;
        launches += ws->last_launches;
    }
    tk.mark(1);
    const size_t total = stats[CNT_ALN];
    if ((rc = result_reserve(res, b.n_items, total))) return rc;
    CU(cudaStreamWaitEvent(ix->d2h, ws->ev1, 0));
    CU(cudaMemcpyAsync(res->n_aln, ws->n_aln_dev, (size_t)b.n_items * sizeof(int32_t), cudaMemcpyDeviceToHost, ix->d2h));
    CU(cudaMemcpyAsync(res->aln_off, ws->aln_off_dev, (size_t)b.n_items * sizeof(uint64_t), cudaMemcpyDeviceToHost, ix->d2h));
    if (total) CU(cudaMemcpyAsync(res->aln, ws->aln_dev, total * sizeof(hsa_aln1_t), cudaMemcpyDeviceToHost, ix->d2h));
    CU(cudaStreamSynchronize(ix->d2h));
    tk.mark(2);
    res->n_items = b.n_items; res->n_aln_total = total;
    res->occ_lookups = stats[CNT_LOOKUPS]; res->n_strict = stats[CNT_STRICT];
    res->pops = stats[CNT_POPS]; res->steps = stats[CNT_STEPS];
    res->kernel_ms = ms_total; res->kernel_launches = launches;
    ws->last_launches = launches;
    return HSA_OK;
}

extern "C" int hsa_job_wait(hsa_job_t *job, hsa_result_t *res)
{
    if (!job || !res) return fail(HSA_E_ARG, "null argument");
    const int rc = job_finish(job, res);
    job_release(job);
    return rc;
}

static int upload_reads(hsa_job *j, const uint8_t *codes, const uint64_t *off, const uint32_t *len, size_t n, uint32_t *max_len_out)
{
    hsa_workspace *ws = j->ws;
    size_t bytes = 0; uint32_t ml = 0, mn = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; ++i) {
        const size_t e = off[i] + len[i];
        if (e > bytes) bytes = e;
        if (len[i] > ml) ml = len[i];
        if (len[i] < mn) mn = len[i];
    }
    if (mn == 0) return fail(HSA_E_ARG, "empty read (len == 0)");
    if (ml > 4095) return fail(HSA_E_ARG, "reads longer than 4095 bases are not supported");
    int rc;
    if ((rc = ensure(ws->codes_dev, ws->codes_cap, bytes + 16))) return rc;
    if (n + 1 > ws->reads_cap || !ws->off_dev) {
        cudaFree(ws->off_dev); cudaFree(ws->len_dev); ws->off_dev = nullptr; ws->len_dev = nullptr;
        size_t c = n + n / 8 + 16;
        CU(cudaMalloc((void **)&ws->off_dev, c * sizeof(uint64_t)));
        CU(cudaMalloc((void **)&ws->len_dev, c * sizeof(uint32_t)));
        ws->reads_cap = c;
    }
    cudaStream_t s = j->ix->h2d;
    CU(cudaMemcpyAsync(ws->codes_dev, codes, bytes, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(ws->off_dev, off, n * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(ws->len_dev, len, n * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    *max_len_out = ml;
    return HSA_OK;
}

// option table (+ length -> option index map) to the device through the workspace's pinned staging: no host sync
static int upload_opts(hsa_workspace *ws, const std::vector<hsa_gap_opt_t> &opts, uint32_t max_len, Batch *bt,
                       const std::vector<uint16_t> *len2opt, cudaStream_t stream)
{
    uint32_t *n_buckets = &bt->n_buckets;
    bt->max_seed_len = 0;
    for (const hsa_gap_opt_t &o : opts)
        if (o.seed_len > 0 && (uint32_t)o.seed_len < max_len) bt->max_seed_len = std::max(bt->max_seed_len, (uint32_t)o.seed_len);
    int rc;
    const uint32_t st = ws->stage_next;
    ws->stage_next = (st + 1) % hsa_workspace::OPT_STAGES;
    if (!ws->stage_ev[st]) CU(cudaEventCreateWithFlags(&ws->stage_ev[st], cudaEventDisableTiming));
    if (ws->stage_used[st]) CU(cudaEventSynchronize(ws->stage_ev[st]));     // the copy that last read this stage has run
    if (opts.size() > ws->opts_stage_cap[st] || !ws->opts_stage[st]) {
        cudaFreeHost(ws->opts_stage[st]); ws->opts_stage[st] = nullptr;
        CU(cudaHostAlloc((void **)&ws->opts_stage[st], (opts.size() + 16) * sizeof(DevOpt), cudaHostAllocDefault));
        ws->opts_stage_cap[st] = opts.size() + 16;
    }
    ws->opts_host = ws->opts_stage[st];
    *n_buckets = 1;
    bt->scores_positive = true;
    for (size_t i = 0; i < opts.size(); ++i) {
        if ((rc = check_opt(opts[i], max_len, n_buckets))) return rc;
        to_devopt(opts[i], ws->opts_host[i]);
        if (opts[i].s_mm < 1 || opts[i].s_gapo < 1 || opts[i].s_gape < 1) bt->scores_positive = false;
    }
    if ((rc = ensure(ws->opts_dev, ws->opts_cap, opts.size()))) return rc;
    CU(cudaMemcpyAsync(ws->opts_dev, ws->opts_host, opts.size() * sizeof(DevOpt), cudaMemcpyHostToDevice, stream));
    if (len2opt) {
        if (len2opt->size() > ws->l2o_stage_cap[st] || !ws->l2o_stage[st]) {
            cudaFreeHost(ws->l2o_stage[st]); ws->l2o_stage[st] = nullptr;
            CU(cudaHostAlloc((void **)&ws->l2o_stage[st], (len2opt->size() + 64) * sizeof(uint16_t), cudaHostAllocDefault));
            ws->l2o_stage_cap[st] = len2opt->size() + 64;
        }
        memcpy(ws->l2o_stage[st], len2opt->data(), len2opt->size() * sizeof(uint16_t));
        if ((rc = ensure(ws->len2opt_dev, ws->len2opt_cap, len2opt->size()))) return rc;
        CU(cudaMemcpyAsync(ws->len2opt_dev, ws->l2o_stage[st], len2opt->size() * sizeof(uint16_t), cudaMemcpyHostToDevice, stream));
    }
    CU(cudaEventRecord(ws->stage_ev[st], stream));
    ws->stage_used[st] = true;
    return HSA_OK;
}

// per-read option resolution of bwa_cal_sa_reg_gap for one caller opt (bwtaln.c:260-261, 273-276, 330-332)
static int resolve_whole_opts(const hsa_gap_opt_t *opt, int keep_gape, const std::vector<uint32_t> &lens, uint32_t max_len,
                              std::vector<hsa_gap_opt_t> &opts, std::vector<uint16_t> &len2opt, int32_t *filter_max_n)
{
    len2opt.assign((size_t)max_len + 1, 0);
    opts.clear();
    for (uint32_t L : lens) {
        hsa_gap_opt_t o = *opt;
        if (!keep_gape) o.mode &= ~HSA_MODE_GAPE;
        if (opt->fnr > 0.0) o.max_diff = hsa_cal_maxdiff((int)L, 0.02, opt->fnr);
        o.seed_len = opt->seed_len < (int)L ? opt->seed_len : 0x7fffffff;
        size_t j = 0;                                   // lengths that resolve to the same options share a slot
        for (; j < opts.size(); ++j) if (opts[j].max_diff == o.max_diff && opts[j].seed_len == o.seed_len) break;
        if (j == opts.size()) {
            if (opts.size() >= 1024) return fail(HSA_E_ARG, "too many distinct resolved option sets");
            opts.push_back(o);
        }
        len2opt[L] = (uint16_t)j;
    }
    *filter_max_n = opt->fnr > 0.0 ? hsa_cal_maxdiff((int)max_len, 0.02, opt->fnr) : opt->max_diff;
    if (opts.empty()) { opts.push_back(*opt); if (opts[0].max_diff < 0) opts[0].max_diff = 0; }
    return HSA_OK;
}

static int empty_result(hsa_result_t *res) { res->n_items = 0; res->n_aln_total = 0; return HSA_OK; }

extern "C" int hsa_match_gap_batch(const hsa_index_t *ix, const uint8_t *codes, size_t codes_bytes,
                                   const hsa_task_t *tasks, size_t n_tasks,
                                   const hsa_gap_opt_t *opts, size_t n_opts, hsa_result_t *res)
{
    if (!ix || !res || (n_tasks && (!codes || !tasks || !opts))) return fail(HSA_E_ARG, "null argument");
    if (n_tasks > 0xFFFFFFF0ull) return fail(HSA_E_ARG, "too many tasks");
    uint32_t max_len = 0;
    for (size_t i = 0; i < n_tasks; ++i) {
        const hsa_task_t &t = tasks[i];
        if (t.opt_idx >= n_opts) return fail(HSA_E_ARG, "task.opt_idx out of range");
        if (t.read_off > codes_bytes || (uint64_t)t.read_len > codes_bytes - t.read_off || (uint64_t)t.sub_off + t.len > t.read_len ||
            (uint64_t)t.wsrc_off + t.len > t.read_len)
            return fail(HSA_E_ARG, "task window outside its read / codes buffer");
        if (t.seed_mode > HSA_SEED_ALIAS || t.strand > 1) return fail(HSA_E_ARG, "bad task.seed_mode / strand");
        if (t.len == 0) return fail(HSA_E_ARG, "empty task (len == 0)");
        if (t.seed_mode == HSA_SEED_TAIL && (opts[t.opt_idx].seed_len <= 0 || (uint32_t)opts[t.opt_idx].seed_len >= t.len))
            return fail(HSA_E_ARG, "HSA_SEED_TAIL needs 0 < opt.seed_len < task.len (bwtaln.c:344)");
        max_len = std::max(max_len, std::max(t.len, t.read_len));
    }
    if (n_tasks == 0) return empty_result(res);
    hsa_job *j; int rc;
    if ((rc = job_begin(ix, &j))) return rc;
    hsa_workspace *ws = j->ws; Batch &b = j->b;
    std::vector<hsa_gap_opt_t> ov(opts, opts + n_opts);
    if ((rc = upload_opts(ws, ov, max_len, &b, nullptr, j->ix->h2d)) ||
        (rc = ensure(ws->codes_dev, ws->codes_cap, codes_bytes + 16)) ||
        (rc = ensure(ws->tasks_dev, ws->tasks_cap, n_tasks))) { job_release(j); return rc; }
    if (cudaMemcpyAsync(ws->codes_dev, codes, codes_bytes, cudaMemcpyHostToDevice, j->ix->h2d) != cudaSuccess ||
        cudaMemcpyAsync(ws->tasks_dev, tasks, n_tasks * sizeof(Task), cudaMemcpyHostToDevice, j->ix->h2d) != cudaSuccess) {
        job_release(j); return fail(HSA_E_CUDA, "H2D copy of the task batch failed");
    }
    b.kind = KIND_TASKS; b.n_groups = (uint32_t)n_tasks; b.n_items = (uint32_t)n_tasks; b.max_len = max_len;
    b.n_opts = (uint32_t)n_opts; b.codes = ws->codes_dev; b.tasks = ws->tasks_dev;
    if ((rc = job_launch(j))) { job_release(j); return rc; }
    return hsa_job_wait(j, res);
}

// bwt_match_gap itself (bwtgap.h:26), one call: the caller's width arrays are arguments (in/out), exactly as in bwt_aux_t.
extern "C" int hsa_match_gap_call(const hsa_index_t *ix, const uint8_t *seq, uint32_t len, int strand, hsa_width_t *width_back,
                                  hsa_width_t *width_seed, const hsa_gap_opt_t *opt, int *n_aln_out, hsa_aln1_t **aln_out)
{
    if (!ix || !seq || !len || !width_back || !opt || !n_aln_out || !aln_out) return fail(HSA_E_ARG, "null / empty argument");
    if (len > 4095) return fail(HSA_E_ARG, "sequences longer than 4095 bases are not supported");
    uint32_t seed_mode = HSA_SEED_NONE;
    if (width_seed == width_back) {
        if (opt->seed_len != (int)len) return fail(HSA_E_ARG, "width_seed aliasing width_back needs opt->seed_len == len (bwtgap.c:802, :809)");
        seed_mode = HSA_SEED_ALIAS;
    } else if (width_seed && opt->seed_len > 0 && (uint32_t)opt->seed_len < len) seed_mode = HSA_SEED_TAIL;   // bwtaln.c:344-346
    // (width_seed with seed_len >= len and no aliasing: the reference reads it out of bounds, SURVEY.md hazard 3 -> unseeded)
    hsa_job *j; int rc;
    if ((rc = job_begin(ix, &j))) return rc;
    hsa_workspace *ws = j->ws; Batch &b = j->b;
    struct Rel { hsa_job *j; bool armed; ~Rel() { if (armed) job_release(j); } } rel{j, true};
    std::vector<hsa_gap_opt_t> ov(1, *opt);
    if ((rc = upload_opts(ws, ov, len, &b, nullptr, j->ix->h2d)) || (rc = ensure(ws->codes_dev, ws->codes_cap, (size_t)len + 16)) ||
        (rc = ensure(ws->tasks_dev, ws->tasks_cap, 1)) || (rc = ensure(ws->width_out_dev, ws->width_out_cap, (size_t)len + 1))) return rc;
    Task t; memset(&t, 0, sizeof(t));
    t.read_off = 0; t.read_len = len; t.strand = 0; t.sub_off = 0; t.len = len; t.wsrc_off = 0; t.seed_mode = seed_mode; t.opt_idx = 0;
    CU(cudaMemcpyAsync(ws->codes_dev, seq, len, cudaMemcpyHostToDevice, j->ix->h2d));
    CU(cudaMemcpyAsync(ws->tasks_dev, &t, sizeof(t), cudaMemcpyHostToDevice, j->ix->h2d));
    CU(cudaStreamSynchronize(j->ix->h2d));                       // `t` lives on this stack frame
    b.kind = KIND_TASKS; b.n_groups = 1; b.n_items = 1; b.max_len = len; b.n_opts = 1;
    b.codes = ws->codes_dev; b.tasks = ws->tasks_dev; b.width_out = ws->width_out_dev;
    // the caller's widths as the item's row (layout: hsa_core.cuh Params::rows)
    Params P; Variant v; uint32_t seed_cap;
    if ((rc = batch_params(ws, b, P, v, seed_cap))) return rc;
    std::vector<uint8_t> row(P.row_stride, 0);
    uint32_t *w = reinterpret_cast<uint32_t *>(row.data());
    uint32_t w_prev = 0xFFFFFFFFu;
    bool has_n = false;
    for (uint32_t i = 0; i <= len; ++i) {
        w[i] = width_back[i].w;
        uint8_t byte = bound_byte((uint32_t)width_back[i].bid, width_back[i].w, w_prev);
        if (i < len) { if (seq[i] > 3) has_n = true; else byte |= (uint8_t)(seq[i] << BB_BASE_SHIFT); }   // the base of step i
        row[P.row_bid_off + i] = byte;
        w_prev = width_back[i].w;
    }
    reinterpret_cast<uint32_t *>(row.data() + P.row_tail_off)[1] = has_n ? ROW_FLAG_HAS_N : 0u;
    if (seed_mode == HSA_SEED_TAIL) {
        w_prev = 0xFFFFFFFFu;
        for (uint32_t i = 0; i <= (uint32_t)opt->seed_len; ++i) {
            row[P.row_seed_off + i] = bound_byte((uint32_t)width_seed[i].bid, width_seed[i].w, w_prev);
            w_prev = width_seed[i].w;
        }
    }
    b.rows_host = row.data(); b.rows_host_bytes = row.size();
    if ((rc = job_launch(j))) return rc;
    hsa_result_t res; memset(&res, 0, sizeof(res));
    rel.armed = false;
    rc = hsa_job_wait(j, &res);                                  // releases the job
    if (rc) { hsa_result_free(&res); return rc; }
    // results as the reference returns them: a malloc-family array of at least 10 zero-filled entries (bwtgap.c:137-138)
    const int n = res.n_aln[0];
    hsa_aln1_t *out = (hsa_aln1_t *)calloc((size_t)(n < 10 ? 10 : n), sizeof(hsa_aln1_t));
    if (!out) { hsa_result_free(&res); return fail(HSA_E_NOMEM, "calloc failed"); }
    if (n) memcpy(out, res.aln + res.aln_off[0], (size_t)n * sizeof(hsa_aln1_t));
    for (int q = 0; q < n; ++q) out[q].strand = (uint32_t)strand & 3u;       // p->strand = aux->strand, bwtgap.c:235
    hsa_result_free(&res);
    // width_back after gap_shadow (bids beyond the device's 5-bit field were never touched by it: keep the caller's)
    std::vector<u32x2> wb((size_t)len + 1);
    CU(cudaSetDevice(ix->device));
    if (cudaMemcpy(wb.data(), ws->width_out_dev, wb.size() * sizeof(u32x2), cudaMemcpyDeviceToHost) != cudaSuccess) {
        free(out); return fail(HSA_E_CUDA, "D2H copy of the widths failed");
    }
    for (uint32_t i = 0; i <= len; ++i) {
        width_back[i].w = wb[i].x;
        if (wb[i].y < BB_BID) width_back[i].bid = (int)wb[i].y;
    }
    *n_aln_out = n; *aln_out = out;
    return HSA_OK;
}

extern "C" int hsa_whole_reads_submit(const hsa_index_t *ix, const uint8_t *codes, const uint64_t *off, const uint32_t *len,
                                      size_t n_reads, const hsa_gap_opt_t *opt, int keep_gape, hsa_job_t **job)
{
    if (!ix || !opt || !job || !n_reads || !codes || !off || !len) return fail(HSA_E_ARG, "null / empty argument");
    if (n_reads > 0x2AAAAAA0ull) return fail(HSA_E_ARG, "too many reads in one batch");
    hsa_job *j; int rc; uint32_t max_len = 0;
    if ((rc = job_begin(ix, &j))) return rc;
    hsa_workspace *ws = j->ws; Batch &b = j->b;
    if ((rc = upload_reads(j, codes, off, len, n_reads, &max_len))) { job_release(j); return rc; }
    std::vector<uint8_t> seen((size_t)max_len + 1, 0);
    for (size_t i = 0; i < n_reads; ++i) seen[len[i]] = 1;
    std::vector<uint32_t> lens;
    for (uint32_t L = 0; L <= max_len; ++L) if (seen[L]) lens.push_back(L);
    std::vector<hsa_gap_opt_t> opts; std::vector<uint16_t> l2o;
    if ((rc = resolve_whole_opts(opt, keep_gape, lens, max_len, opts, l2o, &b.filter_max_n)) ||
        (rc = upload_opts(ws, opts, max_len, &b, &l2o, j->ix->h2d))) { job_release(j); return rc; }
    b.kind = KIND_WHOLE; b.n_groups = (uint32_t)n_reads; b.n_items = (uint32_t)n_reads; b.max_len = max_len;
    b.n_opts = (uint32_t)opts.size(); b.codes = ws->codes_dev; b.read_off = ws->off_dev; b.read_len = ws->len_dev;
    b.ragged = lens.size() > 1;
    if ((rc = job_launch(j))) { job_release(j); return rc; }
    *job = j;
    return HSA_OK;
}

extern "C" int hsa_whole_reads(const hsa_index_t *ix, const uint8_t *codes, const uint64_t *off, const uint32_t *len,
                               size_t n_reads, const hsa_gap_opt_t *opt, int keep_gape, hsa_result_t *res)
{
    if (!res || !opt) return fail(HSA_E_ARG, "null argument");
    if (n_reads == 0) return empty_result(res);
    hsa_job_t *j; int rc;
    HostTick tk;
    if ((rc = hsa_whole_reads_submit(ix, codes, off, len, n_reads, opt, keep_gape, &j))) return rc;
    tk.mark(0);
    return hsa_job_wait(j, res);
}

extern "C" int hsa_splice_seeds_submit(const hsa_index_t *ix, const uint8_t *codes, const uint64_t *off, const uint32_t *len,
                                       size_t n_reads, const hsa_gap_opt_t *opt, hsa_job_t **job)
{
    if (!ix || !opt || !job || !n_reads || !codes || !off || !len) return fail(HSA_E_ARG, "null / empty argument");
    if (n_reads > 0x2AAAAAA0ull) return fail(HSA_E_ARG, "too many reads in one batch");
    hsa_job *j; int rc; uint32_t max_len = 0;
    if ((rc = job_begin(ix, &j))) return rc;
    hsa_workspace *ws = j->ws; Batch &b = j->b;
    if ((rc = upload_reads(j, codes, off, len, n_reads, &max_len))) { job_release(j); return rc; }
    hsa_gap_opt_t so = *opt;                             // bwtgap.c:769-774
    so.mode &= ~HSA_MODE_GAPE; so.max_gapo = 0; so.max_gape = 0; so.max_diff = opt->max_seed_diff;
    std::vector<hsa_gap_opt_t> opts(1, so);
    if ((rc = upload_opts(ws, opts, max_len, &b, nullptr, j->ix->h2d))) { job_release(j); return rc; }
    b.kind = KIND_SEEDS; b.n_groups = (uint32_t)n_reads; b.n_items = (uint32_t)n_reads * 6; b.max_len = max_len;
    b.n_opts = 1; b.codes = ws->codes_dev; b.read_off = ws->off_dev; b.read_len = ws->len_dev;
    if ((rc = job_launch(j))) { job_release(j); return rc; }
    *job = j;
    return HSA_OK;
}

extern "C" int hsa_splice_seeds(const hsa_index_t *ix, const uint8_t *codes, const uint64_t *off, const uint32_t *len,
                                size_t n_reads, const hsa_gap_opt_t *opt, hsa_result_t *res)
{
    if (!res || !opt) return fail(HSA_E_ARG, "null argument");
    if (n_reads == 0) return empty_result(res);
    hsa_job_t *j; int rc;
    if ((rc = hsa_splice_seeds_submit(ix, codes, off, len, n_reads, opt, &j))) return rc;
    return hsa_job_wait(j, res);
}

extern "C" int hsa_cal_width_batch(const hsa_index_t *ix, const uint8_t *codes, const uint64_t *off,
                                   const uint32_t *len, size_t n, int type, hsa_width_t *width_out, int *bid_out)
{
    if (type != 1) return fail(HSA_E_ARG, "only bwt_cal_width type 1 (forward search on rev_bwt) is on the GPU path");
    if (!width_out || !bid_out) return fail(HSA_E_ARG, "null argument");
    if (n == 0) return HSA_OK;
    if (!ix || !codes || !off || !len) return fail(HSA_E_ARG, "null argument");
    hsa_job *j; uint32_t max_len; int rc;
    if ((rc = job_begin(ix, &j))) return rc;
    hsa_workspace *ws = j->ws;
    rc = upload_reads(j, codes, off, len, n, &max_len);
    if (!rc && (cudaEventRecord(j->h2d_done, j->ix->h2d) != cudaSuccess ||
                cudaStreamWaitEvent(ix->stream, j->h2d_done, 0) != cudaSuccess)) rc = fail(HSA_E_CUDA, "event ordering failed");
    struct Rel { hsa_job *j; ~Rel() { cudaStreamSynchronize(j->ix->stream); job_release(j); } } rel{j};
    if (rc) return rc;
    size_t total = 0;
    for (size_t i = 0; i < n; ++i) total = std::max(total, (size_t)off[i] + i + len[i] + 1);
    hsa_gap_opt_t o; hsa_gap_opt_default(&o); o.max_diff = 0;
    std::vector<hsa_gap_opt_t> opts(1, o);
    Batch b;
    if ((rc = upload_opts(ws, opts, max_len, &b, nullptr, ix->stream))) return rc;
    if ((rc = ensure(ws->width_out_dev, ws->width_out_cap, total))) return rc;
    cudaFree(ws->bid_dev); ws->bid_dev = nullptr;
    CU(cudaMalloc((void **)&ws->bid_dev, n * sizeof(int32_t)));
    if ((rc = ensure(ws->n_aln_dev, ws->items_cap, n + 1))) return rc;
    if ((rc = ensure(ws->aln_off_dev, ws->items2_cap, n + 1))) return rc;
    if ((rc = ensure(ws->aln_dev, ws->aln_cap, 1024))) return rc;
    b.kind = KIND_WIDTH; b.n_groups = (uint32_t)n; b.n_items = (uint32_t)n; b.max_len = max_len; b.n_opts = 1;
    b.codes = ws->codes_dev; b.read_off = ws->off_dev; b.read_len = ws->len_dev;
    b.n_aln = ws->n_aln_dev; b.aln_off = ws->aln_off_dev; b.aln = ws->aln_dev; b.aln_cap = ws->aln_cap / 9;
    b.width_out = ws->width_out_dev; b.bid_out = ws->bid_dev;
    uint64_t stats[CNT_N]; float ms;
    if ((rc = run_batch(ws, b, ix->stream, true, stats, &ms))) return rc;
    CU(cudaMemcpyAsync(width_out, ws->width_out_dev, total * sizeof(hsa_width_t), cudaMemcpyDeviceToHost, ix->stream));
    CU(cudaMemcpyAsync(bid_out, ws->bid_dev, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ix->stream));
    CU(cudaStreamSynchronize(ix->stream));
    return HSA_OK;
}

// ---------------------------------------------------------------------------------------------- device-resident
extern "C" int hsa_whole_reads_device(const hsa_index_t *ix, hsa_workspace_t *ws, const uint8_t *codes_dev,
                                      const uint64_t *off_dev, const uint32_t *len_dev, size_t n_reads,
                                      const uint32_t *lens_present, size_t n_lens_present,
                                      const hsa_gap_opt_t *opt, int keep_gape, int32_t *n_aln_dev, uint64_t *aln_off_dev,
                                      hsa_aln1_t *aln_dev, size_t aln_capacity, uint64_t *stats_dev, void *stream)
{
    if (!ix || !ws || ws->idx != ix || !opt || !lens_present || !n_lens_present) return fail(HSA_E_ARG, "bad argument");
    if (n_reads == 0) return HSA_OK;
    if (n_reads > 0xFFFFFFF0ull) return fail(HSA_E_ARG, "too many reads in one batch");
    CU(cudaSetDevice(ix->device));
    int rc;
    std::vector<uint32_t> lens(lens_present, lens_present + n_lens_present);
    uint32_t max_len = *std::max_element(lens.begin(), lens.end());
    std::vector<hsa_gap_opt_t> opts; std::vector<uint16_t> l2o;
    Batch b;
    cudaStream_t s = (cudaStream_t)stream;          // NULL = the legacy default stream, as everywhere in CUDA
    if ((rc = resolve_whole_opts(opt, keep_gape, lens, max_len, opts, l2o, &b.filter_max_n))) return rc;
    if ((rc = upload_opts(ws, opts, max_len, &b, &l2o, s))) return rc;
    b.kind = KIND_WHOLE; b.n_groups = (uint32_t)n_reads; b.n_items = (uint32_t)n_reads; b.max_len = max_len;
    b.n_opts = (uint32_t)opts.size(); b.codes = codes_dev; b.read_off = off_dev; b.read_len = len_dev;
    b.ragged = n_lens_present > 1;
    b.n_aln = n_aln_dev; b.aln_off = aln_off_dev; b.aln = reinterpret_cast<uint32_t *>(aln_dev); b.aln_cap = aln_capacity;
    uint64_t stats[CNT_N]; float ms = 0;
    if ((rc = run_batch(ws, b, s, false, stats, &ms))) return rc;
    if (stats_dev) {
        CU(cudaMemcpyAsync(stats_dev, ws->counters, CNT_N * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s));
        // word 7: searches the queued stages left unprocessed: what the cooperative kernel handed on or, when that
        // stage could not be queued (HSA_B200_COOP=0, non-positive scores, reads > 2000 bases), every heavy search
        CU(cudaMemcpyAsync(stats_dev + 7, ws->counters + (ws->heavy_enqueued ? (int)CNT_STRICT2 : (int)CNT_STRICT), sizeof(unsigned long long),
                           cudaMemcpyDeviceToDevice, s));
    }
    ws->last_stream = s; ws->last_valid = true; ws->last_n = n_reads; ws->last_aln_cap = aln_capacity;
    return HSA_OK;
}

// Completion check of the last hsa_whole_reads_device call on this workspace: waits for its stream, reads the statistics
// and fails unless every read was searched to the end and every hit fits the caller's arena.
extern "C" int hsa_workspace_check(hsa_workspace_t *ws, uint64_t stats_out[8])
{
    if (!ws || !ws->last_valid) return fail(HSA_E_ARG, "no device-resident call to check on this workspace");
    CU(cudaSetDevice(ws->idx->device));
    CU(cudaStreamSynchronize(ws->last_stream));
    unsigned long long cnt[CNT_ALLOC];
    CU(cudaMemcpy(cnt, ws->counters, sizeof(cnt), cudaMemcpyDeviceToHost));
    const uint64_t heavy = cnt[CNT_STRICT];
    uint64_t left = ws->heavy_enqueued ? cnt[CNT_STRICT2] + (heavy > ws->heavy_cap ? heavy - ws->heavy_cap : 0) : heavy;
    if (stats_out) {
        for (int i = 0; i < CNT_N; ++i) stats_out[i] = cnt[i];
        stats_out[7] = left;
    }
    char msg[256];
    if (cnt[CNT_BAD]) return fail(HSA_E_ARG, "a score exceeded the bucket table (internal sizing error)");
    if (cnt[CNT_ALN] > ws->last_aln_cap) {
        snprintf(msg, sizeof(msg), "hit arena too small: %llu hits, capacity %llu (results incomplete)", cnt[CNT_ALN],
                 (unsigned long long)ws->last_aln_cap);
        return fail(HSA_E_CAPACITY, msg);
    }
    if (left) {
        snprintf(msg, sizeof(msg), "%llu of %llu searches were left unprocessed by the queued stages (%llu heavy, cooperative stage %s): "
                 "run this batch through hsa_whole_reads, which finishes them with the large-capacity kernel",
                 (unsigned long long)left, (unsigned long long)ws->last_n, (unsigned long long)heavy, ws->heavy_enqueued ? "queued" : "not usable");
        return fail(HSA_E_CAPACITY, msg);
    }
    return HSA_OK;
}

// ---------------------------------------------------------------------------------------------- roofline probe
// Variants, all at full occupancy over the same buffer: [0] four dependent chains per thread (two 16-byte loads per
// sector), [1..3] 4 / 8 / 16 independent 256-bit loads in flight per thread, [4] the sa_kernel-shaped mix (a 256-bit load
// and a 4-byte load per pair, 8 pairs in flight), [5] the LF-walk shape (one data-dependent chain per thread, geometric
// chain length, 4-byte terminal load).  gbs_out[i] = sectors * 32 B / time of variant i, best of three runs.
enum { PROBE_VARIANTS = 6 };
extern "C" int hsa_random_sector_probe_ex(int device, size_t footprint_bytes, int iters, double *gbs_out, int n_out)
{
    if (!gbs_out || n_out < 1 || footprint_bytes < 4096 || iters < 1) return fail(HSA_E_ARG, "bad argument");
    CU(cudaSetDevice(device));
    apply_l2_fetch_granularity();
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    uint4 *buf = nullptr; unsigned long long *sink = nullptr;
    uint64_t n_sectors = footprint_bytes / 32;
    CU(cudaMalloc((void **)&buf, n_sectors * 32));
    CU(cudaMalloc((void **)&sink, 8));
    CU(cudaMemset(buf, 0x5A, n_sectors * 32));
    CU(cudaMemset(sink, 0, 8));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    const int block = 256;
    struct V { const void *fn; double per_iter; } vs[PROBE_VARIANTS] = {
        {(const void *)probe_kernel<4>, 4}, {(const void *)probe_kernel_mlp<4>, 4}, {(const void *)probe_kernel_mlp<8>, 8},
        {(const void *)probe_kernel_mlp<16>, 16}, {(const void *)probe_kernel_mix<8>, 16}, {(const void *)probe_kernel_walk, 1.125}};
    for (int v = 0; v < PROBE_VARIANTS && v < n_out; ++v) {
        int occ = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, vs[v].fn, block, 0));
        const int grid = sms * std::max(occ, 1);
        int it_warm = 4, it_run = vs[v].per_iter < 2 ? iters * 8 : iters;      // the single-chain variant does one load per iteration
        void *aw[] = {(void *)&buf, (void *)&n_sectors, (void *)&it_warm, (void *)&sink};
        void *ar[] = {(void *)&buf, (void *)&n_sectors, (void *)&it_run, (void *)&sink};
        CU(cudaLaunchKernel(vs[v].fn, dim3(grid), dim3(block), aw, 0, nullptr));       // warm-up (pulls an L2-sized set in)
        double best = 0;
        for (int rep = 0; rep < 3; ++rep) {
            CU(cudaEventRecord(e0));
            CU(cudaLaunchKernel(vs[v].fn, dim3(grid), dim3(block), ar, 0, nullptr));
            CU(cudaEventRecord(e1));
            CU(cudaEventSynchronize(e1));
            float ms = 0;
            CU(cudaEventElapsedTime(&ms, e0, e1));
            const double bytes = (double)grid * block * vs[v].per_iter * (double)it_run * 32.0;
            best = std::max(best, bytes / (ms * 1e-3) / 1e9);
        }
        gbs_out[v] = best;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(buf); cudaFree(sink);
    return HSA_OK;
}

extern "C" int hsa_random_sector_probe(int device, size_t footprint_bytes, int iters, double *gbs_out)
{
    if (!gbs_out) return fail(HSA_E_ARG, "bad argument");
    double v[PROBE_VARIANTS] = {0, 0, 0, 0, 0, 0};
    int rc = hsa_random_sector_probe_ex(device, footprint_bytes, iters, v, PROBE_VARIANTS);
    if (rc) return rc;
    *gbs_out = *std::max_element(v, v + PROBE_VARIANTS);
    return HSA_OK;
}

#include "hsa_sam_api.inl"

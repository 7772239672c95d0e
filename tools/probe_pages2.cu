// Follow-up to probe_pages.cu (the random-sector rate past 256 MB is a function of the number of 2 MB pages touched, not of
// the bytes): WHERE does the translation limit sit and what does it charge for?
//   T1 private pages per block (every SM works inside its own <= 32 pages, L2-resident footprint, span 9 GiB)
//      vs the same pages shared by all threads                                -> per-SM TLB or one shared structure?
//   T2 all lanes of a warp in ONE random page per instruction                  -> charged per lane or per distinct page?
//   T3 every sector loaded twice (two instructions, same address)             -> are L1 hits charged?
//   T4 a fraction of the loads into a compact 128 MB region, the rest over 1.5 GiB -> what a compact hot table would buy
//   T5 random 32-byte stores                                                    -> same limit for stores?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/probe_pages2 tools/probe_pages2.cu ; tools/probe_pages2
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <algorithm>
#include <cuda_runtime.h>
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31);
}
__global__ void fill_kernel(uint4 *buf, uint64_t n16)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t a = mix64(2 * i + 1), b = mix64(2 * i + 2);
        buf[i] = make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b, (uint32_t)(b >> 32));
    }
}
__device__ __forceinline__ void ld256(const void *p, uint32_t v[8])
{
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(p));
}
enum Mode { SHARED_PAGES = 0, BLOCK_PAGES = 1, WARP_PAGE = 2, DUP = 3, MIX = 4, STORE = 5, PLAIN = 6 };
// sectors are 32 bytes; a page is 65536 sectors.  chunk_sect: sectors used at the start of every page (T1, T2).
template <int MODE>
__global__ void probe(uint4 *buf, uint64_t n_pages, uint32_t chunk_sect, uint32_t pages_per_block, uint64_t n_sectors,
                      uint64_t hot_sectors, uint32_t hot_permille, int iters, unsigned long long *sink)
{
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t s = mix64(tid + 777), ws = mix64((tid >> 5) + 31337);
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        uint64_t sec;
        if (MODE == SHARED_PAGES) {
            sec = __umul64hi(s, n_pages) * 65536ull + (uint32_t)(s & 0xffffffu) % chunk_sect;
        } else if (MODE == BLOCK_PAGES) {
            const uint64_t pg = (uint64_t)blockIdx.x * pages_per_block + (uint32_t)(s >> 40) % pages_per_block;
            sec = pg * 65536ull + (uint32_t)(s & 0xffffffu) % chunk_sect;
        } else if (MODE == WARP_PAGE) {
            ws = ws * 6364136223846793005ull + 1442695040888963407ull;
            sec = __umul64hi(ws, n_pages) * 65536ull + (uint32_t)(s & 0xffffffu) % chunk_sect;
        } else if (MODE == MIX) {
            const bool hot = (uint32_t)(s >> 20) % 1000u < hot_permille;
            sec = hot ? __umul64hi(s, hot_sectors) : hot_sectors + __umul64hi(s, n_sectors);
        } else {
            sec = __umul64hi(s, n_sectors);
        }
        uint32_t x;
        if (MODE == STORE) {
            x = (uint32_t)s;
            asm volatile("st.global.v8.u32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" :: "l"(buf + 2 * sec), "r"(x) : "memory");
        } else {
            uint32_t v[8]; ld256(buf + 2 * sec, v); x = v[0] ^ v[7];
            if (MODE == DUP) { uint32_t w[8]; ld256(buf + 2 * sec, w); x ^= w[3]; }
        }
        acc += x;
        s = s * 6364136223846793005ull + 1442695040888963407ull + (MODE == STORE ? 0ull : ((uint64_t)x << 32));
    }
    if (acc == 0xDEADBEEFu) atomicAdd(sink, 1ull);
}

static int g_sms; static cudaEvent_t e0, e1; static unsigned long long *g_sink;
template <int MODE>
static double run(uint4 *buf, uint64_t n_pages, uint32_t chunk_sect, uint32_t ppb, uint64_t n_sectors, uint64_t hot_sectors,
                  uint32_t hot_permille, int blocks_per_sm)
{
    const int block = 256, grid = g_sms * blocks_per_sm;
    int iters = (int)std::max<double>(8, 2.5e8 / ((double)grid * block)), warm = 4;
    probe<MODE><<<grid, block>>>(buf, n_pages, chunk_sect, ppb, n_sectors, hot_sectors, hot_permille, warm, g_sink);
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        CU(cudaEventRecord(e0));
        probe<MODE><<<grid, block>>>(buf, n_pages, chunk_sect, ppb, n_sectors, hot_sectors, hot_permille, iters, g_sink);
        CU(cudaEventRecord(e1)); CU(cudaEventSynchronize(e1));
        float ms = 0; CU(cudaEventElapsedTime(&ms, e0, e1));
        best = std::max(best, (double)grid * block * (double)iters / (ms * 1e-3) / 1e9);
    }
    return best;
}

int main()
{
    CU(cudaSetDevice(0));
    CU(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, 0));
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    CU(cudaMalloc((void **)&g_sink, 8)); CU(cudaMemset(g_sink, 0, 8));
    const size_t MB = 1ull << 20;
    const int bps = 8;                                            // 8 blocks of 256 threads per SM: full occupancy
    const uint32_t ppb = 4;                                       // T1: 4 private pages per block = 32 per SM
    const uint64_t n_pages = (uint64_t)g_sms * bps * ppb;         // 4736 pages = 9.25 GiB
    const size_t big = n_pages * 2 * MB;
    uint4 *buf = nullptr;
    CU(cudaMalloc((void **)&buf, big));
    fill_kernel<<<g_sms * 8, 256>>>(buf, big / 16); CU(cudaDeviceSynchronize());
    for (uint32_t chunk_kb : {16, 64}) {
        const uint32_t cs = chunk_kb * 1024 / 32;
        const double a = run<SHARED_PAGES>(buf, n_pages, cs, ppb, 0, 0, 0, bps);
        const double b = run<BLOCK_PAGES>(buf, n_pages, cs, ppb, 0, 0, 0, bps);
        const double c = run<WARP_PAGE>(buf, n_pages, cs, ppb, 0, 0, 0, bps);
        printf("{\"exp\": \"T1T2\", \"pages\": %llu, \"chunk_kb\": %u, \"footprint_mb\": %.0f, \"shared_pages_gsect\": %.2f, "
               "\"pages_private_to_block_gsect\": %.2f, \"one_page_per_warp_instruction_gsect\": %.2f}\n",
               (unsigned long long)n_pages, chunk_kb, (double)n_pages * chunk_kb / 1024, a, b, c);
        fflush(stdout);
    }
    // T1b: private pages per block, more pages per SM (does the per-SM reach end at 128?)
    for (uint32_t p : {1, 2, 4, 8, 16, 32}) {
        if ((uint64_t)g_sms * bps * p > n_pages) {                // fewer blocks per SM for the large ones
            const int b2 = (int)(n_pages / ((uint64_t)g_sms * p));
            if (b2 < 1) continue;
            const double b = run<BLOCK_PAGES>(buf, n_pages, 512, p, 0, 0, 0, b2);
            printf("{\"exp\": \"T1b\", \"pages_per_block\": %u, \"blocks_per_sm\": %d, \"pages_per_sm\": %u, \"gsect\": %.2f}\n", p, b2, p * b2, b);
        } else {
            const double b = run<BLOCK_PAGES>(buf, n_pages, 512, p, 0, 0, 0, bps);
            printf("{\"exp\": \"T1b\", \"pages_per_block\": %u, \"blocks_per_sm\": %d, \"pages_per_sm\": %u, \"gsect\": %.2f}\n", p, bps, p * bps, b);
        }
        fflush(stdout);
    }
    // T3: duplicate loads, contiguous 1.5 GiB and 64 MB
    for (size_t mb : {64, 1536}) {
        const double a = run<PLAIN>(buf, 0, 0, 0, mb * MB / 32, 0, 0, bps);
        const double d = run<DUP>(buf, 0, 0, 0, mb * MB / 32, 0, 0, bps);
        printf("{\"exp\": \"T3\", \"footprint_mb\": %zu, \"single_load_gsect\": %.2f, \"each_sector_loaded_twice_gsect\": %.2f}\n", mb, a, d);
        fflush(stdout);
    }
    // T4: hot fraction into 128 MB, the rest over 1478 MB (one direction of the 3.1 Gb index)
    for (uint32_t pm : {0, 50, 100, 200, 300, 500, 1000}) {
        const double a = run<MIX>(buf, 0, 0, 0, 1478 * MB / 32, 128 * MB / 32, pm, bps);
        printf("{\"exp\": \"T4\", \"hot_region_mb\": 128, \"cold_region_mb\": 1478, \"hot_permille\": %u, \"gsect\": %.2f}\n", pm, a);
        fflush(stdout);
    }
    // T5: stores
    for (size_t mb : {64, 192, 1536, 4096}) {
        const double a = run<STORE>(buf, 0, 0, 0, mb * MB / 32, 0, 0, bps);
        printf("{\"exp\": \"T5\", \"footprint_mb\": %zu, \"random_32B_stores_gsect\": %.2f}\n", mb, a);
        fflush(stdout);
    }
    return 0;
}

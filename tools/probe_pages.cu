// Is the "HBM random-sector rate" (38-43 G sectors/s past 0.5 GiB, against 280 G/s from L2) a DRAM limit or an address-
// translation limit?  The TLB reach measured on this architecture is 128 entries x 2 MB = 256 MB; the footprint sweep of
// probe_sweep.cu falls off a cliff between 0.25 and 0.5 GiB that the L2 hit rate alone does not explain.  Three experiments:
//   1. a fine footprint sweep across the 256 MB mark (cudaMalloc, dependent chains + 8 independent loads per thread);
//   2. the SAME number of bytes spread over MORE pages: the footprint cut into chunks of C bytes, one chunk per 2 MB page;
//   3. the same sweep over memory obtained in other ways (VMM handle mapped at a 512 MB-aligned address, stream-ordered pool).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/probe_pages tools/probe_pages.cu -lcuda ; tools/probe_pages
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda.h>
#include <cuda_runtime.h>
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)
#define DR(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char *s_ = nullptr; cuGetErrorString(r_, &s_); fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, s_ ? s_ : "?"); exit(1); } } while (0)

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
// sector index s of the logical footprint -> address: chunks of 2^chunk_log sectors, chunk c starts at c * stride_sectors
__device__ __forceinline__ uint64_t place(uint64_t s, int chunk_log, uint64_t stride_sectors)
{
    return (s >> chunk_log) * stride_sectors + (s & ((1ull << chunk_log) - 1));
}
template <int CHAINS>
__global__ void chain_kernel(const uint4 *buf, uint64_t n_sectors, int chunk_log, uint64_t stride_sectors, int iters,
                             unsigned long long *sink)
{
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t s[CHAINS];
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s[c] = mix64(tid * CHAINS + c + 12345);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            const uint64_t u = place(__umul64hi(s[c], n_sectors), chunk_log, stride_sectors);
            uint32_t v[8];
            ld256(buf + u * 2, v);
            const uint32_t x = v[0] ^ v[7];
            acc += x;
            s[c] = s[c] * 6364136223846793005ull + 1442695040888963407ull + ((uint64_t)x << 32);
        }
    }
    if (acc == 0xDEADBEEFu) atomicAdd(sink, 1ull);
}
template <int UNROLL>
__global__ void mlp_kernel(const uint4 *buf, uint64_t n_sectors, int chunk_log, uint64_t stride_sectors, int iters,
                           unsigned long long *sink)
{
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t s = mix64(tid + 999);
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        uint32_t x[UNROLL];
#pragma unroll
        for (int c = 0; c < UNROLL; ++c) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            const uint64_t sec = place(__umul64hi(s, n_sectors), chunk_log, stride_sectors);
            uint32_t v[8]; ld256(buf + 2 * sec, v); x[c] = v[0] ^ v[7];
        }
#pragma unroll
        for (int c = 0; c < UNROLL; ++c) acc += x[c];
    }
    if (acc == 0xDEADBEEFu) atomicAdd(sink, 1ull);
}

static int g_sms = 0;
static cudaEvent_t e0, e1;
static unsigned long long *g_sink = nullptr;

// best of 3 runs of both shapes at full occupancy; returns G sectors/s of the better shape
static double measure(const uint4 *buf, uint64_t n_sectors, int chunk_log, uint64_t stride_sectors, double *chain_out, double *mlp_out)
{
    const int block = 256;
    double res[2] = {0, 0};
    for (int which = 0; which < 2; ++which) {
        const void *fn = which == 0 ? (const void *)chain_kernel<2> : (const void *)mlp_kernel<8>;
        const int spi = which == 0 ? 2 : 8;
        int occ = 0; CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, block, 0));
        const int grid = g_sms * occ;
        int iters = (int)std::max<double>(8, 2.5e8 / ((double)grid * block * spi)), warm = 4;
        void *aw[] = {(void *)&buf, (void *)&n_sectors, (void *)&chunk_log, (void *)&stride_sectors, (void *)&warm, (void *)&g_sink};
        void *ar[] = {(void *)&buf, (void *)&n_sectors, (void *)&chunk_log, (void *)&stride_sectors, (void *)&iters, (void *)&g_sink};
        CU(cudaLaunchKernel(fn, dim3(grid), dim3(block), aw, 0, nullptr));
        for (int rep = 0; rep < 3; ++rep) {
            CU(cudaEventRecord(e0));
            CU(cudaLaunchKernel(fn, dim3(grid), dim3(block), ar, 0, nullptr));
            CU(cudaEventRecord(e1)); CU(cudaEventSynchronize(e1));
            float ms = 0; CU(cudaEventElapsedTime(&ms, e0, e1));
            res[which] = std::max(res[which], (double)grid * block * spi * (double)iters / (ms * 1e-3) / 1e9);
        }
    }
    *chain_out = res[0]; *mlp_out = res[1];
    return std::max(res[0], res[1]);
}

int main()
{
    CU(cudaSetDevice(0)); CU(cudaFree(0));
    CU(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, 0));
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    CU(cudaMalloc((void **)&g_sink, 8)); CU(cudaMemset(g_sink, 0, 8));
    const size_t MB = 1ull << 20;
    const size_t big = 8192 * MB;
    uint4 *buf = nullptr;
    CU(cudaMalloc((void **)&buf, big));
    fill_kernel<<<g_sms * 8, 256>>>(buf, big / 16); CU(cudaDeviceSynchronize());
    printf("{\"exp\": \"ptr\", \"cudaMalloc_8GiB_address\": \"%p\"}\n", (void *)buf);

    // 1. fine footprint sweep, contiguous
    for (size_t mb : {32, 64, 96, 128, 160, 192, 224, 256, 288, 320, 384, 448, 512, 640, 768, 1024, 1536, 2048, 3072, 4096, 8192}) {
        double c, m; measure(buf, mb * MB / 32, 40, 0, &c, &m);
        printf("{\"exp\": \"footprint\", \"alloc\": \"cudaMalloc\", \"footprint_mb\": %zu, \"pages_2mb\": %zu, \"chain2_gsect\": %.2f, \"mlp8_gsect\": %.2f}\n",
               mb, mb / 2, c, m);
        fflush(stdout);
    }
    // 2. same bytes, more pages: chunk of C bytes at the start of every S-byte stride
    struct Sp { size_t foot_mb; size_t chunk_kb; size_t stride_kb; };
    const Sp sps[] = {
        {96, 2048, 2048}, {96, 1024, 2048}, {96, 512, 2048}, {96, 256, 2048}, {96, 128, 2048}, {96, 64, 2048}, {96, 32, 2048},
        {192, 2048, 2048}, {192, 1024, 2048}, {192, 512, 2048}, {192, 256, 2048}, {192, 128, 2048}, {192, 64, 2048},
        {512, 2048, 2048}, {512, 1024, 2048}, {512, 512, 2048}, {512, 256, 2048},
        {1024, 2048, 2048}, {1024, 1024, 2048}, {1024, 512, 2048}, {1024, 256, 2048},
        // 64 KB-page question: the same chunks at a 64 KB stride (one chunk per 64 KB "small page")
        {96, 32, 64}, {96, 16, 64}, {192, 32, 64}, {192, 16, 64},
    };
    for (const Sp &sp : sps) {
        const uint64_t n_sectors = sp.foot_mb * MB / 32;
        int chunk_log = 0; while ((32ull << chunk_log) < sp.chunk_kb * 1024) ++chunk_log;
        const uint64_t stride_sectors = sp.stride_kb * 1024 / 32;
        const uint64_t chunks = n_sectors >> chunk_log;
        if (chunks * stride_sectors * 32 > big) continue;
        double c, m; measure(buf, n_sectors, chunk_log, stride_sectors, &c, &m);
        printf("{\"exp\": \"spread\", \"footprint_mb\": %zu, \"chunk_kb\": %zu, \"stride_kb\": %zu, \"span_mb\": %.0f, \"pages_2mb_touched\": %.0f, "
               "\"chain2_gsect\": %.2f, \"mlp8_gsect\": %.2f}\n", sp.foot_mb, sp.chunk_kb, sp.stride_kb,
               (double)chunks * stride_sectors * 32 / MB, std::max(1.0, (double)chunks * stride_sectors * 32 / (2.0 * MB)) * (sp.stride_kb >= 2048 ? 1 : 1), c, m);
        fflush(stdout);
    }
    CU(cudaFree(buf));

    // 3. other ways to get memory
    CUdevice dev; DR(cuDeviceGet(&dev, 0));
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = 0;
    size_t gmin = 0, grec = 0;
    DR(cuMemGetAllocationGranularity(&gmin, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
    DR(cuMemGetAllocationGranularity(&grec, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    printf("{\"exp\": \"vmm_granularity\", \"minimum\": %zu, \"recommended\": %zu}\n", gmin, grec);
    const size_t vsz = 4096 * MB;
    for (size_t align_mb : {2, 512, 0}) {                 // 0: one 512 MB handle per 512 MB of address space
        CUdeviceptr va = 0;
        const size_t align = (align_mb ? align_mb : 512) * MB;
        if (cuMemAddressReserve(&va, vsz, align, 0, 0) != CUDA_SUCCESS) { printf("{\"exp\": \"vmm\", \"align_mb\": %zu, \"error\": \"reserve\"}\n", align_mb); continue; }
        std::vector<CUmemGenericAllocationHandle> hs;
        const size_t piece = align_mb ? vsz : 512 * MB;
        bool ok = true;
        for (size_t off = 0; off < vsz && ok; off += piece) {
            CUmemGenericAllocationHandle h;
            if (cuMemCreate(&h, piece, &prop, 0) != CUDA_SUCCESS) { ok = false; break; }
            hs.push_back(h);
            if (cuMemMap(va + off, piece, 0, h, 0) != CUDA_SUCCESS) { ok = false; break; }
        }
        CUmemAccessDesc ad = {}; ad.location = prop.location; ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        if (ok && cuMemSetAccess(va, vsz, &ad, 1) != CUDA_SUCCESS) ok = false;
        if (!ok) { printf("{\"exp\": \"vmm\", \"align_mb\": %zu, \"error\": \"create/map\"}\n", align_mb); continue; }
        uint4 *vb = (uint4 *)va;
        fill_kernel<<<g_sms * 8, 256>>>(vb, vsz / 16); CU(cudaDeviceSynchronize());
        for (size_t mb : {192, 512, 1536, 4096}) {
            double c, m; measure(vb, mb * MB / 32, 40, 0, &c, &m);
            printf("{\"exp\": \"footprint\", \"alloc\": \"vmm_align%zuMB%s\", \"address\": \"%p\", \"footprint_mb\": %zu, \"chain2_gsect\": %.2f, \"mlp8_gsect\": %.2f}\n",
                   align_mb ? align_mb : 512, align_mb ? "" : "_512MBhandles", (void *)vb, mb, c, m);
            fflush(stdout);
        }
        DR(cuMemUnmap(va, vsz));
        for (auto h : hs) DR(cuMemRelease(h));
        DR(cuMemAddressFree(va, vsz));
    }
    {   // stream-ordered pool
        uint4 *pb = nullptr;
        CU(cudaMallocAsync((void **)&pb, vsz, 0)); CU(cudaStreamSynchronize(0));
        fill_kernel<<<g_sms * 8, 256>>>(pb, vsz / 16); CU(cudaDeviceSynchronize());
        for (size_t mb : {192, 512, 1536, 4096}) {
            double c, m; measure(pb, mb * MB / 32, 40, 0, &c, &m);
            printf("{\"exp\": \"footprint\", \"alloc\": \"cudaMallocAsync\", \"address\": \"%p\", \"footprint_mb\": %zu, \"chain2_gsect\": %.2f, \"mlp8_gsect\": %.2f}\n",
                   (void *)pb, mb, c, m);
            fflush(stdout);
        }
        CU(cudaFreeAsync(pb, 0)); CU(cudaStreamSynchronize(0));
    }
    {   // managed memory, prefetched to the device (UVM builds its own page tables and may pick other page sizes)
        uint4 *mb_ = nullptr;
        if (cudaMallocManaged((void **)&mb_, vsz) == cudaSuccess) {
            cudaMemLocation loc = {}; loc.type = cudaMemLocationTypeDevice; loc.id = 0;
            cudaMemAdvise(mb_, vsz, cudaMemAdviseSetPreferredLocation, loc);
            cudaMemPrefetchAsync(mb_, vsz, loc, 0, 0);
            CU(cudaDeviceSynchronize());
            fill_kernel<<<g_sms * 8, 256>>>(mb_, vsz / 16); CU(cudaDeviceSynchronize());
            for (size_t mb : {192, 512, 1536, 4096}) {
                double c, m; measure(mb_, mb * MB / 32, 40, 0, &c, &m);
                printf("{\"exp\": \"footprint\", \"alloc\": \"cudaMallocManaged+prefetch\", \"address\": \"%p\", \"footprint_mb\": %zu, \"chain2_gsect\": %.2f, \"mlp8_gsect\": %.2f}\n",
                       (void *)mb_, mb, c, m);
                fflush(stdout);
            }
            CU(cudaFree(mb_));
        }
    }
    return 0;
}

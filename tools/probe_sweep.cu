// Random-sector probe sweep: which access shape / occupancy / footprint gives the highest
// random 32-byte-sector rate on this GPU, so that the roofline denominator is an upper bound for the shipped kernels.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/probe_sweep tools/probe_sweep.cu ; tools/probe_sweep
// Prints one JSON line per (footprint, variant): sectors/s, GB/s at 32 B per sector, units/s.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
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
// CHAINS dependent chains per thread; each step loads SECT adjacent 32-byte sectors of one aligned unit of SECT*32 bytes;
// the next unit of a chain is a hash of the data loaded (like an LF walk).
template <int CHAINS, int SECT>
__global__ void chain_kernel(const uint4 *buf, uint64_t n_units, int iters, unsigned long long *sink)
{
    extern __shared__ char dummy[];
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t s[CHAINS];
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s[c] = mix64(tid * CHAINS + c + 12345);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            const uint64_t u = __umul64hi(s[c], n_units);
            uint32_t x = 0;
#pragma unroll
            for (int k = 0; k < SECT; ++k) {
                uint32_t v[8];
                ld256(buf + (u * SECT + k) * 2, v);
                x ^= v[0] ^ v[7];
            }
            acc += x;
            s[c] = s[c] * 6364136223846793005ull + 1442695040888963407ull + ((uint64_t)x << 32);
        }
    }
    if (acc == 0xDEADBEEFu) atomicAdd(sink, 1ull);
}
// independent loads: UNROLL in flight per thread, addresses from an LCG
template <int UNROLL, int BYTES>
__global__ void mlp_kernel(const uint4 *buf, uint64_t n_sectors, int iters, unsigned long long *sink)
{
    extern __shared__ char dummy[];
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t s = mix64(tid + 999);
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        uint32_t x[UNROLL];
#pragma unroll
        for (int c = 0; c < UNROLL; ++c) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            const uint64_t sec = __umul64hi(s, n_sectors);
            if (BYTES == 32) { uint32_t v[8]; ld256(buf + 2 * sec, v); x[c] = v[0] ^ v[7]; }
            else if (BYTES == 16) { uint4 a = __ldg(buf + 2 * sec + (c & 1)); x[c] = a.x ^ a.w; }
            else { x[c] = __ldg(reinterpret_cast<const uint32_t *>(buf + 2 * sec) + (c & 7)); }
        }
#pragma unroll
        for (int c = 0; c < UNROLL; ++c) acc += x[c];
    }
    if (acc == 0xDEADBEEFu) atomicAdd(sink, 1ull);
}

// SA-walk shape: one dependent chain per thread whose length is geometric (p = 1/8); a finished chain ends with a 4-byte
// load from a second array and the thread takes the next start from a work queue (warp-aggregated atomic)
__global__ void __launch_bounds__(256) walk_kernel(const uint4 *buf, uint64_t n_units, int iters, unsigned long long *sink)
{
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t s = mix64(tid + 4242);
    uint32_t acc = 0;
    const uint32_t *tail = reinterpret_cast<const uint32_t *>(buf + 2 * n_units);      // second array behind the first
    for (int it = 0; it < iters; ++it) {
        const uint64_t u = __umul64hi(s, n_units);
        uint32_t v[8];
        ld256(buf + u * 2, v);
        const uint32_t x = v[0] ^ v[7];
        s = s * 6364136223846793005ull + 1442695040888963407ull + ((uint64_t)x << 32);
        if ((x & 7u) == 0u) {                                   // chain ends: terminal 4-byte load, new start
            acc += __ldg(tail + (__umul64hi(s, n_units) * 8));
            s = mix64(s + it);
        }
        acc += x;
    }
    if (acc == 0xDEADBEEFu) atomicAdd(sink, 1ull);
}

struct Variant { const char *name; const void *fn; int sect_per_iter; int unit_sect; int threads_per_sm; };

int main(int argc, char **argv)
{
    int dev = 0; CU(cudaSetDevice(dev));
    int sms = 0; CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const double gib[] = {0.03125, 0.25, 0.5, 1.0, 1.4437, 2.0, 2.887, 4.0, 8.0};   // 1.4437 GiB = one direction of the 3.1 Gb index
    const size_t max_bytes = (size_t)(8.0 * (1ull << 30));
    uint4 *buf = nullptr; unsigned long long *sink = nullptr;
    CU(cudaMalloc((void **)&buf, max_bytes)); CU(cudaMalloc((void **)&sink, 8)); CU(cudaMemset(sink, 0, 8));
    fill_kernel<<<sms * 8, 256>>>(buf, max_bytes / 16); CU(cudaDeviceSynchronize());
    std::vector<Variant> vs;
    for (int t : {256, 512, 1024, 2048}) {
        vs.push_back({"chain1x32B", (const void *)chain_kernel<1, 1>, 1, 1, t});
        vs.push_back({"chain2x32B", (const void *)chain_kernel<2, 1>, 2, 1, t});
        vs.push_back({"chain4x32B", (const void *)chain_kernel<4, 1>, 4, 1, t});
    }
    for (int t : {1024, 2048}) {
        vs.push_back({"chain1x64B", (const void *)chain_kernel<1, 2>, 2, 2, t});
        vs.push_back({"chain2x64B", (const void *)chain_kernel<2, 2>, 4, 2, t});
        vs.push_back({"chain2x128B", (const void *)chain_kernel<2, 4>, 8, 4, t});
        vs.push_back({"mlp4x32B", (const void *)mlp_kernel<4, 32>, 4, 1, t});
        vs.push_back({"mlp8x32B", (const void *)mlp_kernel<8, 32>, 8, 1, t});
        vs.push_back({"mlp8x16B", (const void *)mlp_kernel<8, 16>, 8, 1, t});
        vs.push_back({"mlp8x4B", (const void *)mlp_kernel<8, 4>, 8, 1, t});
    }
    cudaEvent_t e0, e1; CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    const int block = 256;
    if (argc > 1) {
        // carve-out sweep: no dynamic shared memory, explicit PreferredSharedMemoryCarveout (percent), full occupancy
        struct CV { const char *name; const void *fn; int spi; int half; } cvs[] = {
            {"chain1x32B", (const void *)chain_kernel<1, 1>, 1, 0}, {"chain2x32B", (const void *)chain_kernel<2, 1>, 2, 0},
            {"mlp4x32B", (const void *)mlp_kernel<4, 32>, 4, 0}, {"mlp8x32B", (const void *)mlp_kernel<8, 32>, 8, 0},
            {"walk(sa-like)", (const void *)walk_kernel, 1, 1}};
        for (double g : {0.03125, 1.4437, 2.887}) {
            for (const CV &v : cvs) for (int carve : {-1, 0, 25, 50, 75, 100}) {
                const uint64_t n_sectors = (uint64_t)(g * (1ull << 30)) / 32;
                uint64_t n_units = v.half ? n_sectors / 2 : n_sectors;
                CU(cudaFuncSetAttribute(v.fn, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
                int occ = 0; CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, v.fn, block, 0));
                const int grid = sms * occ;
                int iters = (int)std::max<double>(8, 3e8 / ((double)grid * block * v.spi)), warm = 4;
                void *aw[] = {(void *)&buf, (void *)&n_units, (void *)&warm, (void *)&sink};
                void *ar[] = {(void *)&buf, (void *)&n_units, (void *)&iters, (void *)&sink};
                CU(cudaLaunchKernel(v.fn, dim3(grid), dim3(block), aw, 0, nullptr));
                double best = 0;
                for (int rep = 0; rep < 3; ++rep) {
                    CU(cudaEventRecord(e0));
                    CU(cudaLaunchKernel(v.fn, dim3(grid), dim3(block), ar, 0, nullptr));
                    CU(cudaEventRecord(e1)); CU(cudaEventSynchronize(e1));
                    float ms = 0; CU(cudaEventElapsedTime(&ms, e0, e1));
                    const double sect = (double)grid * block * (double)iters * (v.half ? 1.125 : v.spi);   // walk: + 1/8 tail sector per step
                    best = std::max(best, sect / (ms * 1e-3));
                }
                printf("{\"sweep\": \"carveout\", \"footprint_gib\": %.4f, \"variant\": \"%s\", \"carveout_pct\": %d, \"resident_threads_per_sm\": %d, "
                       "\"gsectors_per_s\": %.2f, \"gbs_32B\": %.1f}\n", g, v.name, carve, occ * block, best / 1e9, best * 32 / 1e9);
                fflush(stdout);
            }
        }
        return 0;
    }
    for (double g : gib) {
        const uint64_t n_sectors = (uint64_t)(g * (1ull << 30)) / 32;
        for (const Variant &v : vs) {
            // cap resident blocks per SM with dynamic shared memory: 227 KB / blocks
            const int blocks_per_sm = v.threads_per_sm / block;
            size_t smem = std::min<size_t>(200 * 1024 / blocks_per_sm, 100 * 1024);
            smem = (smem / 1024) * 1024;
            CU(cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int occ = 0; CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, v.fn, block, smem));
            occ = std::min(occ, blocks_per_sm);
            const int grid = sms * occ;
            uint64_t n_units = n_sectors / v.unit_sect;
            // aim at ~3e8 sectors per run
            int iters = (int)std::max<double>(8, 3e8 / ((double)grid * block * v.sect_per_iter));
            int warm = 4;
            void *aw[] = {(void *)&buf, (void *)&n_units, (void *)&warm, (void *)&sink};
            void *ar[] = {(void *)&buf, (void *)&n_units, (void *)&iters, (void *)&sink};
            CU(cudaLaunchKernel(v.fn, dim3(grid), dim3(block), aw, smem, nullptr));
            double best = 0;
            for (int rep = 0; rep < 3; ++rep) {
                CU(cudaEventRecord(e0));
                CU(cudaLaunchKernel(v.fn, dim3(grid), dim3(block), ar, smem, nullptr));
                CU(cudaEventRecord(e1)); CU(cudaEventSynchronize(e1));
                float ms = 0; CU(cudaEventElapsedTime(&ms, e0, e1));
                best = std::max(best, (double)grid * block * v.sect_per_iter * (double)iters / (ms * 1e-3));
            }
            printf("{\"footprint_gib\": %.4f, \"variant\": \"%s\", \"threads_per_sm\": %d, \"resident_threads_per_sm\": %d, "
                   "\"gsectors_per_s\": %.2f, \"gbs_32B\": %.1f, \"gunits_per_s\": %.2f, \"unit_bytes\": %d}\n",
                   g, v.name, v.threads_per_sm, occ * block, best / 1e9, best * 32 / 1e9, best / v.unit_sect / 1e9, v.unit_sect * 32);
            fflush(stdout);
        }
    }
    return 0;
}

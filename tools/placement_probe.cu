// placement_probe.cu -- does the random-sector read rate depend on WHERE a table sits in HBM?
// (measurement tooling)  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o placement_probe placement_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void __launch_bounds__(256) gather(const char *table, uint64_t n_units, uint32_t per, unsigned long long *sink)
{
    uint64_t x = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
    uint64_t acc = 0;
    for (uint32_t it = 0; it < per; it += 4) {
        uint64_t v[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            x = x * 6364136223846793005ull + 1442695040888963407ull;
            const char *p = table + __umul64hi(x ^ (x >> 29), n_units) * 32;
            uint64_t a, b, c, d;
            asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
            v[m] = a ^ b ^ c ^ d;
        }
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 4; ++m) acc = (acc ^ v[m]) * 0x9E3779B97F4A7C15ull;
    }
    if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

static double rate(const char *table, size_t bytes, unsigned long long *sink, int sms)
{
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(a));
        gather<<<sms * 8, 256>>>(table, bytes / 32, 512, sink);
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (rep && ms < best) best = ms;
    }
    return (double)sms * 8 * 256 * 512 / (best * 1e-3) / 1e9;
}

int main(int argc, char **argv)
{
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    unsigned long long *sink; CK(cudaMalloc(&sink, 8));
    const size_t GiB = (size_t)1 << 30;
    printf("-- successive 1 GiB allocations (all kept)\n");
    char *t[12];
    for (int i = 0; i < 12; ++i) {
        CK(cudaMalloc(&t[i], GiB)); CK(cudaMemset(t[i], 1, GiB));
        printf("alloc %2d at %p: %.2f G loads/s\n", i, (void *)t[i], rate(t[i], GiB, sink, sms));
    }
    for (int i = 0; i < 12; ++i) CK(cudaFree(t[i]));
    printf("-- 1 GiB windows inside one 16 GiB allocation, by offset\n");
    char *big; CK(cudaMalloc(&big, 16 * GiB)); CK(cudaMemset(big, 1, 16 * GiB));
    for (size_t off = 0; off + GiB <= 16 * GiB; off += GiB / 2) printf("offset %5zu MiB: %.2f G loads/s\n", off >> 20, rate(big + off, GiB, sink, sms));
    printf("-- window size at offset 0 of the 16 GiB allocation\n");
    for (size_t sz = GiB / 4; sz <= 16 * GiB; sz *= 2) printf("size %6zu MiB: %.2f G loads/s\n", sz >> 20, rate(big, sz, sink, sms));
    printf("-- odd offsets (1 GiB window)\n");
    for (size_t off : {(size_t)2 << 20, (size_t)34 << 20, (size_t)130 << 20, (size_t)258 << 20, (size_t)770 << 20})
        printf("offset %5zu MiB: %.2f G loads/s\n", off >> 20, rate(big + off, GiB, sink, sms));
    return 0;
}

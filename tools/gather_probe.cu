// gather_probe.cu -- how much DRAM traffic does one random sector read cost on B200, and does
// the kind of load change it?  Measurement tooling for DESIGN.md (not part of the product).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o gather_probe gather_probe.cu
//   run:   ./gather_probe [table_MiB]          (events: G loads/s per variant)
//   ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum ./gather_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t next_addr(uint64_t &x, uint64_t n_units)
{
    x = x * 6364136223846793005ull + 1442695040888963407ull;
    return __umul64hi(x ^ (x >> 29), n_units);
}

template <int V> __device__ __forceinline__ uint64_t load_variant(const char *p, char *smem_slot, uint64_t *bar)
{
    uint64_t a = 0, b = 0, c = 0, d = 0;
    if (V == 0) asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if (V == 1) { asm volatile("ld.global.nc.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
                  asm volatile("ld.global.nc.v2.u64 {%0,%1}, [%2];" : "=l"(c), "=l"(d) : "l"(p + 16)); }
    if (V == 2) asm volatile("ld.global.nc.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
    if (V == 3) asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(a) : "l"(p));
    if (V == 4) asm volatile("ld.global.cv.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
    if (V == 5) asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
    if (V == 6) asm volatile("ld.global.L2::64B.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
    if (V == 7) asm volatile("ld.global.L1::no_allocate.L2::evict_first.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if (V == 8) asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
    if (V == 9) asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
    if (V == 10) { // LDGSTS 16 B, L2 only
        uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_slot);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(p));
    }
    if (V == 11) asm volatile("ld.global.nc.L1::no_allocate.L2::evict_last.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if (V == 13) { asm volatile("ld.global.nc.L2::64B.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
                   asm volatile("ld.global.nc.L2::64B.v2.u64 {%0,%1}, [%2];" : "=l"(c), "=l"(d) : "l"(p + 16)); }
    if (V == 14) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if (V == 15) asm volatile("ld.global.nc.L2::128B.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
    if (V == 12) asm volatile("ld.global.nc.L1::evict_last.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
    return a ^ b ^ c ^ d;
}

template <int V, int MLP> __global__ void __launch_bounds__(256) gather(const char *table, uint64_t n_units, uint32_t per, unsigned long long *sink)
{
    __shared__ __align__(16) char smem[256 * MLP * 16];
    uint64_t x = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
    uint64_t acc = 0;
    for (uint32_t it = 0; it < per; it += MLP) {
        uint64_t v[MLP];
#pragma unroll
        for (int m = 0; m < MLP; ++m) v[m] = load_variant<V>(table + next_addr(x, n_units) * 32, smem + (threadIdx.x * MLP + m) * 16, nullptr);
        if (V == 10) { asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
            for (int m = 0; m < MLP; ++m) v[m] = *reinterpret_cast<uint64_t *>(smem + (threadIdx.x * MLP + m) * 16); }
        __syncwarp();
#pragma unroll
        for (int m = 0; m < MLP; ++m) acc = (acc ^ v[m]) * 0x9E3779B97F4A7C15ull;
    }
    if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

// TMA bulk copy of one 32-byte sector per thread-iteration (issued by every thread, one mbarrier per warp-iteration)
template <int MLP> __global__ void __launch_bounds__(256) gather_bulk(const char *table, uint64_t n_units, uint32_t per, unsigned long long *sink)
{
    __shared__ __align__(128) char smem[256 * MLP * 32];
    __shared__ __align__(8) uint64_t bar;
    uint64_t x = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
    uint64_t acc = 0;
    const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 256;" :: "r"(bar_s));
    __syncthreads();
    uint32_t phase = 0;
    for (uint32_t it = 0; it < per; it += MLP) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_s), "r"(32 * MLP) : "memory");
#pragma unroll
        for (int m = 0; m < MLP; ++m) {
            uint32_t s = (uint32_t)__cvta_generic_to_shared(smem + (threadIdx.x * MLP + m) * 32);
            const char *p = table + next_addr(x, n_units) * 32;
            asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], 32, [%2];" :: "r"(s), "l"(p), "r"(bar_s) : "memory");
        }
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar_s), "r"(phase) : "memory");
        phase ^= 1;
#pragma unroll
        for (int m = 0; m < MLP; ++m) acc = (acc ^ *reinterpret_cast<uint64_t *>(smem + (threadIdx.x * MLP + m) * 32)) * 0x9E3779B97F4A7C15ull;
        __syncthreads();
    }
    if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

template <int V> static void run(const char *name, const char *table, uint64_t n_units, unsigned long long *sink, int sms)
{
    const int blocks = sms * 8;
    const uint32_t per = 1024;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(a));
        if (V == 100) gather_bulk<4><<<blocks, 256>>>(table, n_units, per, sink);
        else gather<(V == 100 ? 0 : V), 4><<<blocks, 256>>>(table, n_units, per, sink);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (rep && ms < best) best = ms;
    }
    printf("%-44s %8.2f G loads/s  (%.3f ms)\n", name, (double)blocks * 256 * per / (best * 1e-3) / 1e9, best);
}

int main(int argc, char **argv)
{
    size_t mib = argc > 1 ? atol(argv[1]) : 1024;
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    char *table; unsigned long long *sink;
    CK(cudaMalloc(&table, mib << 20)); CK(cudaMemset(table, 1, mib << 20)); CK(cudaMalloc(&sink, 8));
    uint64_t n_units = (mib << 20) / 32;
    printf("table %zu MiB, %d SMs, 4 loads in flight per thread, 2048 threads/SM\n", mib, sms);
    run<0>("V0  ld.nc.L1::no_allocate.v4.u64 (32 B)", table, n_units, sink, sms);
    run<1>("V1  2 x ld.nc.v2.u64 (32 B)", table, n_units, sink, sms);
    run<2>("V2  ld.nc.v2.u64 (16 B)", table, n_units, sink, sms);
    run<3>("V3  ld.nc.u64 (8 B)", table, n_units, sink, sms);
    run<4>("V4  ld.cv.v2.u64 (16 B)", table, n_units, sink, sms);
    run<5>("V5  ld.cg.v2.u64 (16 B)", table, n_units, sink, sms);
    run<6>("V6  ld.L2::64B.v2.u64 (16 B)", table, n_units, sink, sms);
    run<7>("V7  ld.L1::no_allocate.L2::evict_first.v4.u64 (32 B)", table, n_units, sink, sms);
    run<8>("V8  ld.volatile.v2.u64 (16 B)", table, n_units, sink, sms);
    run<9>("V9  ld.relaxed.gpu.v2.u64 (16 B)", table, n_units, sink, sms);
    run<10>("V10 cp.async.cg 16 B (LDGSTS)", table, n_units, sink, sms);
    run<11>("V11 ld.nc.L2::evict_last.v4.u64 (32 B)", table, n_units, sink, sms);
    run<12>("V12 ld.nc.L1::evict_last.v2.u64 (16 B)", table, n_units, sink, sms);
    run<13>("V13 2 x ld.nc.L2::64B.v2.u64 (32 B)", table, n_units, sink, sms);
    run<14>("V14 ld.nc.L1::no_allocate.L2::64B.v4.u64 (32 B)", table, n_units, sink, sms);
    run<15>("V15 ld.nc.L2::128B.v2.u64 (16 B)", table, n_units, sink, sms);
    run<100>("V100 cp.async.bulk 32 B (TMA)", table, n_units, sink, sms);
    return 0;
}

// debug microbenchmark: raw streaming bandwidth of the per-warp TMA ring (cp.async.bulk + mbarrier) used by the decode kernels
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void * p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t * bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t * bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t * bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void * d, const void * s, uint32_t bytes, uint64_t * bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(smem_u32(d)), "l"(s), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// every SM streams a contiguous region of `per_sm` bytes, chunks dealt round-robin to its warps; `spin` cycles of fake work per chunk
__global__ void __launch_bounds__(512, 1) ring_kernel(const uint8_t * buf, size_t per_sm, int chunk, int depth, int spin, unsigned long long * out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar[16][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (lane == 0) for (int s = 0; s < depth; s++) mbar_init(&bar[warp][s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const uint8_t * base = buf + (size_t)blockIdx.x * per_sm;
    const int nchunks = (int)(per_sm / chunk);
    uint8_t * ring = smem + (size_t)warp * depth * chunk;
    int issued = warp, ni = 0;
    if (lane == 0) for (int s = 0; s < depth && issued < nchunks; s++, issued += nw, ni++) { mbar_expect_tx(&bar[warp][s], chunk); bulk_g2s(ring + s * chunk, base + (size_t)issued * chunk, chunk, &bar[warp][s]); }
    unsigned acc = 0; int nc = 0;
    for (int c = warp; c < nchunks; c += nw, nc++) {
        const int pos = nc % depth;
        mbar_wait(&bar[warp][pos], (nc / depth) & 1);
        acc += ring[pos * chunk + lane * 16];
        if (spin) { const long long t0 = clock64(); while (clock64() - t0 < spin) { } }
        __syncwarp();
        if (lane == 0 && issued < nchunks) { mbar_expect_tx(&bar[warp][pos], chunk); bulk_g2s(ring + pos * chunk, base + (size_t)issued * chunk, chunk, &bar[warp][pos]); issued += nw; }
    }
    if (acc == 0xffffffffu) out[0] = acc;
}
__global__ void ldg_kernel(const uint4 * buf, size_t n, unsigned long long * out) {
    unsigned acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { uint4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(buf + i)); acc += v.x ^ v.w; }
    if (acc == 0x12345u) out[0] = acc;
}
int main() {
    const size_t total = (size_t)4 << 30;
    uint8_t * buf; unsigned long long * out;
    cudaMalloc(&buf, total); cudaMalloc(&out, 64); cudaMemset(buf, 1, total);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    auto run_ring = [&](int warps, int chunk, int depth, int spin, size_t per_sm) {
        const size_t smem = (size_t)warps * depth * chunk;
        if (smem > 220 * 1024) return;
        per_sm = per_sm / chunk * chunk;
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            // walk the 4 GB buffer in launches of sms*per_sm bytes (like phases), back to back
            size_t off = 0; int launches = 0;
            while (off + (size_t)sms * per_sm <= total && launches < 64) { ring_kernel<<<sms, warps * 32, smem>>>(buf + off, per_sm, chunk, depth, spin, out); off += (size_t)sms * per_sm; launches++; }
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const float gbs = (float)off / ms / 1e6f;
            if (ms < best) { best = ms; printf("ring warps=%2d chunk=%5d depth=%d spin=%4d per_sm=%8zu KB  launches=%2d  %.0f GB/s (%.1f us/launch)\n", warps, chunk, depth, spin, per_sm >> 10, launches, gbs, ms * 1e3f / launches); }
        }
    };
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0); ldg_kernel<<<sms * 8, 512>>>((const uint4 *)buf, total / 16, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); printf("ldg.128 stream of 4 GB: %.0f GB/s\n", (float)total / ms / 1e6f);
    }
    const size_t big = (size_t)24 << 20;   // 24 MB per SM per launch: steady state
    run_ring(10, 9216, 2, 0, big); run_ring(10, 9216, 2, 1500, big); run_ring(10, 9216, 2, 3000, big);
    run_ring(16, 4608, 2, 0, big); run_ring(16, 6720, 2, 0, big); run_ring(10, 4608, 4, 0, big); run_ring(10, 2304, 8, 0, big); run_ring(5, 18432, 2, 0, big);
    // phase-sized launches: 440 KB per SM (gate/up), 170 KB (QKV)
    run_ring(10, 9216, 2, 0, 442368); run_ring(10, 9216, 2, 1500, 442368); run_ring(10, 9216, 2, 0, 165888); run_ring(10, 9216, 2, 1500, 165888);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}

// simt.h — TEST INFRASTRUCTURE.  A minimal SIMT emulation so that CUDA kernel SOURCE (the wide path's kernels: llama-box_b200/csrc/
// mmvq_ext_kernels.cuh, fattn_ext_kernels.cuh, with the warp quantisers of actquant.cuh) runs on the CPU:
//   * one OS thread per CUDA thread of a block (reused across the grid); blocks of a grid run one after the other,
//   * __syncthreads = a block-wide barrier, __shfl_*_sync = a per-warp exchange buffer between two warp-wide barriers,
//   * __shared__ = static storage (valid because blocks run sequentially), dynamic shared memory = one buffer per launch,
//   * threadIdx / blockIdx = thread-local, gridDim / blockDim = per launch; round-to-nearest intrinsics = plain IEEE operations
//     (compile with -ffp-contract=off).
// Nothing here is fast; grids in tests/test_kernel_simt.py are a handful of blocks.  Include AFTER the CUDA headers it overrides.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#include <barrier>
#include <cmath>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

namespace simt {
struct Warp { std::barrier<> bar{32}; uint32_t buf[32]; };
struct Launch { dim3 grid, block; std::barrier<> * cta = nullptr; std::vector<std::unique_ptr<Warp>> warps; uint8_t * dyn = nullptr; };
inline Launch g;
struct TL { uint3 tid, bid; };
inline thread_local TL tl;

template <typename T> inline T shfl(T v, int src_lane) {
    static_assert(sizeof(T) == 4, "32-bit shuffles only");
    Warp & w = *g.warps[tl.tid.x >> 5];
    const int lane = tl.tid.x & 31;
    std::memcpy(&w.buf[lane], &v, 4);
    w.bar.arrive_and_wait();
    T r; std::memcpy(&r, &w.buf[src_lane & 31], 4);
    w.bar.arrive_and_wait();
    return r;
}
// run kernel(args...) for every thread of every block; blockDim.x must be a multiple of 32
template <typename F> void launch(dim3 grid, dim3 block, size_t dyn_bytes, F body) {
    std::vector<uint8_t> dyn(dyn_bytes + 64);
    g.grid = grid; g.block = block;
    g.dyn = (uint8_t *)(((uintptr_t)dyn.data() + 15) & ~(uintptr_t)15);
    g.warps.clear();
    for (unsigned w = 0; w < block.x / 32; w++) g.warps.emplace_back(new Warp());
    // one OS thread per CUDA thread of a block, reused for every block of the grid: the blocks run one after the other with a block-wide barrier in
    // between (so static __shared__ storage is never shared by two blocks).  A thread that leaves the kernel early simply waits at that barrier — valid
    // for kernels whose early exits are either uniform over the block or not followed by __syncthreads (true of every kernel run here, and a CUDA rule).
    std::barrier<> cta((std::ptrdiff_t)block.x);
    g.cta = &cta;
    std::vector<std::thread> th;
    for (unsigned t = 0; t < block.x; t++) th.emplace_back([=, &cta] {
        for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++) {
            tl.tid = { t, 0, 0 }; tl.bid = { bx, by, 0 };
            body();
            cta.arrive_and_wait();
        }
    });
    for (auto & x : th) x.join();
}
} // namespace simt

// ---- the CUDA surface the kernels use -----------------------------------------------------------------------------------------
#undef __shared__
#define __shared__ static
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __global__
#define __global__
#define threadIdx (simt::tl.tid)
#define blockIdx  (simt::tl.bid)
#define gridDim   (simt::g.grid)
#define blockDim  (simt::g.block)
#define B200_DYN_SMEM(name) uint8_t * name = simt::g.dyn

inline void __syncthreads() { simt::g.cta->arrive_and_wait(); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int o) { return simt::shfl(v, (int)((simt::tl.tid.x & 31) ^ (unsigned)o)); }
template <typename T> inline T __shfl_sync(unsigned, T v, int src) { return simt::shfl(v, src); }
inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline int   __float2int_rn(float a) { return (int)lrintf(a); }

// the device helpers of common.cuh that sit behind #ifdef __CUDACC__
inline void pdl_wait() {}
inline void pdl_trigger() {}
inline float warp_sum(float v) { for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; }
inline float warp_max(float v) { for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }
// common.cuh's silu_x86 (behind #ifdef __CUDACC__ there): the x86 vector polynomial of ggml_v_expf with an FMA at every step — restated for the emulation
inline float silu_x86(float x) {
    const float nx = 0.0f - x, r = 0x1.8p23f;
    const float z = fmaf(nx, 0x1.715476p+0f, r), n = z - r;
    const float b = fmaf(-n, 0x1.7f7d1cp-20f, fmaf(-n, 0x1.62e4p-1f, nx));
    const float u = b * b;
    const float j = fmaf(fmaf(fmaf(0x1.0e4020p-7f, b, 0x1.573e2ep-5f), u, fmaf(0x1.555e66p-3f, b, 0x1.fffdb6p-2f)), u, fmaf(0x1.ffffecp-1f, b, 1.0f));
    float e = ldexpf(j, (int)n);
    if (fabsf(n) > 192.0f) e = n <= 0.0f ? 0.0f : INFINITY;
    return x / (1.0f + e);
}


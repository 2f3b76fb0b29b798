// mmvq_ext_kernels.cuh — device code of the wide matvec / MUL_MAT_ID / quantised GET_ROWS kernels (launch code: mmvq_ext.cu).
// Kept in a header so that tests/hostsim/kernsim.cpp can run THE SAME kernel source on the CPU under a small SIMT emulation
// (one OS thread per CUDA thread, barriers for __syncthreads / warp shuffles): shared-memory layout, prologue quantiser, row loops,
// reductions and expert indexing are exercised without a GPU (tests/test_kernel_simt.py).
#pragma once
#include "common.cuh"

#include "actquant.cuh"
#include "actquant_ext.cuh"
#include "extfmt.cuh"

#ifndef B200_DYN_SMEM
#  define B200_DYN_SMEM(name) extern __shared__ __align__(16) uint8_t name[]
#endif

namespace {

struct ExtArgs {
    const uint8_t * W; int64_t row_bytes, nb_layout, expert_stride;          // weights; expert_stride (bytes) only with ids
    const int32_t * ids; int64_t ids_tok_stride; int32_t n_used, n_expert;   // MUL_MAT_ID routing (ids == nullptr: plain MUL_MAT)
    const float * x; int64_t x_col_stride, x_slot_stride, n_b1;              // activations (floats); slot stride / n_b1 only with ids
    float * dst; int64_t dst_col_stride, dst_slot_stride;
    const float * bias; const float * residual; int64_t res_col_stride;
    int64_t m, k, prob0;                                                     // prob0: first problem of this launch (grid.y chunking)
    int32_t cols_per_cta, _pad; int64_t ncols;                               // plain MUL_MAT: a problem = a GROUP of up to 8 columns sharing one pass over the weights
};

constexpr int EXT_MAX_COLS = 8;

constexpr int EXT_WARPS = 8;

__host__ __device__ inline int64_t ext_kp(int64_t k) { return (k + 255) & ~(int64_t)255; }
// shared-memory column: [qs kp][d][s (family 1)][bs]
__host__ __device__ inline int64_t ext_off_d(int64_t kp) { return kp; }
__host__ __device__ inline int64_t ext_off_s(int fam, int64_t kp) { return ext_off_d(kp) + (fam ? kp / 32 * 4 : align16(kp / 256 * 4)); }
__host__ __device__ inline int64_t ext_off_bs(int fam, int64_t kp) { return ext_off_s(fam, kp) + (fam ? kp / 32 * 4 : 0); }
__host__ __device__ inline int64_t ext_smem_bytes(int fam, int64_t kp) { return ext_off_bs(fam, kp) + (fam ? kp / 32 * 2 : kp / 16 * 2); }

template <int T>
__global__ void __launch_bounds__(EXT_WARPS * 32) ext_mmv_kernel(const ExtArgs a) {
    B200_DYN_SMEM(ext_smem);
    constexpr int FAM = (T == XF_Q2_K || T == XF_Q3_K || T == XF_Q4_K || T == XF_Q5_K || T == XF_Q6_K || T == XF_IQ4_XS) ? 0 : 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t k = a.k, kp = ext_kp(k);
    const int64_t col_bytes = ext_smem_bytes(FAM, kp);            // a multiple of 16

    pdl_wait();                                                    // activations and expert ids come from earlier kernels

    const int64_t p = a.prob0 + blockIdx.y;
    const uint8_t * W = a.W; const float * x; float * dst; const float * resid = nullptr;
    int nc = 1;                                                    // columns this CTA carries
    if (a.ids) {
        const int64_t tok = p / a.n_used, slot = p % a.n_used;
        const int e = a.ids[tok * a.ids_tok_stride + slot];
        if (e < 0 || e >= a.n_expert) return;                      // uniform over the CTA (the reference asserts the range: ggml-cpu.c:1496)
        W  += (int64_t)e * a.expert_stride;
        x   = a.x + tok * a.x_col_stride + (slot % a.n_b1) * a.x_slot_stride;
        dst = a.dst + tok * a.dst_col_stride + slot * a.dst_slot_stride;
    } else {
        const int64_t c0 = p * a.cols_per_cta;
        nc  = (int)(a.ncols - c0 < a.cols_per_cta ? a.ncols - c0 : a.cols_per_cta);
        x   = a.x + c0 * a.x_col_stride;
        dst = a.dst + c0 * a.dst_col_stride;
        if (a.residual) resid = a.residual + c0 * a.res_col_stride;
    }

    // ---- prologue: the columns of this CTA, quantised like the oracle, into shared memory (elements past k: zero)
    for (int64_t w = warp; w < (kp / 256) * nc; w += EXT_WARPS) {
        const int64_t c = w % (kp / 256); const int j = (int)(w / (kp / 256));
        uint8_t * col = ext_smem + j * col_bytes;
        int8_t  * qs = (int8_t *)col;
        float   * ad = (float *)(col + ext_off_d(kp));
        float   * as = (float *)(col + ext_off_s(FAM, kp));
        int16_t * bs = (int16_t *)(col + ext_off_bs(FAM, kp));
        float v[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
        const int64_t e0 = c * 256 + lane * 8;
        if (e0 < k) {                                              // k % 32 == 0: a lane's 8 elements are all inside or all outside
            const float4 * px = (const float4 *)(x + j * a.x_col_stride + e0);
            const float4 f0 = px[0], f1 = px[1];
            v[0] = f0.x; v[1] = f0.y; v[2] = f0.z; v[3] = f0.w; v[4] = f1.x; v[5] = f1.y; v[6] = f1.z; v[7] = f1.w;
        }
        if (FAM == 0) { ActOut o; o.qs = qs; o.d = ad; o.bs = bs; warp_quant_q8K(v, o, c, lane); }
        else          warp_quant_q8_01(v, qs, ad, as, bs, c, lane);
    }
    __syncthreads();

    const int64_t nsub = k / 32;
    for (int64_t row = (int64_t)blockIdx.x * EXT_WARPS + warp; row < a.m; row += (int64_t)gridDim.x * EXT_WARPS) {
        const uint8_t * wr = W + row * a.row_bytes;
        float acc[EXT_MAX_COLS];
#pragma unroll
        for (int j = 0; j < EXT_MAX_COLS; j++) acc[j] = 0.0f;
        for (int64_t u = lane; u < nsub; u += 32) {
#pragma unroll
            for (int j = 0; j < EXT_MAX_COLS; j++) {
                if (j < nc) {                                      // uniform over the CTA; the weight bytes of sub-block u come from L1 after the first column
                    const uint8_t * col = ext_smem + j * col_bytes;
                    XfAct A; A.qs = (const int8_t *)col; A.d = (const float *)(col + ext_off_d(kp)); A.s = (const float *)(col + ext_off_s(FAM, kp)); A.bs = (const int16_t *)(col + ext_off_bs(FAM, kp));
                    acc[j] += xf_sub_dot<T>(wr, a.nb_layout, u, A);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < EXT_MAX_COLS; j++) {
            if (j < nc) {
                const float r0 = warp_sum(acc[j]);
                if (lane == 0) {
                    float r = r0;
                    if (a.bias)  r += a.bias[row];
                    if (resid)   r += resid[j * a.res_col_stride + row];
                    dst[j * a.dst_col_stride + row] = r;
                }
            }
        }
    }
}

// one thread per 32-element sub-block of a gathered row
template <int T>
__global__ void __launch_bounds__(128) ext_get_rows_kernel(const uint8_t * __restrict__ src, int64_t row_stride, int64_t nb_layout, int64_t nrows,
                                                           const int32_t * __restrict__ ids, float * __restrict__ dst, int64_t dst_row_stride, int64_t ncols, int64_t id0) {
    pdl_wait();
    const int64_t u = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (u >= ncols / 32) return;
    const int64_t r = id0 + blockIdx.y;
    const int64_t id = ids[r];
    float y[32];
    if (id >= 0 && id < nrows) xf_sub_dequant<T>(src + id * row_stride, nb_layout, u, y);
    else {
#pragma unroll
        for (int i = 0; i < 32; i++) y[i] = 0.0f;
    }
    float4 * o = (float4 *)(dst + r * dst_row_stride + 32 * u);
#pragma unroll
    for (int i = 0; i < 8; i++) o[i] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
}

} // namespace

// glue_ext_kernels.cuh — device code of the mixture-of-experts router glue (launch code: glue_ext.cu): what build_moe_ffn
// (llama.cpp/src/llama-graph.cpp:820-1010) emits around its three MUL_MAT_IDs — the f32 router matmul, SOFT_MAX over the expert logits,
// ARGSORT (top-k), the 3-D GET_ROWS that picks the selected probabilities, SUM_ROWS / DIV (weight normalisation), the broadcast MUL by the
// expert weights and the ADDs over strided expert slices.  All tiny (n_expert x n_tokens): latency-bound, written for exactness against
// the CPU oracle (ggml-cpu/ops.cpp) and for being obviously right; one thread per row where the oracle's arithmetic is sequential
// (double-precision sums, the exchange sort whose tie order ARGSORT must reproduce).  Runs under tests/hostsim/simt.h on the CPU too.
#pragma once
#include "common.cuh"

namespace {

struct BinArgs { const float * a, * b; float * d; int64_t ne[4], a_nb[4], b_ne[4], b_nb[4], d_nb[4]; };     // strides in BYTES (ggml's nb[])

// d[i] = a[i] op b[i mod b_ne] with ggml's broadcasting (binary-ops.cpp): OP 0 add, 1 mul, 2 div; OP 3 = d[i] = a[i] (CONT of a strided view)
template <int OP>
__global__ void __launch_bounds__(256) bin_strided_kernel(const BinArgs A) {
    pdl_wait();
    const int64_t total = A.ne[0] * A.ne[1] * A.ne[2] * A.ne[3];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i0 = i % A.ne[0], r1 = i / A.ne[0], i1 = r1 % A.ne[1], r2 = r1 / A.ne[1], i2 = r2 % A.ne[2], i3 = r2 / A.ne[2];
        const float va = *(const float *)((const char *)A.a + i0 * A.a_nb[0] + i1 * A.a_nb[1] + i2 * A.a_nb[2] + i3 * A.a_nb[3]);
        float vb = 0.0f;
        if (OP != 3) vb = *(const float *)((const char *)A.b + (i0 % A.b_ne[0]) * A.b_nb[0] + (i1 % A.b_ne[1]) * A.b_nb[1] + (i2 % A.b_ne[2]) * A.b_nb[2] + (i3 % A.b_ne[3]) * A.b_nb[3]);
        float r;
        if (OP == 0) r = __fadd_rn(va, vb); else if (OP == 1) r = __fmul_rn(va, vb); else if (OP == 2) r = __fdiv_rn(va, vb); else r = va;
        *(float *)((char *)A.d + i0 * A.d_nb[0] + i1 * A.d_nb[1] + i2 * A.d_nb[2] + i3 * A.d_nb[3]) = r;
    }
}

// SOFT_MAX of rows without mask / sinks (ggml-cpu/ops.cpp:5685-5800, vec.cpp ggml_vec_soft_max_f32): w = x * scale, max, exp(w - max),
// the sum accumulated in double, every element scaled by (float)(1 / sum).  One thread per row.
// With a mask (attention without -fa, llama-graph.cpp build_attn_mha): row r = (token i1, head i2) of [n_kv, n_tok, n_head]; w += slope(head) * mask[i1][i]
// (mask f32 or f16, one row per token, shared by the heads; ALiBi slope as ops.cpp:5719-5738).
struct SoftMaxMask { const void * mask; int is_f16; int64_t row_stride /* elements */, rows_per_head; float max_bias, m0, m1; uint32_t n_head_log2; };
__device__ __forceinline__ float soft_max_w(const float * xr, int64_t i, float scale, const SoftMaxMask & M, const void * mrow, float slope) {
    float w = __fmul_rn(xr[i], scale);
    if (mrow) w = __fadd_rn(w, __fmul_rn(slope, M.is_f16 ? __half2float(((const __half *)mrow)[i]) : ((const float *)mrow)[i]));
    return w;
}
__global__ void __launch_bounds__(128) soft_max_rows_kernel(const float * __restrict__ x, int64_t x_rs, float * __restrict__ y, int64_t y_rs, int64_t ncols, int64_t nrows, float scale, const SoftMaxMask M) {
    pdl_wait();
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    const float * xr = x + r * x_rs; float * yr = y + r * y_rs;
    const void * mrow = nullptr; float slope = 1.0f;
    if (M.mask) {
        const int64_t i1 = r % M.rows_per_head, h = r / M.rows_per_head;
        mrow = (const char *)M.mask + i1 * M.row_stride * (M.is_f16 ? 2 : 4);
        if (M.max_bias > 0.0f) slope = (uint32_t)h < M.n_head_log2 ? powf(M.m0, (float)(h + 1)) : powf(M.m1, (float)(2 * (h - M.n_head_log2) + 1));
    }
    float mx = -INFINITY;
    for (int64_t i = 0; i < ncols; i++) mx = fmaxf(mx, soft_max_w(xr, i, scale, M, mrow, slope));
    double sum = 0.0;
    for (int64_t i = 0; i < ncols; i++) { const float v = expf(__fsub_rn(soft_max_w(xr, i, scale, M, mrow, slope), mx)); yr[i] = v; sum += (double)v; }
    const float inv = (float)(1.0 / sum);
    for (int64_t i = 0; i < ncols; i++) yr[i] = __fmul_rn(yr[i], inv);
}

// ARGSORT of rows: the reference's exchange sort, verbatim in behaviour (ggml-cpu/ops.cpp:8127-8146), so that ties — equal router
// probabilities — come out in the same order and TOP_K picks the same experts.  One thread per row; order 0 = ascending, 1 = descending.
__global__ void __launch_bounds__(128) argsort_rows_kernel(const float * __restrict__ x, int64_t x_rs, int32_t * __restrict__ idx, int64_t idx_rs, int64_t ncols, int64_t nrows, int order) {
    pdl_wait();
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    const float * xr = x + r * x_rs; int32_t * d = idx + r * idx_rs;
    for (int64_t j = 0; j < ncols; j++) d[j] = (int32_t)j;
    for (int64_t j = 0; j < ncols; j++)
        for (int64_t k = j + 1; k < ncols; k++) {
            const float vj = xr[d[j]], vk = xr[d[k]];
            if (order == 0 ? vj > vk : vj < vk) { const int32_t t = d[j]; d[j] = d[k]; d[k] = t; }
        }
}

// SUM_ROWS (ggml-cpu/ops.cpp sum_rows -> ggml_vec_sum_f32: sequential sum in double, stored as float).  One thread per row.
__global__ void __launch_bounds__(128) sum_rows_kernel(const float * __restrict__ x, int64_t x_rs, float * __restrict__ y, int64_t ncols, int64_t nrows) {
    pdl_wait();
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    double s = 0.0;
    for (int64_t i = 0; i < ncols; i++) s += (double)x[r * x_rs + i];
    y[r] = (float)s;
}

// GET_ROWS f32 with batched ids (ggml.c:3620-3650): dst[:, i, b] = src[:, ids[i, b], b]; strides in floats / int32s
__global__ void __launch_bounds__(256) get_rows_f32_3d_kernel(const float * __restrict__ src, int64_t s_rs, int64_t s_bs, int64_t n_src_rows, const int32_t * __restrict__ ids, int64_t id_bs,
                                                              float * __restrict__ dst, int64_t d_rs, int64_t d_bs, int64_t ncols, int64_t n_ids, int64_t n_batch) {
    pdl_wait();
    const int64_t total = ncols * n_ids * n_batch;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = i % ncols, r = (i / ncols) % n_ids, b = i / (ncols * n_ids);
        const int64_t id = ids[b * id_bs + r];
        dst[b * d_bs + r * d_rs + c] = (id >= 0 && id < n_src_rows) ? src[b * s_bs + id * s_rs + c] : 0.0f;
    }
}

// MUL_MAT with a small f32 weight matrix (the router: ffn_gate_inp [n_embd, n_expert]): one warp per output element
__global__ void __launch_bounds__(256) mul_mat_f32_kernel(const float * __restrict__ W, int64_t w_rs, const float * __restrict__ x, int64_t x_cs, float * __restrict__ dst, int64_t d_cs,
                                                          int64_t m, int64_t k, int64_t ncols) {
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int64_t o = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (o >= m * ncols) return;                              // whole warps leave together
    const int64_t r = o % m, c = o / m;
    float acc = 0.0f;
    for (int64_t i = lane; i < k; i += 32) acc = fmaf(W[r * w_rs + i], x[c * x_cs + i], acc);
    acc = warp_sum(acc);
    if (lane == 0) dst[c * d_cs + r] = acc;
}

// Unary ops of gating variants (ggml-cpu/unary-ops.cpp, vec.h): OP 0 SCALE (x * s, or fma(x, s, b) when b != 0: ops.cpp:4815-4850), 1 SILU (x / (1 + exp(-x)), the x86
// vector polynomial of common.cuh silu_x86), 2 SIGMOID (1 / (1 + expf(-x))).  Contiguous f32.
template <int OP>
__global__ void __launch_bounds__(256) unary_kernel(const float * __restrict__ x, float * __restrict__ y, int64_t n, float s, float b) {
    pdl_wait();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = x[i];
        float r;
        if (OP == 0) r = b == 0.0f ? __fmul_rn(v, s) : fmaf(v, s, b);
        else if (OP == 1) r = silu_x86(v);
        else r = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-v)));
        y[i] = r;
    }
}

// ---- attention without -fa (llama-graph.cpp build_attn_mha, non-flash branch): KQ = K^T Q and KQV = V^T softmax(KQ) are batched MUL_MATs with f16
// src0 views of the KV cache (permuted / transposed, GQA-broadcast over dim 2); the CPU oracle rounds the f32 operand to f16 (vec_dot_type of F16,
// ggml-cpu.c:209-303) and accumulates in f32 (ggml_vec_dot_f16).  One warp per output element; strides in BYTES.
struct MMF16Args { const char * A; int64_t a_nb1, a_nb2; const char * B; int64_t b_nb1, b_nb2; char * D; int64_t d_nb1, d_nb2; int64_t m, n, nbatch, k, r2; };
__global__ void __launch_bounds__(256) mul_mat_f16_kernel(const MMF16Args a) {
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int64_t o = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (o >= a.m * a.n * a.nbatch) return;                      // whole warps leave together
    const int64_t i0 = o % a.m, i1 = (o / a.m) % a.n, i2 = o / (a.m * a.n);
    const __half * ar = (const __half *)(a.A + i0 * a.a_nb1 + (i2 / a.r2) * a.a_nb2);
    const float  * br = (const float *)(a.B + i1 * a.b_nb1 + i2 * a.b_nb2);
    float acc = 0.0f;
    for (int64_t i = lane; i < a.k; i += 32) acc = fmaf(__half2float(ar[i]), __half2float(__float2half_rn(br[i])), acc);
    acc = warp_sum(acc);
    if (lane == 0) *(float *)(a.D + i0 * 4 + i1 * a.d_nb1 + i2 * a.d_nb2) = acc;
}

// SET_ROWS whose rows are single elements (the transposed V cache of attention without -fa, llama-kv-cache-unified.cpp:1157-1167): dst[ids[i]] = src[i]
__global__ void __launch_bounds__(256) scatter_rows1_kernel(const float * __restrict__ src, const int64_t * __restrict__ ids, void * __restrict__ dst, int dst_f16, int64_t n, int64_t n_dst) {
    pdl_wait();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t id = ids[i];
        if (id < 0 || id >= n_dst) continue;
        if (dst_f16) ((__half *)dst)[id] = __float2half_rn(src[i]); else ((float *)dst)[id] = src[i];
    }
}

} // namespace

// glue_ext.cu — launch code of the mixture-of-experts router glue (kernels: glue_ext_kernels.cuh; part of the wide path, GGML_B200_WIDE=1).
// Replaces, for the shapes build_moe_ffn emits, ggml-cuda's softmax.cu (soft_max_f32 without mask), argsort.cu (k_argsort_f32_i32),
// sumrows.cu, binbcast.cu's strided / broadcast cases, getrows.cu's batched f32 case and the cuBLAS route of the f32 router matmul
// (ggml-cuda.cu:1227-1330).  With these a Mixtral-style FFN never leaves the device between its RMS_NORM and the residual ADD.
#include "glue_ext_kernels.cuh"

static int ge_grid(int64_t work, int threads) {
    int64_t g = (work + threads - 1) / threads;
    const int64_t cap = (int64_t)b200_sm_count() * 8;
    if (g > cap) g = cap;
    return (int)(g < 1 ? 1 : g);
}
#define GE_DEV() do { if (b200_device_count() <= 0) { b200_set_error("no CUDA device"); return B200_ERR_CUDA; } } while (0)

extern "C" int b200_binary_strided(int op, const float * a, const int64_t * a_nb, const float * b, const int64_t * b_ne, const int64_t * b_nb,
                                   float * dst, const int64_t * ne, const int64_t * d_nb, void * stream) {
    GE_DEV();
    if (op < 0 || op > 3 || !a || !b || !dst || !a_nb || !b_ne || !b_nb || !ne || !d_nb) { b200_set_error("binary_strided: bad arguments"); return B200_ERR_INVALID; }
    BinArgs A; A.a = a; A.b = b; A.d = dst;
    for (int i = 0; i < 4; i++) {
        A.ne[i] = ne[i]; A.a_nb[i] = a_nb[i]; A.b_ne[i] = b_ne[i]; A.b_nb[i] = b_nb[i]; A.d_nb[i] = d_nb[i];
        if (ne[i] <= 0 || b_ne[i] <= 0 || ne[i] % b_ne[i] != 0 || (a_nb[i] & 3) || (b_nb[i] & 3) || (d_nb[i] & 3)) { b200_set_error("binary_strided: shapes must broadcast, strides must be multiples of 4 bytes"); return B200_ERR_INVALID; }
    }
    if ((((uintptr_t)a | (uintptr_t)b | (uintptr_t)dst) & 3)) { b200_set_error("binary_strided: pointers must be 4-byte aligned"); return B200_ERR_INVALID; }
    const int g = ge_grid(ne[0] * ne[1] * ne[2] * ne[3], 256);
    if (op == 0) bin_strided_kernel<0><<<g, 256, 0, (cudaStream_t)stream>>>(A);
    else if (op == 1) bin_strided_kernel<1><<<g, 256, 0, (cudaStream_t)stream>>>(A);
    else if (op == 2) bin_strided_kernel<2><<<g, 256, 0, (cudaStream_t)stream>>>(A);
    else bin_strided_kernel<3><<<g, 256, 0, (cudaStream_t)stream>>>(A);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_soft_max_rows(const float * x, int64_t x_row_stride, float * y, int64_t y_row_stride, int64_t ncols, int64_t nrows, float scale, void * stream) {
    GE_DEV();
    if (!x || !y || ncols <= 0 || nrows <= 0 || (((uintptr_t)x | (uintptr_t)y) & 3)) { b200_set_error("soft_max_rows: bad arguments"); return B200_ERR_INVALID; }
    SoftMaxMask M = {};
    soft_max_rows_kernel<<<(unsigned)((nrows + 127) / 128), 128, 0, (cudaStream_t)stream>>>(x, x_row_stride, y, y_row_stride, ncols, nrows, scale, M);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_soft_max_mask(const float * x, float * y, const void * mask, int mask_is_f16, int64_t mask_row_stride, int64_t ncols, int64_t n_tok, int64_t n_head,
                                  float scale, float max_bias, void * stream) {
    GE_DEV();
    if (!x || !y || !mask || ncols <= 0 || n_tok <= 0 || n_head <= 0 || (((uintptr_t)x | (uintptr_t)y) & 3) || ((uintptr_t)mask & (mask_is_f16 ? 1 : 3))) { b200_set_error("soft_max_mask: bad arguments"); return B200_ERR_INVALID; }
    SoftMaxMask M = {};
    M.mask = mask; M.is_f16 = mask_is_f16 ? 1 : 0; M.row_stride = mask_row_stride; M.rows_per_head = n_tok; M.max_bias = max_bias;
    M.n_head_log2 = 1u << (uint32_t)floor(log2((double)n_head));
    M.m0 = powf(2.0f, -max_bias / (float)M.n_head_log2); M.m1 = powf(2.0f, -(max_bias / 2.0f) / (float)M.n_head_log2);
    const int64_t nrows = n_tok * n_head;
    soft_max_rows_kernel<<<(unsigned)((nrows + 127) / 128), 128, 0, (cudaStream_t)stream>>>(x, ncols, y, ncols, ncols, nrows, scale, M);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_mul_mat_f16(const void * A, int64_t a_nb1, int64_t a_nb2, int64_t a_ne2, const float * B, int64_t b_nb1, int64_t b_nb2, float * dst, int64_t d_nb1, int64_t d_nb2,
                                int64_t m, int64_t n, int64_t n_batch, int64_t k, void * stream) {
    GE_DEV();
    if (!A || !B || !dst || m <= 0 || n <= 0 || n_batch <= 0 || k <= 0 || a_ne2 <= 0 || n_batch % a_ne2 != 0 || ((uintptr_t)A & 1) || ((a_nb1 | a_nb2) & 1) ||
        (((uintptr_t)B | (uintptr_t)dst) & 3) || ((b_nb1 | b_nb2 | d_nb1 | d_nb2) & 3) || m * n * n_batch > ((int64_t)1 << 31)) { b200_set_error("mul_mat_f16: bad arguments"); return B200_ERR_INVALID; }
    MMF16Args a; a.A = (const char *)A; a.a_nb1 = a_nb1; a.a_nb2 = a_nb2; a.B = (const char *)B; a.b_nb1 = b_nb1; a.b_nb2 = b_nb2; a.D = (char *)dst; a.d_nb1 = d_nb1; a.d_nb2 = d_nb2;
    a.m = m; a.n = n; a.nbatch = n_batch; a.k = k; a.r2 = n_batch / a_ne2;
    const int64_t warps = m * n * n_batch;
    mul_mat_f16_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_scatter_rows1(const float * src, const int64_t * ids, void * dst, int dst_type, int64_t n, int64_t n_dst, void * stream) {
    GE_DEV();
    if (!src || !ids || !dst || n <= 0 || n_dst <= 0 || (dst_type != B200_TYPE_F16 && dst_type != B200_TYPE_F32) || ((uintptr_t)src & 3) || ((uintptr_t)ids & 7) || ((uintptr_t)dst & (dst_type == B200_TYPE_F16 ? 1 : 3))) {
        b200_set_error("scatter_rows1: bad arguments"); return B200_ERR_INVALID; }
    scatter_rows1_kernel<<<ge_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(src, ids, dst, dst_type == B200_TYPE_F16 ? 1 : 0, n, n_dst);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_argsort_rows(const float * x, int64_t x_row_stride, int32_t * idx, int64_t idx_row_stride, int64_t ncols, int64_t nrows, int descending, void * stream) {
    GE_DEV();
    if (!x || !idx || ncols <= 0 || ncols > 4096 || nrows <= 0 || (((uintptr_t)x | (uintptr_t)idx) & 3)) { b200_set_error("argsort_rows: bad arguments (at most 4096 columns)"); return B200_ERR_INVALID; }
    argsort_rows_kernel<<<(unsigned)((nrows + 127) / 128), 128, 0, (cudaStream_t)stream>>>(x, x_row_stride, idx, idx_row_stride, ncols, nrows, descending ? 1 : 0);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_sum_rows(const float * x, int64_t x_row_stride, float * y, int64_t ncols, int64_t nrows, void * stream) {
    GE_DEV();
    if (!x || !y || ncols <= 0 || nrows <= 0 || (((uintptr_t)x | (uintptr_t)y) & 3)) { b200_set_error("sum_rows: bad arguments"); return B200_ERR_INVALID; }
    sum_rows_kernel<<<(unsigned)((nrows + 127) / 128), 128, 0, (cudaStream_t)stream>>>(x, x_row_stride, y, ncols, nrows);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_get_rows_f32_batched(const float * src, int64_t src_row_stride, int64_t src_batch_stride, int64_t n_src_rows, const int32_t * ids, int64_t ids_batch_stride,
                                         float * dst, int64_t dst_row_stride, int64_t dst_batch_stride, int64_t ncols, int64_t n_ids, int64_t n_batch, void * stream) {
    GE_DEV();
    if (!src || !ids || !dst || ncols <= 0 || n_ids <= 0 || n_batch <= 0 || n_src_rows <= 0 || (((uintptr_t)src | (uintptr_t)ids | (uintptr_t)dst) & 3)) { b200_set_error("get_rows_f32_batched: bad arguments"); return B200_ERR_INVALID; }
    get_rows_f32_3d_kernel<<<ge_grid(ncols * n_ids * n_batch, 256), 256, 0, (cudaStream_t)stream>>>(src, src_row_stride, src_batch_stride, n_src_rows, ids, ids_batch_stride, dst, dst_row_stride, dst_batch_stride, ncols, n_ids, n_batch);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_mul_mat_f32(const float * W, int64_t w_row_stride, const float * x, int64_t x_col_stride, float * dst, int64_t dst_col_stride, int64_t m, int64_t k, int64_t ncols, void * stream) {
    GE_DEV();
    if (!W || !x || !dst || m <= 0 || k <= 0 || ncols <= 0 || (((uintptr_t)W | (uintptr_t)x | (uintptr_t)dst) & 3)) { b200_set_error("mul_mat_f32: bad arguments"); return B200_ERR_INVALID; }
    const int64_t warps = m * ncols;
    mul_mat_f32_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(W, w_row_stride, x, x_col_stride, dst, dst_col_stride, m, k, ncols);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_unary(int op, const float * x, float * y, int64_t n, float s, float b, void * stream) {
    GE_DEV();
    if (op < 0 || op > 2 || !x || !y || n <= 0 || (((uintptr_t)x | (uintptr_t)y) & 3)) { b200_set_error("unary: bad arguments (op 0 scale, 1 silu, 2 sigmoid)"); return B200_ERR_INVALID; }
    const int g = ge_grid(n, 256);
    if (op == 0) unary_kernel<0><<<g, 256, 0, (cudaStream_t)stream>>>(x, y, n, s, b);
    else if (op == 1) unary_kernel<1><<<g, 256, 0, (cudaStream_t)stream>>>(x, y, n, s, b);
    else unary_kernel<2><<<g, 256, 0, (cudaStream_t)stream>>>(x, y, n, s, b);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

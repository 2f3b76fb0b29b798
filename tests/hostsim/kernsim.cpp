// kernsim.cpp — TEST INFRASTRUCTURE.  Runs the SOURCE of the wide path's CUDA kernels on the CPU under tests/hostsim/simt.h and exposes
// one C entry point per kernel with the argument lists of the product's C-ABI (b200_mul_mat_vec_wide / b200_mul_mat_id / b200_get_rows_q /
// b200_set_rows_q4_0 / b200_flash_attn_q4_0), host pointers instead of device pointers.  The launch geometry mirrors mmvq_ext.cu /
// fattn_ext.cu (grid = (row groups, problems), 256 / 128 threads) with a small row-group count.
#define XF_CHECK_ALIGN 1
#include "simt.h"

#include "../../llama-box_b200/csrc/mmvq_ext_kernels.cuh"
#include "../../llama-box_b200/csrc/fattn_ext_kernels.cuh"
#include "../../llama-box_b200/csrc/glue_ext_kernels.cuh"

void b200_set_error(const char *, ...) {}
extern "C" { long xf_misaligned = 0; long sim_misaligned(void) { return xf_misaligned; } }

namespace {
// the grouping rule of mmvq_ext.cu (ext_launch_t): columns of a plain MUL_MAT share a CTA, up to 8, while their quantised forms fit shared memory
template <int T> void run_mmv(const ExtArgs & a_in, int64_t n_items, int gx) {
    const int fam = xf_act_family(T);
    const int64_t col = ext_smem_bytes(fam, ext_kp(a_in.k));
    ExtArgs a = a_in;
    int64_t C = 1;
    if (!a.ids) { C = (96 * 1024) / col; if (C > EXT_MAX_COLS) C = EXT_MAX_COLS; if (C > n_items) C = n_items; if (C < 1) C = 1; }
    a.cols_per_cta = (int32_t)C; a.ncols = n_items;
    const int64_t n_prob = a.ids ? n_items : (n_items + C - 1) / C;
    simt::launch(dim3((unsigned)gx, (unsigned)n_prob), dim3(EXT_WARPS * 32), (size_t)(col * C), [a] { ext_mmv_kernel<T>(a); });
}
template <int T> void run_get_rows(const uint8_t * src, int64_t rs, int64_t nb, int64_t nrows, const int32_t * ids, float * dst, int64_t drs, int64_t ncols, int64_t n_ids) {
    simt::launch(dim3((unsigned)((ncols / 32 + 127) / 128), (unsigned)n_ids), dim3(128), 0, [=] { ext_get_rows_kernel<T>(src, rs, nb, nrows, ids, dst, drs, ncols, 0); });
}
int64_t wide_row_bytes(int type, int64_t k) { return k / xf_block_elems(type) * xf_block_bytes(type); }
}

extern "C" {
int sim_mul_mat_vec_wide(int type, const void * W, const float * x, int64_t x_col_stride, float * dst, int64_t dst_col_stride, const float * bias, const float * residual,
                         int64_t m, int64_t k, int64_t ncols, int gx) {
    ExtArgs a = {};
    a.W = (const uint8_t *)W; a.row_bytes = wide_row_bytes(type, k); a.nb_layout = k / xf_block_elems(type);
    a.x = x; a.x_col_stride = x_col_stride; a.dst = dst; a.dst_col_stride = dst_col_stride ? dst_col_stride : m;
    a.bias = bias; a.residual = residual; a.res_col_stride = a.dst_col_stride; a.m = m; a.k = k; a.n_b1 = 1; a.n_used = 1;
    int ok = 0;
    XF_DISPATCH(type, { run_mmv<T>(a, ncols, gx); ok = 1; });
    return ok;
}
int sim_mul_mat_id(int type, const void * as, int64_t expert_stride_bytes, const float * b, int64_t b_tok_stride, int64_t b_slot_stride, int64_t n_b1,
                   const int32_t * ids, int64_t ids_tok_stride, float * dst, int64_t dst_tok_stride, int64_t dst_slot_stride,
                   int64_t m, int64_t k, int64_t n_expert, int64_t n_used, int64_t n_tok, int gx) {
    ExtArgs a = {};
    a.W = (const uint8_t *)as; a.row_bytes = wide_row_bytes(type, k); a.nb_layout = k / xf_block_elems(type); a.expert_stride = expert_stride_bytes;
    a.ids = ids; a.ids_tok_stride = ids_tok_stride; a.n_used = (int32_t)n_used; a.n_expert = (int32_t)n_expert;
    a.x = b; a.x_col_stride = b_tok_stride; a.x_slot_stride = b_slot_stride; a.n_b1 = n_b1;
    a.dst = dst; a.dst_col_stride = dst_tok_stride; a.dst_slot_stride = dst_slot_stride; a.m = m; a.k = k;
    int ok = 0;
    XF_DISPATCH(type, { run_mmv<T>(a, n_tok * n_used, gx); ok = 1; });
    return ok;
}
int sim_get_rows_q(int type, const void * src, int64_t src_row_stride, int64_t nrows, const int32_t * ids, float * dst, int64_t dst_row_stride, int64_t ncols, int64_t n_ids) {
    int ok = 0;
    XF_DISPATCH(type, { run_get_rows<T>((const uint8_t *)src, src_row_stride, ncols / xf_block_elems(type), nrows, ids, dst, dst_row_stride, ncols, n_ids); ok = 1; });
    return ok;
}
void sim_set_rows_q4_0(const float * src, int64_t src_row_stride, const int64_t * ids, void * dst, int64_t dst_row_stride, int64_t ncols, int64_t nrows) {
    const int64_t nblk = ncols / 32;
    simt::launch(dim3((unsigned)((nblk + 127) / 128), (unsigned)nrows), dim3(128), 0, [=] { set_rows_q4_0_kernel(src, src_row_stride, ids, (uint8_t *)dst, dst_row_stride, nblk); });
}
void sim_flash_attn_q4_0(const float * q, int64_t q_tok_stride, int64_t q_head_stride, const void * k, int64_t k_row_stride, int64_t k_head_stride,
                         const void * v, int64_t v_row_stride, int64_t v_head_stride, const void * mask, int64_t mask_row_stride, float * dst,
                         int64_t d, int64_t n_head, int64_t n_head_kv, int64_t n_tok, int64_t n_kv, float scale, float max_bias, float logit_softcap) {
    FaWideArgs a = {};
    a.q = q; a.q_ts = q_tok_stride; a.q_hs = q_head_stride; a.k = (const uint8_t *)k; a.k_rs = k_row_stride; a.k_hs = k_head_stride;
    a.v = (const uint8_t *)v; a.v_rs = v_row_stride; a.v_hs = v_head_stride; a.mask = (const uint16_t *)mask; a.mask_rs = mask_row_stride;
    a.dst = dst; a.n_head = n_head; a.n_head_kv = n_head_kv; a.n_kv = n_kv;
    a.scale = logit_softcap != 0.0f ? scale / logit_softcap : scale; a.max_bias = max_bias; a.softcap = logit_softcap;
    a.nh_log2 = 1u << (uint32_t)floor(log2((double)n_head));
    a.m0 = powf(2.0f, -max_bias / (float)a.nh_log2); a.m1 = powf(2.0f, -(max_bias / 2.0f) / (float)a.nh_log2);
    const dim3 grid((unsigned)n_head, (unsigned)n_tok);
    if (d == 128) simt::launch(grid, dim3(128), 0, [a] { fattn_q4_0_kernel<128>(a); });
    else          simt::launch(grid, dim3(128), 0, [a] { fattn_q4_0_kernel<64>(a); });
}

// ---- mixture-of-experts router glue (glue_ext_kernels.cuh); geometry as in glue_ext.cu with a small grid
void sim_binary_strided(int op, const float * a, const int64_t * a_nb, const float * b, const int64_t * b_ne, const int64_t * b_nb, float * dst, const int64_t * ne, const int64_t * d_nb) {
    BinArgs A; A.a = a; A.b = b; A.d = dst;
    for (int i = 0; i < 4; i++) { A.ne[i] = ne[i]; A.a_nb[i] = a_nb[i]; A.b_ne[i] = b_ne[i]; A.b_nb[i] = b_nb[i]; A.d_nb[i] = d_nb[i]; }
    if (op == 0) simt::launch(dim3(2), dim3(256), 0, [A] { bin_strided_kernel<0>(A); });
    else if (op == 1) simt::launch(dim3(2), dim3(256), 0, [A] { bin_strided_kernel<1>(A); });
    else if (op == 2) simt::launch(dim3(2), dim3(256), 0, [A] { bin_strided_kernel<2>(A); });
    else simt::launch(dim3(2), dim3(256), 0, [A] { bin_strided_kernel<3>(A); });
}
void sim_soft_max_rows(const float * x, int64_t xrs, float * y, int64_t yrs, int64_t ncols, int64_t nrows, float scale) {
    SoftMaxMask M = {};
    simt::launch(dim3((unsigned)((nrows + 127) / 128)), dim3(128), 0, [=] { soft_max_rows_kernel(x, xrs, y, yrs, ncols, nrows, scale, M); });
}
void sim_soft_max_mask(const float * x, float * y, const void * mask, int mask_is_f16, int64_t mask_row_stride, int64_t ncols, int64_t n_tok, int64_t n_head, float scale, float max_bias) {
    SoftMaxMask M = {};
    M.mask = mask; M.is_f16 = mask_is_f16; M.row_stride = mask_row_stride; M.rows_per_head = n_tok; M.max_bias = max_bias;
    M.n_head_log2 = 1u << (uint32_t)floor(log2((double)n_head));
    M.m0 = powf(2.0f, -max_bias / (float)M.n_head_log2); M.m1 = powf(2.0f, -(max_bias / 2.0f) / (float)M.n_head_log2);
    const int64_t nrows = n_tok * n_head;
    simt::launch(dim3((unsigned)((nrows + 127) / 128)), dim3(128), 0, [=] { soft_max_rows_kernel(x, ncols, y, ncols, ncols, nrows, scale, M); });
}
void sim_mul_mat_f16(const void * A, int64_t a_nb1, int64_t a_nb2, int64_t a_ne2, const float * B, int64_t b_nb1, int64_t b_nb2, float * dst, int64_t d_nb1, int64_t d_nb2,
                     int64_t m, int64_t n, int64_t n_batch, int64_t k) {
    MMF16Args a; a.A = (const char *)A; a.a_nb1 = a_nb1; a.a_nb2 = a_nb2; a.B = (const char *)B; a.b_nb1 = b_nb1; a.b_nb2 = b_nb2; a.D = (char *)dst; a.d_nb1 = d_nb1; a.d_nb2 = d_nb2;
    a.m = m; a.n = n; a.nbatch = n_batch; a.k = k; a.r2 = n_batch / a_ne2;
    simt::launch(dim3((unsigned)((m * n * n_batch * 32 + 255) / 256)), dim3(256), 0, [a] { mul_mat_f16_kernel(a); });
}
void sim_scatter_rows1(const float * src, const int64_t * ids, void * dst, int dst_f16, int64_t n, int64_t n_dst) {
    simt::launch(dim3(2), dim3(256), 0, [=] { scatter_rows1_kernel(src, ids, dst, dst_f16, n, n_dst); });
}
void sim_argsort_rows(const float * x, int64_t xrs, int32_t * idx, int64_t irs, int64_t ncols, int64_t nrows, int desc) {
    simt::launch(dim3((unsigned)((nrows + 127) / 128)), dim3(128), 0, [=] { argsort_rows_kernel(x, xrs, idx, irs, ncols, nrows, desc); });
}
void sim_sum_rows(const float * x, int64_t xrs, float * y, int64_t ncols, int64_t nrows) {
    simt::launch(dim3((unsigned)((nrows + 127) / 128)), dim3(128), 0, [=] { sum_rows_kernel(x, xrs, y, ncols, nrows); });
}
void sim_get_rows_f32_batched(const float * src, int64_t srs, int64_t sbs, int64_t nsr, const int32_t * ids, int64_t ibs, float * dst, int64_t drs, int64_t dbs, int64_t ncols, int64_t n_ids, int64_t n_batch) {
    simt::launch(dim3(2), dim3(256), 0, [=] { get_rows_f32_3d_kernel(src, srs, sbs, nsr, ids, ibs, dst, drs, dbs, ncols, n_ids, n_batch); });
}
void sim_mul_mat_f32(const float * W, int64_t wrs, const float * x, int64_t xcs, float * dst, int64_t dcs, int64_t m, int64_t k, int64_t ncols) {
    simt::launch(dim3((unsigned)((m * ncols * 32 + 255) / 256)), dim3(256), 0, [=] { mul_mat_f32_kernel(W, wrs, x, xcs, dst, dcs, m, k, ncols); });
}
void sim_flash_attn_any(int kv_type, const float * q, int64_t q_tok_stride, int64_t q_head_stride, const void * k, int64_t k_row_stride, int64_t k_head_stride,
                        const void * v, int64_t v_row_stride, int64_t v_head_stride, const void * mask, int64_t mask_row_stride, float * dst,
                        int64_t d, int64_t n_head, int64_t n_head_kv, int64_t n_tok, int64_t n_kv, float scale, float max_bias, float logit_softcap) {
    FaWideArgs a = {};
    a.q = q; a.q_ts = q_tok_stride; a.q_hs = q_head_stride; a.k = (const uint8_t *)k; a.k_rs = k_row_stride; a.k_hs = k_head_stride;
    a.v = (const uint8_t *)v; a.v_rs = v_row_stride; a.v_hs = v_head_stride; a.mask = (const uint16_t *)mask; a.mask_rs = mask_row_stride;
    a.dst = dst; a.n_head = n_head; a.n_head_kv = n_head_kv; a.n_kv = n_kv;
    a.scale = logit_softcap != 0.0f ? scale / logit_softcap : scale; a.max_bias = max_bias; a.softcap = logit_softcap;
    a.nh_log2 = 1u << (uint32_t)floor(log2((double)n_head));
    a.m0 = powf(2.0f, -max_bias / (float)a.nh_log2); a.m1 = powf(2.0f, -(max_bias / 2.0f) / (float)a.nh_log2);
    const dim3 grid((unsigned)n_head, (unsigned)n_tok);
    const int D = (int)d;
    if (kv_type == 1)      simt::launch(grid, dim3(128), 0, [a, D] { fattn_any_kernel<1>(a, D); });
    else if (kv_type == 8) simt::launch(grid, dim3(128), 0, [a, D] { fattn_any_kernel<8>(a, D); });
    else                   simt::launch(grid, dim3(128), 0, [a, D] { fattn_any_kernel<2>(a, D); });
}
void sim_unary(int op, const float * x, float * y, int64_t n, float sc, float b) {
    if (op == 0) simt::launch(dim3(2), dim3(256), 0, [=] { unary_kernel<0>(x, y, n, sc, b); });
    else if (op == 1) simt::launch(dim3(2), dim3(256), 0, [=] { unary_kernel<1>(x, y, n, sc, b); });
    else simt::launch(dim3(2), dim3(256), 0, [=] { unary_kernel<2>(x, y, n, sc, b); });
}
}

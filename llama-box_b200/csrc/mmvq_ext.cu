// mmvq_ext.cu — the "wide" kernels: one simple, format-generic matvec that carries
//   * MUL_MAT for the formats the tuned decode kernels (mmvq.cu) do not have — Q4_1, Q5_1, Q2_K, Q3_K, IQ4_NL, IQ4_XS, MXFP4
//     (replaces the rest of ggml-cuda/vecdotq.cuh:75-175,656-1171 behind mul_mat_vec_q, ggml-cuda/mmvq.cu:139-226),
//   * MUL_MAT_ID (mixture-of-experts routing, replaces ggml_cuda_mul_mat_id, ggml-cuda/ggml-cuda.cu:2064-2205 and the ids path of
//     mmvq.cu:163-166) for those formats AND for Q4_0 / Q5_0 / Q8_0 / Q4_K / Q5_K / Q6_K in the library's weight layout: the expert
//     ids are read ON THE DEVICE by the CTA that needs them — no device->host synchronisation per op (the reference copies ids to
//     the host and synchronises the stream, ggml-cuda.cu:2115-2125),
//   * GET_ROWS on quantised tables (replaces k_get_rows, ggml-cuda/getrows.cu:5-67): the token-embedding lookup on the device.
//
// Design (HBM-bound byte/integer work; correctness first, this is the widening stage — the five formats of the BASELINE
// configs keep their tuned kernels):
//   grid = (row groups, problems); a problem is ONE activation column against one weight matrix: column c of a MUL_MAT, or
//   (token, expert slot) of a MUL_MAT_ID.  The CTA quantises its column in the prologue straight into shared memory the way the CPU
//   oracle does (q8_K / q8_0 / q8_1: same warp-level quantisers as mmvq.cu's prologue, actquant.cuh) — no activation buffer in HBM —
//   then each of its 8 warps walks whole weight rows, lanes striding over 32-element sub-blocks (extfmt.cuh: dp4a integer sums, the
//   oracle's scale arithmetic), warp-reduces and writes dst (+ bias, + residual).  Weights are read once per column; the few
//   columns of a decode / verify batch re-read them from L2.
//   The per-format arithmetic lives in extfmt.cuh and is ALSO compiled for the host and checked against the reference on the CPU
//   (tests/test_extfmt_hostsim.py).
#include "mmvq_ext_kernels.cuh"

namespace {

// k / alignment rules of a weight row in the layout these kernels read
bool ext_shape_ok(int type, int64_t k) {
    if (xf_act_family(type) < 0 || k <= 0) return false;
    if (xf_is_ext_only(type)) return k % xf_block_elems(type) == 0;             // ggml layout, rows at least 2-byte aligned for every k
    if (type == XF_Q6_K) return k % 512 == 0;                                   // library layout: 4-byte aligned rows and sections
    return k % 256 == 0;                                                        // library layout of the other tuned formats
}

// n_items: columns of a plain MUL_MAT (grouped up to 8 per CTA while their quantised forms fit shared memory), or (token, slot) pairs of a MUL_MAT_ID
template <int T> int ext_launch_t(const ExtArgs & a_in, int64_t n_items, cudaStream_t st) {
    const int fam = xf_act_family(T);
    const int64_t col = ext_smem_bytes(fam, ext_kp(a_in.k));
    ExtArgs a0 = a_in;
    int64_t C = 1;
    if (!a0.ids) { C = (96 * 1024) / col;      /* two resident CTAs per SM */ if (C > EXT_MAX_COLS) C = EXT_MAX_COLS; if (C > n_items) C = n_items; if (C < 1) C = 1; }
    a0.cols_per_cta = (int32_t)C; a0.ncols = n_items;
    const int64_t n_prob = a0.ids ? n_items : (n_items + C - 1) / C;
    const int64_t smem = col * C;
    if (col > 200 * 1024) { b200_set_error("wide matvec: k = %lld does not fit shared memory", (long long)a0.k); return B200_ERR_UNSUPPORTED; }
    if (smem > 48 * 1024) {
        static bool attr[64];
        int dev = 0; cudaGetDevice(&dev);
        if (!attr[dev & 63]) { B200_CUDA(cudaFuncSetAttribute(ext_mmv_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr[dev & 63] = true; }
    }
    int64_t gx = (a0.m + 31) / 32;                                              // >= 4 rows per warp before a CTA's prologue is amortised
    const int64_t cap = (int64_t)b200_sm_count() * 4;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    for (int64_t p0 = 0; p0 < n_prob; p0 += 32768) {
        ExtArgs a = a0; a.prob0 = p0;
        const int64_t ny = n_prob - p0 < 32768 ? n_prob - p0 : 32768;
        ext_mmv_kernel<T><<<dim3((unsigned)gx, (unsigned)ny), EXT_WARPS * 32, (size_t)smem, st>>>(a);
        B200_LAUNCH_CHECK();
    }
    return B200_OK;
}
int ext_launch(int type, const ExtArgs & a, int64_t n_prob, cudaStream_t st) {
    int s = B200_ERR_UNSUPPORTED;
    XF_DISPATCH(type, s = ext_launch_t<T>(a, n_prob, st));
    return s;
}

template <int T> void ext_get_rows_launch_t(dim3 grid, cudaStream_t st, const uint8_t * src, int64_t row_stride, int64_t nb, int64_t nrows, const int32_t * ids,
                                            float * dst, int64_t dst_row_stride, int64_t ncols, int64_t id0) {
    ext_get_rows_kernel<T><<<grid, 128, 0, st>>>(src, row_stride, nb, nrows, ids, dst, dst_row_stride, ncols, id0);
}

bool al16(const void * p) { return ((uintptr_t)p & 15) == 0; }

} // namespace

extern "C" int b200_wide_type_supported(int type) { return xf_act_family(type) >= 0 ? 1 : 0; }
extern "C" int b200_wide_shape_supported(int type, int64_t k) { return ext_shape_ok(type, k) ? 1 : 0; }
extern "C" int64_t b200_wide_row_bytes(int type, int64_t k) { const int be = xf_block_elems(type); return be ? k / be * xf_block_bytes(type) : 0; }

extern "C" int b200_mul_mat_vec_wide(int type, const void * W, const float * x, int64_t x_col_stride, float * dst, int64_t dst_col_stride,
                                     const float * bias, const float * residual, int64_t m, int64_t k, int64_t ncols, void * stream) {
    if (b200_device_count() <= 0) { b200_set_error("no CUDA device"); return B200_ERR_CUDA; }
    if (!ext_shape_ok(type, k)) { b200_set_error("wide matvec: type %d with k = %lld is not supported", type, (long long)k); return B200_ERR_UNSUPPORTED; }
    if (!W || !x || !dst || m <= 0 || ncols <= 0 || !al16(W) || !al16(x) || (x_col_stride & 3) || ((uintptr_t)dst & 3)) { b200_set_error("wide matvec: bad arguments (W, x 16-byte aligned; x column stride a multiple of 4 floats)"); return B200_ERR_INVALID; }
    ExtArgs a = {};
    a.W = (const uint8_t *)W; a.row_bytes = b200_wide_row_bytes(type, k); a.nb_layout = k / xf_block_elems(type);
    a.x = x; a.x_col_stride = x_col_stride; a.dst = dst; a.dst_col_stride = dst_col_stride ? dst_col_stride : m;
    a.bias = bias; a.residual = residual; a.res_col_stride = a.dst_col_stride; a.m = m; a.k = k; a.n_b1 = 1; a.n_used = 1;
    return ext_launch(type, a, ncols, (cudaStream_t)stream);
}

extern "C" int b200_mul_mat_id(int type, const void * as, int64_t expert_stride_bytes, const float * b, int64_t b_tok_stride, int64_t b_slot_stride, int64_t n_b1,
                               const int32_t * ids, int64_t ids_tok_stride, float * dst, int64_t dst_tok_stride, int64_t dst_slot_stride,
                               int64_t m, int64_t k, int64_t n_expert, int64_t n_used, int64_t n_tok, void * stream) {
    if (b200_device_count() <= 0) { b200_set_error("no CUDA device"); return B200_ERR_CUDA; }
    if (!ext_shape_ok(type, k)) { b200_set_error("mul_mat_id: type %d with k = %lld is not supported", type, (long long)k); return B200_ERR_UNSUPPORTED; }
    if (!as || !b || !ids || !dst || m <= 0 || n_expert <= 0 || n_used <= 0 || n_tok <= 0 || n_b1 <= 0 || n_expert > INT32_MAX || n_used > INT32_MAX ||
        !al16(as) || !al16(b) || (b_tok_stride & 3) || (b_slot_stride & 3) || (expert_stride_bytes & 15) || ((uintptr_t)dst & 3)) {
        b200_set_error("mul_mat_id: bad arguments (as, b 16-byte aligned; b strides multiples of 4 floats; expert stride a multiple of 16 bytes)"); return B200_ERR_INVALID;
    }
    ExtArgs a = {};
    a.W = (const uint8_t *)as; a.row_bytes = b200_wide_row_bytes(type, k); a.nb_layout = k / xf_block_elems(type); a.expert_stride = expert_stride_bytes;
    a.ids = ids; a.ids_tok_stride = ids_tok_stride; a.n_used = (int32_t)n_used; a.n_expert = (int32_t)n_expert;
    a.x = b; a.x_col_stride = b_tok_stride; a.x_slot_stride = b_slot_stride; a.n_b1 = n_b1;
    a.dst = dst; a.dst_col_stride = dst_tok_stride; a.dst_slot_stride = dst_slot_stride; a.m = m; a.k = k;
    return ext_launch(type, a, n_tok * n_used, (cudaStream_t)stream);
}

extern "C" int b200_get_rows_q(int type, const void * src, int64_t src_row_stride, int64_t nrows, const int32_t * ids, float * dst, int64_t dst_row_stride,
                               int64_t ncols, int64_t n_ids, void * stream) {
    if (b200_device_count() <= 0) { b200_set_error("no CUDA device"); return B200_ERR_CUDA; }
    if (!ext_shape_ok(type, ncols)) { b200_set_error("get_rows: type %d with %lld columns is not supported", type, (long long)ncols); return B200_ERR_UNSUPPORTED; }
    if (!src || !ids || !dst || nrows <= 0 || n_ids <= 0 || !al16(src) || !al16(dst) || (dst_row_stride & 3) || (src_row_stride & 1)) { b200_set_error("get_rows: bad arguments"); return B200_ERR_INVALID; }
    const int64_t nb = ncols / xf_block_elems(type);
    const unsigned gx = (unsigned)((ncols / 32 + 127) / 128);
    for (int64_t i0 = 0; i0 < n_ids; i0 += 32768) {
        const unsigned ny = (unsigned)(n_ids - i0 < 32768 ? n_ids - i0 : 32768);
        const dim3 grid(gx, ny);
        XF_DISPATCH(type, ext_get_rows_launch_t<T>(grid, (cudaStream_t)stream, (const uint8_t *)src, src_row_stride, nb, nrows, ids, dst, dst_row_stride, ncols, i0));
        B200_LAUNCH_CHECK();
    }
    return B200_OK;
}

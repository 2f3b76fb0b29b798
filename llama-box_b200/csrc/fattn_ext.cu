// fattn_ext.cu — KV cache type q4_0 (llama-box `-ctk q4_0 -ctv q4_0`): SET_ROWS into a Q4_0 cache and FLASH_ATTN_EXT over
// Q4_0 K / V (replaces k_set_rows_quant<block_q4_0> + quantize_f32_q4_0_block, ggml-cuda/set-rows.cu:13-52 / cpy-utils.cuh:25-52, and the
// q4_0-q4_0 instances of flash_attn_vec_ext, ggml-cuda/fattn.cu:184 / fattn-vec-f32.cuh).  Part of the wide path (GGML_B200_WIDE=1): the
// F16 / Q8_0 caches of the BASELINE configs keep their tuned kernels (fattn.cu, fattn_tc.cu).
//
// Numerics follow the CPU oracle (ggml-cpu/ops.cpp:8252-8404): the query row is quantised to q8_0 (x86 quantiser) and dotted with the K
// blocks in integers; s = dot * scale (+ softcap) + slope * mask; one-pass online softmax in f32; V blocks are expanded to f32 and
// accumulated in f32.  The cache keeps ggml's native 18-byte blocks.  Per-block arithmetic: extfmt.cuh (host-checked).
//
// One CTA per (head, query token), 4 warps, warp w takes KV positions w, w + 4, ...: lanes 0..D/32-1 dot one K block each, every lane owns
// D/32 output elements; the four partial softmax states are merged through shared memory.  HBM-bound byte work — simple on purpose.
#include "fattn_ext_kernels.cuh"

extern "C" int b200_set_rows_q4_0(const float * src, int64_t src_row_stride, const int64_t * ids, void * dst, int64_t dst_row_stride, int64_t ncols, int64_t nrows, void * stream) {
    if (b200_device_count() <= 0) { b200_set_error("no CUDA device"); return B200_ERR_CUDA; }
    if (!src || !ids || !dst || ncols <= 0 || ncols % 32 != 0 || nrows <= 0 || nrows > 65535 || ((uintptr_t)src & 15) || (src_row_stride & 3) || ((uintptr_t)dst & 1) || (dst_row_stride & 1)) {
        b200_set_error("set_rows q4_0: bad arguments (ncols a multiple of 32, src rows 16-byte aligned, at most 65535 rows per call)"); return B200_ERR_INVALID;
    }
    const int64_t nblk = ncols / 32;
    set_rows_q4_0_kernel<<<dim3((unsigned)((nblk + 127) / 128), (unsigned)nrows), 128, 0, (cudaStream_t)stream>>>(src, src_row_stride, ids, (uint8_t *)dst, dst_row_stride, nblk);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_flash_attn_q4_0(const float * q, int64_t q_tok_stride, int64_t q_head_stride, const void * k, int64_t k_row_stride, int64_t k_head_stride,
                                    const void * v, int64_t v_row_stride, int64_t v_head_stride, const void * mask, int64_t mask_row_stride, float * dst,
                                    int64_t d, int64_t n_head, int64_t n_head_kv, int64_t n_tok, int64_t n_kv, float scale, float max_bias, float logit_softcap, void * stream) {
    if (b200_device_count() <= 0) { b200_set_error("no CUDA device"); return B200_ERR_CUDA; }
    if (d != 64 && d != 128) { b200_set_error("flash_attn q4_0: head size %lld unsupported", (long long)d); return B200_ERR_UNSUPPORTED; }
    if (!q || !k || !v || !dst || n_head <= 0 || n_head_kv <= 0 || n_head % n_head_kv != 0 || n_tok <= 0 || n_tok > 65535 || n_kv <= 0 ||
        ((uintptr_t)q & 15) || (q_tok_stride & 3) || (q_head_stride & 3) || (((uintptr_t)k | (uintptr_t)v) & 1) || ((k_row_stride | k_head_stride | v_row_stride | v_head_stride) & 1)) {
        b200_set_error("flash_attn q4_0: bad arguments"); return B200_ERR_INVALID;
    }
    FaWideArgs a = {};
    a.q = q; a.q_ts = q_tok_stride; a.q_hs = q_head_stride; a.k = (const uint8_t *)k; a.k_rs = k_row_stride; a.k_hs = k_head_stride;
    a.v = (const uint8_t *)v; a.v_rs = v_row_stride; a.v_hs = v_head_stride; a.mask = (const uint16_t *)mask; a.mask_rs = mask_row_stride;
    a.dst = dst; a.n_head = n_head; a.n_head_kv = n_head_kv; a.n_kv = n_kv;
    a.scale = logit_softcap != 0.0f ? scale / logit_softcap : scale; a.max_bias = max_bias; a.softcap = logit_softcap;
    a.nh_log2 = 1u << (uint32_t)floor(log2((double)n_head));
    a.m0 = powf(2.0f, -max_bias / (float)a.nh_log2); a.m1 = powf(2.0f, -(max_bias / 2.0f) / (float)a.nh_log2);
    const dim3 grid((unsigned)n_head, (unsigned)n_tok);
    if (d == 128) fattn_q4_0_kernel<128><<<grid, 128, 0, (cudaStream_t)stream>>>(a);
    else          fattn_q4_0_kernel<64><<<grid, 128, 0, (cudaStream_t)stream>>>(a);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

// any head size (multiple of 32, at most 256) over an F16 / Q8_0 / Q4_0 cache — the shapes the tuned kernels do not carry (they have 64 and 128)
extern "C" int b200_flash_attn_any(int kv_type, const float * q, int64_t q_tok_stride, int64_t q_head_stride, const void * k, int64_t k_row_stride, int64_t k_head_stride,
                                   const void * v, int64_t v_row_stride, int64_t v_head_stride, const void * mask, int64_t mask_row_stride, float * dst,
                                   int64_t d, int64_t n_head, int64_t n_head_kv, int64_t n_tok, int64_t n_kv, float scale, float max_bias, float logit_softcap, void * stream) {
    if (b200_device_count() <= 0) { b200_set_error("no CUDA device"); return B200_ERR_CUDA; }
    if (d <= 0 || d % 32 != 0 || d > 256 || (kv_type != B200_TYPE_F16 && kv_type != B200_TYPE_Q8_0 && kv_type != B200_TYPE_Q4_0)) {
        b200_set_error("flash_attn_any: head size %lld / kv type %d unsupported", (long long)d, kv_type); return B200_ERR_UNSUPPORTED; }
    if (!q || !k || !v || !dst || n_head <= 0 || n_head_kv <= 0 || n_head % n_head_kv != 0 || n_tok <= 0 || n_tok > 65535 || n_kv <= 0 ||
        ((uintptr_t)q & 15) || (q_tok_stride & 3) || (q_head_stride & 3) || (((uintptr_t)k | (uintptr_t)v) & 1) || ((k_row_stride | k_head_stride | v_row_stride | v_head_stride) & 1)) {
        b200_set_error("flash_attn_any: bad arguments"); return B200_ERR_INVALID;
    }
    FaWideArgs a = {};
    a.q = q; a.q_ts = q_tok_stride; a.q_hs = q_head_stride; a.k = (const uint8_t *)k; a.k_rs = k_row_stride; a.k_hs = k_head_stride;
    a.v = (const uint8_t *)v; a.v_rs = v_row_stride; a.v_hs = v_head_stride; a.mask = (const uint16_t *)mask; a.mask_rs = mask_row_stride;
    a.dst = dst; a.n_head = n_head; a.n_head_kv = n_head_kv; a.n_kv = n_kv;
    a.scale = logit_softcap != 0.0f ? scale / logit_softcap : scale; a.max_bias = max_bias; a.softcap = logit_softcap;
    a.nh_log2 = 1u << (uint32_t)floor(log2((double)n_head));
    a.m0 = powf(2.0f, -max_bias / (float)a.nh_log2); a.m1 = powf(2.0f, -(max_bias / 2.0f) / (float)a.nh_log2);
    const dim3 grid((unsigned)n_head, (unsigned)n_tok);
    if (kv_type == B200_TYPE_F16)       fattn_any_kernel<1><<<grid, 128, 0, (cudaStream_t)stream>>>(a, (int)d);
    else if (kv_type == B200_TYPE_Q8_0) fattn_any_kernel<8><<<grid, 128, 0, (cudaStream_t)stream>>>(a, (int)d);
    else                                fattn_any_kernel<2><<<grid, 128, 0, (cudaStream_t)stream>>>(a, (int)d);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

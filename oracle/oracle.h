/*
 * oracle.h — CPU restatement of the reference's ggml-cpu algorithms for the GGUF-quantized
 * decode/prefill hot path.  TEST INFRASTRUCTURE ONLY: nothing in the product path
 * (llama-box_b200/csrc, the ggml plugin, bench.py's timed GPU region) may call into this file.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every function here against the
 * unmodified reference (oracle/_ref/libggml-cpu.so + libggml-base.so, compiled from
 * /root/reference/llama.cpp by oracle/Makefile) and against the committed fixtures in
 * tests/golden/ that oracle/make_golden.py generated from that reference build.
 *
 * All file:line citations are relative to /root/reference/llama.cpp/.
 */
#ifndef B200_ORACLE_H
#define B200_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* weight / cache type ids: numerically equal to enum ggml_type (ggml/include/ggml.h:377-418) */
enum orc_type {
    ORC_F32  = 0,
    ORC_F16  = 1,
    ORC_Q4_0 = 2,
    ORC_Q4_1 = 3,
    ORC_Q5_0 = 6,
    ORC_Q5_1 = 7,
    ORC_Q8_0 = 8,
    ORC_Q8_1 = 9,
    ORC_Q2_K = 10,
    ORC_Q3_K = 11,
    ORC_Q4_K = 12,
    ORC_Q5_K = 13,
    ORC_Q6_K = 14,
    ORC_Q8_K = 15,
    ORC_IQ4_NL = 20,
    ORC_IQ4_XS = 23,
    ORC_MXFP4 = 39,
};

/* block geometry (ggml/src/ggml-common.h:170-175,219-224,295-344) */
int64_t orc_block_elems(int type);          /* 32 or 256 (1 for f32/f16) */
int64_t orc_block_bytes(int type);          /* 18, 34, 144, 176, 210, 292 ... */
int64_t orc_row_bytes(int type, int64_t k); /* k / elems * bytes */

/* IEEE half <-> float, round-to-nearest-even (ggml/src/ggml-impl.h fp16 helpers / F16C) */
float    orc_fp16_to_fp32(uint16_t h);
uint16_t orc_fp32_to_fp16(float f);

/* activation quantisers as the x86 CPU backend runs them
 *   q8_0: ggml-cpu/arch/x86/quants.c:290-360 (id = 127/max, round-to-nearest-even)
 *   q8_K: ggml-quants.c:2555-2592 (iscale = -127/max, nearest_int, bsums) */
void orc_quantize_row_q8_0(const float *x, void *y, int64_t k);
void orc_quantize_row_q8_K(const float *x, void *y, int64_t k);

/* weight de-quantisers (ggml-quants.c dequantize_row_q4_0/q8_0/q4_K/q5_K/q6_K) */
void orc_dequantize_row(int type, const void *x, float *y, int64_t k);

/* dot products (ggml-cpu/quants.c:115-149,305-333,550-758): w is `type`, a is the matching
 * vec_dot_type (q8_0 for Q4_0/Q8_0, q8_K for K-quants; ggml-cpu/ggml-cpu.c:209-303) */
float orc_vec_dot(int type, int64_t k, const void *w, const void *a);

/* MUL_MAT (ggml-cpu/ggml-cpu.c:1202-1394): dst[n][m] = sum_k W[m][k] * X[n][k]
 *   W: m rows of `type` blocks, row stride orc_row_bytes(type,k); X: n rows of k f32; dst f32 */
void orc_mul_mat(int type, const void *W, const float *X, float *dst, int64_t m, int64_t n, int64_t k);

/* RMS_NORM (ggml-cpu/ops.cpp:4138-4185) with the optional fused MUL by a weight row
 * (llama-graph.cpp:605-619); w may be NULL */
void orc_rms_norm(const float *x, const float *w, float *y, int64_t ncols, int64_t nrows, float eps);

/* ROPE (ggml-cpu/ops.cpp:6049-6300): x is [n_tok][n_head][head_dim] f32 contiguous,
 * mode 0 = NORM pairs (2i,2i+1), mode 2 = NEOX pairs (i, i+n_dims/2) */
void orc_rope(const float *x, float *y, const int32_t *pos, const float *freq_factors,
              int64_t head_dim, int64_t n_head, int64_t n_tok,
              int n_dims, int mode, int n_ctx_orig, float freq_base, float freq_scale,
              float ext_factor, float attn_factor, float beta_fast, float beta_slow);

/* SET_ROWS (ggml-cpu/ops.cpp:5359-5415): dst row ids[i] <- from_float(src row i);
 * dst_type F32, F16, Q8_0 or Q4_0; dst_row_stride in bytes */
void orc_set_rows(const float *src, const int64_t *ids, void *dst, int dst_type,
                  int64_t ncols, int64_t nrows, int64_t dst_row_stride);

/* FLASH_ATTN_EXT (ggml-cpu/ops.cpp:8169-8405), one sequence:
 *   q   [n_head][n_tok][dk] f32 given with byte strides (q_nb1 between tokens, q_nb2 between heads)
 *   k,v [n_head_kv][n_kv] rows of kv_type (F16, Q8_0 or Q4_0) with byte strides (nb1 row, nb2 head)
 *   mask f16 [n_tok_pad][n_kv] (row stride n_kv halves) or NULL
 *   dst [n_tok][n_head][dv] f32 contiguous (the op's permuted output, ggml.c:4814-4858) */
void orc_flash_attn_ext(const void *q, int64_t q_nb1, int64_t q_nb2,
                        const void *k, int64_t k_nb1, int64_t k_nb2,
                        const void *v, int64_t v_nb1, int64_t v_nb2,
                        const uint16_t *mask, float *dst,
                        int kv_type, int64_t dk, int64_t dv, int64_t n_head, int64_t n_head_kv,
                        int64_t n_tok, int64_t n_kv, float scale, float max_bias, float logit_softcap);

/* ---- oracle_ext.c: the formats / ops of SURVEY.md §8 rows f2-f4 (orc_dequantize_row / orc_vec_dot / orc_mul_mat route to them) ----
 * Q4_1, Q5_1, Q2_K, Q3_K, IQ4_NL, IQ4_XS, MXFP4 (ggml-common.h:176-300,414-428) */
int   orc_is_ext_type(int type);
int   orc_act_type(int weight_type);                       /* vec_dot_type: ORC_Q8_0 / ORC_Q8_1 / ORC_Q8_K (ggml-cpu/ggml-cpu.c:209-303) */
void  orc_quantize_row_q8_1(const float *x, void *y, int64_t k);   /* ggml-cpu/arch/x86/quants.c:388-492 */
/* Q4_0 as ggml's from_float writes it (KV cache type q4_0): ggml-quants.c quantize_row_q4_0_ref (ggml-cpu/quants.c:25-27 calls it) */
void  orc_quantize_row_q4_0(const float *x, void *y, int64_t k);
int   orc_dequantize_row_ext(int type, const void *x, float *y, int64_t k);
float orc_vec_dot_ext(int type, int64_t k, const void *w, const void *a);
/* MUL_MAT_ID (ggml-cpu/ggml-cpu.c:1400-1620): dst[t][s] = as[ids[t][s]] * b[t][s % n_b1];  as [n_expert][m] rows, b [n_tok][n_b1][k],
 * ids i32 [n_tok][ids_stride], dst [n_tok][n_used][m] */
void  orc_mul_mat_id(int type, const void *as, const float *b, const int32_t *ids, float *dst,
                     int64_t m, int64_t k, int64_t n_expert, int64_t n_used, int64_t n_tok, int64_t n_b1, int64_t ids_stride);
/* GET_ROWS on a quantised table (ggml-cpu/ops.cpp get_rows_q) */
void  orc_get_rows_q(int type, const void *src, const int32_t *ids, float *dst, int64_t ncols, int64_t n_ids);

/* mixture-of-experts router glue (llama-graph.cpp build_moe_ffn): SOFT_MAX without mask (ggml-cpu/ops.cpp:5685-5800), ARGSORT with the reference's
 * tie order (ops.cpp:8110-8147), SUM_ROWS (double accumulation).  The broadcasting ADD / MUL / DIV and the batched GET_ROWS of that graph are
 * exact f32 element operations; tests state them with numpy. */
void orc_soft_max_rows(const float *x, float *y, int64_t ncols, int64_t nrows, float scale);
void orc_argsort_rows(const float *x, int32_t *idx, int64_t ncols, int64_t nrows, int descending);
void orc_sum_rows(const float *x, float *y, int64_t ncols, int64_t nrows);

/* attention without -fa (llama-graph.cpp build_attn_mha): batched MUL_MAT with an f16 src0 broadcast over dim 2 (the f32 operand rounded to f16, f32
 * accumulation: ggml-cpu.c:1202-1394), SOFT_MAX with mask + ALiBi slopes (ggml-cpu/ops.cpp:5685-5800).  Strides in bytes / mask row stride in elements. */
void orc_mul_mat_f16(const void *A, int64_t a_nb1, int64_t a_nb2, int64_t a_ne2, const float *B, int64_t b_nb1, int64_t b_nb2,
                     float *dst, int64_t d_nb1, int64_t d_nb2, int64_t m, int64_t n, int64_t n_batch, int64_t k);
void orc_soft_max_mask(const float *x, float *y, const void *mask, int mask_is_f16, int64_t mask_row_stride, int64_t ncols, int64_t n_tok, int64_t n_head,
                       float scale, float max_bias);

/* SCALE / SILU / SIGMOID (op 0 / 1 / 2) as gating variants use them (ggml-cpu/ops.cpp:4815-4850, vec.h:574,691) */
void orc_unary(int op, const float *x, float *y, int64_t n, float s, float b);

/* glue (ggml-cpu/vec.h:691, ops.cpp swiglu / binary-ops.cpp / get_rows / cpy) */
void orc_swiglu(const float *gate, const float *up, float *y, int64_t n);
void orc_add(const float *a, const float *b, float *y, int64_t ncols, int64_t nrows, int64_t b_rows);
void orc_mul(const float *a, const float *b, float *y, int64_t ncols, int64_t nrows, int64_t b_rows);
void orc_get_rows_f32(const float *src, const int32_t *ids, float *dst, int64_t ncols, int64_t n_ids);
void orc_cpy_f32_f16(const float *src, uint16_t *dst, int64_t n);

#ifdef __cplusplus
}
#endif
#endif

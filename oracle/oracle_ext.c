/*
 * oracle_ext.c — the C oracle for the formats and ops of SURVEY.md §8 rows f2 / f3 / f4: Q4_1, Q5_1, Q2_K, Q3_K, IQ4_NL,
 * IQ4_XS, MXFP4 weights (de-quantisation, dot products against the reference's vec_dot_type), the q8_1 activation quantiser,
 * MUL_MAT_ID (mixture-of-experts routing) and GET_ROWS on quantised tables.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 * Written from the algorithm; every function cites the reference lines (relative to /root/reference/llama.cpp/) it follows.
 * Pinned by tests/test_oracle_vs_ref.py (live reference build) and tests/test_oracle_golden.py (tests/golden/ext_*.bin).
 */
#define _GNU_SOURCE
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#pragma pack(push, 1)
typedef struct { uint16_t d, m; uint8_t qs[16]; } blk_q4_1;                                  /* 20  ggml-common.h:176-187 */
typedef struct { uint16_t d, m; uint8_t qh[4]; uint8_t qs[16]; } blk_q5_1;                   /* 24  ggml-common.h:205-217 */
typedef struct { uint16_t d; int8_t qs[32]; } blk_q8_0;                                      /* 34  */
typedef struct { uint16_t d, s; int8_t qs[32]; } blk_q8_1;                                   /* 36  ggml-common.h:226-237 */
typedef struct { uint8_t e; uint8_t qs[16]; } blk_mxfp4;                                     /* 17  ggml-common.h:190-194 */
typedef struct { uint8_t scales[16]; uint8_t qs[64]; uint16_t d, dmin; } blk_q2_K;           /* 84  ggml-common.h:262-273 */
typedef struct { uint8_t hmask[32]; uint8_t qs[64]; uint8_t scales[12]; uint16_t d; } blk_q3_K; /* 110 ggml-common.h:280-286 */
typedef struct { uint16_t d; uint8_t qs[16]; } blk_iq4_nl;                                   /* 18  ggml-common.h:416-419 */
typedef struct { uint16_t d; uint16_t scales_h; uint8_t scales_l[4]; uint8_t qs[128]; } blk_iq4_xs; /* 136 ggml-common.h:422-427 */
typedef struct { float d; int8_t qs[256]; int16_t bsums[16]; } blk_q8_K;                     /* 292 */
#pragma pack(pop)
_Static_assert(sizeof(blk_q4_1) == 20 && sizeof(blk_q5_1) == 24 && sizeof(blk_q8_1) == 36 && sizeof(blk_mxfp4) == 17 &&
               sizeof(blk_q2_K) == 84 && sizeof(blk_q3_K) == 110 && sizeof(blk_iq4_nl) == 18 && sizeof(blk_iq4_xs) == 136, "block sizes");

/* ggml-common.h:1088-1096 */
static const int8_t KV_IQ4NL[16] = { -127, -104, -83, -65, -49, -35, -22, -10, 1, 13, 25, 38, 53, 69, 89, 113 };
static const int8_t KV_MXFP4[16] = { 0, 1, 2, 3, 4, 6, 8, 12, 0, -1, -2, -3, -4, -6, -8, -12 };

/* ggml-impl.h:451-470: E8M0 exponent byte -> 2^(x-128) */
static float e8m0_half(uint8_t x) {
    uint32_t bits = x < 2 ? (0x00200000u << x) : ((uint32_t)(x - 1) << 23);
    float f; memcpy(&f, &bits, 4); return f;
}

int orc_is_ext_type(int t) { return t == ORC_Q4_1 || t == ORC_Q5_1 || t == ORC_Q2_K || t == ORC_Q3_K || t == ORC_IQ4_NL || t == ORC_IQ4_XS || t == ORC_MXFP4; }
/* vec_dot_type (ggml-cpu/ggml-cpu.c:209-303) */
int orc_act_type(int t) {
    switch (t) {
        case ORC_Q4_0: case ORC_Q5_0: case ORC_Q8_0: case ORC_IQ4_NL: case ORC_MXFP4: return ORC_Q8_0;
        case ORC_Q4_1: case ORC_Q5_1: return ORC_Q8_1;
        default: return ORC_Q8_K;
    }
}

/* q8_1 as the x86 CPU backend computes it (ggml-cpu/arch/x86/quants.c:388-492, AVX2 branch): the same int8 values as its q8_0
 * (multiplier 127/max, round half to even); d = max/127 stored as f16; s = f16(d_f32 * sum of the quants). */
void orc_quantize_row_q8_1(const float *x, void *vy, int64_t k) {
    blk_q8_1 *y = (blk_q8_1 *)vy;
    for (int64_t b = 0; b < k / 32; b++) {
        float amax = 0.0f;
        for (int j = 0; j < 32; j++) { float a = fabsf(x[b*32 + j]); if (a > amax) amax = a; }
        const float d  = amax / 127.0f;
        const float id = amax != 0.0f ? 127.0f / amax : 0.0f;
        int sum = 0;
        for (int j = 0; j < 32; j++) { int q = (int)lrintf(x[b*32 + j] * id); y[b].qs[j] = (int8_t)q; sum += q; }
        y[b].d = orc_fp32_to_fp16(d);
        y[b].s = orc_fp32_to_fp16(d * (float)sum);
    }
}

/* Q4_0 from f32 (ggml-quants.c quantize_row_q4_0_ref): the element of largest magnitude (first one on ties) maps to -8; values are
 * truncated after adding 8.5, clamped to 15 */
void orc_quantize_row_q4_0(const float *x, void *vy, int64_t k) {
    uint8_t *y = (uint8_t *)vy;
    for (int64_t b = 0; b < k / 32; b++, y += 18) {
        float amax = 0.0f, mx = 0.0f;
        for (int j = 0; j < 32; j++) { const float v = x[b*32 + j]; if (amax < fabsf(v)) { amax = fabsf(v); mx = v; } }
        const float d = mx / -8.0f, id = d != 0.0f ? 1.0f / d : 0.0f;
        const uint16_t dh = orc_fp32_to_fp16(d);
        memcpy(y, &dh, 2);
        for (int j = 0; j < 16; j++) {
            const float x0 = x[b*32 + j] * id, x1 = x[b*32 + 16 + j] * id;
            int q0 = (int)(int8_t)(x0 + 8.5f), q1 = (int)(int8_t)(x1 + 8.5f);
            if (q0 > 15) q0 = 15;
            if (q1 > 15) q1 = 15;
            y[2 + j] = (uint8_t)(q0 | (q1 << 4));
        }
    }
}

/* Q3_K: the 16 six-bit scales of a super-block, unpacked (ggml-quants.c:1143-1148) */
static void q3K_scales(const uint8_t *sc12, int8_t *out16) {
    uint32_t aux[4]; memcpy(aux, sc12, 12);
    const uint32_t k1 = 0x03030303u, k2 = 0x0f0f0f0fu, tmp = aux[2];
    aux[2] = ((aux[0] >> 4) & k2) | (((tmp >> 4) & k1) << 4);
    aux[3] = ((aux[1] >> 4) & k2) | (((tmp >> 6) & k1) << 4);
    aux[0] = (aux[0] & k2) | (((tmp >> 0) & k1) << 4);
    aux[1] = (aux[1] & k2) | (((tmp >> 2) & k1) << 4);
    memcpy(out16, aux, 16);
}
/* integer values of a Q2_K / Q3_K super-block in element order (ggml-quants.c:784-816, 1128-1178) */
static void q2K_ints(const blk_q2_K *w, int *q) {
    for (int n = 0; n < 2; n++) for (int j = 0; j < 4; j++) for (int l = 0; l < 32; l++) q[128*n + 32*j + l] = (w->qs[32*n + l] >> (2*j)) & 3;
}
static void q3K_ints(const blk_q3_K *w, int *q) {
    for (int n = 0; n < 2; n++) for (int j = 0; j < 4; j++) for (int l = 0; l < 32; l++)
        q[128*n + 32*j + l] = ((w->qs[32*n + l] >> (2*j)) & 3) - ((w->hmask[l] & (1 << (4*n + j))) ? 0 : 4);
}
static int iq4xs_ls(const blk_iq4_xs *w, int ib) { return ((w->scales_l[ib/2] >> (4*(ib%2))) & 0xf) | (((w->scales_h >> (2*ib)) & 3) << 4); }

/* de-quantisers (ggml-quants.c:327-345 q4_1, 374-398 q5_1, 417-434 mxfp4, 784-816 q2_K, 1128-1178 q3_K, 2512-2528 iq4_nl, 2530-2552 iq4_xs) */
int orc_dequantize_row_ext(int type, const void *vx, float *y, int64_t k) {
    switch (type) {
        case ORC_Q4_1: {
            const blk_q4_1 *x = (const blk_q4_1 *)vx;
            for (int64_t b = 0; b < k/32; b++) {
                const float d = orc_fp16_to_fp32(x[b].d), m = orc_fp16_to_fp32(x[b].m);
                for (int j = 0; j < 16; j++) { y[b*32 + j] = (float)(x[b].qs[j] & 15) * d + m; y[b*32 + j + 16] = (float)(x[b].qs[j] >> 4) * d + m; }
            }
            return 1;
        }
        case ORC_Q5_1: {
            const blk_q5_1 *x = (const blk_q5_1 *)vx;
            for (int64_t b = 0; b < k/32; b++) {
                const float d = orc_fp16_to_fp32(x[b].d), m = orc_fp16_to_fp32(x[b].m);
                uint32_t qh; memcpy(&qh, x[b].qh, 4);
                for (int j = 0; j < 16; j++) {
                    const int h0 = ((qh >> j) << 4) & 0x10, h1 = (qh >> (j + 12)) & 0x10;
                    y[b*32 + j] = (float)((x[b].qs[j] & 15) | h0) * d + m; y[b*32 + j + 16] = (float)((x[b].qs[j] >> 4) | h1) * d + m;
                }
            }
            return 1;
        }
        case ORC_MXFP4: {
            const blk_mxfp4 *x = (const blk_mxfp4 *)vx;
            for (int64_t b = 0; b < k/32; b++) {
                const float d = e8m0_half(x[b].e);
                for (int j = 0; j < 16; j++) { y[b*32 + j] = (float)KV_MXFP4[x[b].qs[j] & 15] * d; y[b*32 + j + 16] = (float)KV_MXFP4[x[b].qs[j] >> 4] * d; }
            }
            return 1;
        }
        case ORC_IQ4_NL: {
            const blk_iq4_nl *x = (const blk_iq4_nl *)vx;
            for (int64_t b = 0; b < k/32; b++) {
                const float d = orc_fp16_to_fp32(x[b].d);
                for (int j = 0; j < 16; j++) { y[b*32 + j] = d * (float)KV_IQ4NL[x[b].qs[j] & 15]; y[b*32 + j + 16] = d * (float)KV_IQ4NL[x[b].qs[j] >> 4]; }
            }
            return 1;
        }
        case ORC_Q2_K: {
            const blk_q2_K *x = (const blk_q2_K *)vx; int q[256];
            for (int64_t b = 0; b < k/256; b++, y += 256) {
                const float d = orc_fp16_to_fp32(x[b].d), mn = orc_fp16_to_fp32(x[b].dmin);
                q2K_ints(&x[b], q);
                for (int g = 0; g < 16; g++) {
                    const float dl = d * (float)(x[b].scales[g] & 15), ml = mn * (float)(x[b].scales[g] >> 4);
                    for (int l = 0; l < 16; l++) y[16*g + l] = dl * (float)q[16*g + l] - ml;
                }
            }
            return 1;
        }
        case ORC_Q3_K: {
            const blk_q3_K *x = (const blk_q3_K *)vx; int q[256]; int8_t sc[16];
            for (int64_t b = 0; b < k/256; b++, y += 256) {
                const float d = orc_fp16_to_fp32(x[b].d);
                q3K_ints(&x[b], q); q3K_scales(x[b].scales, sc);
                for (int g = 0; g < 16; g++) {
                    const float dl = d * (float)(sc[g] - 32);
                    for (int l = 0; l < 16; l++) y[16*g + l] = dl * (float)q[16*g + l];
                }
            }
            return 1;
        }
        case ORC_IQ4_XS: {
            const blk_iq4_xs *x = (const blk_iq4_xs *)vx;
            for (int64_t b = 0; b < k/256; b++, y += 256) {
                const float d = orc_fp16_to_fp32(x[b].d);
                for (int ib = 0; ib < 8; ib++) {
                    const float dl = d * (float)(iq4xs_ls(&x[b], ib) - 32);
                    for (int j = 0; j < 16; j++) { y[32*ib + j] = dl * (float)KV_IQ4NL[x[b].qs[16*ib + j] & 15]; y[32*ib + j + 16] = dl * (float)KV_IQ4NL[x[b].qs[16*ib + j] >> 4]; }
                }
            }
            return 1;
        }
        default: return 0;
    }
}

/* dot products: the generic functions' arithmetic (ggml-cpu/quants.c:152-186 q4_1, 262-303 q5_1, 188-217 mxfp4, 1108-1135 iq4_nl,
 * 419-469 q2_K, 471-548 q3_K, 1137-1183 iq4_xs); `a` is the weight type's vec_dot_type (orc_act_type) */
float orc_vec_dot_ext(int type, int64_t k, const void *vw, const void *va) {
    float acc = 0.0f;
    switch (type) {
        case ORC_Q4_1: {
            const blk_q4_1 *w = (const blk_q4_1 *)vw; const blk_q8_1 *a = (const blk_q8_1 *)va;
            for (int64_t b = 0; b < k/32; b++) {
                int s = 0;
                for (int j = 0; j < 16; j++) s += (w[b].qs[j] & 15) * a[b].qs[j] + (w[b].qs[j] >> 4) * a[b].qs[j + 16];
                acc += (orc_fp16_to_fp32(w[b].d) * orc_fp16_to_fp32(a[b].d)) * (float)s + orc_fp16_to_fp32(w[b].m) * orc_fp16_to_fp32(a[b].s);
            }
            return acc;
        }
        case ORC_Q5_1: {
            const blk_q5_1 *w = (const blk_q5_1 *)vw; const blk_q8_1 *a = (const blk_q8_1 *)va;
            for (int64_t b = 0; b < k/32; b++) {
                uint32_t qh; memcpy(&qh, w[b].qh, 4);
                int s = 0;
                for (int j = 0; j < 16; j++) {
                    const int h0 = ((qh >> j) << 4) & 0x10, h1 = (qh >> (j + 12)) & 0x10;
                    s += ((w[b].qs[j] & 15) | h0) * a[b].qs[j] + ((w[b].qs[j] >> 4) | h1) * a[b].qs[j + 16];
                }
                acc += (orc_fp16_to_fp32(w[b].d) * orc_fp16_to_fp32(a[b].d)) * (float)s + orc_fp16_to_fp32(w[b].m) * orc_fp16_to_fp32(a[b].s);
            }
            return acc;
        }
        case ORC_MXFP4: case ORC_IQ4_NL: {
            const blk_q8_0 *a = (const blk_q8_0 *)va;
            for (int64_t b = 0; b < k/32; b++) {
                const uint8_t *qs; float dw; const int8_t *kv;
                if (type == ORC_MXFP4) { const blk_mxfp4 *w = (const blk_mxfp4 *)vw + b; qs = w->qs; dw = e8m0_half(w->e); kv = KV_MXFP4; }
                else                   { const blk_iq4_nl *w = (const blk_iq4_nl *)vw + b; qs = w->qs; dw = orc_fp16_to_fp32(w->d); kv = KV_IQ4NL; }
                int s = 0;
                for (int j = 0; j < 16; j++) s += a[b].qs[j] * kv[qs[j] & 15] + a[b].qs[j + 16] * kv[qs[j] >> 4];
                acc += (orc_fp16_to_fp32(a[b].d) * dw) * (float)s;
            }
            return acc;
        }
        case ORC_Q2_K: {
            const blk_q2_K *w = (const blk_q2_K *)vw; const blk_q8_K *a = (const blk_q8_K *)va; int q[256];
            for (int64_t b = 0; b < k/256; b++) {
                q2K_ints(&w[b], q);
                int summs = 0, isum = 0;
                for (int g = 0; g < 16; g++) {
                    summs += a[b].bsums[g] * (w[b].scales[g] >> 4);
                    int t = 0;
                    for (int l = 0; l < 16; l++) t += a[b].qs[16*g + l] * q[16*g + l];
                    isum += (w[b].scales[g] & 15) * t;
                }
                const float dall = a[b].d * orc_fp16_to_fp32(w[b].d), dmin = a[b].d * orc_fp16_to_fp32(w[b].dmin);
                acc += dall * (float)isum - dmin * (float)summs;
            }
            return acc;
        }
        case ORC_Q3_K: {
            /* the generic code keeps 8 float lanes (element index mod 8) and folds them at the end */
            const blk_q3_K *w = (const blk_q3_K *)vw; const blk_q8_K *a = (const blk_q8_K *)va; int q[256]; int8_t sc[16];
            float lanes[8] = {0};
            for (int64_t b = 0; b < k/256; b++) {
                q3K_ints(&w[b], q); q3K_scales(w[b].scales, sc);
                int32_t li[8] = {0};
                for (int g = 0; g < 16; g++) for (int l = 0; l < 16; l++) li[l & 7] += (sc[g] - 32) * (a[b].qs[16*g + l] * q[16*g + l]);
                const float d = orc_fp16_to_fp32(w[b].d) * a[b].d;
                for (int l = 0; l < 8; l++) lanes[l] += d * (float)li[l];
            }
            for (int l = 0; l < 8; l++) acc += lanes[l];
            return acc;
        }
        case ORC_IQ4_XS: {
            const blk_iq4_xs *w = (const blk_iq4_xs *)vw; const blk_q8_K *a = (const blk_q8_K *)va;
            for (int64_t b = 0; b < k/256; b++) {
                const float d4d8 = orc_fp16_to_fp32(w[b].d) * a[b].d;
                for (int ib = 0; ib < 8; ib++) {
                    const float dl = d4d8 * (float)(iq4xs_ls(&w[b], ib) - 32);
                    int s = 0;
                    for (int j = 0; j < 16; j++) s += a[b].qs[32*ib + j] * KV_IQ4NL[w[b].qs[16*ib + j] & 15] + a[b].qs[32*ib + j + 16] * KV_IQ4NL[w[b].qs[16*ib + j] >> 4];
                    acc += dl * (float)s;
                }
            }
            return acc;
        }
        default: return NAN;
    }
}

/* MUL_MAT_ID (ggml-cpu/ggml-cpu.c:1400-1620; contract ggml.c:3064-3106):
 *   as  [n_expert][m] rows of `type` (k elements), b f32 [n_tok][n_b1][k] with n_b1 = 1 (shared) or n_used,
 *   ids i32 [n_tok][ids_stride] (the first n_used entries of each row are used), dst f32 [n_tok][n_used][m]:
 *   dst[t][s] = as[ids[t][s]] * b[t][s % n_b1] */
void orc_mul_mat_id(int type, const void *as, const float *b, const int32_t *ids, float *dst,
                    int64_t m, int64_t k, int64_t n_expert, int64_t n_used, int64_t n_tok, int64_t n_b1, int64_t ids_stride) {
    const int64_t expert_bytes = m * orc_row_bytes(type, k);
    for (int64_t t = 0; t < n_tok; t++)
        for (int64_t s = 0; s < n_used; s++) {
            const int32_t e = ids[t * ids_stride + s];
            if (e < 0 || e >= n_expert) continue;
            orc_mul_mat(type, (const uint8_t *)as + (int64_t)e * expert_bytes, b + (t * n_b1 + s % n_b1) * k, dst + (t * n_used + s) * m, m, 1, k);
        }
}

/* GET_ROWS on a quantised table (ggml-cpu/ops.cpp ggml_compute_forward_get_rows_q: de-quantise row ids[i]) */
void orc_get_rows_q(int type, const void *src, const int32_t *ids, float *dst, int64_t ncols, int64_t n_ids) {
    const int64_t rb = orc_row_bytes(type, ncols);
    for (int64_t i = 0; i < n_ids; i++) orc_dequantize_row(type, (const uint8_t *)src + (int64_t)ids[i] * rb, dst + i * ncols, ncols);
}

/* ---- mixture-of-experts router glue (llama.cpp/src/llama-graph.cpp build_moe_ffn) ---- */

/* SOFT_MAX of rows without mask (ggml-cpu/ops.cpp:5685-5800; vec.cpp ggml_vec_soft_max_f32): w = x*scale, exp(w - max) summed in double,
 * scaled by (float)(1/sum).  The reference evaluates exp with expf or with its vectorised polynomial depending on build and row length;
 * both are within 2 ulp of this expf. */
void orc_soft_max_rows(const float *x, float *y, int64_t ncols, int64_t nrows, float scale) {
    for (int64_t r = 0; r < nrows; r++) {
        const float *xr = x + r*ncols; float *yr = y + r*ncols;
        float mx = -INFINITY;
        for (int64_t i = 0; i < ncols; i++) { const float w = xr[i] * scale; if (w > mx) mx = w; }
        double sum = 0.0;
        for (int64_t i = 0; i < ncols; i++) { const float v = expf(xr[i] * scale - mx); yr[i] = v; sum += (double)v; }
        const float inv = (float)(1.0 / sum);
        for (int64_t i = 0; i < ncols; i++) yr[i] *= inv;
    }
}
/* ARGSORT (ggml-cpu/ops.cpp:8110-8147): the reference's exchange sort, whose result for equal keys is part of the contract (TOP_K) */
void orc_argsort_rows(const float *x, int32_t *idx, int64_t ncols, int64_t nrows, int descending) {
    for (int64_t r = 0; r < nrows; r++) {
        const float *xr = x + r*ncols; int32_t *d = idx + r*ncols;
        for (int64_t j = 0; j < ncols; j++) d[j] = (int32_t)j;
        for (int64_t j = 0; j < ncols; j++) for (int64_t k = j + 1; k < ncols; k++)
            if (descending ? xr[d[j]] < xr[d[k]] : xr[d[j]] > xr[d[k]]) { const int32_t t = d[j]; d[j] = d[k]; d[k] = t; }
    }
}
/* SUM_ROWS (ggml-cpu/ops.cpp sum_rows_f32 -> ggml_vec_sum_f32): sequential double sum, stored as float */
void orc_sum_rows(const float *x, float *y, int64_t ncols, int64_t nrows) {
    for (int64_t r = 0; r < nrows; r++) { double s = 0.0; for (int64_t i = 0; i < ncols; i++) s += (double)x[r*ncols + i]; y[r] = (float)s; }
}

/* ---- attention without -fa (llama-graph.cpp build_attn_mha, non-flash branch) ---- */

/* batched MUL_MAT with an f16 src0 (ggml-cpu/ggml-cpu.c:1202-1394): src1 rows are converted to f16 (vec_dot_type of F16), products accumulated in f32
 * (ggml_vec_dot_f16); src0 is broadcast over dim 2 (i02 = i12 / r2).  Strides in bytes. */
void orc_mul_mat_f16(const void *A, int64_t a_nb1, int64_t a_nb2, int64_t a_ne2, const float *B, int64_t b_nb1, int64_t b_nb2,
                     float *dst, int64_t d_nb1, int64_t d_nb2, int64_t m, int64_t n, int64_t n_batch, int64_t k) {
    const int64_t r2 = n_batch / a_ne2;
    for (int64_t i2 = 0; i2 < n_batch; i2++) for (int64_t i1 = 0; i1 < n; i1++) for (int64_t i0 = 0; i0 < m; i0++) {
        const uint16_t *ar = (const uint16_t *)((const char *)A + i0*a_nb1 + (i2/r2)*a_nb2);
        const float *br = (const float *)((const char *)B + i1*b_nb1 + i2*b_nb2);
        float acc = 0.0f;
        for (int64_t i = 0; i < k; i++) acc += orc_fp16_to_fp32(ar[i]) * orc_fp16_to_fp32(orc_fp32_to_fp16(br[i]));
        *(float *)((char *)dst + i0*4 + i1*d_nb1 + i2*d_nb2) = acc;
    }
}
/* SOFT_MAX with a mask and ALiBi slopes (ggml-cpu/ops.cpp:5685-5800): x [ncols, n_tok, n_head] contiguous, mask one row per token (f32 or f16) */
void orc_soft_max_mask(const float *x, float *y, const void *mask, int mask_is_f16, int64_t mask_row_stride, int64_t ncols, int64_t n_tok, int64_t n_head,
                       float scale, float max_bias) {
    const uint32_t nh_log2 = 1u << (uint32_t)floor(log2((double)n_head));
    const float m0 = powf(2.0f, -max_bias / (float)nh_log2), m1 = powf(2.0f, -(max_bias / 2.0f) / (float)nh_log2);
    for (int64_t h = 0; h < n_head; h++) for (int64_t t = 0; t < n_tok; t++) {
        const float slope = max_bias > 0.0f ? ((uint32_t)h < nh_log2 ? powf(m0, (float)(h + 1)) : powf(m1, (float)(2*(h - nh_log2) + 1))) : 1.0f;
        const float *xr = x + (h*n_tok + t)*ncols; float *yr = y + (h*n_tok + t)*ncols;
        float mx = -INFINITY;
        for (int64_t i = 0; i < ncols; i++) {
            const float mv = mask_is_f16 ? orc_fp16_to_fp32(((const uint16_t *)mask)[t*mask_row_stride + i]) : ((const float *)mask)[t*mask_row_stride + i];
            yr[i] = xr[i] * scale + slope * mv;
            if (yr[i] > mx) mx = yr[i];
        }
        double sum = 0.0;
        for (int64_t i = 0; i < ncols; i++) { const float v = expf(yr[i] - mx); yr[i] = v; sum += (double)v; }
        const float inv = (float)(1.0 / sum);
        for (int64_t i = 0; i < ncols; i++) yr[i] *= inv;
    }
}

/* SCALE (ggml-cpu/ops.cpp:4815-4850: x*s, or the FMA x*s + b of ggml_vec_mad1_f32 when b != 0), SILU and SIGMOID in their scalar forms
 * (ggml-cpu/vec.h:574,691: expf; the vector body of SILU uses the polynomial exp of orc_swiglu: both within 2 ulp) */
void orc_unary(int op, const float *x, float *y, int64_t n, float s, float b) {
    for (int64_t i = 0; i < n; i++) {
        if (op == 0) y[i] = b == 0.0f ? x[i] * s : fmaf(x[i], s, b);
        else if (op == 1) y[i] = x[i] / (1.0f + expf(-x[i]));
        else y[i] = 1.0f / (1.0f + expf(-x[i]));
    }
}


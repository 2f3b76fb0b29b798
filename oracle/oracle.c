/*
 * oracle.c — plain-C restatement of the reference's ggml-cpu algorithms for the hot path.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Written from the algorithm, not copied: each
 * function cites the reference file:line (relative to /root/reference/llama.cpp/) it follows.
 * Parity: PINNED against oracle/_ref (the unmodified reference build) by
 * tests/test_oracle_vs_ref.py and against tests/golden/ fixtures by tests/test_oracle_golden.py.
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off (no FMA contraction: the reference's scalar
 * code paths round every product).
 */
#define _GNU_SOURCE
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define QK  32
#define QKK 256

/* ---- block layouts (ggml/src/ggml-common.h:170-175,219-224,295-344); byte-packed ---- */
#pragma pack(push, 1)
typedef struct { uint16_t d; uint8_t qs[16]; } blk_q4_0;                                   /* 18  */
typedef struct { uint16_t d; uint8_t qh[4]; uint8_t qs[16]; } blk_q5_0;                    /* 22  (ggml-common.h:186-191) */
typedef struct { uint16_t d; int8_t  qs[32]; } blk_q8_0;                                   /* 34  */
typedef struct { uint16_t d, dmin; uint8_t sc[12]; uint8_t qs[128]; } blk_q4_K;            /* 144 */
typedef struct { uint16_t d, dmin; uint8_t sc[12]; uint8_t qh[32]; uint8_t qs[128]; } blk_q5_K; /* 176 */
typedef struct { uint8_t ql[128]; uint8_t qh[64]; int8_t sc[16]; uint16_t d; } blk_q6_K;   /* 210 */
typedef struct { float d; int8_t qs[256]; int16_t bsums[16]; } blk_q8_K;                   /* 292 */
#pragma pack(pop)

_Static_assert(sizeof(blk_q4_0) == 18,  "q4_0");
_Static_assert(sizeof(blk_q5_0) == 22,  "q5_0");
_Static_assert(sizeof(blk_q8_0) == 34,  "q8_0");
_Static_assert(sizeof(blk_q4_K) == 144, "q4_K");
_Static_assert(sizeof(blk_q5_K) == 176, "q5_K");
_Static_assert(sizeof(blk_q6_K) == 210, "q6_K");
_Static_assert(sizeof(blk_q8_K) == 292, "q8_K");

int64_t orc_block_elems(int type) {
    switch (type) {
        case ORC_F32: case ORC_F16: return 1;
        case ORC_Q4_0: case ORC_Q5_0: case ORC_Q8_0: case ORC_Q4_1: case ORC_Q5_1: case ORC_Q8_1: case ORC_IQ4_NL: case ORC_MXFP4: return QK;
        default: return QKK;
    }
}
int64_t orc_block_bytes(int type) {
    switch (type) {
        case ORC_F32:  return 4;
        case ORC_F16:  return 2;
        case ORC_Q4_0: return 18;
        case ORC_Q5_0: return 22;
        case ORC_Q8_0: return 34;
        case ORC_Q4_K: return 144;
        case ORC_Q5_K: return 176;
        case ORC_Q6_K: return 210;
        case ORC_Q8_K: return 292;
        case ORC_Q4_1: return 20;  case ORC_Q5_1: return 24;  case ORC_Q8_1: return 36;   /* ggml-common.h:176-237 */
        case ORC_Q2_K: return 84;  case ORC_Q3_K: return 110;                               /* ggml-common.h:262-286 */
        case ORC_IQ4_NL: return 18; case ORC_IQ4_XS: return 136; case ORC_MXFP4: return 17; /* ggml-common.h:190-194,414-428 */
        default: return 0;
    }
}
int64_t orc_row_bytes(int type, int64_t k) { return k / orc_block_elems(type) * orc_block_bytes(type); }

/* ---- fp16 (IEEE binary16, RNE; what F16C / ggml_compute_fp32_to_fp16 produce) ---- */
float orc_fp16_to_fp32(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    const uint32_t exp  = (h >> 10) & 0x1f;
    const uint32_t man  = h & 0x3ffu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else { /* subnormal: normalise */
            int e = -1;
            uint32_t m = man;
            do { e++; m <<= 1; } while ((m & 0x400u) == 0);
            bits = sign | (uint32_t)(127 - 15 - e) << 23 | (m & 0x3ffu) << 13;
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | man << 13;
    } else {
        bits = sign | (exp + 112) << 23 | man << 13;
    }
    float f; memcpy(&f, &bits, 4); return f;
}

uint16_t orc_fp32_to_fp16(float f) {
    uint32_t x; memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7fffffffu;
    if (x >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | (x > 0x7f800000u ? 0x200u | ((x >> 13) & 0x3ffu) : 0));
    if (x >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);       /* rounds to >= 65520 -> inf */
    if (x < 0x33000001u)  return (uint16_t)sign;                   /* < 2^-25 (or == 2^-25: ties to even 0) */
    const int e = (int)(x >> 23) - 127;
    uint32_t man = (x & 0x7fffffu) | 0x800000u;
    int shift;
    uint32_t base;
    if (e < -14) { shift = 13 + (-14 - e); base = 0; }             /* subnormal half */
    else         { shift = 13; base = (uint32_t)(e + 15) << 10; man &= 0x7fffffu; }
    uint32_t q   = man >> shift;
    uint32_t rem = man & ((1u << shift) - 1);
    uint32_t half = 1u << (shift - 1);
    if (rem > half || (rem == half && (q & 1))) q++;
    return (uint16_t)(sign | (base + q));                          /* carry propagates into exponent */
}

/* ---- activation quantisers ---- */

/* ggml-cpu/arch/x86/quants.c:290-360: per 32: d = max|x|/127 (stored f16), multiplier 127/max,
 * round half to even, no clamp needed (|x*id| <= 127). */
void orc_quantize_row_q8_0(const float *x, void *vy, int64_t k) {
    blk_q8_0 *y = (blk_q8_0 *)vy;
    for (int64_t b = 0; b < k / QK; b++) {
        float amax = 0.0f;
        for (int j = 0; j < QK; j++) { float a = fabsf(x[b*QK + j]); if (a > amax) amax = a; }
        const float d  = amax / 127.0f;
        const float id = amax != 0.0f ? 127.0f / amax : 0.0f;
        y[b].d = orc_fp32_to_fp16(d);
        for (int j = 0; j < QK; j++) y[b].qs[j] = (int8_t)lrintf(x[b*QK + j] * id); /* RNE in default mode */
    }
}

/* ggml-quants.c:2555-2592 (x86 calls this reference impl: arch/x86/quants.c:493-495) */
void orc_quantize_row_q8_K(const float *x, void *vy, int64_t k) {
    blk_q8_K *y = (blk_q8_K *)vy;
    for (int64_t b = 0; b < k / QKK; b++, x += QKK) {
        float amax = 0.0f, vmax = 0.0f;            /* vmax keeps the sign of the first largest |x| */
        for (int j = 0; j < QKK; j++) { float a = fabsf(x[j]); if (a > amax) { amax = a; vmax = x[j]; } }
        if (amax == 0.0f) {
            y[b].d = 0.0f; memset(y[b].qs, 0, QKK);
            /* bsums are left untouched by the reference; zero them so the block is deterministic
               (the dot products multiply them by d == 0 anyway) */
            memset(y[b].bsums, 0, sizeof(y[b].bsums));
            continue;
        }
        const float iscale = -127.0f / vmax;
        for (int j = 0; j < QKK; j++) {
            int v = (int)lrintf(iscale * x[j]);    /* nearest_int(): round half to even */
            y[b].qs[j] = (int8_t)(v > 127 ? 127 : v);
        }
        for (int g = 0; g < 16; g++) {
            int s = 0;
            for (int j = 0; j < 16; j++) s += y[b].qs[g*16 + j];
            y[b].bsums[g] = (int16_t)s;
        }
        y[b].d = 1.0f / iscale;
    }
}

/* ---- 6-bit packed scale/min pairs of q4_K / q5_K (ggml-quants.c:703-711) ---- */
static void k4_scale_min(int j, const uint8_t *p, int *sc, int *mn) {
    if (j < 4) { *sc = p[j] & 63;                          *mn = p[j + 4] & 63; }
    else       { *sc = (p[j + 4] & 15) | ((p[j - 4] >> 6) << 4); *mn = (p[j + 4] >> 4) | ((p[j] >> 6) << 4); }
}

/* unpack one super-block's quants to integers in element order */
static void q4K_ints(const blk_q4_K *w, int *q) {          /* ggml-quants.c dequantize_row_q4_K */
    for (int c = 0; c < 4; c++)
        for (int l = 0; l < 32; l++) {
            q[c*64 + l]      = w->qs[c*32 + l] & 15;
            q[c*64 + 32 + l] = w->qs[c*32 + l] >> 4;
        }
}
static void q5K_ints(const blk_q5_K *w, int *q) {          /* ggml-quants.c dequantize_row_q5_K */
    for (int c = 0; c < 4; c++)
        for (int l = 0; l < 32; l++) {
            q[c*64 + l]      = (w->qs[c*32 + l] & 15) + (((w->qh[l] >> (2*c))     & 1) << 4);
            q[c*64 + 32 + l] = (w->qs[c*32 + l] >> 4) + (((w->qh[l] >> (2*c + 1)) & 1) << 4);
        }
}
static void q6K_ints(const blk_q6_K *w, int *q) {          /* ggml-quants.c dequantize_row_q6_K */
    for (int h = 0; h < 2; h++)
        for (int l = 0; l < 32; l++) {
            const uint8_t *ql = w->ql + 64*h, *qh = w->qh + 32*h;
            q[128*h + l]      = ((ql[l]      & 15) | (((qh[l] >> 0) & 3) << 4)) - 32;
            q[128*h + 32 + l] = ((ql[l + 32] & 15) | (((qh[l] >> 2) & 3) << 4)) - 32;
            q[128*h + 64 + l] = ((ql[l]      >> 4) | (((qh[l] >> 4) & 3) << 4)) - 32;
            q[128*h + 96 + l] = ((ql[l + 32] >> 4) | (((qh[l] >> 6) & 3) << 4)) - 32;
        }
}

void orc_dequantize_row(int type, const void *vx, float *y, int64_t k) {
    if (orc_dequantize_row_ext(type, vx, y, k)) return;            /* oracle_ext.c */
    if (type == ORC_F32) { memcpy(y, vx, (size_t)k*4); return; }
    if (type == ORC_F16) { for (int64_t i = 0; i < k; i++) y[i] = orc_fp16_to_fp32(((const uint16_t *)vx)[i]); return; }
    if (type == ORC_Q4_0) {                                /* ggml-quants.c dequantize_row_q4_0 */
        const blk_q4_0 *x = (const blk_q4_0 *)vx;
        for (int64_t b = 0; b < k/QK; b++) {
            const float d = orc_fp16_to_fp32(x[b].d);
            for (int j = 0; j < 16; j++) {
                y[b*QK + j]      = (float)((x[b].qs[j] & 15) - 8) * d;
                y[b*QK + j + 16] = (float)((x[b].qs[j] >> 4) - 8) * d;
            }
        }
        return;
    }
    if (type == ORC_Q5_0) {                                /* ggml-quants.c dequantize_row_q5_0: fifth bit j / j + 16 of qh */
        const blk_q5_0 *x = (const blk_q5_0 *)vx;
        for (int64_t b = 0; b < k/QK; b++) {
            const float d = orc_fp16_to_fp32(x[b].d);
            uint32_t qh; memcpy(&qh, x[b].qh, 4);
            for (int j = 0; j < 16; j++) {
                const int xh0 = ((qh >> (j + 0)) << 4) & 0x10, xh1 = (qh >> (j + 12)) & 0x10;
                y[b*QK + j]      = (float)(((x[b].qs[j] & 15) | xh0) - 16) * d;
                y[b*QK + j + 16] = (float)(((x[b].qs[j] >> 4) | xh1) - 16) * d;
            }
        }
        return;
    }
    if (type == ORC_Q8_0) {                                /* ggml-quants.c dequantize_row_q8_0 */
        const blk_q8_0 *x = (const blk_q8_0 *)vx;
        for (int64_t b = 0; b < k/QK; b++) {
            const float d = orc_fp16_to_fp32(x[b].d);
            for (int j = 0; j < QK; j++) y[b*QK + j] = (float)x[b].qs[j] * d;
        }
        return;
    }
    int q[QKK];
    for (int64_t b = 0; b < k/QKK; b++, y += QKK) {
        if (type == ORC_Q4_K || type == ORC_Q5_K) {
            const uint8_t *sc; float d, dmin;
            if (type == ORC_Q4_K) { const blk_q4_K *w = (const blk_q4_K *)vx + b; q4K_ints(w, q); sc = w->sc; d = orc_fp16_to_fp32(w->d); dmin = orc_fp16_to_fp32(w->dmin); }
            else                  { const blk_q5_K *w = (const blk_q5_K *)vx + b; q5K_ints(w, q); sc = w->sc; d = orc_fp16_to_fp32(w->d); dmin = orc_fp16_to_fp32(w->dmin); }
            for (int g = 0; g < 8; g++) {
                int s, m; k4_scale_min(g, sc, &s, &m);
                const float ds = d * (float)s, ms = dmin * (float)m;
                for (int l = 0; l < 32; l++) y[g*32 + l] = ds * (float)q[g*32 + l] - ms;
            }
        } else if (type == ORC_Q6_K) {
            const blk_q6_K *w = (const blk_q6_K *)vx + b; q6K_ints(w, q);
            const float d = orc_fp16_to_fp32(w->d);
            for (int g = 0; g < 16; g++)
                for (int l = 0; l < 16; l++) y[g*16 + l] = d * (float)w->sc[g] * (float)q[g*16 + l];
        } else if (type == ORC_Q8_K) {
            const blk_q8_K *w = (const blk_q8_K *)vx + b;
            for (int l = 0; l < QKK; l++) y[l] = w->d * (float)w->qs[l];
        }
    }
}

/* ---- dot products ---- */

static float dot_q4_0(int64_t k, const blk_q4_0 *w, const blk_q8_0 *a) {   /* ggml-cpu/quants.c:115-149 */
    float acc = 0.0f;
    for (int64_t b = 0; b < k/QK; b++) {
        int s = 0;
        for (int j = 0; j < 16; j++)
            s += ((w[b].qs[j] & 15) - 8) * a[b].qs[j] + ((w[b].qs[j] >> 4) - 8) * a[b].qs[j + 16];
        acc += (float)s * orc_fp16_to_fp32(w[b].d) * orc_fp16_to_fp32(a[b].d);
    }
    return acc;
}
static float dot_q5_0(int64_t k, const blk_q5_0 *w, const blk_q8_0 *a) {   /* ggml-cpu/quants.c ggml_vec_dot_q5_0_q8_0_generic; x86: arch/x86/quants.c:845-925 */
    float acc = 0.0f;
    for (int64_t b = 0; b < k/QK; b++) {
        uint32_t qh; memcpy(&qh, w[b].qh, 4);
        int s = 0;
        for (int j = 0; j < 16; j++) {
            const int xh0 = ((qh >> (j + 0)) << 4) & 0x10, xh1 = (qh >> (j + 12)) & 0x10;
            s += (((w[b].qs[j] & 15) | xh0) - 16) * a[b].qs[j] + (((w[b].qs[j] >> 4) | xh1) - 16) * a[b].qs[j + 16];
        }
        acc += (orc_fp16_to_fp32(w[b].d) * orc_fp16_to_fp32(a[b].d)) * (float)s;
    }
    return acc;
}
static float dot_q8_0(int64_t k, const blk_q8_0 *w, const blk_q8_0 *a) {   /* ggml-cpu/quants.c:305-333 */
    float acc = 0.0f;
    for (int64_t b = 0; b < k/QK; b++) {
        int s = 0;
        for (int j = 0; j < QK; j++) s += w[b].qs[j] * a[b].qs[j];
        acc += (float)s * (orc_fp16_to_fp32(w[b].d) * orc_fp16_to_fp32(a[b].d));
    }
    return acc;
}

/* K-quants (ggml-cpu/quants.c:550-758): the generic code keeps 8 float lanes (element index
 * mod 8), adds d * int32 lane sums per super-block, subtracts dmin * sum(bsums*mins) from a
 * scalar, then folds the lanes in order 0..7.  We keep that association. */
static float dot_kquant(int type, int64_t k, const void *vw, const blk_q8_K *a) {
    float lanes[8] = {0};
    float acc = 0.0f;
    int q[QKK];
    for (int64_t b = 0; b < k/QKK; b++) {
        int32_t li[8] = {0};
        float d, dmin = 0.0f; int minsum = 0;
        if (type == ORC_Q6_K) {
            const blk_q6_K *w = (const blk_q6_K *)vw + b; q6K_ints(w, q);
            for (int g = 0; g < 16; g++)
                for (int l = 0; l < 16; l++) li[l & 7] += (int32_t)w->sc[g] * (q[g*16 + l] * a[b].qs[g*16 + l]);
            d = orc_fp16_to_fp32(w->d) * a[b].d;
        } else {
            const uint8_t *sc;
            if (type == ORC_Q4_K) { const blk_q4_K *w = (const blk_q4_K *)vw + b; q4K_ints(w, q); sc = w->sc; d = orc_fp16_to_fp32(w->d); dmin = orc_fp16_to_fp32(w->dmin); }
            else                  { const blk_q5_K *w = (const blk_q5_K *)vw + b; q5K_ints(w, q); sc = w->sc; d = orc_fp16_to_fp32(w->d); dmin = orc_fp16_to_fp32(w->dmin); }
            for (int g = 0; g < 8; g++) {
                int s, m; k4_scale_min(g, sc, &s, &m);
                minsum += (a[b].bsums[2*g] + a[b].bsums[2*g + 1]) * m;
                for (int l = 0; l < 32; l++) li[l & 7] += s * (q[g*32 + l] * a[b].qs[g*32 + l]);
            }
            d    = d * a[b].d;
            dmin = dmin * a[b].d;
        }
        for (int l = 0; l < 8; l++) lanes[l] += d * (float)li[l];
        if (type != ORC_Q6_K) acc -= dmin * (float)minsum;
    }
    for (int l = 0; l < 8; l++) acc += lanes[l];
    return acc;
}

float orc_vec_dot(int type, int64_t k, const void *w, const void *a) {
    switch (type) {
        case ORC_Q4_0: return dot_q4_0(k, (const blk_q4_0 *)w, (const blk_q8_0 *)a);
        case ORC_Q5_0: return dot_q5_0(k, (const blk_q5_0 *)w, (const blk_q8_0 *)a);
        case ORC_Q8_0: return dot_q8_0(k, (const blk_q8_0 *)w, (const blk_q8_0 *)a);
        case ORC_Q4_K: case ORC_Q5_K: case ORC_Q6_K: return dot_kquant(type, k, w, (const blk_q8_K *)a);
        default: return orc_vec_dot_ext(type, k, w, a);            /* oracle_ext.c; NAN for unknown types */
    }
}

/* ggml-cpu/ggml-cpu.c:1202-1394: quantise each src1 row to the weight's vec_dot_type, then one
 * vec_dot per (row of W, row of X). */
void orc_mul_mat(int type, const void *W, const float *X, float *dst, int64_t m, int64_t n, int64_t k) {
    const int at = orc_act_type(type);                             /* vec_dot_type: q8_0, q8_1 or q8_K */
    const int64_t abytes = orc_row_bytes(at, k);
    const int64_t wbytes = orc_row_bytes(type, k);
    uint8_t *aq = (uint8_t *)malloc((size_t)abytes);
    for (int64_t j = 0; j < n; j++) {
        if (at == ORC_Q8_0) orc_quantize_row_q8_0(X + j*k, aq, k); else if (at == ORC_Q8_1) orc_quantize_row_q8_1(X + j*k, aq, k); else orc_quantize_row_q8_K(X + j*k, aq, k);
        for (int64_t i = 0; i < m; i++)
            dst[j*m + i] = orc_vec_dot(type, k, (const uint8_t *)W + i*wbytes, aq);
    }
    free(aq);
}

/* ---- RMS_NORM (+MUL) — ggml-cpu/ops.cpp:4164-4183: double accumulation of float squares ---- */
void orc_rms_norm(const float *x, const float *w, float *y, int64_t ncols, int64_t nrows, float eps) {
    for (int64_t r = 0; r < nrows; r++) {
        const float *xr = x + r*ncols; float *yr = y + r*ncols;
        double sum = 0.0;
        for (int64_t i = 0; i < ncols; i++) sum += (double)(xr[i] * xr[i]);
        const float mean  = (float)(sum / (double)ncols);
        const float scale = 1.0f / sqrtf(mean + eps);
        for (int64_t i = 0; i < ncols; i++) {
            float v = xr[i] * scale;
            yr[i] = w ? v * w[i] : v;
        }
    }
}

/* ---- ROPE — ggml-cpu/ops.cpp:6049-6100 (yarn, cache init), 6150-6330 (apply); ggml.c:4082-4095 ---- */
static float yarn_corr_dim(int n_dims, int n_ctx_orig, float n_rot, float base) {
    return (float)n_dims * logf((float)n_ctx_orig / (n_rot * 2.0f * (float)M_PI)) / (2.0f * logf(base));
}
void orc_rope(const float *x, float *y, const int32_t *pos, const float *ff,
              int64_t hd, int64_t n_head, int64_t n_tok,
              int n_dims, int mode, int n_ctx_orig, float freq_base, float freq_scale,
              float ext_factor, float attn_factor, float beta_fast, float beta_slow) {
    const float theta_scale = powf(freq_base, -2.0f / (float)n_dims);
    float lo = floorf(yarn_corr_dim(n_dims, n_ctx_orig, beta_fast, freq_base));
    float hi = ceilf (yarn_corr_dim(n_dims, n_ctx_orig, beta_slow, freq_base));
    if (lo < 0.0f) lo = 0.0f;
    if (hi > (float)(n_dims - 1)) hi = (float)(n_dims - 1);
    const int neox = (mode & 2) != 0;
    float *cs = (float *)malloc(sizeof(float) * (size_t)hd);
    for (int64_t t = 0; t < n_tok; t++) {
        float theta = (float)pos[t];                       /* theta_base, then *= theta_scale per pair */
        for (int64_t i0 = 0; i0 < hd; i0 += 2) {
            const float f = ff ? ff[i0/2] : 1.0f;
            const float extrap = theta / f;
            const float interp = freq_scale * extrap;
            float th = interp, ms = attn_factor;
            if (ext_factor != 0.0f) {
                float r = ((float)(i0/2) - lo) / fmaxf(0.001f, hi - lo);
                float ramp = (1.0f - fminf(1.0f, fmaxf(0.0f, r))) * ext_factor;
                th = interp * (1.0f - ramp) + extrap * ramp;
                ms *= 1.0f + 0.1f * logf(1.0f / freq_scale);
            }
            cs[i0] = cosf(th) * ms; cs[i0 + 1] = sinf(th) * ms;
            theta *= theta_scale;
        }
        for (int64_t h = 0; h < n_head; h++) {
            const float *s = x + (t*n_head + h)*hd; float *d = y + (t*n_head + h)*hd;
            for (int64_t i0 = 0; i0 < n_dims; i0 += 2) {
                const float c = cs[i0], sn = cs[i0 + 1];
                const int64_t a = neox ? i0/2 : i0, b = neox ? i0/2 + n_dims/2 : i0 + 1;
                const float x0 = s[a], x1 = s[b];
                d[a] = x0*c - x1*sn;
                d[b] = x0*sn + x1*c;
            }
            for (int64_t i0 = n_dims; i0 < hd; i0++) d[i0] = s[i0];
        }
    }
    free(cs);
}

/* ---- SET_ROWS — ggml-cpu/ops.cpp:5359-5415, from_float of the destination type ---- */
void orc_set_rows(const float *src, const int64_t *ids, void *dst, int dst_type,
                  int64_t ncols, int64_t nrows, int64_t stride) {
    for (int64_t r = 0; r < nrows; r++) {
        uint8_t *drow = (uint8_t *)dst + ids[r]*stride;
        if (dst_type == ORC_F16)       orc_cpy_f32_f16(src + r*ncols, (uint16_t *)drow, ncols);
        else if (dst_type == ORC_Q8_0) orc_quantize_row_q8_0(src + r*ncols, drow, ncols);
        else if (dst_type == ORC_Q4_0) orc_quantize_row_q4_0(src + r*ncols, drow, ncols);   /* oracle_ext.c */
        else if (dst_type == ORC_F32)  memcpy(drow, src + r*ncols, (size_t)ncols*4);
    }
}

/* ---- FLASH_ATTN_EXT — ggml-cpu/ops.cpp:8169-8405: one-pass online softmax per (head, token) ----
 * Q row is converted to K's vec_dot_type (f16 for F16 K, q8_0 for Q8_0 K); s = dot*scale (+softcap)
 * + slope*mask; F16 V accumulates in an fp16 vector (ggml_vec_mad_f16 / ggml_vec_scale_f16 round
 * every element to half each step), quantised V is expanded to f32 and accumulated in f32. */
void orc_flash_attn_ext(const void *q, int64_t q_nb1, int64_t q_nb2,
                        const void *k, int64_t k_nb1, int64_t k_nb2,
                        const void *v, int64_t v_nb1, int64_t v_nb2,
                        const uint16_t *mask, float *dst,
                        int kv_type, int64_t dk, int64_t dv, int64_t n_head, int64_t n_head_kv,
                        int64_t n_tok, int64_t n_kv, float scale, float max_bias, float softcap) {
    if (softcap != 0.0f) scale /= softcap;
    const uint32_t nh_log2 = 1u << (uint32_t)floor(log2((double)n_head));
    const float m0 = powf(2.0f, -(max_bias)        / (float)nh_log2);
    const float m1 = powf(2.0f, -(max_bias / 2.0f) / (float)nh_log2);
    const int64_t gq = n_head / n_head_kv;
    float    *acc32 = (float *)malloc(sizeof(float) * (size_t)dv);
    float    *v32   = (float *)malloc(sizeof(float) * (size_t)dv);
    uint16_t *acc16 = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)dv);
    uint8_t  *qq    = (uint8_t *)malloc((size_t)dk * 4);
    for (int64_t h = 0; h < n_head; h++)
    for (int64_t t = 0; t < n_tok; t++) {
        const float slope = max_bias > 0.0f
            ? ((uint32_t)h < nh_log2 ? powf(m0, (float)(h + 1)) : powf(m1, (float)(2*(h - nh_log2) + 1))) : 1.0f;
        const float *qrow = (const float *)((const uint8_t *)q + t*q_nb1 + h*q_nb2);
        if (kv_type == ORC_F16) orc_cpy_f32_f16(qrow, (uint16_t *)qq, dk);
        else                    orc_quantize_row_q8_0(qrow, qq, dk);
        float S = 0.0f, M = -INFINITY;
        if (kv_type == ORC_F16) memset(acc16, 0, (size_t)dv*2); else memset(acc32, 0, (size_t)dv*4);
        const uint16_t *mrow = mask ? mask + t*n_kv : NULL;
        const int64_t hk = h / gq;
        for (int64_t c = 0; c < n_kv; c++) {
            const float mv = mrow ? slope * orc_fp16_to_fp32(mrow[c]) : 0.0f;
            if (mv == -INFINITY) continue;
            const uint8_t *krow = (const uint8_t *)k + c*k_nb1 + hk*k_nb2;
            const uint8_t *vrow = (const uint8_t *)v + c*v_nb1 + hk*v_nb2;
            float s;
            if (kv_type == ORC_F16) {
                /* ggml_vec_dot_f16: products of half values accumulated in f32 (SIMD order) */
                float a = 0.0f;
                for (int64_t i = 0; i < dk; i++)
                    a += orc_fp16_to_fp32(((const uint16_t *)krow)[i]) * orc_fp16_to_fp32(((const uint16_t *)qq)[i]);
                s = a;
            } else {
                s = orc_vec_dot(kv_type, dk, krow, qq);        /* Q8_0 / Q4_0 K against the q8_0 form of the Q row */
            }
            s *= scale;
            if (softcap != 0.0f) s = softcap * tanhf(s);
            s += mv;
            const float Mold = M;
            float ms = 1.0f, vs = 1.0f;
            if (s > M) { M = s; ms = expf(Mold - M); } else { vs = expf(s - M); }
            if (kv_type == ORC_F16) {
                if (s > Mold)
                    for (int64_t i = 0; i < dv; i++) acc16[i] = orc_fp32_to_fp16(orc_fp16_to_fp32(acc16[i]) * ms);
                for (int64_t i = 0; i < dv; i++)
                    acc16[i] = orc_fp32_to_fp16(orc_fp16_to_fp32(acc16[i]) + orc_fp16_to_fp32(((const uint16_t *)vrow)[i]) * vs);
            } else {
                if (s > Mold) for (int64_t i = 0; i < dv; i++) acc32[i] *= ms;
                orc_dequantize_row(kv_type, vrow, v32, dv);
                for (int64_t i = 0; i < dv; i++) acc32[i] += v32[i] * vs;
            }
            S = S*ms + vs;
        }
        if (kv_type == ORC_F16) for (int64_t i = 0; i < dv; i++) acc32[i] = orc_fp16_to_fp32(acc16[i]);
        const float inv = 1.0f / S;
        float *o = dst + (t*n_head + h)*dv;
        for (int64_t i = 0; i < dv; i++) o[i] = acc32[i] * inv;
    }
    free(acc32); free(v32); free(acc16); free(qq);
}

/* ---- glue ---- */
/* exp() as the x86 CPU backend evaluates it inside silu: ggml_v_expf (ggml-cpu/vec.h:785-810 AVX512,
 * :828-860 AVX2 — same constants, FMA at every step), restated with scalar fmaf so that it is bit-identical
 * for the finite range silu needs (|n| <= 126). */
static float v_expf(float x) {
    const float r = 0x1.8p23f;
    const float z = fmaf(x, 0x1.715476p+0f, r);
    const float n = z - r;
    const float b = fmaf(-n, 0x1.7f7d1cp-20f, fmaf(-n, 0x1.62e4p-1f, x));
    const float u = b * b;
    const float j = fmaf(fmaf(fmaf(0x1.0e4020p-7f, b, 0x1.573e2ep-5f), u, fmaf(0x1.555e66p-3f, b, 0x1.fffdb6p-2f)), u, fmaf(0x1.ffffecp-1f, b, 1.0f));
    if (fabsf(n) > 192.0f) return n <= 0.0f ? 0.0f : INFINITY;
    return ldexpf(j, (int)n);
}
/* ggml_vec_swiglu_f32 (ggml-cpu/vec.cpp:260-282): silu(x) = x / (1 + v_expf(-x)), then * up; rows on this path are
 * multiples of 16 so the vector body covers every element */
void orc_swiglu(const float *gate, const float *up, float *y, int64_t n) {
    for (int64_t i = 0; i < n; i++) y[i] = (gate[i] / (1.0f + v_expf(0.0f - gate[i]))) * up[i];
}
void orc_add(const float *a, const float *b, float *y, int64_t ncols, int64_t nrows, int64_t b_rows) {
    for (int64_t r = 0; r < nrows; r++)
        for (int64_t i = 0; i < ncols; i++) y[r*ncols + i] = a[r*ncols + i] + b[(r % b_rows)*ncols + i];
}
void orc_mul(const float *a, const float *b, float *y, int64_t ncols, int64_t nrows, int64_t b_rows) {
    for (int64_t r = 0; r < nrows; r++)
        for (int64_t i = 0; i < ncols; i++) y[r*ncols + i] = a[r*ncols + i] * b[(r % b_rows)*ncols + i];
}
void orc_get_rows_f32(const float *src, const int32_t *ids, float *dst, int64_t ncols, int64_t n_ids) {
    for (int64_t r = 0; r < n_ids; r++) memcpy(dst + r*ncols, src + (int64_t)ids[r]*ncols, (size_t)ncols*4);
}
void orc_cpy_f32_f16(const float *src, uint16_t *dst, int64_t n) {
    for (int64_t i = 0; i < n; i++) dst[i] = orc_fp32_to_fp16(src[i]);
}

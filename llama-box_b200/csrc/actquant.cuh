// actquant.cuh — warp-level activation quantisers shared by quantize.cu and the matvec prologue (mmvq.cu).
#pragma once
#include "common.cuh"

struct ActOut { int8_t * qs; float * d; int16_t * bs; };

__device__ __forceinline__ ActOut act_sections(void * act, int kind, int64_t k, int64_t col) {
    uint8_t * base = (uint8_t *)act + col * act_col_bytes(kind, k);
    ActOut o;
    o.qs = (int8_t *)base;
    o.d  = (float *)(base + act_d_off(kind, k));
    o.bs = (int16_t *)(base + act_bsum_off(kind, k));
    return o;
}

// quantise 8 values per lane (a warp covers elements [256*blk, 256*blk+256)) as q8_K
__device__ __forceinline__ void warp_quant_q8K(const float (&v)[8], ActOut o, int64_t blk, int lane) {
    float am = 0.0f; int ai = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) { float a = fabsf(v[j]); if (a > am) { am = a; ai = lane * 8 + j; } }
    // first index of the largest |x| (the reference scans sequentially with a strict '>')
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) {
        float am2 = __shfl_xor_sync(0xffffffffu, am, o2);
        int   ai2 = __shfl_xor_sync(0xffffffffu, ai, o2);
        if (am2 > am || (am2 == am && ai2 < ai)) { am = am2; ai = ai2; }
    }
    float mine = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; j++) if ((ai & 7) == j) mine = v[j];
    const float maxv = __shfl_sync(0xffffffffu, mine, ai >> 3);

    int q[8]; int s = 0;
    float d = 0.0f;
    if (am != 0.0f) {
        const float iscale = __fdiv_rn(-127.0f, maxv);
#pragma unroll
        for (int j = 0; j < 8; j++) { int t = __float2int_rn(__fmul_rn(iscale, v[j])); q[j] = t > 127 ? 127 : t; s += q[j]; }
        d = __fdiv_rn(1.0f, iscale);
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++) q[j] = 0;
    }
    uint2 pk;
    pk.x = (q[0] & 0xff) | ((q[1] & 0xff) << 8) | ((q[2] & 0xff) << 16) | ((q[3] & 0xff) << 24);
    pk.y = (q[4] & 0xff) | ((q[5] & 0xff) << 8) | ((q[6] & 0xff) << 16) | ((q[7] & 0xff) << 24);
    *(uint2 *)(o.qs + blk * 256 + lane * 8) = pk;
    s += __shfl_xor_sync(0xffffffffu, s, 1);                     // 16-element group sums
    if ((lane & 1) == 0) o.bs[blk * 16 + (lane >> 1)] = (int16_t)s;
    if (lane == 0) o.d[blk] = d;
}

// same 256 elements as 8 q8_0 blocks (4 lanes per block)
__device__ __forceinline__ void warp_quant_q80(const float (&v)[8], ActOut o, int64_t blk256, int lane) {
    float am = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; j++) am = fmaxf(am, fabsf(v[j]));
    am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, 1));
    am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, 2));
    const float d  = __fdiv_rn(am, 127.0f);
    const float id = am != 0.0f ? __fdiv_rn(127.0f, am) : 0.0f;
    int q[8]; int s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) { q[j] = __float2int_rn(__fmul_rn(v[j], id)); s += q[j]; }
    uint2 pk;
    pk.x = (q[0] & 0xff) | ((q[1] & 0xff) << 8) | ((q[2] & 0xff) << 16) | ((q[3] & 0xff) << 24);
    pk.y = (q[4] & 0xff) | ((q[5] & 0xff) << 8) | ((q[6] & 0xff) << 16) | ((q[7] & 0xff) << 24);
    *(uint2 *)(o.qs + blk256 * 256 + lane * 8) = pk;
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if ((lane & 3) == 0) {
        const int64_t b = blk256 * 8 + (lane >> 2);
        o.d[b]  = __half2float(__float2half_rn(d));             // the oracle stores d as f16
        o.bs[b] = (int16_t)s;
    }
}


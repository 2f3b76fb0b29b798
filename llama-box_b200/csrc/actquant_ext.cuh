// actquant_ext.cuh — the q8_0 / q8_1 warp quantiser of the wide kernels (mmvq_ext.cu prologue, fattn_ext.cu query row).
#pragma once
#include "common.cuh"

// 256 elements (8 per lane) as 8 blocks of q8_0 / q8_1 in the x86 CPU backend's arithmetic (ggml-cpu/arch/x86/quants.c:290-492): the same
// int8 values for both; d = f16(max/127); s = f16((max/127) * sum) (q8_1); bs = sum (the -8 / -16 offsets of Q4_0 / Q5_0 use it)
__device__ __forceinline__ void warp_quant_q8_01(const float (&v)[8], int8_t * qs, float * ad, float * as, int16_t * bs, int64_t blk256, int lane) {
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
    *(uint2 *)(qs + blk256 * 256 + lane * 8) = pk;
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if ((lane & 3) == 0) {
        const int64_t b = blk256 * 8 + (lane >> 2);
        ad[b] = __half2float(__float2half_rn(d));
        as[b] = __half2float(__float2half_rn(__fmul_rn(d, (float)s)));
        bs[b] = (int16_t)s;
    }
}


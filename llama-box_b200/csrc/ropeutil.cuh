// ropeutil.cuh — device helpers shared by rope.cu and the persistent decode kernel (decode_mk.cu).
#pragma once
#include "common.cuh"
#include "mk.h"
#include <math.h>

typedef MkRope RopeDev;
RopeDev rope_host_params(const b200_rope_params * p);     // rope.cu (host)

#ifdef __CUDACC__
// cs[2i] = cos, cs[2i+1] = sin for pair i of this token
__device__ __forceinline__ void rope_table(float * cs, int32_t pos, const float * ff, const RopeDev & rp, int tid, int nthreads) {
    for (int i = tid; i < rp.n_dims / 2; i += nthreads) {
        float theta = (float)pos;
        for (int j = 0; j < i; j++) theta = __fmul_rn(theta, rp.theta_scale);
        const float extrap = ff ? __fdiv_rn(theta, ff[i]) : theta;
        const float interp = __fmul_rn(rp.freq_scale, extrap);
        float th = interp;
        if (rp.ext_factor != 0.0f) {
            const float y = __fdiv_rn((float)i - rp.corr_lo, fmaxf(0.001f, rp.corr_hi - rp.corr_lo));
            const float ramp = __fmul_rn(1.0f - fminf(1.0f, fmaxf(0.0f, y)), rp.ext_factor);
            th = __fadd_rn(__fmul_rn(interp, 1.0f - ramp), __fmul_rn(extrap, ramp));
        }
        cs[2 * i]     = __fmul_rn(cosf(th), rp.mscale);
        cs[2 * i + 1] = __fmul_rn(sinf(th), rp.mscale);
    }
}

__device__ __forceinline__ void rope_pair(const float * s, float * d, int i, const float * cs, const RopeDev & rp) {
    const int a = rp.neox ? i : 2 * i, b = rp.neox ? i + rp.n_dims / 2 : 2 * i + 1;
    const float c = cs[2 * i], sn = cs[2 * i + 1];
    const float x0 = s[a], x1 = s[b];
    d[a] = __fsub_rn(__fmul_rn(x0, c), __fmul_rn(x1, sn));
    d[b] = __fadd_rn(__fmul_rn(x0, sn), __fmul_rn(x1, c));
}

// ---- row converters -----------------------------------------------------------------------------
// 8 consecutive floats per lane -> destination row of `type` at element offset e (multiple of 8)
__device__ __forceinline__ void store8(void * drow, int type, int64_t e, const float (&v)[8], int lane, bool active = true) {
    if (type == B200_TYPE_F32) {
        if (!active) return;
        *(float4 *)((float *)drow + e)     = make_float4(v[0], v[1], v[2], v[3]);
        *(float4 *)((float *)drow + e + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else if (type == B200_TYPE_F16) {
        if (!active) return;
        uint4 pk;
        pk.x = f2h_rn(v[0]) | ((uint32_t)f2h_rn(v[1]) << 16); pk.y = f2h_rn(v[2]) | ((uint32_t)f2h_rn(v[3]) << 16);
        pk.z = f2h_rn(v[4]) | ((uint32_t)f2h_rn(v[5]) << 16); pk.w = f2h_rn(v[6]) | ((uint32_t)f2h_rn(v[7]) << 16);
        *(uint4 *)((uint16_t *)drow + e) = pk;
    } else { // Q8_0, native 34-byte blocks; 4 lanes per block (all 32 lanes of the warp must call)
        float am = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; j++) am = fmaxf(am, fabsf(v[j]));
        am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, 1));
        am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, 2));
        const float d  = __fdiv_rn(am, 127.0f);
        const float id = am != 0.0f ? __fdiv_rn(127.0f, am) : 0.0f;
        int q[8];
#pragma unroll
        for (int j = 0; j < 8; j++) q[j] = __float2int_rn(__fmul_rn(v[j], id));
        if (!active) return;
        uint8_t * blk = (uint8_t *)drow + (e / 32) * 34;
        uint16_t * o = (uint16_t *)(blk + 2 + (e % 32));
        o[0] = (uint16_t)((q[0] & 0xff) | ((q[1] & 0xff) << 8)); o[1] = (uint16_t)((q[2] & 0xff) | ((q[3] & 0xff) << 8));
        o[2] = (uint16_t)((q[4] & 0xff) | ((q[5] & 0xff) << 8)); o[3] = (uint16_t)((q[6] & 0xff) | ((q[7] & 0xff) << 8));
        if ((lane & 3) == 0) *(uint16_t *)blk = f2h_rn(d);
    }
}

#endif

// fattn_ext_kernels.cuh — device code of the q4_0 KV-cache kernels (launch code: fattn_ext.cu); a header for the same reason as
// mmvq_ext_kernels.cuh: tests/hostsim/kernsim.cpp runs this source on the CPU under a SIMT emulation.
#pragma once
#include "common.cuh"

#include "actquant_ext.cuh"
#include "extfmt.cuh"

namespace {

struct FaWideArgs {
    const float * q; int64_t q_ts, q_hs;                       // floats
    const uint8_t * k; int64_t k_rs, k_hs; const uint8_t * v; int64_t v_rs, v_hs;   // bytes
    const uint16_t * mask; int64_t mask_rs;                    // halves
    float * dst; int64_t n_head, n_head_kv, n_kv;
    float scale, max_bias, softcap, m0, m1; uint32_t nh_log2;
};

template <int D>
__global__ void __launch_bounds__(128) fattn_q4_0_kernel(const FaWideArgs a) {
    constexpr int EPL = D / 32;                                // output elements per lane
    __shared__ __align__(16) int8_t s_qs[256];
    __shared__ float s_ad[8], s_as[8]; __shared__ int16_t s_bs[8];
    __shared__ float s_m[4], s_l[4], s_acc[4][D];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t h = blockIdx.x, t = blockIdx.y, hk = h / (a.n_head / a.n_head_kv);
    pdl_wait();
    if (warp == 0) {                                           // the query row as q8_0 (elements past D: zero)
        float qv[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
        if (lane * 8 < D) {
            const float4 * pq = (const float4 *)(a.q + t * a.q_ts + h * a.q_hs + lane * 8);
            const float4 f0 = pq[0], f1 = pq[1];
            qv[0] = f0.x; qv[1] = f0.y; qv[2] = f0.z; qv[3] = f0.w; qv[4] = f1.x; qv[5] = f1.y; qv[6] = f1.z; qv[7] = f1.w;
        }
        warp_quant_q8_01(qv, s_qs, s_ad, s_as, s_bs, 0, lane);
    }
    __syncthreads();

    float slope = 1.0f;
    if (a.max_bias > 0.0f) slope = (uint32_t)h < a.nh_log2 ? powf(a.m0, (float)(h + 1)) : powf(a.m1, (float)(2 * (h - a.nh_log2) + 1));
    float M = -INFINITY, S = 0.0f, acc[EPL];
#pragma unroll
    for (int i = 0; i < EPL; i++) acc[i] = 0.0f;
    const uint16_t * mrow = a.mask ? a.mask + t * a.mask_rs : nullptr;
    for (int64_t c = warp; c < a.n_kv; c += 4) {
        float mv = 0.0f;
        if (mrow) { mv = slope * xf_h2f(mrow[c]); if (mv == -INFINITY) continue; }        // uniform over the warp
        const uint8_t * kb = a.k + c * a.k_rs + hk * a.k_hs;
        float part = 0.0f;
        if (lane < EPL) part = xf_q4_0n_dot(kb + lane * 18, s_qs + lane * 32, s_ad[lane], (int)s_bs[lane]);
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        float s = __shfl_sync(0xffffffffu, part, 0) * a.scale;
        if (a.softcap != 0.0f) s = a.softcap * tanhf(s);
        s += mv;
        const float Mold = M;
        float ms = 1.0f, vs = 1.0f;
        if (s > M) { M = s; ms = expf(Mold - M); } else vs = expf(s - M);
        const uint8_t * vb = a.v + c * a.v_rs + hk * a.v_hs;
#pragma unroll
        for (int i = 0; i < EPL; i++) {
            const int e = lane * EPL + i;
            acc[i] = acc[i] * ms + xf_q4_0n_value(vb + (e >> 5) * 18, e & 31) * vs;
        }
        S = S * ms + vs;
    }
    if (lane == 0) { s_m[warp] = M; s_l[warp] = S; }
#pragma unroll
    for (int i = 0; i < EPL; i++) s_acc[warp][lane * EPL + i] = acc[i];
    __syncthreads();
    for (int e = threadIdx.x; e < D; e += 128) {
        float Mx = fmaxf(fmaxf(s_m[0], s_m[1]), fmaxf(s_m[2], s_m[3]));
        float num = 0.0f, den = 0.0f;
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const float f = s_m[w] == -INFINITY ? 0.0f : expf(s_m[w] - Mx);
            num += s_acc[w][e] * f; den += s_l[w] * f;
        }
        a.dst[(t * a.n_head + h) * D + e] = num / den;
    }
}

// Any head size that is a multiple of 32 up to 256 (Gemma 256, Phi 96, StableLM 80 ... — the tuned kernels of fattn.cu / fattn_tc.cu carry 64 and 128) over an
// F16, Q8_0 or Q4_0 cache in ggml's native layout.  Same structure as above with a run-time D: lane l owns output elements l, l + 32, ...; the K dot is
// per block for the quantised caches (q8_0 query, integer sums) and element-strided for F16 (query rounded to f16 like the oracle's vec_dot_type, f32 sums).
// KVT: 1 = F16, 8 = Q8_0, 2 = Q4_0.
template <int KVT>
__global__ void __launch_bounds__(128) fattn_any_kernel(const FaWideArgs a, const int D) {
    __shared__ __align__(16) int8_t s_qs[256];
    __shared__ float s_ad[8], s_as[8]; __shared__ int16_t s_bs[8];
    __shared__ float s_qf[256];
    __shared__ float s_m[4], s_l[4], s_acc[4][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nblk = D / 32;
    const int64_t h = blockIdx.x, t = blockIdx.y, hk = h / (a.n_head / a.n_head_kv);
    const float * qrow = a.q + t * a.q_ts + h * a.q_hs;
    pdl_wait();
    if (KVT == 1) {
        for (int e = threadIdx.x; e < D; e += 128) s_qf[e] = __half2float(__float2half_rn(qrow[e]));
    } else if (warp == 0) {
        float qv[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
        if (lane * 8 < D) {
            const float4 * pq = (const float4 *)(qrow + lane * 8);
            const float4 f0 = pq[0], f1 = pq[1];
            qv[0] = f0.x; qv[1] = f0.y; qv[2] = f0.z; qv[3] = f0.w; qv[4] = f1.x; qv[5] = f1.y; qv[6] = f1.z; qv[7] = f1.w;
        }
        warp_quant_q8_01(qv, s_qs, s_ad, s_as, s_bs, 0, lane);
    }
    __syncthreads();

    float slope = 1.0f;
    if (a.max_bias > 0.0f) slope = (uint32_t)h < a.nh_log2 ? powf(a.m0, (float)(h + 1)) : powf(a.m1, (float)(2 * (h - a.nh_log2) + 1));
    float M = -INFINITY, S = 0.0f, acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = 0.0f;
    const uint16_t * mrow = a.mask ? a.mask + t * a.mask_rs : nullptr;
    for (int64_t c = warp; c < a.n_kv; c += 4) {
        float mv = 0.0f;
        if (mrow) { mv = slope * xf_h2f(mrow[c]); if (mv == -INFINITY) continue; }        // uniform over the warp
        const uint8_t * kb = a.k + c * a.k_rs + hk * a.k_hs;
        float part = 0.0f;
        if (KVT == 1) {
            const __half * kr = (const __half *)kb;
            for (int i = lane; i < D; i += 32) part = fmaf(__half2float(kr[i]), s_qf[i], part);
        } else if (lane < nblk) {
            part = KVT == 8 ? xf_q8_0n_dot(kb + lane * 34, s_qs + lane * 32, s_ad[lane]) : xf_q4_0n_dot(kb + lane * 18, s_qs + lane * 32, s_ad[lane], (int)s_bs[lane]);
        }
        float s = warp_sum(part) * a.scale;
        if (a.softcap != 0.0f) s = a.softcap * tanhf(s);
        s += mv;
        const float Mold = M;
        float ms = 1.0f, vs = 1.0f;
        if (s > M) { M = s; ms = expf(Mold - M); } else vs = expf(s - M);
        const uint8_t * vb = a.v + c * a.v_rs + hk * a.v_hs;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (i < nblk) {
                const float v = KVT == 1 ? __half2float(((const __half *)vb)[lane + 32 * i]) : (KVT == 8 ? xf_q8_0n_value(vb + i * 34, lane) : xf_q4_0n_value(vb + i * 18, lane));
                acc[i] = acc[i] * ms + v * vs;
            }
        }
        S = S * ms + vs;
    }
    if (lane == 0) { s_m[warp] = M; s_l[warp] = S; }
#pragma unroll
    for (int i = 0; i < 8; i++) if (i < nblk) s_acc[warp][lane + 32 * i] = acc[i];
    __syncthreads();
    for (int e = threadIdx.x; e < D; e += 128) {
        const float Mx = fmaxf(fmaxf(s_m[0], s_m[1]), fmaxf(s_m[2], s_m[3]));
        float num = 0.0f, den = 0.0f;
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const float f = s_m[w] == -INFINITY ? 0.0f : expf(s_m[w] - Mx);
            num += s_acc[w][e] * f; den += s_l[w] * f;
        }
        a.dst[(t * a.n_head + h) * D + e] = num / den;
    }
}

// one thread per destination block
__global__ void __launch_bounds__(128) set_rows_q4_0_kernel(const float * __restrict__ src, int64_t src_rs, const int64_t * __restrict__ ids, uint8_t * __restrict__ dst, int64_t dst_rs, int64_t nblk) {
    pdl_wait();
    const int64_t b = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (b >= nblk) return;
    const int64_t r = blockIdx.y;
    float x[32];
    const float4 * p = (const float4 *)(src + r * src_rs + b * 32);
#pragma unroll
    for (int i = 0; i < 8; i++) { const float4 f = p[i]; x[4 * i] = f.x; x[4 * i + 1] = f.y; x[4 * i + 2] = f.z; x[4 * i + 3] = f.w; }
    uint8_t blk[18];
    xf_q4_0_quantize_block(x, blk);
    uint16_t * o = (uint16_t *)(dst + ids[r] * dst_rs + b * 18);
#pragma unroll
    for (int i = 0; i < 9; i++) o[i] = (uint16_t)blk[2 * i] | ((uint16_t)blk[2 * i + 1] << 8);
}

} // namespace

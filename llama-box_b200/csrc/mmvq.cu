// mmvq.cu — decode matvec on GGUF-quantised weights (sm_100a), ncols <= 8.
//
// Replaces ggml_cuda_mul_mat_vec_q / mul_mat_vec_q<type,ncols_dst> (ggml-cuda/mmvq.cu:139-226,
// 500-570) and vec_dot_q*_q8_1 (vecdotq.cuh:579-839).  Differences by design:
//   * activations are quantised like the CPU oracle (q8_K / RNE q8_0), so per-block integer sums
//     equal the oracle's bit for bit (ggml-cpu/quants.c:115-149,305-333,550-758);
//   * every weight load is a 16-byte streaming load (rows repacked at load time where the GGUF
//     block size is not a multiple of 16: see repack.cu) — the reference uses 2-/4-byte loads;
//   * the quantised activation vector is staged once per CTA into shared memory with one bulk
//     async copy (TMA engine, mbarrier completion) and read with conflict-free LDS.128;
//   * one warp owns TWO rows at a time so activation registers are reused and 4+ independent
//     16-byte HBM loads per lane are in flight; several matrices (QKV, gate+up) share a launch;
//   * epilogues fused: bias, residual add, SwiGLU (gate row x up row).
// HBM-bound: algorithmic bytes = m * row_bytes (weights) + act + dst.
#include "common.cuh"

#define MMV_WARPS 8
#define MMV_MAX_MATS 4

enum { MMV_MODE_PLAIN = 0, MMV_MODE_SWIGLU = 1 };

struct MmvMat {
    const uint8_t * W;
    float *         dst;
    const float *   bias;
    const float *   residual;
    int64_t         m;
    int64_t         dst_col_stride;
    int32_t         type;
    int32_t         pair0;      // first row-pair index of this matrix in the launch
};
struct MmvArgs {
    MmvMat  mat[MMV_MAX_MATS];
    const uint8_t * act[2];     // [kind]: 0 = q8_K act buffer, 1 = q8_0 act buffer (may be null)
    int64_t k;
    int32_t n_mats;
    int32_t total_pairs;
    int32_t mode;
    int32_t act_bytes[2];       // bytes to stage per kind (ncols * col_bytes), 0 if unused
};

// ---- per-type weight registers for one 32-element chunk ------------------------------------
template <int T> struct WReg;
template <> struct WReg<B200_TYPE_Q4_0> { uint4 qs; uint16_t d; };
template <> struct WReg<B200_TYPE_Q8_0> { uint4 q0, q1; uint16_t d; };
template <> struct WReg<B200_TYPE_Q4_K> { uint4 hdr, qs; };
template <> struct WReg<B200_TYPE_Q5_K> { uint4 hdr, qh, qs; };
template <> struct WReg<B200_TYPE_Q6_K> { uint4 ql, qh; uint32_t sc; uint16_t d; };

// chunk i of a row: 32 consecutive "lane units" (see per-type mapping below)
template <int T> __device__ __forceinline__ void wload(WReg<T> & w, const uint8_t * row, int64_t nb, int i);

template <> __device__ __forceinline__ void wload<B200_TYPE_Q4_0>(WReg<B200_TYPE_Q4_0> & w, const uint8_t * row, int64_t nb, int i) {
    w.qs = ldg_stream16(row + (int64_t)i * 16);
    w.d  = ldg_nc16(row + nb * 16 + (int64_t)i * 2);
}
template <> __device__ __forceinline__ void wload<B200_TYPE_Q8_0>(WReg<B200_TYPE_Q8_0> & w, const uint8_t * row, int64_t nb, int i) {
    w.q0 = ldg_stream16(row + (int64_t)i * 32);
    w.q1 = ldg_stream16(row + (int64_t)i * 32 + 16);
    w.d  = ldg_nc16(row + nb * 32 + (int64_t)i * 2);
}
template <> __device__ __forceinline__ void wload<B200_TYPE_Q4_K>(WReg<B200_TYPE_Q4_K> & w, const uint8_t * row, int64_t nb, int i) {
    const uint8_t * b = row + (int64_t)(i >> 3) * 144;
    w.hdr = __ldg((const uint4 *)b);                       // d, dmin, 12 scale bytes: shared by 8 lanes
    w.qs  = ldg_stream16(b + 16 + (i & 7) * 16);
}
template <> __device__ __forceinline__ void wload<B200_TYPE_Q5_K>(WReg<B200_TYPE_Q5_K> & w, const uint8_t * row, int64_t nb, int i) {
    const uint8_t * b = row + (int64_t)(i >> 3) * 176;
    w.hdr = __ldg((const uint4 *)b);
    w.qh  = __ldg((const uint4 *)(b + 16 + (i & 1) * 16));  // high bits of l = 16h .. 16h+15, shared by 4 lanes
    w.qs  = ldg_stream16(b + 48 + (i & 7) * 16);
}
template <> __device__ __forceinline__ void wload<B200_TYPE_Q6_K>(WReg<B200_TYPE_Q6_K> & w, const uint8_t * row, int64_t nb, int i) {
    const int sb = i >> 3, j = i & 7, hh = j >> 2, ii = j & 3;
    w.ql = ldg_stream16(row + (int64_t)sb * 128 + hh * 64 + ii * 16);
    w.qh = ldg_stream16(row + nb * 128 + (int64_t)sb * 64 + hh * 32 + (ii & 1) * 16);
    // scales 8hh+ii (low-nibble part) and 8hh+ii+4 (high-nibble part)
    const uint8_t * sc = row + nb * 192 + (int64_t)sb * 16 + hh * 8 + ii;
    w.sc = (uint32_t)__ldg(sc) | ((uint32_t)__ldg(sc + 4) << 8);
    w.d  = ldg_nc16(row + nb * 208 + (int64_t)sb * 2);
}

// ---- activation view in shared memory --------------------------------------------------------
struct ActView { const uint8_t * qs; const float * d; const int16_t * bs; };

__device__ __forceinline__ int dot16(const uint4 & a, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
    int s = dp4a_us(w0, (int)a.x, 0);
    s = dp4a_us(w1, (int)a.y, s);
    s = dp4a_us(w2, (int)a.z, s);
    return dp4a_us(w3, (int)a.w, s);
}
__device__ __forceinline__ int dot16s(const uint4 & a, const uint4 & w) {
    int s = dp4a_s((int)w.x, (int)a.x, 0);
    s = dp4a_s((int)w.y, (int)a.y, s);
    s = dp4a_s((int)w.z, (int)a.z, s);
    return dp4a_s((int)w.w, (int)a.w, s);
}

// 6-bit scale / min of sub-block `is` from the q4_K/q5_K header regs (ggml-quants.c:703-711)
__device__ __forceinline__ void k4_scale_min(const uint4 & hdr, int is, int & sc, int & mn) {
    const int e = (is & 3) * 8;
    const int sc_lo = (hdr.y >> e) & 63, mn_lo = (hdr.z >> e) & 63;
    const int sc_hi = ((hdr.w >> e) & 15) | (((hdr.y >> (e + 6)) & 3) << 4);
    const int mn_hi = ((hdr.w >> (e + 4)) & 15) | (((hdr.z >> (e + 6)) & 3) << 4);
    sc = is < 4 ? sc_lo : sc_hi;
    mn = is < 4 ? mn_lo : mn_hi;
}

// Each lane handles 32 elements as two 16-element parts; the order in which the two parts are
// read from shared memory is swapped on half of the lanes of every quarter-warp so that each
// LDS.128 touches 8 distinct 16-byte bank groups (no conflicts).
template <int T> __device__ __forceinline__ float wdot(const WReg<T> & w, const ActView & a, int i);

template <> __device__ __forceinline__ float wdot<B200_TYPE_Q4_0>(const WReg<B200_TYPE_Q4_0> & w, const ActView & a, int i) {
    const int sw = (i >> 2) & 1;                                  // part order swap
    const uint4 A = *(const uint4 *)(a.qs + i * 32 + sw * 16);
    const uint4 B = *(const uint4 *)(a.qs + i * 32 + (sw ^ 1) * 16);
    const int shA = sw * 4, shB = (sw ^ 1) * 4;                   // low nibbles <-> elements 0..15
    int s = dot16(A, (w.qs.x >> shA) & 0x0F0F0F0Fu, (w.qs.y >> shA) & 0x0F0F0F0Fu, (w.qs.z >> shA) & 0x0F0F0F0Fu, (w.qs.w >> shA) & 0x0F0F0F0Fu);
    s    += dot16(B, (w.qs.x >> shB) & 0x0F0F0F0Fu, (w.qs.y >> shB) & 0x0F0F0F0Fu, (w.qs.z >> shB) & 0x0F0F0F0Fu, (w.qs.w >> shB) & 0x0F0F0F0Fu);
    s -= 8 * (int)a.bs[i];
    return __fmul_rn(__fmul_rn((float)s, h2f(w.d)), a.d[i]);
}
template <> __device__ __forceinline__ float wdot<B200_TYPE_Q8_0>(const WReg<B200_TYPE_Q8_0> & w, const ActView & a, int i) {
    const int sw = (i >> 2) & 1;
    const uint4 A = *(const uint4 *)(a.qs + i * 32 + sw * 16);
    const uint4 B = *(const uint4 *)(a.qs + i * 32 + (sw ^ 1) * 16);
    const int s = sw ? dot16s(A, w.q1) + dot16s(B, w.q0) : dot16s(A, w.q0) + dot16s(B, w.q1);
    return __fmul_rn((float)s, __fmul_rn(h2f(w.d), a.d[i]));
}
template <> __device__ __forceinline__ float wdot<B200_TYPE_Q4_K>(const WReg<B200_TYPE_Q4_K> & w, const ActView & a, int i) {
    const int sb = i >> 3, j = i & 7, c = j >> 1, h = j & 1, sw = c >> 1;
    const int is0 = 2 * c + sw, is1 = 2 * c + (sw ^ 1);           // sub-block (32 elems) of each part
    const uint8_t * base = a.qs + sb * 256 + h * 16;
    const uint4 A = *(const uint4 *)(base + is0 * 32);
    const uint4 B = *(const uint4 *)(base + is1 * 32);
    const int shA = (is0 & 1) * 4, shB = (is1 & 1) * 4;
    const int sA = dot16(A, (w.qs.x >> shA) & 0x0F0F0F0Fu, (w.qs.y >> shA) & 0x0F0F0F0Fu, (w.qs.z >> shA) & 0x0F0F0F0Fu, (w.qs.w >> shA) & 0x0F0F0F0Fu);
    const int sB = dot16(B, (w.qs.x >> shB) & 0x0F0F0F0Fu, (w.qs.y >> shB) & 0x0F0F0F0Fu, (w.qs.z >> shB) & 0x0F0F0F0Fu, (w.qs.w >> shB) & 0x0F0F0F0Fu);
    int scA, mnA, scB, mnB;
    k4_scale_min(w.hdr, is0, scA, mnA);
    k4_scale_min(w.hdr, is1, scB, mnB);
    const int isum = scA * sA + scB * sB;
    const int imin = mnA * (int)a.bs[sb * 16 + is0 * 2 + h] + mnB * (int)a.bs[sb * 16 + is1 * 2 + h];
    const float da = a.d[sb];
    const float dw = h2f((uint16_t)(w.hdr.x & 0xffff)), dm = h2f((uint16_t)(w.hdr.x >> 16));
    return __fmul_rn(dw, da) * (float)isum - __fmul_rn(dm, da) * (float)imin;
}
template <> __device__ __forceinline__ float wdot<B200_TYPE_Q5_K>(const WReg<B200_TYPE_Q5_K> & w, const ActView & a, int i) {
    const int sb = i >> 3, j = i & 7, c = j >> 1, h = j & 1, sw = c >> 1;
    const int is0 = 2 * c + sw, is1 = 2 * c + (sw ^ 1);
    const uint8_t * base = a.qs + sb * 256 + h * 16;
    const uint4 A = *(const uint4 *)(base + is0 * 32);
    const uint4 B = *(const uint4 *)(base + is1 * 32);
    const int shA = (is0 & 1) * 4, shB = (is1 & 1) * 4;
#define Q5W(q, hb, sh, is) ((((q) >> (sh)) & 0x0F0F0F0Fu) | ((((hb) >> (is)) & 0x01010101u) << 4))
    const int sA = dot16(A, Q5W(w.qs.x, w.qh.x, shA, is0), Q5W(w.qs.y, w.qh.y, shA, is0), Q5W(w.qs.z, w.qh.z, shA, is0), Q5W(w.qs.w, w.qh.w, shA, is0));
    const int sB = dot16(B, Q5W(w.qs.x, w.qh.x, shB, is1), Q5W(w.qs.y, w.qh.y, shB, is1), Q5W(w.qs.z, w.qh.z, shB, is1), Q5W(w.qs.w, w.qh.w, shB, is1));
#undef Q5W
    int scA, mnA, scB, mnB;
    k4_scale_min(w.hdr, is0, scA, mnA);
    k4_scale_min(w.hdr, is1, scB, mnB);
    const int isum = scA * sA + scB * sB;
    const int imin = mnA * (int)a.bs[sb * 16 + is0 * 2 + h] + mnB * (int)a.bs[sb * 16 + is1 * 2 + h];
    const float da = a.d[sb];
    const float dw = h2f((uint16_t)(w.hdr.x & 0xffff)), dm = h2f((uint16_t)(w.hdr.x >> 16));
    return __fmul_rn(dw, da) * (float)isum - __fmul_rn(dm, da) * (float)imin;
}
template <> __device__ __forceinline__ float wdot<B200_TYPE_Q6_K>(const WReg<B200_TYPE_Q6_K> & w, const ActView & a, int i) {
    const int sb = i >> 3, j = i & 7, hh = j >> 2, ii = j & 3, sw = hh;
    // part p (0: low nibbles, 1: high nibbles) covers elements 128hh + 16ii + 64p .. +15,
    // 2-bit highs at qh bit 2*(ii/2) + 4p, scale 8hh + ii + 4p
    const int p0 = sw, p1 = sw ^ 1;
    const uint8_t * base = a.qs + sb * 256 + hh * 128 + ii * 16;
    const uint4 A = *(const uint4 *)(base + p0 * 64);
    const uint4 B = *(const uint4 *)(base + p1 * 64);
    const int hs0 = (ii >> 1) * 2 + p0 * 4, hs1 = (ii >> 1) * 2 + p1 * 4;
#define Q6W(q, hb, p, hs) ((((q) >> ((p) * 4)) & 0x0F0F0F0Fu) | ((((hb) >> (hs)) & 0x03030303u) << 4))
    int sA = dot16(A, Q6W(w.ql.x, w.qh.x, p0, hs0), Q6W(w.ql.y, w.qh.y, p0, hs0), Q6W(w.ql.z, w.qh.z, p0, hs0), Q6W(w.ql.w, w.qh.w, p0, hs0));
    int sB = dot16(B, Q6W(w.ql.x, w.qh.x, p1, hs1), Q6W(w.ql.y, w.qh.y, p1, hs1), Q6W(w.ql.z, w.qh.z, p1, hs1), Q6W(w.ql.w, w.qh.w, p1, hs1));
#undef Q6W
    const int g = sb * 16 + hh * 8 + ii;                          // 16-element group of part 0/1: g + 4p
    sA -= 32 * (int)a.bs[g + 4 * p0];
    sB -= 32 * (int)a.bs[g + 4 * p1];
    const int scA = (int)(int8_t)((w.sc >> (8 * p0)) & 0xff), scB = (int)(int8_t)((w.sc >> (8 * p1)) & 0xff);
    const int isum = scA * sA + scB * sB;
    return __fmul_rn(h2f(w.d), a.d[sb]) * (float)isum;
}

// ---- one warp, R=2 rows, all chunks of the rows -----------------------------------------------
template <int T, int NCOLS>
__device__ __forceinline__ void warp_rows(const uint8_t * row0, const uint8_t * row1, int64_t nb, int nchunks,
                                          const uint8_t * act_s, int64_t k, int kind, int lane, float (&acc)[2][NCOLS]) {
    const int64_t colb = act_col_bytes(kind, k);
    const int64_t doff = act_d_off(kind, k), boff = act_bsum_off(kind, k);
#pragma unroll
    for (int c = 0; c < NCOLS; c++) { acc[0][c] = 0.0f; acc[1][c] = 0.0f; }
    WReg<T> w0, w1, n0, n1;
    int i = lane;
    if (i < nchunks) { wload<T>(w0, row0, nb, i); wload<T>(w1, row1, nb, i); }
#pragma unroll 1
    for (; i < nchunks; i += 32) {
        const int inext = i + 32;
        if (inext < nchunks) { wload<T>(n0, row0, nb, inext); wload<T>(n1, row1, nb, inext); }   // prefetch next iteration
#pragma unroll
        for (int c = 0; c < NCOLS; c++) {
            ActView a;
            a.qs = act_s + c * colb;
            a.d  = (const float *)(act_s + c * colb + doff);
            a.bs = (const int16_t *)(act_s + c * colb + boff);
            acc[0][c] += wdot<T>(w0, a, i);
            acc[1][c] += wdot<T>(w1, a, i);
        }
        w0 = n0; w1 = n1;
    }
#pragma unroll
    for (int c = 0; c < NCOLS; c++) { acc[0][c] = warp_sum(acc[0][c]); acc[1][c] = warp_sum(acc[1][c]); }
}

template <int NCOLS, int TT>
__device__ __forceinline__ void rows_dispatch(int type, const uint8_t * row0, const uint8_t * row1, int64_t k,
                                              const uint8_t * act_s0, const uint8_t * act_s1, int lane, float (&acc)[2][NCOLS]) {
    const int nchunks = (int)(k / 32);
    if (TT >= 0) type = TT;                       // uniform-type launch: the switch folds away
    switch (type) {
        case B200_TYPE_Q4_K: warp_rows<B200_TYPE_Q4_K, NCOLS>(row0, row1, k / 256, nchunks, act_s0, k, 0, lane, acc); break;
        case B200_TYPE_Q5_K: warp_rows<B200_TYPE_Q5_K, NCOLS>(row0, row1, k / 256, nchunks, act_s0, k, 0, lane, acc); break;
        case B200_TYPE_Q6_K: warp_rows<B200_TYPE_Q6_K, NCOLS>(row0, row1, k / 256, nchunks, act_s0, k, 0, lane, acc); break;
        case B200_TYPE_Q4_0: warp_rows<B200_TYPE_Q4_0, NCOLS>(row0, row1, k / 32,  nchunks, act_s1, k, 1, lane, acc); break;
        default:             warp_rows<B200_TYPE_Q8_0, NCOLS>(row0, row1, k / 32,  nchunks, act_s1, k, 1, lane, acc); break;
    }
}

template <int NCOLS, int TT, int MODE>
__global__ void __launch_bounds__(MMV_WARPS * 32, (NCOLS <= 4 ? 2 : 1)) mmvq_kernel(const __grid_constant__ MmvArgs args) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    __syncthreads();
    // everything above overlaps the previous kernel's tail (PDL); activations are its output
    pdl_wait();
    uint8_t * act_s0 = smem;
    uint8_t * act_s1 = smem + args.act_bytes[0];
    if (tid == 0) {
        mbar_expect_tx(&bar, (uint32_t)(args.act_bytes[0] + args.act_bytes[1]));
        if (args.act_bytes[0]) bulk_g2s(act_s0, args.act[0], (uint32_t)args.act_bytes[0], &bar);
        if (args.act_bytes[1]) bulk_g2s(act_s1, args.act[1], (uint32_t)args.act_bytes[1], &bar);
    }
    mbar_wait(&bar, 0);

    const int64_t k = args.k;
    for (int p = blockIdx.x * MMV_WARPS + warp; p < args.total_pairs; p += gridDim.x * MMV_WARPS) {
        float acc[2][NCOLS];
        if (MODE == MMV_MODE_SWIGLU) {
            // pair p = (gate row p, up row p)
            const MmvMat & g = args.mat[0]; const MmvMat & u = args.mat[1];
            const int64_t rbg = type_block_bytes(g.type) * (k / type_block_elems(g.type));
            const int64_t rbu = type_block_bytes(u.type) * (k / type_block_elems(u.type));
            if (TT >= 0 || g.type == u.type) {
                rows_dispatch<NCOLS, TT>(g.type, g.W + p * rbg, u.W + p * rbu, k, act_s0, act_s1, lane, acc);
            } else {
                float t[2][NCOLS];
                rows_dispatch<NCOLS, TT>(g.type, g.W + p * rbg, g.W + p * rbg, k, act_s0, act_s1, lane, t);
#pragma unroll
                for (int c = 0; c < NCOLS; c++) acc[0][c] = t[0][c];
                rows_dispatch<NCOLS, TT>(u.type, u.W + p * rbu, u.W + p * rbu, k, act_s0, act_s1, lane, t);
#pragma unroll
                for (int c = 0; c < NCOLS; c++) acc[1][c] = t[0][c];
            }
            if (lane == 0) {
#pragma unroll
                for (int c = 0; c < NCOLS; c++) {
                    const float gv = acc[0][c];
                    g.dst[c * g.dst_col_stride + p] = (gv / (1.0f + expf(-gv))) * acc[1][c];   // ggml-cpu/vec.h:691
                }
            }
        } else {
            MmvMat M = args.mat[0];
#pragma unroll
            for (int q = 1; q < MMV_MAX_MATS; q++) if (q < args.n_mats && p >= args.mat[q].pair0) M = args.mat[q];
            const int64_t r0 = (int64_t)(p - M.pair0) * 2;
            const bool two = r0 + 1 < M.m;
            const int64_t rb = type_block_bytes(M.type) * (k / type_block_elems(M.type));
            const uint8_t * row0 = M.W + r0 * rb;
            rows_dispatch<NCOLS, TT>(M.type, row0, two ? row0 + rb : row0, k, act_s0, act_s1, lane, acc);
            if (lane < 2 && (lane == 0 || two)) {
                const int64_t r = r0 + lane;
#pragma unroll
                for (int c = 0; c < NCOLS; c++) {
                    float v = lane == 0 ? acc[0][c] : acc[1][c];
                    if (M.bias)     v += M.bias[r];
                    if (M.residual) v += M.residual[c * M.dst_col_stride + r];
                    M.dst[c * M.dst_col_stride + r] = v;
                }
            }
        }
    }
    pdl_trigger();
}

// ---- host ------------------------------------------------------------------------------------
static bool mmv_type_ok(int t) {
    return t == B200_TYPE_Q4_0 || t == B200_TYPE_Q8_0 || t == B200_TYPE_Q4_K || t == B200_TYPE_Q5_K || t == B200_TYPE_Q6_K;
}
// k for which every row start and every repacked section is 16-byte aligned
static bool mmv_k_ok(int t, int64_t k) {
    if (k <= 0 || k % 256 != 0) return false;
    if (t == B200_TYPE_Q6_K) return k % 2048 == 0;
    return true;
}

template <int NCOLS, int TT, int MODE> static int mmv_launch_ntm(const MmvArgs & a, size_t smem, int grid, cudaStream_t st) {
    static bool attr[64] = { false };
    int dev = 0; cudaGetDevice(&dev);
    if (!attr[dev & 63]) { cudaFuncSetAttribute(mmvq_kernel<NCOLS, TT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr[dev & 63] = true; }
    mmvq_kernel<NCOLS, TT, MODE><<<grid, MMV_WARPS * 32, smem, st>>>(a);
    B200_LAUNCH_CHECK();
    return B200_OK;
}
template <int NCOLS, int TT> static int mmv_launch_nt(const MmvArgs & a, size_t smem, int grid, cudaStream_t st) {
    return a.mode == MMV_MODE_SWIGLU ? mmv_launch_ntm<NCOLS, TT, MMV_MODE_SWIGLU>(a, smem, grid, st)
                                     : mmv_launch_ntm<NCOLS, TT, MMV_MODE_PLAIN>(a, smem, grid, st);
}
template <int NCOLS> static int mmv_launch_n(const MmvArgs & a, size_t smem, int grid, cudaStream_t st) {
    int tt = a.mat[0].type;
    for (int i = 1; i < a.n_mats; i++) if (a.mat[i].type != tt) tt = -1;
    switch (tt) {
        case B200_TYPE_Q4_0: return mmv_launch_nt<NCOLS, B200_TYPE_Q4_0>(a, smem, grid, st);
        case B200_TYPE_Q8_0: return mmv_launch_nt<NCOLS, B200_TYPE_Q8_0>(a, smem, grid, st);
        case B200_TYPE_Q4_K: return mmv_launch_nt<NCOLS, B200_TYPE_Q4_K>(a, smem, grid, st);
        case B200_TYPE_Q5_K: return mmv_launch_nt<NCOLS, B200_TYPE_Q5_K>(a, smem, grid, st);
        case B200_TYPE_Q6_K: return mmv_launch_nt<NCOLS, B200_TYPE_Q6_K>(a, smem, grid, st);
        default:             return mmv_launch_nt<NCOLS, -1>(a, smem, grid, st);
    }
}

static int mmv_launch(MmvArgs & a, int64_t ncols, cudaStream_t st) {
    if (ncols < 1 || ncols > 8) { b200_set_error("mmvq: ncols must be 1..8"); return B200_ERR_INVALID; }
    bool need[2] = { false, false };
    for (int i = 0; i < a.n_mats; i++) {
        const MmvMat & M = a.mat[i];
        if (!mmv_type_ok(M.type)) { b200_set_error("mmvq: unsupported weight type %d", M.type); return B200_ERR_UNSUPPORTED; }
        if (!mmv_k_ok(M.type, a.k)) { b200_set_error("mmvq: k=%lld not supported for type %d (needs k%%256==0, Q6_K k%%2048==0)", (long long)a.k, M.type); return B200_ERR_UNSUPPORTED; }
        if (((uintptr_t)M.W & 15) || !M.W || !M.dst || M.m <= 0) { b200_set_error("mmvq: weights must be 16-byte aligned / non-null"); return B200_ERR_INVALID; }
        need[b200_act_kind_for(M.type)] = true;
    }
    size_t smem = 0;
    for (int kd = 0; kd < 2; kd++) {
        a.act_bytes[kd] = 0;
        if (need[kd]) {
            if (!a.act[kd] || ((uintptr_t)a.act[kd] & 15)) { b200_set_error("mmvq: missing/unaligned act buffer of kind %d", kd); return B200_ERR_INVALID; }
            a.act_bytes[kd] = (int32_t)(ncols * act_col_bytes(kd, a.k));
            smem += a.act_bytes[kd];
        } else a.act[kd] = nullptr;
    }
    if (smem > 200 * 1024) { b200_set_error("mmvq: activations (%zu bytes) exceed shared memory", smem); return B200_ERR_UNSUPPORTED; }
    const int sms = b200_sm_count();
    int ctas_per_sm = (int)((200 * 1024) / (smem + 1024));
    if (ctas_per_sm > 8) ctas_per_sm = 8;            // 8 CTAs x 256 threads = 2048 threads / SM
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    int grid = (a.total_pairs + MMV_WARPS - 1) / MMV_WARPS;
    if (grid > sms * ctas_per_sm) grid = sms * ctas_per_sm;
    if (grid < 1) grid = 1;
    switch (ncols) {
        case 1: return mmv_launch_n<1>(a, smem, grid, st);
        case 2: return mmv_launch_n<2>(a, smem, grid, st);
        case 3: return mmv_launch_n<3>(a, smem, grid, st);
        case 4: return mmv_launch_n<4>(a, smem, grid, st);
        case 5: return mmv_launch_n<5>(a, smem, grid, st);
        case 6: return mmv_launch_n<6>(a, smem, grid, st);
        case 7: return mmv_launch_n<7>(a, smem, grid, st);
        default: return mmv_launch_n<8>(a, smem, grid, st);
    }
}

extern "C" int b200_mul_mat_vec_q(int type, const void * W, const void * act, float * dst, int64_t dst_col_stride,
                                  const float * bias, const float * residual, int64_t m, int64_t k, int64_t ncols, void * stream) {
    MmvArgs a = {};
    a.mat[0] = { (const uint8_t *)W, dst, bias, residual, m, dst_col_stride, type, 0 };
    a.n_mats = 1; a.k = k; a.mode = MMV_MODE_PLAIN;
    a.total_pairs = (int32_t)((m + 1) / 2);
    const int kind = b200_act_kind_for(type);
    if (kind < 0) { b200_set_error("mmvq: unsupported weight type %d", type); return B200_ERR_UNSUPPORTED; }
    a.act[kind] = (const uint8_t *)act;
    return mmv_launch(a, ncols, (cudaStream_t)stream);
}

extern "C" int b200_mul_mat_vec_q_multi(const b200_mmv_desc * descs, int n_mats, const void * act_q8K, const void * act_q80,
                                        int64_t k, int64_t ncols, void * stream) {
    if (n_mats < 1 || n_mats > MMV_MAX_MATS || !descs) { b200_set_error("mmvq_multi: n_mats must be 1..%d", MMV_MAX_MATS); return B200_ERR_INVALID; }
    MmvArgs a = {};
    int32_t pairs = 0;
    for (int i = 0; i < n_mats; i++) {
        a.mat[i] = { (const uint8_t *)descs[i].W, descs[i].dst, descs[i].bias, nullptr, descs[i].m, descs[i].m, descs[i].type, pairs };
        pairs += (int32_t)((descs[i].m + 1) / 2);
    }
    a.n_mats = n_mats; a.k = k; a.mode = MMV_MODE_PLAIN; a.total_pairs = pairs;
    a.act[0] = (const uint8_t *)act_q8K; a.act[1] = (const uint8_t *)act_q80;
    return mmv_launch(a, ncols, (cudaStream_t)stream);
}

extern "C" int b200_mul_mat_vec_q_swiglu(int type_gate, const void * Wg, int type_up, const void * Wu,
                                         const void * act_q8K, const void * act_q80, float * dst,
                                         int64_t m, int64_t k, int64_t ncols, void * stream) {
    MmvArgs a = {};
    a.mat[0] = { (const uint8_t *)Wg, dst, nullptr, nullptr, m, m, type_gate, 0 };
    a.mat[1] = { (const uint8_t *)Wu, dst, nullptr, nullptr, m, m, type_up, 0 };
    a.n_mats = 2; a.k = k; a.mode = MMV_MODE_SWIGLU; a.total_pairs = (int32_t)m;
    a.act[0] = (const uint8_t *)act_q8K; a.act[1] = (const uint8_t *)act_q80;
    return mmv_launch(a, ncols, (cudaStream_t)stream);
}

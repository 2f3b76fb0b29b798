// mmvq.cu — decode matvec on GGUF-quantised weights (sm_100a), ncols <= 8.
//
// Replaces ggml_cuda_mul_mat_vec_q / mul_mat_vec_q<type,ncols_dst> (ggml-cuda/mmvq.cu:139-226,
// 500-570) and vec_dot_q*_q8_1 (vecdotq.cuh:579-839).  Design (HBM-bound; bytes = m * row_bytes):
//   * ONE persistent CTA per SM (16 warps).  Every warp runs its own TMA pipeline: lane 0 issues
//     1-D bulk async copies (cp.async.bulk, mbarrier complete_tx) of the next row-pair segments of the
//     weight matrix into a private shared-memory ring, so the bytes in flight per SM are set by the
//     ring size (~128 KB), not by registers or occupancy — that is what saturates HBM3e.  The
//     reference instead issues 2-/4-byte LDGs from 128-thread blocks, one row per block.
//   * weights never depend on the previous kernel, so the ring is primed BEFORE griddepcontrol.wait:
//     under programmatic dependent launch the weight stream of this matvec starts while the previous
//     kernel is still draining.  Only then is the quantised activation vector staged (one bulk copy
//     per CTA) and consumed with conflict-free LDS.128.
//   * activations are quantised like the CPU oracle (q8_K / RNE q8_0), so per-block integer sums
//     equal the oracle's bit for bit (ggml-cpu/quants.c:115-149,305-333,550-758);
//   * a warp owns TWO rows at a time (activation registers reused); several matrices (QKV, gate+up)
//     share a launch; epilogues fused: bias, residual add, SwiGLU (gate row x up row).
// Rows of Q4_0/Q8_0/Q6_K are repacked at load time so every part of a segment is 16-byte aligned
// (repack.cu); Q4_K/Q5_K keep the GGUF layout.
#include "common.cuh"

#define MMV_WARPS 16
#define MMV_MAX_MATS 4
#define MMV_MAX_STAGES 8

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
    int32_t act_bytes[2];       // bytes to stage per kind (ncols * col_bytes), 0 if unused
    int32_t segc;               // chunks (32 elements) per segment
    int32_t nseg;               // segments per row
    int32_t stages;             // ring depth per warp
    int32_t slot_bytes;         // bytes of one ring slot (two row segments of the widest type in the launch)
};

// ---- per-type geometry of a row segment [c0, c0+segc) chunks ------------------------------------
// bytes one row contributes to a ring slot
__host__ __device__ inline int seg_row_bytes(int type, int segc) {
    switch (type) {
        case B200_TYPE_Q4_0: return segc * 18;
        case B200_TYPE_Q8_0: return segc * 34;
        case B200_TYPE_Q4_K: return segc / 8 * 144;
        case B200_TYPE_Q5_K: return segc / 8 * 176;
        default:             return segc / 8 * 210;     // Q6_K
    }
}

// lane 0 only: copy the segment `seg` of one (repacked) row into smem at `dst`, completing on `bar`
__device__ __forceinline__ void issue_row(int type, uint8_t * dst, const uint8_t * row, int64_t nb, int seg, int segc, uint64_t * bar) {
    const int64_t c0 = (int64_t)seg * segc;
    switch (type) {
        case B200_TYPE_Q4_K: bulk_g2s(dst, row + c0 / 8 * 144, segc / 8 * 144, bar); break;
        case B200_TYPE_Q5_K: bulk_g2s(dst, row + c0 / 8 * 176, segc / 8 * 176, bar); break;
        case B200_TYPE_Q4_0:
            bulk_g2s(dst, row + c0 * 16, segc * 16, bar);
            bulk_g2s(dst + segc * 16, row + nb * 16 + c0 * 2, segc * 2, bar);
            break;
        case B200_TYPE_Q8_0:
            bulk_g2s(dst, row + c0 * 32, segc * 32, bar);
            bulk_g2s(dst + segc * 32, row + nb * 32 + c0 * 2, segc * 2, bar);
            break;
        default: { // Q6_K: ql | qh | sc | d
            const int64_t s0 = c0 / 8; const int ns = segc / 8;
            bulk_g2s(dst,            row + s0 * 128,            ns * 128, bar);
            bulk_g2s(dst + ns * 128, row + nb * 128 + s0 * 64,  ns * 64,  bar);
            bulk_g2s(dst + ns * 192, row + nb * 192 + s0 * 16,  ns * 16,  bar);
            bulk_g2s(dst + ns * 208, row + nb * 208 + s0 * 2,   ns * 2,   bar);
        }
    }
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

// ---- one 32-element chunk: weights from the ring slot (local chunk lc), activations chunk i -----
// Each lane handles 32 elements as two 16-element parts; the order in which the two parts are read
// from the activation vector is swapped on half of the lanes of every quarter-warp so that each
// LDS.128 touches 8 distinct 16-byte bank groups (no conflicts).
template <int T> struct WChunk;

template <> struct WChunk<B200_TYPE_Q4_0> {
    uint4 qs; uint16_t d;
    __device__ __forceinline__ void load(const uint8_t * s, int segc, int lc) { qs = *(const uint4 *)(s + lc * 16); d = *(const uint16_t *)(s + segc * 16 + lc * 2); }
    __device__ __forceinline__ float dot(const ActView & a, int i) const {
        const int sw = (i >> 2) & 1;
        const uint4 A = *(const uint4 *)(a.qs + i * 32 + sw * 16);
        const uint4 B = *(const uint4 *)(a.qs + i * 32 + (sw ^ 1) * 16);
        const int shA = sw * 4, shB = (sw ^ 1) * 4;                   // low nibbles <-> elements 0..15
        int s = dot16(A, (qs.x >> shA) & 0x0F0F0F0Fu, (qs.y >> shA) & 0x0F0F0F0Fu, (qs.z >> shA) & 0x0F0F0F0Fu, (qs.w >> shA) & 0x0F0F0F0Fu);
        s    += dot16(B, (qs.x >> shB) & 0x0F0F0F0Fu, (qs.y >> shB) & 0x0F0F0F0Fu, (qs.z >> shB) & 0x0F0F0F0Fu, (qs.w >> shB) & 0x0F0F0F0Fu);
        s -= 8 * (int)a.bs[i];
        return __fmul_rn(__fmul_rn((float)s, h2f(d)), a.d[i]);          // ggml-cpu/quants.c:146
    }
};
template <> struct WChunk<B200_TYPE_Q8_0> {
    uint4 q0, q1; uint16_t d;
    __device__ __forceinline__ void load(const uint8_t * s, int segc, int lc) {
        q0 = *(const uint4 *)(s + lc * 32); q1 = *(const uint4 *)(s + lc * 32 + 16); d = *(const uint16_t *)(s + segc * 32 + lc * 2);
    }
    __device__ __forceinline__ float dot(const ActView & a, int i) const {
        const int sw = (i >> 2) & 1;
        const uint4 A = *(const uint4 *)(a.qs + i * 32 + sw * 16);
        const uint4 B = *(const uint4 *)(a.qs + i * 32 + (sw ^ 1) * 16);
        const int s = sw ? dot16s(A, q1) + dot16s(B, q0) : dot16s(A, q0) + dot16s(B, q1);
        return __fmul_rn((float)s, __fmul_rn(h2f(d), a.d[i]));           // ggml-cpu/quants.c:330
    }
};
template <> struct WChunk<B200_TYPE_Q4_K> {
    uint4 hdr, qs;
    __device__ __forceinline__ void load(const uint8_t * s, int, int lc) { const uint8_t * b = s + (lc >> 3) * 144; hdr = *(const uint4 *)b; qs = *(const uint4 *)(b + 16 + (lc & 7) * 16); }
    __device__ __forceinline__ float dot(const ActView & a, int i) const {
        const int sb = i >> 3, j = i & 7, c = j >> 1, h = j & 1, sw = c >> 1;
        const int is0 = 2 * c + sw, is1 = 2 * c + (sw ^ 1);           // sub-block (32 elems) of each part
        const uint8_t * base = a.qs + sb * 256 + h * 16;
        const uint4 A = *(const uint4 *)(base + is0 * 32);
        const uint4 B = *(const uint4 *)(base + is1 * 32);
        const int shA = (is0 & 1) * 4, shB = (is1 & 1) * 4;
        const int sA = dot16(A, (qs.x >> shA) & 0x0F0F0F0Fu, (qs.y >> shA) & 0x0F0F0F0Fu, (qs.z >> shA) & 0x0F0F0F0Fu, (qs.w >> shA) & 0x0F0F0F0Fu);
        const int sB = dot16(B, (qs.x >> shB) & 0x0F0F0F0Fu, (qs.y >> shB) & 0x0F0F0F0Fu, (qs.z >> shB) & 0x0F0F0F0Fu, (qs.w >> shB) & 0x0F0F0F0Fu);
        int scA, mnA, scB, mnB;
        k4_scale_min(hdr, is0, scA, mnA);
        k4_scale_min(hdr, is1, scB, mnB);
        const int isum = scA * sA + scB * sB;
        const int imin = mnA * (int)a.bs[sb * 16 + is0 * 2 + h] + mnB * (int)a.bs[sb * 16 + is1 * 2 + h];
        const float da = a.d[sb];
        const float dw = h2f((uint16_t)(hdr.x & 0xffff)), dm = h2f((uint16_t)(hdr.x >> 16));
        return __fmul_rn(dw, da) * (float)isum - __fmul_rn(dm, da) * (float)imin;   // ggml-cpu/quants.c:615-620
    }
};
template <> struct WChunk<B200_TYPE_Q5_K> {
    uint4 hdr, qh, qs;
    __device__ __forceinline__ void load(const uint8_t * s, int, int lc) {
        const uint8_t * b = s + (lc >> 3) * 176; hdr = *(const uint4 *)b; qh = *(const uint4 *)(b + 16 + (lc & 1) * 16); qs = *(const uint4 *)(b + 48 + (lc & 7) * 16);
    }
    __device__ __forceinline__ float dot(const ActView & a, int i) const {
        const int sb = i >> 3, j = i & 7, c = j >> 1, h = j & 1, sw = c >> 1;
        const int is0 = 2 * c + sw, is1 = 2 * c + (sw ^ 1);
        const uint8_t * base = a.qs + sb * 256 + h * 16;
        const uint4 A = *(const uint4 *)(base + is0 * 32);
        const uint4 B = *(const uint4 *)(base + is1 * 32);
        const int shA = (is0 & 1) * 4, shB = (is1 & 1) * 4;
#define Q5W(q, hb, sh, is) ((((q) >> (sh)) & 0x0F0F0F0Fu) | ((((hb) >> (is)) & 0x01010101u) << 4))
        const int sA = dot16(A, Q5W(qs.x, qh.x, shA, is0), Q5W(qs.y, qh.y, shA, is0), Q5W(qs.z, qh.z, shA, is0), Q5W(qs.w, qh.w, shA, is0));
        const int sB = dot16(B, Q5W(qs.x, qh.x, shB, is1), Q5W(qs.y, qh.y, shB, is1), Q5W(qs.z, qh.z, shB, is1), Q5W(qs.w, qh.w, shB, is1));
#undef Q5W
        int scA, mnA, scB, mnB;
        k4_scale_min(hdr, is0, scA, mnA);
        k4_scale_min(hdr, is1, scB, mnB);
        const int isum = scA * sA + scB * sB;
        const int imin = mnA * (int)a.bs[sb * 16 + is0 * 2 + h] + mnB * (int)a.bs[sb * 16 + is1 * 2 + h];
        const float da = a.d[sb];
        const float dw = h2f((uint16_t)(hdr.x & 0xffff)), dm = h2f((uint16_t)(hdr.x >> 16));
        return __fmul_rn(dw, da) * (float)isum - __fmul_rn(dm, da) * (float)imin;
    }
};
template <> struct WChunk<B200_TYPE_Q6_K> {
    uint4 ql, qh; uint32_t sc; uint16_t d;
    __device__ __forceinline__ void load(const uint8_t * s, int segc, int lc) {
        const int ns = segc >> 3, sb = lc >> 3, j = lc & 7, hh = j >> 2, ii = j & 3;
        ql = *(const uint4 *)(s + sb * 128 + hh * 64 + ii * 16);
        qh = *(const uint4 *)(s + ns * 128 + sb * 64 + hh * 32 + (ii & 1) * 16);
        const uint8_t * scp = s + ns * 192 + sb * 16 + hh * 8 + ii;     // scales 8hh+ii and 8hh+ii+4
        sc = (uint32_t)scp[0] | ((uint32_t)scp[4] << 8);
        d = *(const uint16_t *)(s + ns * 208 + sb * 2);
    }
    __device__ __forceinline__ float dot(const ActView & a, int i) const {
        const int sb = i >> 3, j = i & 7, hh = j >> 2, ii = j & 3, sw = hh;
        // part p (0: low nibbles, 1: high nibbles) covers elements 128hh + 16ii + 64p .. +15,
        // 2-bit highs at qh bit 2*(ii/2) + 4p, scale 8hh + ii + 4p
        const int p0 = sw, p1 = sw ^ 1;
        const uint8_t * base = a.qs + sb * 256 + hh * 128 + ii * 16;
        const uint4 A = *(const uint4 *)(base + p0 * 64);
        const uint4 B = *(const uint4 *)(base + p1 * 64);
        const int hs0 = (ii >> 1) * 2 + p0 * 4, hs1 = (ii >> 1) * 2 + p1 * 4;
#define Q6W(q, hb, p, hs) ((((q) >> ((p) * 4)) & 0x0F0F0F0Fu) | ((((hb) >> (hs)) & 0x03030303u) << 4))
        int sA = dot16(A, Q6W(ql.x, qh.x, p0, hs0), Q6W(ql.y, qh.y, p0, hs0), Q6W(ql.z, qh.z, p0, hs0), Q6W(ql.w, qh.w, p0, hs0));
        int sB = dot16(B, Q6W(ql.x, qh.x, p1, hs1), Q6W(ql.y, qh.y, p1, hs1), Q6W(ql.z, qh.z, p1, hs1), Q6W(ql.w, qh.w, p1, hs1));
#undef Q6W
        const int g = sb * 16 + hh * 8 + ii;                          // 16-element group of part p: g + 4p
        sA -= 32 * (int)a.bs[g + 4 * p0];
        sB -= 32 * (int)a.bs[g + 4 * p1];
        const int scA = (int)(int8_t)((sc >> (8 * p0)) & 0xff), scB = (int)(int8_t)((sc >> (8 * p1)) & 0xff);
        return __fmul_rn(h2f(d), a.d[sb]) * (float)(scA * sA + scB * sB);   // ggml-cpu/quants.c:752-754
    }
};

// both rows of a slot, all chunks of the segment, all columns
template <int T, int NCOLS>
__device__ __forceinline__ void slot_dot(const uint8_t * s0, const uint8_t * s1, int segc, int seg, const uint8_t * act_s, int64_t k, int kind, int lane, float (&acc)[2][NCOLS]) {
    const int64_t colb = act_col_bytes(kind, k);
    const int64_t doff = act_d_off(kind, k), boff = act_bsum_off(kind, k);
#pragma unroll 2
    for (int lc = lane; lc < segc; lc += 32) {
        WChunk<T> w0, w1;
        w0.load(s0, segc, lc); w1.load(s1, segc, lc);
        const int i = seg * segc + lc;
#pragma unroll
        for (int c = 0; c < NCOLS; c++) {
            ActView a;
            a.qs = act_s + c * colb;
            a.d  = (const float *)(act_s + c * colb + doff);
            a.bs = (const int16_t *)(act_s + c * colb + boff);
            acc[0][c] += w0.dot(a, i);
            acc[1][c] += w1.dot(a, i);
        }
    }
}

template <int NCOLS>
__device__ __forceinline__ void slot_one_type(int t, const uint8_t * s0, const uint8_t * s1, int segc, int seg, const uint8_t * a0, const uint8_t * a1,
                                              int64_t k, int lane, float (&acc)[2][NCOLS]) {
    switch (t) {
        case B200_TYPE_Q4_K: slot_dot<B200_TYPE_Q4_K, NCOLS>(s0, s1, segc, seg, a0, k, 0, lane, acc); break;
        case B200_TYPE_Q5_K: slot_dot<B200_TYPE_Q5_K, NCOLS>(s0, s1, segc, seg, a0, k, 0, lane, acc); break;
        case B200_TYPE_Q6_K: slot_dot<B200_TYPE_Q6_K, NCOLS>(s0, s1, segc, seg, a0, k, 0, lane, acc); break;
        case B200_TYPE_Q4_0: slot_dot<B200_TYPE_Q4_0, NCOLS>(s0, s1, segc, seg, a1, k, 1, lane, acc); break;
        default:             slot_dot<B200_TYPE_Q8_0, NCOLS>(s0, s1, segc, seg, a1, k, 1, lane, acc); break;
    }
}

template <int NCOLS, int TT>
__device__ __forceinline__ void slot_dispatch(int t0, int t1, const uint8_t * s0, const uint8_t * s1, int segc, int seg, const uint8_t * a0, const uint8_t * a1,
                                              int64_t k, int lane, float (&acc)[2][NCOLS]) {
    if (TT >= 0) {
        switch (TT) {       // uniform-type launch: a single path is compiled
            case B200_TYPE_Q4_K: slot_dot<B200_TYPE_Q4_K, NCOLS>(s0, s1, segc, seg, a0, k, 0, lane, acc); break;
            case B200_TYPE_Q5_K: slot_dot<B200_TYPE_Q5_K, NCOLS>(s0, s1, segc, seg, a0, k, 0, lane, acc); break;
            case B200_TYPE_Q6_K: slot_dot<B200_TYPE_Q6_K, NCOLS>(s0, s1, segc, seg, a0, k, 0, lane, acc); break;
            case B200_TYPE_Q4_0: slot_dot<B200_TYPE_Q4_0, NCOLS>(s0, s1, segc, seg, a1, k, 1, lane, acc); break;
            default:             slot_dot<B200_TYPE_Q8_0, NCOLS>(s0, s1, segc, seg, a1, k, 1, lane, acc); break;
        }
    } else if (t0 == t1) {
        slot_one_type<NCOLS>(t0, s0, s1, segc, seg, a0, a1, k, lane, acc);
    } else {
        // SwiGLU pair with different gate / up types: one row each (second accumulator of each call unused)
        float t[2][NCOLS];
#pragma unroll
        for (int c = 0; c < NCOLS; c++) { t[0][c] = 0.0f; t[1][c] = 0.0f; }
        slot_one_type<NCOLS>(t0, s0, s0, segc, seg, a0, a1, k, lane, t);
#pragma unroll
        for (int c = 0; c < NCOLS; c++) { acc[0][c] += t[0][c]; t[0][c] = 0.0f; t[1][c] = 0.0f; }
        slot_one_type<NCOLS>(t1, s1, s1, segc, seg, a0, a1, k, lane, t);
#pragma unroll
        for (int c = 0; c < NCOLS; c++) acc[1][c] += t[0][c];
    }
}

struct PairInfo { const uint8_t * row0, * row1; int t0, t1; int64_t nb0, nb1; };

template <int MODE>
__device__ __forceinline__ PairInfo pair_info(const MmvArgs & args, int p, MmvMat & M, int64_t & r0, bool & two) {
    PairInfo pi;
    const int64_t k = args.k;
    if (MODE == MMV_MODE_SWIGLU) {
        const MmvMat & g = args.mat[0]; const MmvMat & u = args.mat[1];
        pi.t0 = g.type; pi.t1 = u.type;
        pi.nb0 = k / type_block_elems(g.type); pi.nb1 = k / type_block_elems(u.type);
        pi.row0 = g.W + (int64_t)p * pi.nb0 * type_block_bytes(g.type);
        pi.row1 = u.W + (int64_t)p * pi.nb1 * type_block_bytes(u.type);
        M = g; r0 = p; two = true;
    } else {
        M = args.mat[0];
#pragma unroll
        for (int q = 1; q < MMV_MAX_MATS; q++) if (q < args.n_mats && p >= args.mat[q].pair0) M = args.mat[q];
        r0 = (int64_t)(p - M.pair0) * 2;
        two = r0 + 1 < M.m;
        pi.t0 = pi.t1 = M.type;
        pi.nb0 = pi.nb1 = k / type_block_elems(M.type);
        const int64_t rb = pi.nb0 * type_block_bytes(M.type);
        pi.row0 = M.W + r0 * rb;
        pi.row1 = two ? pi.row0 + rb : pi.row0;
    }
    return pi;
}

template <int NCOLS, int TT, int MODE>
__global__ void __launch_bounds__(MMV_WARPS * 32, 1) mmvq_kernel(const __grid_constant__ MmvArgs args) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t act_bar;
    __shared__ __align__(8) uint64_t full_bar[MMV_WARPS][MMV_MAX_STAGES];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = args.stages, segc = args.segc, nseg = args.nseg;

    uint8_t * act_s0 = smem;
    uint8_t * act_s1 = smem + args.act_bytes[0];
    uint8_t * ring   = smem + ((args.act_bytes[0] + args.act_bytes[1] + 127) & ~127) + (size_t)warp * S * args.slot_bytes;

    if (tid == 0) mbar_init(&act_bar, 1);
    if (lane == 0) for (int s = 0; s < S; s++) mbar_init(&full_bar[warp][s], 1);
    mbar_fence_init();
    __syncthreads();

    // this warp's work: pairs gw, gw + TW, ...; each pair = nseg units (one ring slot each)
    const int TW = gridDim.x * MMV_WARPS;
    const int gw = warp * gridDim.x + blockIdx.x;                 // consecutive pairs -> consecutive SMs
    const int my_pairs = gw < args.total_pairs ? (args.total_pairs - gw + TW - 1) / TW : 0;
    const int n_units = my_pairs * nseg;

    auto issue = [&](int t) {                                     // lane 0 only
        const int p = gw + (t / nseg) * TW, seg = t % nseg, slot = t % S;
        MmvMat M; int64_t r0; bool two;
        const PairInfo pi = pair_info<MODE>(args, p, M, r0, two);
        uint8_t * dst = ring + (size_t)slot * args.slot_bytes;
        const int b0 = seg_row_bytes(pi.t0, segc), b1 = seg_row_bytes(pi.t1, segc);
        mbar_expect_tx(&full_bar[warp][slot], (uint32_t)(b0 + b1));
        issue_row(pi.t0, dst, pi.row0, pi.nb0, seg, segc, &full_bar[warp][slot]);
        issue_row(pi.t1, dst + (args.slot_bytes >> 1), pi.row1, pi.nb1, seg, segc, &full_bar[warp][slot]);
    };

    // prime the ring: weights do not depend on the previous kernel (PDL overlap)
    if (lane == 0) for (int t = 0; t < S && t < n_units; t++) issue(t);
    pdl_trigger();
    pdl_wait();
    if (tid == 0) {
        mbar_expect_tx(&act_bar, (uint32_t)(args.act_bytes[0] + args.act_bytes[1]));
        if (args.act_bytes[0]) bulk_g2s(act_s0, args.act[0], (uint32_t)args.act_bytes[0], &act_bar);
        if (args.act_bytes[1]) bulk_g2s(act_s1, args.act[1], (uint32_t)args.act_bytes[1], &act_bar);
    }
    mbar_wait(&act_bar, 0);

    const int64_t k = args.k;
    float acc[2][NCOLS];
    for (int t = 0; t < n_units; t++) {
        const int pi_idx = t / nseg, seg = t % nseg, slot = t % S;
        const int p = gw + pi_idx * TW;
        if (seg == 0) {
#pragma unroll
            for (int c = 0; c < NCOLS; c++) { acc[0][c] = 0.0f; acc[1][c] = 0.0f; }
        }
        MmvMat M; int64_t r0; bool two;
        const PairInfo pi = pair_info<MODE>(args, p, M, r0, two);
        mbar_wait(&full_bar[warp][slot], (uint32_t)((t / S) & 1));
        const uint8_t * s0 = ring + (size_t)slot * args.slot_bytes;
        slot_dispatch<NCOLS, TT>(pi.t0, pi.t1, s0, s0 + (args.slot_bytes >> 1), segc, seg, act_s0, act_s1, k, lane, acc);
        __syncwarp();                                             // every lane is done reading the slot
        if (lane == 0 && t + S < n_units) issue(t + S);           // refill it
        if (seg == nseg - 1) {
#pragma unroll
            for (int c = 0; c < NCOLS; c++) { acc[0][c] = warp_sum(acc[0][c]); acc[1][c] = warp_sum(acc[1][c]); }
            if (MODE == MMV_MODE_SWIGLU) {
                if (lane == 0) {
#pragma unroll
                    for (int c = 0; c < NCOLS; c++) {
                        M.dst[c * M.dst_col_stride + r0] = __fmul_rn(silu_x86(acc[0][c]), acc[1][c]);   // ggml-cpu/vec.cpp:260-282
                    }
                }
            } else if (lane < 2 && (lane == 0 || two)) {
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
    if (!attr[dev & 63]) { cudaFuncSetAttribute(mmvq_kernel<NCOLS, TT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024); attr[dev & 63] = true; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(MMV_WARPS * 32); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = b200_pdl_enabled() ? 1 : 0;
    int s = b200_check(cudaLaunchKernelEx(&cfg, mmvq_kernel<NCOLS, TT, MODE>, a), "mmvq launch");
    if (s != B200_OK) return s;
    b200_count_launch();
    return B200_OK;
}
template <int NCOLS, int TT> static int mmv_launch_nt(const MmvArgs & a, int mode, size_t smem, int grid, cudaStream_t st) {
    return mode == MMV_MODE_SWIGLU ? mmv_launch_ntm<NCOLS, TT, MMV_MODE_SWIGLU>(a, smem, grid, st)
                                   : mmv_launch_ntm<NCOLS, TT, MMV_MODE_PLAIN>(a, smem, grid, st);
}
template <int NCOLS> static int mmv_launch_n(const MmvArgs & a, int mode, size_t smem, int grid, cudaStream_t st) {
    int tt = a.mat[0].type;
    for (int i = 1; i < a.n_mats; i++) if (a.mat[i].type != tt) tt = -1;
    switch (tt) {
        case B200_TYPE_Q4_0: return mmv_launch_nt<NCOLS, B200_TYPE_Q4_0>(a, mode, smem, grid, st);
        case B200_TYPE_Q8_0: return mmv_launch_nt<NCOLS, B200_TYPE_Q8_0>(a, mode, smem, grid, st);
        case B200_TYPE_Q4_K: return mmv_launch_nt<NCOLS, B200_TYPE_Q4_K>(a, mode, smem, grid, st);
        case B200_TYPE_Q5_K: return mmv_launch_nt<NCOLS, B200_TYPE_Q5_K>(a, mode, smem, grid, st);
        case B200_TYPE_Q6_K: return mmv_launch_nt<NCOLS, B200_TYPE_Q6_K>(a, mode, smem, grid, st);
        default:             return mmv_launch_nt<NCOLS, -1>(a, mode, smem, grid, st);
    }
}

static int mmv_launch(MmvArgs & a, int mode, int64_t ncols, cudaStream_t st) {
    if (ncols < 1 || ncols > 8) { b200_set_error("mmvq: ncols must be 1..8"); return B200_ERR_INVALID; }
    bool need[2] = { false, false };
    bool has_q6 = false;
    for (int i = 0; i < a.n_mats; i++) {
        const MmvMat & M = a.mat[i];
        if (!mmv_type_ok(M.type)) { b200_set_error("mmvq: unsupported weight type %d", M.type); return B200_ERR_UNSUPPORTED; }
        if (!mmv_k_ok(M.type, a.k)) { b200_set_error("mmvq: k=%lld not supported for type %d (needs k%%256==0, Q6_K k%%2048==0)", (long long)a.k, M.type); return B200_ERR_UNSUPPORTED; }
        if (((uintptr_t)M.W & 15) || !M.W || !M.dst || M.m <= 0) { b200_set_error("mmvq: weights must be 16-byte aligned / non-null"); return B200_ERR_INVALID; }
        need[b200_act_kind_for(M.type)] = true;
        has_q6 |= M.type == B200_TYPE_Q6_K;
    }
    size_t act = 0;
    for (int kd = 0; kd < 2; kd++) {
        a.act_bytes[kd] = 0;
        if (need[kd]) {
            if (!a.act[kd] || ((uintptr_t)a.act[kd] & 15)) { b200_set_error("mmvq: missing/unaligned act buffer of kind %d", kd); return B200_ERR_INVALID; }
            a.act_bytes[kd] = (int32_t)(ncols * act_col_bytes(kd, a.k));
            act += a.act_bytes[kd];
        } else a.act[kd] = nullptr;
    }
    // segment = the largest divisor of the row's chunk count that is a multiple of 8 and <= 64
    // (64 chunks = 2048 elements: Q6_K needs exactly that for 16-byte aligned scale/d parts)
    const int nchunks = (int)(a.k / 32);
    int segc = 8;
    for (int c = 64; c >= 8; c -= 8) if (nchunks % c == 0) { segc = c; break; }
    if (has_q6 && segc != 64) { b200_set_error("mmvq: Q6_K rows need k %% 2048 == 0"); return B200_ERR_UNSUPPORTED; }
    a.segc = segc; a.nseg = nchunks / segc;
    int slot = 0;
    for (int i = 0; i < a.n_mats; i++) { const int b = 2 * seg_row_bytes(a.mat[i].type, segc); if (b > slot) slot = b; }
    a.slot_bytes = (slot + 255) & ~255;
    const size_t budget = 216 * 1024;
    const size_t act_al = (act + 127) & ~(size_t)127;
    if (act_al + (size_t)MMV_WARPS * 2 * a.slot_bytes > budget) { b200_set_error("mmvq: activations (%zu bytes) leave no room for the weight ring", act); return B200_ERR_UNSUPPORTED; }
    int stages = (int)((budget - act_al) / ((size_t)MMV_WARPS * a.slot_bytes));
    if (stages > MMV_MAX_STAGES) stages = MMV_MAX_STAGES;
    a.stages = stages;
    const size_t smem = act_al + (size_t)MMV_WARPS * stages * a.slot_bytes;
    const int sms = b200_sm_count();
    int grid = (a.total_pairs + MMV_WARPS - 1) / MMV_WARPS;
    if (grid > sms) grid = sms;
    if (grid < 1) grid = 1;
    switch (ncols) {
        case 1: return mmv_launch_n<1>(a, mode, smem, grid, st);
        case 2: return mmv_launch_n<2>(a, mode, smem, grid, st);
        case 3: return mmv_launch_n<3>(a, mode, smem, grid, st);
        case 4: return mmv_launch_n<4>(a, mode, smem, grid, st);
        case 5: return mmv_launch_n<5>(a, mode, smem, grid, st);
        case 6: return mmv_launch_n<6>(a, mode, smem, grid, st);
        case 7: return mmv_launch_n<7>(a, mode, smem, grid, st);
        default: return mmv_launch_n<8>(a, mode, smem, grid, st);
    }
}

extern "C" int b200_mul_mat_vec_q(int type, const void * W, const void * act, float * dst, int64_t dst_col_stride,
                                  const float * bias, const float * residual, int64_t m, int64_t k, int64_t ncols, void * stream) {
    MmvArgs a = {};
    a.mat[0] = { (const uint8_t *)W, dst, bias, residual, m, dst_col_stride, type, 0 };
    a.n_mats = 1; a.k = k;
    a.total_pairs = (int32_t)((m + 1) / 2);
    const int kind = b200_act_kind_for(type);
    if (kind < 0) { b200_set_error("mmvq: unsupported weight type %d", type); return B200_ERR_UNSUPPORTED; }
    a.act[kind] = (const uint8_t *)act;
    return mmv_launch(a, MMV_MODE_PLAIN, ncols, (cudaStream_t)stream);
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
    a.n_mats = n_mats; a.k = k; a.total_pairs = pairs;
    a.act[0] = (const uint8_t *)act_q8K; a.act[1] = (const uint8_t *)act_q80;
    return mmv_launch(a, MMV_MODE_PLAIN, ncols, (cudaStream_t)stream);
}

extern "C" int b200_mul_mat_vec_q_swiglu(int type_gate, const void * Wg, int type_up, const void * Wu,
                                         const void * act_q8K, const void * act_q80, float * dst,
                                         int64_t m, int64_t k, int64_t ncols, void * stream) {
    MmvArgs a = {};
    a.mat[0] = { (const uint8_t *)Wg, dst, nullptr, nullptr, m, m, type_gate, 0 };
    a.mat[1] = { (const uint8_t *)Wu, dst, nullptr, nullptr, m, m, type_up, 0 };
    a.n_mats = 2; a.k = k; a.total_pairs = (int32_t)m;
    a.act[0] = (const uint8_t *)act_q8K; a.act[1] = (const uint8_t *)act_q80;
    return mmv_launch(a, MMV_MODE_SWIGLU, ncols, (cudaStream_t)stream);
}

// mk.h — phase program of the persistent decode kernel (decode_mk.cu), shared with the executor (executor.cu).
//
// A batch-1 decode step of a llama-family graph (llama-model.cpp:5980-6110) is a fixed chain of
//   [rms_norm ->] quantised matvec(s) [+bias] [+residual] [SwiGLU]      (MK_MMV)
//   rope(q,k) -> KV-cache store -> flash attention                       (MK_ATTN)
// The executor compiles that chain into an array of MkPhase records; ONE kernel launch walks the array with a
// grid-wide barrier between phases, while every warp keeps streaming the weights of the *next* matvec phase into
// its shared-memory ring (weights never depend on activations), so HBM stays busy across phase boundaries instead
// of draining and refilling at every kernel boundary (the reference launches ~10 kernels per layer:
// ggml-cuda.cu:2207-2493).
#pragma once
#include <stdint.h>

#define MK_MAX_MATS 4
#define MK_MMV  1
#define MK_ATTN 2

struct MkRope { int32_t n_dims, neox; float theta_scale, freq_scale, ext_factor, mscale, corr_lo, corr_hi; };

struct MkMat {
    const uint8_t * W;          // repacked weight matrix [m][rb]
    float *         dst;        // [m] (SwiGLU: the shared output of the gate/up pair)
    const float *   bias;       // [m] or null
    const float *   residual;   // [m] or null
    int64_t         rb;         // row bytes
    int32_t         m;          // rows (even)
    int32_t         type;       // B200_TYPE_Q4_K | B200_TYPE_Q6_K
};

struct MkMmv {
    int32_t n_mats, swiglu;     // swiglu: mat[0] = gate, mat[1] = up, same m / type
    int32_t act_source;         // 1: quantise x;  2: rms_norm(x) * norm_w, then quantise
    int32_t k;                  // row length (multiple of 2048)
    float   eps; int32_t pad;
    const float * x;            // [k] f32
    const float * norm_w;       // [k] or null
    MkMat mat[MK_MAX_MATS];
};

struct MkAttn {
    const float * q_src; float * q_dst;            // [n_head][hd] f32: un-roped in, roped out
    const float * k; const float * v;              // [n_head_kv][hd] f32: this token's K (un-roped) and V
    const int32_t * pos; const float * ff;         // position of the token; optional frequency factors
    const int64_t * k_ids; const int64_t * v_ids;  // cache cell of this token
    uint8_t * k_cache; uint8_t * v_cache;          // [cells][n_head_kv][hd] of kv_type
    const uint16_t * mask;                         // f16 [n_kv] or null
    float * dst;                                   // [n_head][hd]
    float * ws; unsigned int * counters;           // split partials / completion counters (zeroed once)
    int64_t k_rs, k_hs, v_rs, v_hs;                // cell / head strides in bytes
    int32_t kv_type, hd, n_head, n_head_kv, n_kv, split_len, n_splits, nh_log2;
    float scale, max_bias, softcap, m0, m1;
    MkRope rp;
};

struct MkPhase {                 // 256 bytes: copied into shared memory with 16 cp.async of 16 bytes
    int32_t kind; int32_t pad;
    union { MkMmv mmv; MkAttn attn; };
    uint8_t pad2[16];
};
static_assert(sizeof(MkPhase) == 256, "MkPhase must be 256 bytes");

// geometry shared by host and device
#define MK_SLOT_BYTES   9216            // one ring slot: 8 rows x 8 super-blocks of Q4_K (4 rows x 8 of Q6_K = 6720)
#define MK_AREG_BYTES   (24 * 1024)     // quantised activation vector / attention scratch
#define MK_MAX_WARPS    10
static inline int64_t mk_act_bytes(int64_t k) { return k + 52 * (k / 256); }   // qs | d f32 | sums16 | sums32

// host entry points (decode_mk.cu)
int  mk_phase_ok_k(int64_t k, int act_source);
int  mk_launch(const MkPhase * dev_prog, int n_phases, unsigned long long * dev_sync, void * stream);

// repack_layout.cuh — where each byte of a ggml weight row lives in the library's row layout (b200_repack_rows, repack.cu).  A header so that
// tests/hostsim can check the numpy model of this layout (tests/refutil.py repack_rows_np, which the wide kernels' CPU tests are built on) against
// the very function the repack kernel runs.
#pragma once
#include <stdint.h>
#include "../../include/b200_ops.h"
#if !defined(__CUDACC__) && !defined(__device__)
#  define __device__
#  define __forceinline__ inline
#endif

// native byte offset -> repacked byte offset, for 2-byte unit `u` of a row with nb blocks
__device__ __forceinline__ int64_t repacked_off(int type, int64_t nb, int64_t off) {
    if (type == B200_TYPE_Q4_0) {
        const int64_t b = off / 18, o = off % 18;
        return o < 2 ? nb * 16 + b * 2 + o : b * 16 + (o - 2);
    } else if (type == B200_TYPE_Q5_0) {                     // d[2] qh[4] qs[16] -> [qs][qh][d]
        const int64_t b = off / 22, o = off % 22;
        if (o < 2) return nb * 20 + b * 2 + o;
        if (o < 6) return nb * 16 + b * 4 + (o - 2);
        return b * 16 + (o - 6);
    } else if (type == B200_TYPE_Q8_0) {
        const int64_t b = off / 34, o = off % 34;
        return o < 2 ? nb * 32 + b * 2 + o : b * 32 + (o - 2);
    } else { // Q6_K: ql[128] qh[64] sc[16] d[2]
        const int64_t b = off / 210, o = off % 210;
        if (o < 128) return b * 128 + o;
        if (o < 192) return nb * 128 + b * 64 + (o - 128);
        if (o < 208) return nb * 192 + b * 16 + (o - 192);
        return nb * 208 + b * 2 + (o - 208);
    }
}


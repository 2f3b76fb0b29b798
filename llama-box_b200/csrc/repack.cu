// repack.cu — load-time, in-place, per-row byte permutation of Q4_0 / Q8_0 / Q6_K weights into a
// 16-byte-aligned structure-of-arrays row (see b200_ops.h).  Plays the role ggml-cpu/repack.cpp
// plays for the CPU backend; the CUDA reference instead keeps the 18/34/210-byte blocks and pays
// with 2-/4-byte loads (ggml-cuda/vecdotq.cuh:18-29).  Row size and offsets do not change.
//
// One CTA per row: the row is read into shared memory (coalesced 2-byte units — every field
// offset in the three formats is even), then written back permuted.
#include "common.cuh"
#include "repack_layout.cuh"

__global__ void __launch_bounds__(256) repack_rows_kernel(uint8_t * rows, int type, int64_t nb, int64_t rb, int inverse) {
    extern __shared__ __align__(16) uint8_t srow[];
    uint8_t * row = rows + (int64_t)blockIdx.x * rb;
    const int64_t units = rb / 2;
    for (int64_t u = threadIdx.x; u < units; u += blockDim.x) ((uint16_t *)srow)[u] = ((const uint16_t *)row)[u];
    __syncthreads();
    for (int64_t u = threadIdx.x; u < units; u += blockDim.x) {
        const int64_t nat = u * 2, rep = repacked_off(type, nb, nat);
        if (!inverse) *(uint16_t *)(row + rep) = *(const uint16_t *)(srow + nat);
        else          *(uint16_t *)(row + nat) = *(const uint16_t *)(srow + rep);
    }
}

// out of place, with row padding: native rows of nb blocks (stride rb) <-> repacked rows of nbp >= nb blocks (stride rbp), pad blocks zero
__global__ void __launch_bounds__(256) repack_pad_kernel(const uint8_t * src, uint8_t * dst, int type, int64_t nb, int64_t rb, int64_t nbp, int64_t rbp, int inverse) {
    const uint8_t * srow = src + (int64_t)blockIdx.x * (inverse ? rbp : rb);
    uint8_t * drow = dst + (int64_t)blockIdx.x * (inverse ? rb : rbp);
    if (!inverse) {
        for (int64_t u = threadIdx.x; u < rbp / 2; u += blockDim.x) ((uint16_t *)drow)[u] = 0;
        __syncthreads();
    }
    for (int64_t u = threadIdx.x; u < rb / 2; u += blockDim.x) {
        const int64_t nat = u * 2, rep = repacked_off(type, nbp, nat);
        if (!inverse) *(uint16_t *)(drow + rep) = *(const uint16_t *)(srow + nat);
        else          *(uint16_t *)(drow + nat) = *(const uint16_t *)(srow + rep);
    }
}

extern "C" int b200_type_is_repacked(int t) { return t == B200_TYPE_Q4_0 || t == B200_TYPE_Q5_0 || t == B200_TYPE_Q8_0 || t == B200_TYPE_Q6_K; }
extern "C" int64_t b200_padded_k(int type, int64_t k) { return padded_k(type, k); }
extern "C" int b200_repack_rows_padded(int type, const void * src, void * dst, int64_t nrows, int64_t k, int inverse, void * stream) {
    if (!type_is_block32(type)) { b200_set_error("repack_padded: only the 32-element block types have padded rows"); return B200_ERR_UNSUPPORTED; }
    if (!src || !dst || src == dst || nrows < 0 || k <= 0 || k % 32 != 0 || (((uintptr_t)src | (uintptr_t)dst) & 1)) { b200_set_error("repack_padded: bad arguments"); return B200_ERR_INVALID; }
    if (nrows == 0) return B200_OK;
    const int64_t nb = k / 32, nbp = padded_k(type, k) / 32, bb = type_block_bytes(type);
    for (int64_t r0 = 0; r0 < nrows; r0 += 1 << 30) {
        const int64_t n = nrows - r0 < (1 << 30) ? nrows - r0 : (1 << 30);
        repack_pad_kernel<<<(unsigned)n, 256, 0, (cudaStream_t)stream>>>((const uint8_t *)src + r0 * (inverse ? nbp : nb) * bb, (uint8_t *)dst + r0 * (inverse ? nb : nbp) * bb, type, nb, nb * bb, nbp, nbp * bb, inverse);
        B200_LAUNCH_CHECK();
    }
    return B200_OK;
}

static int repack_impl(int type, void * rows, int64_t nrows, int64_t k, int inverse, cudaStream_t st) {
    if (!b200_type_is_repacked(type)) return B200_OK;           // Q4_K / Q5_K / F16 / F32 stay native
    const int64_t be = type_block_elems(type);
    if (!rows || nrows < 0 || k <= 0 || k % be != 0) { b200_set_error("repack: k must be a multiple of the block size"); return B200_ERR_INVALID; }
    if (nrows == 0) return B200_OK;
    const int64_t nb = k / be, rb = nb * type_block_bytes(type);
    if (rb > 200 * 1024) { b200_set_error("repack: row of %lld bytes exceeds shared memory", (long long)rb); return B200_ERR_UNSUPPORTED; }
    cudaFuncSetAttribute(repack_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int64_t r0 = 0; r0 < nrows; r0 += 1 << 30) {
        const int64_t n = nrows - r0 < (1 << 30) ? nrows - r0 : (1 << 30);
        repack_rows_kernel<<<(unsigned)n, 256, (size_t)rb, st>>>((uint8_t *)rows + r0 * rb, type, nb, rb, inverse);
        B200_LAUNCH_CHECK();
    }
    return B200_OK;
}

extern "C" int b200_repack_rows(int type, void * rows, int64_t nrows, int64_t k, void * stream) { return repack_impl(type, rows, nrows, k, 0, (cudaStream_t)stream); }
extern "C" int b200_unpack_rows(int type, void * rows, int64_t nrows, int64_t k, void * stream) { return repack_impl(type, rows, nrows, k, 1, (cudaStream_t)stream); }

// extsim.cpp — TEST INFRASTRUCTURE.  Compiles llama-box_b200/csrc/extfmt.cuh — the very functions the CUDA kernels of
// mmvq_ext.cu call per 32-element sub-block — with the host compiler, so their bit manipulation is checked against the
// reference on the CPU (tests/test_extfmt_hostsim.py).  Not linked into any product library.
#include "../../llama-box_b200/csrc/extfmt.cuh"
#include "../../llama-box_b200/csrc/repack_layout.cuh"

extern "C" {
// dot product of one weight row (library layout for Q4_0/Q5_0/Q8_0/Q6_K, ggml layout otherwise) with a quantised column, summed over
// sub-blocks in index order.  nb_layout = blocks per row in the layout.  Returns 0 for a type extfmt.cuh does not know.
int sim_row_dot(int type, const uint8_t * row, int64_t k, int64_t nb_layout, const int8_t * qs, const float * d, const float * s, const int16_t * bs, float * out) {
    XfAct A = { qs, d, s, bs };
    float acc = 0.0f; int ok = 0;
    XF_DISPATCH(type, { for (int64_t u = 0; u < k / 32; u++) acc += xf_sub_dot<T>(row, nb_layout, u, A); ok = 1; });
    *out = acc;
    return ok;
}
int sim_row_dequant(int type, const uint8_t * row, int64_t k, int64_t nb_layout, float * y) {
    int ok = 0;
    XF_DISPATCH(type, { for (int64_t u = 0; u < k / 32; u++) xf_sub_dequant<T>(row, nb_layout, u, y + 32 * u); ok = 1; });
    return ok;
}
// KV cache type q4_0 (ggml's native blocks): quantise a row, dot one row of K blocks with the q8_0 form of a query, de-quantise
void sim_q4_0_quantize_row(const float * x, uint8_t * y, int64_t k) { for (int64_t b = 0; b < k / 32; b++) xf_q4_0_quantize_block(x + 32 * b, y + 18 * b); }
float sim_q4_0n_row_dot(const uint8_t * row, int64_t k, const int8_t * qs, const float * d, const int16_t * bs) {
    float acc = 0.0f;
    for (int64_t b = 0; b < k / 32; b++) acc += xf_q4_0n_dot(row + 18 * b, qs + 32 * b, d[b], bs[b]);
    return acc;
}
void sim_q4_0n_row_dequant(const uint8_t * row, int64_t k, float * y) { for (int64_t e = 0; e < k; e++) y[e] = xf_q4_0n_value(row + 18 * (e / 32), (int)(e % 32)); }
// the library's row layout exactly as the repack kernel computes it (repack.cu: every 2-byte unit of a ggml row moved to repacked_off)
void sim_repack_row(int type, const uint8_t * native, uint8_t * out, int64_t nb, int64_t row_bytes) {
    for (int64_t u = 0; u < row_bytes; u += 2) { const int64_t r = repacked_off(type, nb, u); out[r] = native[u]; out[r + 1] = native[u + 1]; }
}
float sim_h2f(uint16_t h) { return xf_h2f(h); }
int sim_block_bytes(int t) { return xf_block_bytes(t); }
int sim_act_family(int t) { return xf_act_family(t); }
}

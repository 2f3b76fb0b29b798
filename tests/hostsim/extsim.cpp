// extsim.cpp — TEST INFRASTRUCTURE.  Compiles llama-box_b200/csrc/extfmt.cuh — the very functions the CUDA kernels of
// mmvq_ext.cu call per 32-element sub-block — with the host compiler, so their bit manipulation is checked against the
// reference on the CPU (tests/test_extfmt_hostsim.py).  Not linked into any product library.
#include "../../llama-box_b200/csrc/extfmt.cuh"

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
float sim_h2f(uint16_t h) { return xf_h2f(h); }
int sim_block_bytes(int t) { return xf_block_bytes(t); }
int sim_act_family(int t) { return xf_act_family(t); }
}

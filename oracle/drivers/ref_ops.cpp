// ref_ops — TEST INFRASTRUCTURE.  Runs ONE ggml op (as libllama would emit it) on a chosen ggml
// backend and writes the result tensor to a bundle file.  It links the UNMODIFIED reference
// libggml (oracle/_ref) and is used two ways:
//   * --backend CPU                : the reference's ggml-cpu backend = the parity oracle
//                                    (generates tests/golden/*.bin via oracle/make_golden.py)
//   * --backend B200 --plugin x.so : our plugin, through the real ggml_backend C-ABI
//                                    (buffer alloc, set/get_tensor, supports_op, graph_compute)
//
// Bundle format (little endian): u32 magic 'B2TB', u32 n; per tensor: u32 name_len, name,
// i32 ggml_type, i64 ne[4], u64 nbytes, data (contiguous, ggml row layout).
//
// usage: ref_ops --op <name> --in in.bin --out out.bin [--backend CPU] [--plugin lib.so]
//                [--threads N] [key=value ...]
#include "ggml.h"
#include "ggml-backend.h"
#include "ggml-alloc.h"

#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

struct TB { std::string name; int32_t type; int64_t ne[4]; std::vector<uint8_t> data; };

static std::vector<TB> read_bundle(const char * path) {
    std::vector<TB> out;
    FILE * f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    uint32_t magic = 0, n = 0;
    if (fread(&magic, 4, 1, f) != 1 || fread(&n, 4, 1, f) != 1 || magic != 0x42543242u) { fprintf(stderr, "bad bundle\n"); exit(2); }
    for (uint32_t i = 0; i < n; i++) {
        TB t; uint32_t nl; uint64_t nb;
        if (fread(&nl, 4, 1, f) != 1) exit(2);
        t.name.resize(nl);
        if (nl && fread(&t.name[0], 1, nl, f) != nl) exit(2);
        if (fread(&t.type, 4, 1, f) != 1 || fread(t.ne, 8, 4, f) != 4 || fread(&nb, 8, 1, f) != 1) exit(2);
        t.data.resize(nb);
        if (nb && fread(t.data.data(), 1, nb, f) != nb) exit(2);
        out.push_back(std::move(t));
    }
    fclose(f);
    return out;
}

static void write_bundle(const char * path, const std::vector<TB> & ts) {
    FILE * f = fopen(path, "wb");
    if (!f) { fprintf(stderr, "cannot write %s\n", path); exit(2); }
    uint32_t magic = 0x42543242u, n = (uint32_t)ts.size();
    fwrite(&magic, 4, 1, f); fwrite(&n, 4, 1, f);
    for (auto & t : ts) {
        uint32_t nl = (uint32_t)t.name.size(); uint64_t nb = t.data.size();
        fwrite(&nl, 4, 1, f); fwrite(t.name.data(), 1, nl, f);
        fwrite(&t.type, 4, 1, f); fwrite(t.ne, 8, 4, f); fwrite(&nb, 8, 1, f);
        fwrite(t.data.data(), 1, nb, f);
    }
    fclose(f);
}

static std::map<std::string, std::string> kv;
static double P(const char * k, double def) { auto it = kv.find(k); return it == kv.end() ? def : atof(it->second.c_str()); }

int main(int argc, char ** argv) {
    std::string op, in, outp, backend_name = "CPU", plugin;
    int threads = 1;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return std::string(argv[++i]); };
        if      (a == "--op")      op = next();
        else if (a == "--in")      in = next();
        else if (a == "--out")     outp = next();
        else if (a == "--backend") backend_name = next();
        else if (a == "--plugin")  plugin = next();
        else if (a == "--threads") threads = atoi(next().c_str());
        else { auto eq = a.find('='); if (eq == std::string::npos) { fprintf(stderr, "bad arg %s\n", a.c_str()); return 2; } kv[a.substr(0, eq)] = a.substr(eq + 1); }
    }
    ggml_backend_load_all();                       // picks libggml-cpu*.so next to this binary
    if (!plugin.empty() && !ggml_backend_load(plugin.c_str())) { fprintf(stderr, "failed to load plugin %s\n", plugin.c_str()); return 3; }

    ggml_backend_dev_t dev = nullptr;
    for (size_t i = 0; i < ggml_backend_dev_count(); i++) {
        ggml_backend_dev_t d = ggml_backend_dev_get(i);
        if (std::string(ggml_backend_dev_name(d)).rfind(backend_name, 0) == 0) { dev = d; break; }
    }
    if (!dev) { fprintf(stderr, "no device named %s*\n", backend_name.c_str()); return 3; }
    ggml_backend_t be = ggml_backend_dev_init(dev, nullptr);
    if (!be) { fprintf(stderr, "backend init failed\n"); return 3; }
    if (ggml_backend_dev_type(dev) == GGML_BACKEND_DEVICE_TYPE_CPU) {
        auto fn = (ggml_backend_set_n_threads_t) ggml_backend_reg_get_proc_address(ggml_backend_dev_backend_reg(dev), "ggml_backend_set_n_threads");
        if (fn) fn(be, threads);
    }

    std::vector<TB> ins = read_bundle(in.c_str());
    ggml_init_params ip = { ggml_tensor_overhead() * 64 + ggml_graph_overhead(), nullptr, true };
    ggml_context * ctx = ggml_init(ip);
    std::map<std::string, ggml_tensor *> T;
    for (auto & t : ins) {
        ggml_tensor * g = ggml_new_tensor_4d(ctx, (ggml_type)t.type, t.ne[0], t.ne[1], t.ne[2], t.ne[3]);
        ggml_set_name(g, t.name.c_str());
        T[t.name] = g;
    }
    auto need = [&](const char * n) { auto it = T.find(n); if (it == T.end()) { fprintf(stderr, "op %s needs tensor '%s'\n", op.c_str(), n); exit(2); } return it->second; };
    auto opt  = [&](const char * n) -> ggml_tensor * { auto it = T.find(n); return it == T.end() ? nullptr : it->second; };

    ggml_tensor * out = nullptr;
    if (op == "mul_mat") {
        out = ggml_mul_mat(ctx, need("w"), need("x"));
    } else if (op == "mul_mat_id") {
        // as [k, m, n_expert], b [k, n_b1, n_tok] f32, ids i32 [n_used, n_tok] (llama-graph.cpp build_moe_ffn)
        out = ggml_mul_mat_id(ctx, need("w"), need("x"), need("ids"));
    } else if (op == "scale") {
        out = ggml_scale_bias(ctx, need("x"), (float)P("s", 1), (float)P("b", 0));
    } else if (op == "silu") {
        out = ggml_silu(ctx, need("x"));
    } else if (op == "sigmoid") {
        out = ggml_sigmoid(ctx, need("x"));
    } else if (op == "soft_max") {
        out = ggml_soft_max_ext(ctx, need("x"), opt("mask"), (float)P("scale", 1), (float)P("max_bias", 0));
    } else if (op == "argsort") {
        out = ggml_argsort(ctx, need("x"), P("desc", 1) != 0 ? GGML_SORT_ORDER_DESC : GGML_SORT_ORDER_ASC);
    } else if (op == "sum_rows") {
        out = ggml_sum_rows(ctx, need("x"));
    } else if (op == "div") {
        out = ggml_div(ctx, need("a"), need("b"));
    } else if (op == "rms_norm") {
        out = ggml_rms_norm(ctx, need("x"), (float)P("eps", 1e-5));
        if (opt("w")) out = ggml_mul(ctx, out, opt("w"));
    } else if (op == "rope") {
        out = ggml_rope_ext(ctx, need("x"), need("pos"), opt("ff"), (int)P("n_dims", 128), (int)P("mode", 0), (int)P("n_ctx_orig", 8192),
                            (float)P("freq_base", 10000), (float)P("freq_scale", 1), (float)P("ext_factor", 0), (float)P("attn_factor", 1),
                            (float)P("beta_fast", 32), (float)P("beta_slow", 1));
    } else if (op == "set_rows") {
        out = ggml_set_rows(ctx, need("cache"), need("src"), need("ids"));
    } else if (op == "flash_attn") {
        // q_base [dk, n_head, n_tok] permuted to [dk, n_tok, n_head] as llama-graph.cpp:1236 does;
        // k/v: cache tensors [n_embd_gqa, kv_size] viewed as [d, n_head_kv, n_kv] then permuted
        // (llama-kv-cache-unified.cpp:1056-1101, llama-graph.cpp:1237-1238)
        ggml_tensor * q = ggml_permute(ctx, need("q"), 0, 2, 1, 3);
        ggml_tensor * kc = need("k"), * vc = need("v");
        const int64_t dk = (int64_t)P("dk", 128), dv = (int64_t)P("dv", 128), nhkv = (int64_t)P("n_head_kv", 8), n_kv = (int64_t)P("n_kv", kc->ne[1]);
        ggml_tensor * k = ggml_view_3d(ctx, kc, dk, nhkv, n_kv, ggml_row_size(kc->type, dk), ggml_row_size(kc->type, kc->ne[0]), 0);
        ggml_tensor * v = ggml_view_3d(ctx, vc, dv, nhkv, n_kv, ggml_row_size(vc->type, dv), ggml_row_size(vc->type, vc->ne[0]), 0);
        k = ggml_permute(ctx, k, 0, 2, 1, 3);
        v = ggml_permute(ctx, v, 0, 2, 1, 3);
        out = ggml_flash_attn_ext(ctx, q, k, v, opt("mask"), (float)P("scale", 1), (float)P("max_bias", 0), (float)P("softcap", 0));
        ggml_flash_attn_ext_set_prec(out, GGML_PREC_F32);
    } else if (op == "swiglu") {
        out = ggml_swiglu_split(ctx, need("gate"), need("up"));
    } else if (op == "add") {
        out = ggml_add(ctx, need("a"), need("b"));
    } else if (op == "mul") {
        out = ggml_mul(ctx, need("a"), need("b"));
    } else if (op == "get_rows") {
        out = ggml_get_rows(ctx, need("src"), need("ids"));
    } else if (op == "cpy_f16") {
        ggml_tensor * s = need("src");
        ggml_tensor * d = ggml_new_tensor_4d(ctx, GGML_TYPE_F16, s->ne[0], s->ne[1], s->ne[2], s->ne[3]);
        out = ggml_cpy(ctx, s, d);
    } else {
        fprintf(stderr, "unknown op %s\n", op.c_str()); return 2;
    }
    ggml_set_name(out, "dst");
    ggml_cgraph * gf = ggml_new_graph(ctx);
    ggml_build_forward_expand(gf, out);

    if (!ggml_backend_supports_op(be, out)) { fprintf(stderr, "backend %s does not support op %s\n", ggml_backend_name(be), op.c_str()); return 4; }
    ggml_backend_buffer_t buf = ggml_backend_alloc_ctx_tensors(ctx, be);
    if (!buf) { fprintf(stderr, "alloc failed\n"); return 3; }
    for (auto & t : ins) {
        ggml_tensor * g = T[t.name];
        if (ggml_nbytes(g) != t.data.size()) { fprintf(stderr, "tensor %s: %zu bytes given, %zu expected\n", t.name.c_str(), t.data.size(), ggml_nbytes(g)); return 2; }
        ggml_backend_tensor_set(g, t.data.data(), 0, t.data.size());
    }
    int reps = (int)P("reps", 1);
    int64_t t0 = ggml_time_us();
    for (int r = 0; r < reps; r++) {
        if (ggml_backend_graph_compute(be, gf) != GGML_STATUS_SUCCESS) { fprintf(stderr, "graph_compute failed\n"); return 5; }
    }
    ggml_backend_synchronize(be);
    int64_t t1 = ggml_time_us();
    if (reps > 1) printf("{\"op\": \"%s\", \"us_per_run\": %.3f, \"reps\": %d, \"threads\": %d}\n", op.c_str(), (double)(t1 - t0) / reps, reps, threads);

    // SET_ROWS writes into the cache tensor: return the whole cache
    ggml_tensor * res = op == "set_rows" ? T["cache"] : out;
    TB o; o.name = "dst"; o.type = res->type; for (int i = 0; i < 4; i++) o.ne[i] = res->ne[i];
    o.data.resize(ggml_nbytes(res));
    ggml_backend_tensor_get(res, o.data.data(), 0, o.data.size());
    write_bundle(outp.c_str(), { o });

    ggml_backend_buffer_free(buf);
    ggml_free(ctx);
    ggml_backend_free(be);
    return 0;
}

// llama_drv — TEST / MEASUREMENT INFRASTRUCTURE.  A minimal greedy-decode driver over the UNMODIFIED
// reference libllama (oracle/_ref/libllama.so): the *caller* side of the ggml backend boundary,
// doing exactly what llama-box's reconcile loop does per token (httpserver.hpp:3591 llama_decode,
// :4285-4299 sample) with greedy sampling.  Used
//   * with no plug-in and --ngl 0   : the reference's ggml-cpu path (parity oracle for token IDs /
//                                     logits, and the `--impl reference` bench arm);
//   * with --plugin libggml-b200.so : the same libllama driving OUR backend through the C-ABI
//                                     (end-to-end drop-in check and the `e2e` bench number).
//
// usage: llama_drv --model m.gguf [--plugin p.so] [--ngl N] [--threads T] [--ctx C] [--ubatch U]
//                  [--prompt-len P] [--gen G] [--seed S] [--fa] [--ctk f16|q8_0] [--ctv f16|q8_0]
//                  [--ts 1,1,..] [--logits-out file] [--no-repack]
// prints one JSON line: token ids, prefill / decode tokens per second.
#include "llama.h"
#include "ggml-backend.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

static FILE * g_dump = nullptr;
// eval callback (ggml-backend.h:282-289): print a checksum of every f32 node so two backends can be diffed
static bool dump_cb(struct ggml_tensor * t, bool ask, void *) {
    if (ask) return true;
    if (!g_dump || t->type != GGML_TYPE_F32) return true;
    std::vector<float> buf(ggml_nelements(t));
    if (!ggml_is_contiguous(t)) return true;
    ggml_backend_tensor_get(t, buf.data(), 0, ggml_nbytes(t));
    double s = 0, a = 0; for (float v : buf) { s += v; a += v < 0 ? -v : v; }
    fprintf(g_dump, "%-28s %-14s [%lld,%lld,%lld,%lld] sum=%.9g abs=%.9g\n", t->name, ggml_op_name(t->op), (long long)t->ne[0], (long long)t->ne[1], (long long)t->ne[2], (long long)t->ne[3], s, a);
    return true;
}

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char ** argv) {
    std::string model_path, plugin, logits_out, ts;
    int ngl = 0, threads = 8, ctx = 4096, ubatch = 512, prompt_len = 32, gen = 16, seed = 42;
    bool fa = false, no_repack = false;
    ggml_type ctk = GGML_TYPE_F16, ctv = GGML_TYPE_F16;
    auto ty = [](const std::string & s) { return s == "q8_0" ? GGML_TYPE_Q8_0 : GGML_TYPE_F16; };
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return std::string(argv[++i]); };
        if      (a == "--model")      model_path = next();
        else if (a == "--plugin")     plugin = next();
        else if (a == "--ngl")        ngl = atoi(next().c_str());
        else if (a == "--threads")    threads = atoi(next().c_str());
        else if (a == "--ctx")        ctx = atoi(next().c_str());
        else if (a == "--ubatch")     ubatch = atoi(next().c_str());
        else if (a == "--prompt-len") prompt_len = atoi(next().c_str());
        else if (a == "--gen")        gen = atoi(next().c_str());
        else if (a == "--seed")       seed = atoi(next().c_str());
        else if (a == "--fa")         fa = true;
        else if (a == "--no-repack")  no_repack = true;
        else if (a == "--ctk")        ctk = ty(next());
        else if (a == "--ctv")        ctv = ty(next());
        else if (a == "--ts")         ts = next();
        else if (a == "--logits-out") logits_out = next();
        else if (a == "--dump")       { g_dump = fopen(next().c_str(), "w"); }
        else { fprintf(stderr, "unknown arg %s\n", a.c_str()); return 2; }
    }
    llama_log_set([](ggml_log_level lvl, const char * txt, void *) { if (lvl >= GGML_LOG_LEVEL_WARN) fputs(txt, stderr); }, nullptr);
    ggml_backend_load_all();
    if (!plugin.empty() && !ggml_backend_load(plugin.c_str())) { fprintf(stderr, "failed to load plugin %s\n", plugin.c_str()); return 3; }
    llama_backend_init();

    llama_model_params mp = llama_model_default_params();
    mp.n_gpu_layers = ngl;
    mp.use_mmap = true;
    std::vector<float> split(llama_max_devices(), 0.0f);
    if (!ts.empty()) {
        size_t p = 0; int j = 0;
        while (p < ts.size() && j < (int)split.size()) { size_t q = ts.find(',', p); if (q == std::string::npos) q = ts.size(); split[j++] = (float)atof(ts.substr(p, q - p).c_str()); p = q + 1; }
        mp.tensor_split = split.data();
    }
    // parity runs use the plain vec_dot path of the CPU backend (engine_param.hpp:1659-1661 -nr / --no-repack)
    std::vector<ggml_backend_buffer_type_t> no_extra = { nullptr };
    llama_model_tensor_buft_override ov[1] = { { nullptr, nullptr } };
    (void)ov; (void)no_extra;
    mp.use_extra_bufts = !no_repack;

    llama_model * model = llama_model_load_from_file(model_path.c_str(), mp);
    if (!model) { fprintf(stderr, "model load failed\n"); return 4; }
    const llama_vocab * vocab = llama_model_get_vocab(model);
    const int n_vocab = llama_vocab_n_tokens(vocab);

    llama_context_params cp = llama_context_default_params();
    cp.n_ctx = ctx; cp.n_batch = ubatch > 2048 ? ubatch : 2048; cp.n_ubatch = ubatch; cp.n_seq_max = 1;
    cp.n_threads = threads; cp.n_threads_batch = threads;
    cp.flash_attn = fa; cp.type_k = ctk; cp.type_v = ctv; cp.no_perf = true;
    if (g_dump) { cp.cb_eval = dump_cb; cp.cb_eval_user_data = nullptr; }
    llama_context * lctx = llama_init_from_model(model, cp);
    if (!lctx) { fprintf(stderr, "context init failed\n"); return 4; }

    std::mt19937 rng(seed);
    std::vector<llama_token> prompt(prompt_len);
    for (auto & t : prompt) t = (llama_token)(rng() % (uint32_t)n_vocab);

    FILE * lf = logits_out.empty() ? nullptr : fopen(logits_out.c_str(), "wb");
    std::vector<llama_token> out_tokens;

    // prefill
    double t0 = now_s();
    for (int i = 0; i < prompt_len; i += (int)cp.n_batch) {
        int n = std::min((int)cp.n_batch, prompt_len - i);
        if (llama_decode(lctx, llama_batch_get_one(prompt.data() + i, n)) != 0) { fprintf(stderr, "prefill decode failed\n"); return 5; }
    }
    llama_synchronize(lctx);
    double t1 = now_s();
    // greedy decode
    auto pick = [&](const float * lg) { int b = 0; for (int i = 1; i < n_vocab; i++) if (lg[i] > lg[b]) b = i; return (llama_token)b; };
    const float * lg = llama_get_logits_ith(lctx, -1);
    if (lf) fwrite(lg, sizeof(float), n_vocab, lf);
    llama_token tok = pick(lg);
    out_tokens.push_back(tok);
    double t2 = now_s();
    for (int g = 1; g < gen; g++) {
        if (llama_decode(lctx, llama_batch_get_one(&tok, 1)) != 0) { fprintf(stderr, "decode failed at %d\n", g); return 5; }
        lg = llama_get_logits_ith(lctx, -1);
        if (lf) fwrite(lg, sizeof(float), n_vocab, lf);
        tok = pick(lg);
        out_tokens.push_back(tok);
    }
    llama_synchronize(lctx);
    double t3 = now_s();
    if (lf) fclose(lf);

    printf("{\"prompt_len\": %d, \"gen\": %d, \"n_vocab\": %d, \"threads\": %d, \"ngl\": %d, \"prefill_s\": %.6f, \"prefill_tps\": %.3f, \"decode_s\": %.6f, \"decode_tps\": %.3f, \"tokens\": [",
           prompt_len, gen, n_vocab, threads, ngl, t1 - t0, prompt_len / (t1 - t0), t3 - t2, gen > 1 ? (gen - 1) / (t3 - t2) : 0.0);
    for (size_t i = 0; i < out_tokens.size(); i++) printf("%s%d", i ? ", " : "", out_tokens[i]);
    printf("]}\n");
    llama_free(lctx);
    llama_model_free(model);
    llama_backend_free();
    return 0;
}

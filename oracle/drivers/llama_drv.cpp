// llama_drv — TEST / MEASUREMENT INFRASTRUCTURE.  A minimal greedy-decode driver over the UNMODIFIED
// reference libllama (oracle/_ref/libllama.so): the *caller* side of the ggml backend boundary,
// doing exactly what llama-box's reconcile loop does per token (httpserver.hpp:3591 llama_decode,
// :4285-4299 sample) with greedy sampling.  Used
//   * with no plug-in and --ngl 0   : the reference's ggml-cpu path (parity oracle for token IDs /
//                                     logits, and the `--impl reference` bench arm);
//   * with --plugin libggml-b200.so : the same libllama driving OUR backend through the C-ABI
//                                     (end-to-end drop-in check and the `e2e` bench number);
//   * with --plugin libggml-cuda.so : the reference's own GPU backend (oracle/_ref_cuda), the bar to beat.
//
// Two build products of this one file (oracle/Makefile): the `llama_drv` executable, and `libllama_drv.so`
// (same code, -DLLAMA_DRV_LIB) whose extern "C" drv_* entry points bench.py binds with ctypes so that the
// decode loop — and with it libggml-b200.so — runs INSIDE the measuring process.
//
// usage: llama_drv --model m.gguf [--plugin p.so] [--ngl N] [--threads T] [--ctx C] [--ubatch U]
//                  [--prompt-len P] [--gen G] [--seed S] [--fa] [--ctk f16|q8_0] [--ctv f16|q8_0]
//                  [--ts 1,1,..] [--logits-out file] [--no-repack] [--verify-batch B] [--embeddings file] [--repeat R]
// prints one JSON line: token ids, prefill / decode tokens per second (+ inter-device handoff statistics
// when the plug-in exposes them).
#include "llama.h"
#include "ggml-backend.h"

#include <dlfcn.h>

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

struct drv {
    llama_model * model = nullptr;
    llama_context * lctx = nullptr;
    int n_vocab = 0, n_embd = 0;
    int n_past = 0;
    llama_batch batch = {};
    int batch_cap = 0;
    std::string reg_name;
};

static bool g_backends_loaded = false;

extern "C" {

// open a model + context.  ts: "1,1,.." or "" ; ctk/ctv: "f16" | "q8_0".  Returns NULL on failure.
void * drv_open(const char * model_path, const char * plugin, int ngl, const char * ts, int ctx, int ubatch, int threads,
                int fa, const char * ctk, const char * ctv, int no_repack, int embeddings) {
    auto ty = [](const char * s) { return s && !strcmp(s, "q8_0") ? GGML_TYPE_Q8_0 : GGML_TYPE_F16; };
    if (!g_backends_loaded) {
        // warnings and errors only; LLAMA_DRV_LOG_DEBUG=1 lets everything through (e.g. GGML_SCHED_DEBUG's split assignment)
        static const bool all = getenv("LLAMA_DRV_LOG_DEBUG") != nullptr;
        llama_log_set([](ggml_log_level lvl, const char * txt, void *) { if (all || lvl >= GGML_LOG_LEVEL_WARN || lvl == GGML_LOG_LEVEL_CONT) fputs(txt, stderr); }, nullptr);
        ggml_backend_load_all();
        // inside another process (python) the executable's directory holds no backends: also look next to this file's .so
        Dl_info info;
        if (ggml_backend_reg_count() == 0 && dladdr((void *)&now_s, &info) && info.dli_fname) {
            std::string dir = info.dli_fname; const size_t sl = dir.rfind('/'); dir = sl == std::string::npos ? "." : dir.substr(0, sl);
            ggml_backend_load_all_from_path(dir.c_str());
        }
        llama_backend_init();
        g_backends_loaded = true;
    }
    drv * d = new drv();
    if (plugin && *plugin) {
        ggml_backend_reg_t reg = ggml_backend_load(plugin);
        if (!reg) {                                        // already loaded by an earlier drv_open: find it by its devices
            for (size_t i = 0; i < ggml_backend_reg_count(); i++) { const char * nm = ggml_backend_reg_name(ggml_backend_reg_get(i)); if (strcmp(nm, "CPU")) d->reg_name = nm; }
            if (d->reg_name.empty()) { fprintf(stderr, "failed to load plugin %s\n", plugin); delete d; return nullptr; }
        } else d->reg_name = ggml_backend_reg_name(reg);
    }
    llama_model_params mp = llama_model_default_params();
    mp.n_gpu_layers = ngl;
    mp.use_mmap = true;
    std::vector<float> split(llama_max_devices(), 0.0f);
    if (ts && *ts) {
        std::string s = ts; size_t p = 0; int j = 0;
        while (p < s.size() && j < (int)split.size()) { size_t q = s.find(',', p); if (q == std::string::npos) q = s.size(); split[j++] = (float)atof(s.substr(p, q - p).c_str()); p = q + 1; }
        mp.tensor_split = split.data();
    }
    // parity runs use the plain vec_dot path of the CPU backend (engine_param.hpp:1659-1661 -nr / --no-repack)
    mp.use_extra_bufts = !no_repack;
    d->model = llama_model_load_from_file(model_path, mp);
    if (!d->model) { fprintf(stderr, "model load failed\n"); delete d; return nullptr; }
    d->n_vocab = llama_vocab_n_tokens(llama_model_get_vocab(d->model));
    d->n_embd = llama_model_n_embd(d->model);
    llama_context_params cp = llama_context_default_params();
    cp.n_ctx = ctx; cp.n_batch = ubatch > 2048 ? ubatch : 2048; cp.n_ubatch = ubatch; cp.n_seq_max = 1;
    cp.n_threads = threads; cp.n_threads_batch = threads;
    cp.flash_attn = fa != 0; cp.type_k = ty(ctk); cp.type_v = ty(ctv); cp.no_perf = true;
    cp.embeddings = embeddings != 0;
    if (embeddings) cp.pooling_type = LLAMA_POOLING_TYPE_NONE;
    if (g_dump) { cp.cb_eval = dump_cb; cp.cb_eval_user_data = nullptr; }
    d->lctx = llama_init_from_model(d->model, cp);
    if (!d->lctx) { fprintf(stderr, "context init failed\n"); llama_model_free(d->model); delete d; return nullptr; }
    d->batch_cap = (int)cp.n_batch;
    d->batch = llama_batch_init(d->batch_cap, 0, 1);
    return d;
}

int drv_n_vocab(void * h) { return ((drv *)h)->n_vocab; }
int drv_n_embd(void * h)  { return ((drv *)h)->n_embd; }
int drv_n_past(void * h)  { return ((drv *)h)->n_past; }

// llama_decode of n tokens appended at the current position; all_logits != 0 requests an output row for every token
// (speculative verify), otherwise only for the last.  Returns llama_decode's code.
int drv_decode(void * h, const int32_t * tokens, int n, int all_logits) {
    drv * d = (drv *)h;
    if (n > d->batch_cap) return -100;
    llama_batch & b = d->batch;
    b.n_tokens = n;
    for (int i = 0; i < n; i++) { b.token[i] = tokens[i]; b.pos[i] = d->n_past + i; b.n_seq_id[i] = 1; b.seq_id[i][0] = 0; b.logits[i] = (all_logits || i == n - 1) ? 1 : 0; }
    const int rc = llama_decode(d->lctx, b);
    if (rc == 0) d->n_past += n;
    return rc;
}
// logits row of batch token i of the last decode (-1 = last), valid until the next decode (synchronises the backend)
const float * drv_logits(void * h, int i) { return llama_get_logits_ith(((drv *)h)->lctx, i); }
const float * drv_embeddings(void * h, int i) { return llama_get_embeddings_ith(((drv *)h)->lctx, i); }
void drv_sync(void * h) { llama_synchronize(((drv *)h)->lctx); }
// forget the KV cache and start again at position 0 (repeated measurements on one loaded model)
void drv_reset(void * h) { drv * d = (drv *)h; llama_memory_clear(llama_get_memory(d->lctx), true); d->n_past = 0; }
int drv_argmax(const float * lg, int n) { int b = 0; for (int i = 1; i < n; i++) if (lg[i] > lg[b]) b = i; return b; }

// inter-device handoff statistics of the B200 plug-in (reg->get_proc_address); returns 0 when the backend has none
int drv_handoff_stats(void * h, int64_t * copies, int64_t * bytes, double * device_us, double * host_us, int reset) {
    drv * d = (drv *)h;
    if (d->reg_name.empty()) return 0;
    ggml_backend_reg_t reg = ggml_backend_reg_by_name(d->reg_name.c_str());
    if (!reg) return 0;
    typedef void (*stats_fn)(int64_t *, int64_t *, double *, double *); typedef void (*reset_fn)(void);
    stats_fn f = (stats_fn)ggml_backend_reg_get_proc_address(reg, "ggml_b200_handoff_stats");
    if (!f) return 0;
    f(copies, bytes, device_us, host_us);
    if (reset) { reset_fn r = (reset_fn)ggml_backend_reg_get_proc_address(reg, "ggml_b200_handoff_reset"); if (r) r(); }
    return 1;
}
// number of devices the loaded plug-in registered
int drv_n_devices(void * h) {
    drv * d = (drv *)h;
    if (d->reg_name.empty()) return 0;
    ggml_backend_reg_t reg = ggml_backend_reg_by_name(d->reg_name.c_str());
    return reg ? (int)ggml_backend_reg_dev_count(reg) : 0;
}

void drv_close(void * h) {
    drv * d = (drv *)h;
    if (!d) return;
    llama_batch_free(d->batch);
    llama_free(d->lctx);
    llama_model_free(d->model);
    delete d;
}

} // extern "C"

#ifndef LLAMA_DRV_LIB
int main(int argc, char ** argv) {
    std::string model_path, plugin, logits_out, ts, ctk = "f16", ctv = "f16", embd_out;
    int ngl = 0, threads = 8, ctx = 4096, ubatch = 512, prompt_len = 32, gen = 16, seed = 42, verify = 1, repeat = 1;
    bool fa = false, no_repack = false;
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
        else if (a == "--ctk")        ctk = next();
        else if (a == "--ctv")        ctv = next();
        else if (a == "--ts")         ts = next();
        else if (a == "--logits-out") logits_out = next();
        else if (a == "--verify-batch") verify = atoi(next().c_str());
        else if (a == "--embeddings") embd_out = next();
        else if (a == "--repeat")     repeat = atoi(next().c_str());
        else if (a == "--dump")       { g_dump = fopen(next().c_str(), "w"); }
        else { fprintf(stderr, "unknown arg %s\n", a.c_str()); return 2; }
    }
    void * h = drv_open(model_path.c_str(), plugin.c_str(), ngl, ts.c_str(), ctx, ubatch, threads, fa, ctk.c_str(), ctv.c_str(), no_repack, !embd_out.empty());
    if (!h) return 4;
    const int n_vocab = drv_n_vocab(h), n_embd = drv_n_embd(h);

    std::mt19937 rng(seed);
    std::vector<llama_token> prompt(prompt_len);
    for (auto & t : prompt) t = (llama_token)(rng() % (uint32_t)n_vocab);

    FILE * lf = logits_out.empty() ? nullptr : fopen(logits_out.c_str(), "wb");
    FILE * ef = embd_out.empty() ? nullptr : fopen(embd_out.c_str(), "wb");
    std::vector<llama_token> out_tokens;

    // prefill
    double t0 = now_s();
    for (int i = 0; i < prompt_len; i += 2048) {
        int n = std::min(2048, prompt_len - i);
        if (drv_decode(h, prompt.data() + i, n, 0) != 0) { fprintf(stderr, "prefill decode failed\n"); return 5; }
    }
    drv_sync(h);
    double t1 = now_s();
    // greedy decode; with --verify-batch B every step submits B tokens (the sampled one + B-1 pseudo-random "draft" tokens)
    // with an output row for each, as llama-box's speculative verification does (httpserver.hpp, target batch of draft tokens)
    const float * lg = drv_logits(h, -1);
    if (lf) fwrite(lg, sizeof(float), n_vocab, lf);
    if (ef) fwrite(drv_embeddings(h, -1), sizeof(float), n_embd, ef);
    llama_token tok = drv_argmax(lg, n_vocab);
    out_tokens.push_back(tok);
    int64_t c0 = 0, b0 = 0; double du0 = 0, hu0 = 0;
    drv_handoff_stats(h, &c0, &b0, &du0, &hu0, 1);       // reset: count decode-phase handoffs only
    double t2 = now_s();
    std::vector<llama_token> step(verify);
    std::vector<double> run_tps;
    for (int rep = 0; rep < repeat; rep++) {
    const double r0 = now_s();
    for (int g = 1; g < gen; g++) {
        step[0] = tok;
        for (int j = 1; j < verify; j++) step[j] = (llama_token)(rng() % (uint32_t)n_vocab);
        if (drv_decode(h, step.data(), verify, verify > 1) != 0) { fprintf(stderr, "decode failed at %d\n", g); return 5; }
        for (int j = 0; j < verify; j++) {
            lg = drv_logits(h, verify > 1 ? j : -1);
            if (lf) fwrite(lg, sizeof(float), n_vocab, lf);
        }
        if (ef) fwrite(drv_embeddings(h, -1), sizeof(float), n_embd, ef);
        tok = drv_argmax(lg, n_vocab);
        out_tokens.push_back(tok);
    }
    drv_sync(h);
    run_tps.push_back(gen > 1 ? (gen - 1) * verify / (now_s() - r0) : 0.0);
    }
    double t3 = now_s();
    if (lf) fclose(lf);
    if (ef) fclose(ef);
    int64_t hc = 0, hb = 0; double hdu = 0, hhu = 0;
    const int have_stats = drv_handoff_stats(h, &hc, &hb, &hdu, &hhu, 0);

    printf("{\"prompt_len\": %d, \"gen\": %d, \"verify_batch\": %d, \"n_vocab\": %d, \"threads\": %d, \"ngl\": %d, \"prefill_s\": %.6f, \"prefill_tps\": %.3f, \"decode_s\": %.6f, \"decode_tps\": %.3f, ",
           prompt_len, gen, verify, n_vocab, threads, ngl, t1 - t0, prompt_len / (t1 - t0), t3 - t2, gen > 1 ? (double)repeat * (gen - 1) * verify / (t3 - t2) : 0.0);
    printf("\"decode_tps_runs\": [");
    for (size_t i = 0; i < run_tps.size(); i++) printf("%s%.3f", i ? ", " : "", run_tps[i]);
    printf("], ");
    if (have_stats) printf("\"n_devices\": %d, \"handoff\": {\"copies\": %lld, \"bytes\": %lld, \"device_us\": %.3f, \"host_us\": %.3f, \"copies_per_step\": %.3f}, ",
                           drv_n_devices(h), (long long)hc, (long long)hb, hdu, hhu, gen > 1 ? (double)hc / ((double)repeat * (gen - 1)) : 0.0);
    printf("\"tokens\": [");
    for (size_t i = 0; i < out_tokens.size(); i++) printf("%s%d", i ? ", " : "", out_tokens[i]);
    printf("]}\n");
    drv_close(h);
    llama_backend_free();
    return 0;
}
#endif

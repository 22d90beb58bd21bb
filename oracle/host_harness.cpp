// host_harness.cpp — TEST / BENCH INFRASTRUCTURE (built into oracle/_ref/bin/host_harness).
//
// Drives the UNMODIFIED reference host (chatllm's ModelObject / AbstractModel API, src/chat.h:1371-1413, :1006-1056)
// with raw token ids, so that the very same decode loop can be executed
//   * on the reference's CPU backend  (-ngl 0)   -> the oracle / the CPU baseline, and
//   * through the drop-in boundary    (-ngl all) -> our libggml-cuda.so,
// and the logits compared bit-for-bit-ish (1e-3 relative, BASELINE.json) and the per-token wall time read.
// It contains no arithmetic: it calls model->generate_next_token (src/models.cpp:1108-1123 -> run_model :1244-1312).
//
//   host_harness --model M.bin --ggml_dir DIR --ngl all|0 [--threads N] [--max_length L] [--prefill P] [--decode T]
//                [--batch B] [--seed S] [--dump logits.bin] [--feed argmax|seeded]
// Output: one JSON line on stdout.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

// --trace needs the host's BackendContext (a protected member of the model classes); test infrastructure only.
#define protected public
#define private public
#include "chat.h"
#include "backend.h"
#include "models.h"
#include "models_priv.h"
#undef protected
#undef private

using namespace chatllm;

// the host's log sink normally lives in main.cpp (src/main.cpp:984); the harness replaces main.cpp, so it provides one
void log_internal(int level, const char * text) {
    if (level >= 3) fprintf(stderr, "%s", text);  // GGML_LOG_LEVEL_WARN and above
}

// ---- --trace: per-node checksums through the host's eval-observe hook (src/backend.cpp:800-823, :845-851) ----------
static FILE * g_trace = nullptr;
static int g_trace_idx = 0;
static int g_dump_idx = -1;
static FILE * g_full = nullptr;  // --trace_full: raw bytes of every contiguous node output, in order
static std::string g_trace_path;
static void dump_tensor_raw(ggml::tensor * t, const char * tag) {
    if (!t) return;
    std::vector<uint8_t> buf(ggml_nbytes(t));
    ggml_backend_tensor_get(t, buf.data(), 0, buf.size());
    char name[512];
    snprintf(name, sizeof(name), "%s.node%d.%s", g_trace_path.c_str(), g_dump_idx, tag);
    FILE * f = fopen(name, "wb");
    if (f) { fwrite(buf.data(), 1, buf.size(), f); fclose(f); }
    fprintf(g_trace, "# dump %s type=%s ne=[%lld,%lld,%lld,%lld] nb=[%zu,%zu,%zu,%zu]\n", tag, ggml_type_name(t->type), (long long) t->ne[0], (long long) t->ne[1], (long long) t->ne[2], (long long) t->ne[3], t->nb[0], t->nb[1], t->nb[2], t->nb[3]);
}
static bool trace_need(ggml::tensor * t, void *) { return true; }
static bool trace_observe(ggml::tensor * t, void *) {
    if (g_trace_idx == g_dump_idx) { dump_tensor_raw(t, "dst"); dump_tensor_raw(t->src[0], "src0"); dump_tensor_raw(t->src[1], "src1"); }
    const size_t nb = ggml_nbytes(t);
    std::vector<uint8_t> buf(nb);
    ggml_backend_tensor_get(t, buf.data(), 0, nb);
    if (g_full && ggml_is_contiguous(t)) fwrite(buf.data(), 1, nb, g_full);
    double sum = 0, abs = 0;
    const int64_t n = ggml_nelements(t);
    if (ggml_is_contiguous(t)) {
        if (t->type == GGML_TYPE_F32) { const float * p = (const float *) buf.data(); for (int64_t i = 0; i < n; ++i) { if (std::isfinite(p[i])) { sum += p[i]; abs += std::fabs(p[i]); } } }
        else if (t->type == GGML_TYPE_F16) { const ggml_fp16_t * p = (const ggml_fp16_t *) buf.data(); for (int64_t i = 0; i < n; ++i) { float v = ggml_fp16_to_fp32(p[i]); sum += v; abs += std::fabs(v); } }
        else if (t->type == GGML_TYPE_I32) { const int32_t * p = (const int32_t *) buf.data(); for (int64_t i = 0; i < n; ++i) { sum += p[i]; abs += std::abs(p[i]); } }
    }
    fprintf(g_trace, "%d %s %s [%lld,%lld,%lld,%lld] %s sum=%.9g abs=%.9g\n", g_trace_idx++, ggml_op_name(t->op), ggml_type_name(t->type),
            (long long) t->ne[0], (long long) t->ne[1], (long long) t->ne[2], (long long) t->ne[3], ggml_is_contiguous(t) ? "c" : "nc", sum, abs);
    return true;
}
static BackendContext * find_backend_context(AbstractModel * m) {
    if (auto * p = dynamic_cast<ModelProxy *>(m)) m = p->model;
    if (auto * b = dynamic_cast<BaseModelForConditionalGeneration *>(m)) return &b->backend_context;
    return nullptr;
}

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char ** argv) {
    std::string model_path, ggml_dir, ngl = "0", dump, feed = "seeded", trace;
    int threads = 8, max_length = 4352, prefill = 16, decode = 8, batch = 4096, seed = 1, vocab_limit = 0, fake_prefill = 0, skip = 0;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> const char * { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--model") model_path = next();
        else if (a == "--ggml_dir") ggml_dir = next();
        else if (a == "--ngl") ngl = next();
        else if (a == "--threads") threads = atoi(next());
        else if (a == "--max_length") max_length = atoi(next());
        else if (a == "--prefill") prefill = atoi(next());
        else if (a == "--decode") decode = atoi(next());
        else if (a == "--batch") batch = atoi(next());
        else if (a == "--seed") seed = atoi(next());
        else if (a == "--dump") dump = next();
        else if (a == "--feed") feed = next();
        else if (a == "--vocab_limit") vocab_limit = atoi(next());
        else if (a == "--trace") trace = next();
        else if (a == "--fake_prefill") fake_prefill = atoi(next());
        else if (a == "--skip") skip = atoi(next());
        else if (a == "--trace_dump") g_dump_idx = atoi(next());
        else if (a == "--trace_full") g_full = fopen(next(), "wb");
    }
    if (model_path.empty()) { fprintf(stderr, "usage: host_harness --model M --ggml_dir D --ngl all|0 ...\n"); return 2; }

    ComputeManager::init(ggml_dir);
    std::vector<ComputeManager::DeviceInfo> devs;
    ComputeManager::get_devices_info(devs);

    ModelObject::extra_args args(max_length, "", false, threads, batch, "");
    if (ngl != "0") args.model_n_gpu_layers["any"] = ngl;

    const double t_load0 = now_ms();
    ModelObject obj(model_path, args);
    const double t_load = now_ms() - t_load0;
    AbstractModel * model = obj.model.get();

    GenerationConfig gen(max_length, max_length, false, false, 1, 1.0f, 1.0f, threads, "greedy", 1.0f, 1.0f);

    // seeded token stream (LCG), ids in [0, vocab)
    std::vector<float> logits;
    uint64_t state = (uint64_t) seed * 6364136223846793005ULL + 1442695040888963407ULL;
    auto next_id = [&](int vocab) {
        state = state * 6364136223846793005ULL + 1442695040888963407ULL;
        return (int) ((state >> 33) % (uint64_t) vocab);
    };

    // a first 1-token call to learn the vocabulary size (and warm the graph reserve), then rewind
    std::vector<int> ids = {1};
    model->set_n_past(0);
    model->generate_next_token(ids, gen, logits);
    int vocab = (int) logits.size();
    if (vocab_limit > 0 && vocab_limit < vocab) vocab = vocab_limit;

    if (!trace.empty()) {
        g_trace = fopen(trace.c_str(), "w");
        g_trace_path = trace;
        BackendContext * bc = find_backend_context(model);
        if (bc && g_trace) bc->set_eval_observe_callback(trace_need, trace_observe, nullptr);
        else fprintf(stderr, "trace: cannot reach the backend context\n");
    }

    // ---- prefill
    ids.resize(prefill);
    for (int i = 0; i < prefill; ++i) ids[i] = next_id(vocab);
    model->set_n_past(0);
    const double t_p0 = now_ms();
    if (prefill > 0) model->generate_next_token(ids, gen, logits);
    const double t_prefill = now_ms() - t_p0;
    // --fake_prefill N: timing runs only — pretend N positions are already cached (KV buffers hold zeros / whatever the
    // allocator returned); attention is dense so the decode time does not depend on the cache contents.
    if (fake_prefill > 0) prefill = fake_prefill;
    model->set_n_past(prefill);

    FILE * fd = dump.empty() ? nullptr : fopen(dump.c_str(), "wb");
    if (fd && prefill > 0) fwrite(logits.data(), sizeof(float), logits.size(), fd);

    // ---- decode
    std::vector<double> step_ms;
    int n_past = prefill;
    for (int s = 0; s < decode; ++s) {
        int tok;
        if (feed == "argmax" && !logits.empty()) tok = (int) (std::max_element(logits.begin(), logits.end()) - logits.begin());
        else tok = next_id(vocab);
        std::vector<int> one = {tok};
        const double t0 = now_ms();
        model->generate_next_token(one, gen, logits);
        step_ms.push_back(now_ms() - t0);
        n_past++;
        model->set_n_past(n_past);
        if (fd) fwrite(logits.data(), sizeof(float), logits.size(), fd);
    }
    if (fd) fclose(fd);
    if (g_trace) fclose(g_trace);
    if (g_full) fclose(g_full);

    double sum = 0, best = 1e30, sum_skip = 0;
    for (size_t i = 0; i < step_ms.size(); ++i) { const double v = step_ms[i]; sum += v; best = std::min(best, v); if ((int) i >= skip) sum_skip += v; }
    const double mean_skip = (int) step_ms.size() > skip ? sum_skip / (step_ms.size() - skip) : 0;
    std::vector<double> sorted = step_ms;
    std::sort(sorted.begin(), sorted.end());
    const double median = sorted.empty() ? 0 : sorted[sorted.size() / 2];

    printf("{\"model\": \"%s\", \"ngl\": \"%s\", \"devices\": %d, \"device0\": \"%s\", \"threads\": %d, \"vocab\": %d, \"load_ms\": %.1f, "
           "\"prefill_tokens\": %d, \"prefill_ms\": %.3f, \"decode_tokens\": %d, \"decode_ms_total\": %.3f, \"decode_ms_median\": %.4f, "
           "\"decode_ms_min\": %.4f, \"decode_ms_mean_after_skip\": %.4f, \"n_past_end\": %d}\n",
           model_path.c_str(), ngl.c_str(), (int) devs.size(), devs.empty() ? "" : devs[0].name.c_str(), threads, (int) logits.size(), t_load,
           prefill, t_prefill, decode, sum, median, best == 1e30 ? 0 : best, mean_skip, n_past);
    return 0;
}

// oracle/ref_kokoro_driver.cpp -- TEST INFRASTRUCTURE (never linked into the product).
//
// Drives the UNMODIFIED reference Kokoro runner (compiled by oracle/Makefile from
// /root/reference) below the phonemizer, with explicit token ids, and writes the
// results (durations, PCM, optional intermediate graph nodes) to raw float32 files.
// It is the ground truth the numpy restatement (oracle/kokoro_port.py) and the CUDA
// path are pinned against, and the CPU baseline `bench.py --impl reference` times.
//
// The body of run_one() follows kokoro_runner::run / kokoro_duration_runner::run
// (reference src/models/kokoro/model.cpp:1069-1123, 1277-1325) step for step so that
// graph nodes can be flagged as outputs (and therefore survive buffer reuse) before
// the scheduler allocates the graph.  It calls only the reference's own public
// members; no reference source is copied.
//
// usage: kokoro_ref <model.gguf> <tokens.txt> <out_prefix> [--threads N] [--reps R]
//                   [--dump-dur a,b] [--dump-gen a,b] [--list-nodes] [--quiet]
//   tokens.txt : one utterance per line, space separated token ids (incl. BOS/EOS)
//   dump specs : node names (ggml_set_name) or "#<index>" node indices
#include "models/kokoro/model.h"
#include "models/loaders.h"
#include "ggml.h"
#include "ggml-backend.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

using clk = std::chrono::steady_clock;
static double ms_since(clk::time_point t0) { return std::chrono::duration<double, std::milli>(clk::now() - t0).count(); }

static void write_f32(const std::string & path, const float * d, size_t n) {
    FILE * f = fopen(path.c_str(), "wb");
    if (!f) { fprintf(stderr, "cannot write %s\n", path.c_str()); exit(2); }
    fwrite(d, sizeof(float), n, f);
    fclose(f);
}

static std::vector<std::string> split_csv(const std::string & s) {
    std::vector<std::string> out; std::stringstream ss(s); std::string it;
    while (std::getline(ss, it, ',')) if (!it.empty()) out.push_back(it);
    return out;
}

static ggml_tensor * find_node(ggml_cgraph * gf, const std::string & spec) {
    if (!spec.empty() && spec[0] == '#') {
        int idx = atoi(spec.c_str() + 1);
        if (idx < 0) idx += ggml_graph_n_nodes(gf);
        return ggml_graph_node(gf, idx);
    }
    return ggml_graph_get_tensor(gf, spec.c_str());
}

static void list_nodes(ggml_cgraph * gf, const char * tag) {
    int n = ggml_graph_n_nodes(gf);
    for (int i = 0; i < n; i++) {
        ggml_tensor * t = ggml_graph_node(gf, i);
        printf("NODE %s %d op=%d name=%s ne=[%lld,%lld,%lld,%lld]\n", tag, i, (int) t->op, t->name,
               (long long) t->ne[0], (long long) t->ne[1], (long long) t->ne[2], (long long) t->ne[3]);
    }
}

struct dump_req { std::string spec; ggml_tensor * t = nullptr; };

static void dump_all(std::vector<dump_req> & reqs, const std::string & prefix, const char * tag) {
    for (auto & r : reqs) {
        if (!r.t) continue;
        std::vector<float> buf(ggml_nelements(r.t));
        ggml_backend_tensor_get(r.t, buf.data(), 0, ggml_nbytes(r.t));
        std::string nm = r.spec; for (auto & c : nm) if (c == '#') c = 'n';
        write_f32(prefix + "." + tag + "." + nm + ".f32", buf.data(), buf.size());
        printf("DUMP %s %s ne=[%lld,%lld,%lld,%lld]\n", tag, r.spec.c_str(), (long long) r.t->ne[0], (long long) r.t->ne[1],
               (long long) r.t->ne[2], (long long) r.t->ne[3]);
    }
}

int main(int argc, char ** argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s model.gguf tokens.txt out_prefix [opts]\n", argv[0]); return 2; }
    std::string model_path = argv[1], tok_path = argv[2], out_prefix = argv[3];
    int n_threads = 1, reps = 1, warm = -1; bool list = false, quiet = false;
    std::vector<dump_req> ddur, dgen;
    for (int i = 4; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--threads" && i + 1 < argc) n_threads = atoi(argv[++i]);
        else if (a == "--reps" && i + 1 < argc) reps = atoi(argv[++i]);
        else if (a == "--warm" && i + 1 < argc) warm = atoi(argv[++i]);   // leading repetitions excluded from SUMMARY (default: 1 if reps > 1)
        else if (a == "--dump-dur" && i + 1 < argc) { for (auto & s : split_csv(argv[++i])) ddur.push_back({s}); }
        else if (a == "--dump-gen" && i + 1 < argc) { for (auto & s : split_csv(argv[++i])) dgen.push_back({s}); }
        else if (a == "--list-nodes") list = true;
        else if (a == "--quiet") quiet = true;
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }

    std::vector<std::vector<uint32_t>> utts;
    {
        std::ifstream f(tok_path); std::string line;
        while (std::getline(f, line)) {
            std::stringstream ss(line); std::vector<uint32_t> t; long v;
            while (ss >> v) t.push_back((uint32_t) v);
            if (!t.empty()) utts.push_back(t);
        }
    }
    if (utts.empty()) { fprintf(stderr, "no utterances in %s\n", tok_path.c_str()); return 2; }

    auto t_load = clk::now();
    generation_configuration cfg("af_heart", 1, 1.0f, 1.0f, true, "", 0, 1.0f, false);
    auto runner = runner_from_file(model_path.c_str(), n_threads, cfg, true);
    kokoro_runner * kr = static_cast<kokoro_runner *>(runner.get());
    kokoro_model * model = kr->model.get();
    printf("LOAD ms=%.1f threads=%d utterances=%zu\n", ms_since(t_load), n_threads, utts.size());

    double total_audio_s = 0, total_ms = 0, total_dur_ms = 0;
    for (int rep = 0; rep < reps; rep++) {
        for (size_t u = 0; u < utts.size(); u++) {
            auto & toks = utts[u];
            kokoro_ubatch batch; batch.n_tokens = toks.size(); batch.input_tokens = toks.data();
            std::string prefix = out_prefix + ".u" + std::to_string(u);
            auto t0 = clk::now();

            // ---- duration pass (mirrors kokoro_duration_runner::run) ----
            kokoro_duration_runner * dr = kr->drunner;
            kokoro_duration_context * dctx = dr->kctx;
            batch.resp = new kokoro_duration_response;
            std::vector<float> hidden(batch.n_tokens * (model->duration_hidden_size + model->style_half_size));
            std::vector<float> lens(batch.n_tokens);
            ggml_backend_sched_reset(dctx->sched);
            ggml_cgraph * gd = dr->build_kokoro_duration_graph(batch);
            ggml_tensor * lens_t = ggml_graph_node(gd, -1);
            ggml_tensor * hid_t  = ggml_graph_get_tensor(gd, "duration_hidden_states");
            ggml_set_output(hid_t);
            for (auto & r : ddur) { r.t = find_node(gd, r.spec); if (r.t) ggml_set_output(r.t); else fprintf(stderr, "no dur node %s\n", r.spec.c_str()); }
            if (list && rep == 0 && u == 0) list_nodes(gd, "dur");
            ggml_backend_sched_alloc_graph(dctx->sched, gd);
            dr->set_inputs(batch);
            ggml_backend_sched_graph_compute_async(dctx->sched, gd);
            ggml_backend_sched_synchronize(dctx->sched);
            ggml_backend_tensor_get(lens_t, lens.data(), 0, lens.size() * sizeof(float));
            ggml_backend_tensor_get(hid_t, hidden.data(), 0, hidden.size() * sizeof(float));
            if (rep == 0) dump_all(ddur, prefix, "dur");
            ggml_backend_sched_reset(dctx->sched);
            batch.resp->lengths = lens.data(); batch.resp->hidden_states = hidden.data(); batch.resp->n_outputs = batch.n_tokens;
            double dur_ms = ms_since(t0);

            // ---- generation pass (mirrors kokoro_runner::run) ----
            kokoro_context * kctx = kr->kctx;
            ggml_backend_sched_reset(kctx->sched);
            uint32_t total_length = 0;
            for (size_t i = 0; i < batch.n_tokens; i++) total_length += (uint32_t) lens[i];
            size_t n_out = (size_t) total_length * model->up_sampling_factor;
            std::vector<float> pcm(n_out);
            kctx->sequence_length = batch.n_tokens;
            kctx->total_duration  = total_length;
            ggml_cgraph * gg = kr->build_kokoro_graph(batch);
            ggml_tensor * out_t = ggml_graph_node(gg, -1);
            for (auto & r : dgen) { r.t = find_node(gg, r.spec); if (r.t) ggml_set_output(r.t); else fprintf(stderr, "no gen node %s\n", r.spec.c_str()); }
            if (list && rep == 0 && u == 0) list_nodes(gg, "gen");
            ggml_backend_sched_alloc_graph(kctx->sched, gg);
            kr->set_inputs(batch, total_length);
            ggml_backend_sched_graph_compute_async(kctx->sched, gg);
            ggml_backend_sched_synchronize(kctx->sched);
            ggml_backend_tensor_get(out_t, pcm.data(), 0, n_out * sizeof(float));
            if (rep == 0) dump_all(dgen, prefix, "gen");
            ggml_backend_sched_reset(kctx->sched);
            double ms = ms_since(t0);
            delete batch.resp;

            double rms = 0, mx = 0;
            for (float v : pcm) { rms += (double) v * v; if (std::fabs(v) > mx) mx = std::fabs(v); }
            rms = std::sqrt(rms / std::max<size_t>(1, n_out));
            if (rep == 0) {
                write_f32(prefix + ".pcm.f32", pcm.data(), pcm.size());
                write_f32(prefix + ".lens.f32", lens.data(), lens.size());
                write_f32(prefix + ".hidden.f32", hidden.data(), hidden.size());
            }
            if (!quiet) {
                printf("UTT rep=%d u=%zu n_tokens=%zu T=%u samples=%zu ms=%.1f dur_ms=%.1f rms=%.6g max=%.6g\n", rep, u, toks.size(),
                       total_length, n_out, ms, dur_ms, rms, mx);
            }
            if (rep >= (warm >= 0 ? warm : (reps > 1 ? 1 : 0))) { total_audio_s += n_out / 24000.0; total_ms += ms; total_dur_ms += dur_ms; }
        }
    }
    printf("SUMMARY {\"threads\": %d, \"audio_s\": %.4f, \"wall_s\": %.4f, \"duration_graph_s\": %.4f, \"audio_s_per_s\": %.5f}\n", n_threads,
           total_audio_s, total_ms / 1000.0, total_dur_ms / 1000.0, total_ms > 0 ? total_audio_s / (total_ms / 1000.0) : 0.0);
    fflush(stdout);
    (void) runner.release();   // reference destructors are not safe (its own tools leak the runner too: examples/cli/cli.cpp:97)
    _exit(0);
}

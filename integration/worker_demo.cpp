// integration/worker_demo.cpp -- the batch-draining worker of b200_batch_worker.h, run.
//
//   worker_demo selftest
//       no GPU: the reference's own dummy_runner (src/models/dummy) behind a server-shaped queue; checks what the drain promises -- compatible tasks share one forward,
//       incompatible ones keep their place and order, max_batch is respected, timed-out tasks get no response, a runner without generate_batch is driven one by one and
//       every task owns its PCM.  Prints "worker selftest OK".
//   worker_demo stress
//       no GPU: four producers and two workers on one queue (built a second time with -fsanitize=thread: worker_demo_tsan).
//   worker_demo <model.gguf> <prompts.txt> <out_prefix> <max_batch>
//       GPU: all prompts of the file are queued (as the HTTP handlers would), ONE worker thread runs b200::batch_loop over the B200 runner, every response is dumped as
//       <out_prefix>.worker.<i>.f32; then the same prompts go one by one through generate() on a second runner -> <out_prefix>.single.<i>.f32.  Prints the batch sizes
//       and the wall time of both legs.
//
// The task / queue / response-map types below have the SHAPE of examples/server/server.cpp:102-221 (those are defined inside server.cpp, not in a header); they are
// this test's own minimal doubles, not copies: only the members b200_batch_worker.h touches.
#include "models/loaders.h"
#include "models/dummy/model.h"
#include "b200_batch_worker.h"

#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <deque>
#include <fstream>
#include <map>
#include <thread>

enum kind { TTS_KIND, OTHER_KIND };

struct task_t {
    kind                     task = TTS_KIND;
    int                      id   = 0;
    std::string              prompt, model;
    generation_configuration gen_config;
    void *                   response = nullptr;
    size_t                   length   = 0;
    bool                     success  = false;
    float                    sample_rate = 0;
    bool                     expired  = false;
    bool timed_out(int) { return expired; }
};

struct queue_t {
    std::mutex              rw_mutex;
    std::condition_variable condition;
    std::deque<task_t *>    queue;
    bool                    running = true;
    task_t * get_next() {
        std::unique_lock<std::mutex> lock(rw_mutex);
        condition.wait(lock, [&] { return !queue.empty() || !running; });
        if (!running) return nullptr;
        task_t * t = queue.front();
        queue.pop_front();
        return t;
    }
    void push(task_t * t) { std::lock_guard<std::mutex> lock(rw_mutex); queue.push_back(t); condition.notify_one(); }
    void terminate() { std::lock_guard<std::mutex> lock(rw_mutex); running = false; condition.notify_all(); }
};

struct map_t {
    std::mutex              m;
    std::condition_variable updated;
    std::map<int, task_t *> completed;
    std::vector<int>        order;
    void push(task_t * t) { { std::lock_guard<std::mutex> lock(m); completed[t->id] = t; order.push_back(t->id); } updated.notify_all(); }
    void wait_for(size_t n) { std::unique_lock<std::mutex> lock(m); updated.wait(lock, [&] { return completed.size() >= n; }); }
};

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "worker selftest FAILED at %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

static int selftest() {
    dummy_runner runner;
    std::vector<std::unique_ptr<task_t>> tasks;
    auto mk = [&](const char * prompt, const char * voice, kind k = TTS_KIND, const char * model = "m") {
        auto t = std::make_unique<task_t>();
        t->id = (int) tasks.size(); t->prompt = prompt; t->model = model; t->task = k; t->gen_config.voice = voice;
        tasks.push_back(std::move(t));
        return tasks.back().get();
    };
    // queue: a(v1) b(v2) c(v1) OTHER d(v1) e(v1, expired) f(v1, other model) g(v1) h(v1)
    queue_t q; map_t done;
    const char * spec[][2] = { {"a","v1"}, {"bb","v2"}, {"ccc","v1"}, {"-","v1"}, {"dddd","v1"}, {"e","v1"}, {"ff","v1"}, {"gg","v1"}, {"hhh","v1"} };
    for (int i = 0; i < 9; i++) q.push(mk(spec[i][0], spec[i][1], i == 3 ? OTHER_KIND : TTS_KIND, i == 6 ? "m2" : "m"));
    tasks[5]->expired = true;

    // 1. drain semantics, max_batch = 4: head a takes c, d, e (front to back), leaves b, OTHER, f, g, h in order
    task_t * head = q.get_next();
    std::vector<task_t *> batch{ head };
    CHECK(b200::drain_compatible(q, head, TTS_KIND, 4, batch) == 3);
    CHECK(batch.size() == 4 && batch[1]->prompt == "ccc" && batch[2]->prompt == "dddd" && batch[3]->prompt == "e");
    CHECK(q.queue.size() == 5 && q.queue[0]->prompt == "bb" && q.queue[1]->task == OTHER_KIND && q.queue[2]->prompt == "ff" && q.queue[3]->prompt == "gg" && q.queue[4]->prompt == "hhh");

    // 2. one forward for the batch through a generate_batch stand-in that records what it was given; the expired task is dropped without a response
    std::vector<std::vector<std::string>> calls;
    std::vector<std::vector<float>> keep;
    auto fake_batch = [&](tts_generation_runner & r, const std::vector<const char *> & prompts, std::vector<tts_response> & outs, const generation_configuration &) {
        calls.emplace_back(prompts.begin(), prompts.end());
        outs.assign(prompts.size(), tts_response{});
        keep.assign(prompts.size(), {});
        for (size_t i = 0; i < prompts.size(); i++) {
            keep[i].assign(strlen(prompts[i]) * 10, (float) prompts[i][0]);
            outs[i].data = keep[i].data(); outs[i].n_outputs = keep[i].size();
        }
        r.sampling_rate = 24000.0f;
        return true;
    };
    b200::process_batch(batch, runner, done, 300, fake_batch);
    CHECK(calls.size() == 1 && calls[0].size() == 3 && calls[0][0] == "a" && calls[0][1] == "ccc" && calls[0][2] == "dddd");
    CHECK(done.completed.size() == 3 && !done.completed.count(5));
    CHECK(tasks[2]->success && tasks[2]->length == 30 && tasks[2]->sample_rate == 24000.0f && ((float *) tasks[2]->response)[29] == (float) 'c');
    keep.clear();                                                                      // the runner's buffer goes away: the tasks own their PCM
    CHECK(((float *) tasks[4]->response)[39] == (float) 'd');

    // 3. the loop: b alone (v2); OTHER to the other handler; f alone (model m2); g + h together; a runner without generate_batch (the reference's dummy_runner) is
    //    driven one by one and the first task's PCM survives the second generate()
    std::atomic<bool> running{ true };
    std::vector<size_t> sizes;
    int others = 0;
    auto no_batch = [](tts_generation_runner &, const std::vector<const char *> &, std::vector<tts_response> &, const generation_configuration &) { return false; };
    std::thread w([&] {
        b200::batch_loop(running, q, done, 300, 8, TTS_KIND, [&](task_t *) -> tts_generation_runner & { return runner; },
                         [&](task_t * t) { others++; t->success = true; done.push(t); }, no_batch, &sizes);
    });
    done.wait_for(8);
    running = false; q.terminate(); w.join();
    CHECK(others == 1);
    CHECK(sizes.size() == 3 && sizes[0] == 1 && sizes[1] == 1 && sizes[2] == 2);
    CHECK(tasks[7]->length == 2 * 44100 && tasks[8]->length == 3 * 44100 && tasks[7]->sample_rate == 44100.0f);
    tts_response ref{};
    runner.generate("gg", ref, generation_configuration{});
    CHECK(memcmp(ref.data, tasks[7]->response, ref.n_outputs * sizeof(float)) == 0);
    for (auto & t : tasks) if (t->task == TTS_KIND) b200::release(t.get());
    printf("worker selftest OK\n");
    return 0;
}

// worker_demo stress: 4 producers x 100 tasks (3 voices, 2 models, a few non-TTS tasks) against 2 workers draining the same queue with max_batch 8, each with its own
// dummy_runner; every task must be answered exactly once with ITS OWN audio.  The test runs this under ThreadSanitizer (integration/Makefile: worker_demo_tsan).
static int stress() {
    queue_t q; map_t done;
    constexpr int P = 4, N = 100;
    std::vector<std::unique_ptr<task_t>> tasks((size_t) P * N);
    std::atomic<bool> running{ true };
    std::atomic<int> others{ 0 }, forwards{ 0 }, max_seen{ 0 };
    auto fake_batch = [&](tts_generation_runner & r, const std::vector<const char *> & prompts, std::vector<tts_response> & outs, const generation_configuration & cfg) {
        static thread_local std::vector<std::vector<float>> keep;                      // "runner-owned" buffers of this worker, overwritten by its next forward
        keep.assign(prompts.size(), {});
        outs.assign(prompts.size(), tts_response{});
        for (size_t i = 0; i < prompts.size(); i++) {
            keep[i].assign(64, (float) atoi(prompts[i]) + (cfg.voice == "v1" ? 0.25f : cfg.voice == "v2" ? 0.5f : 0.0f));
            outs[i].data = keep[i].data(); outs[i].n_outputs = keep[i].size();
        }
        r.sampling_rate = 24000.0f;
        forwards++;
        int m = max_seen.load(); while ((int) prompts.size() > m && !max_seen.compare_exchange_weak(m, (int) prompts.size())) {}
        return true;
    };
    std::vector<std::thread> workers;
    std::vector<std::unique_ptr<dummy_runner>> runners;
    for (int w = 0; w < 2; w++) runners.push_back(std::make_unique<dummy_runner>());
    for (int w = 0; w < 2; w++) workers.emplace_back([&, w] {
        b200::batch_loop(running, q, done, 300, 8, TTS_KIND, [&, w](task_t *) -> tts_generation_runner & { return *runners[(size_t) w]; },
                         [&](task_t * t) { others++; t->success = true; done.push(t); }, fake_batch);
    });
    std::vector<std::thread> producers;
    for (int p = 0; p < P; p++) producers.emplace_back([&, p] {
        for (int i = 0; i < N; i++) {
            const int id = p * N + i;
            auto t = std::make_unique<task_t>();
            t->id = id; t->prompt = std::to_string(id); t->model = (id % 7 == 0) ? "m2" : "m"; t->task = (id % 23 == 0) ? OTHER_KIND : TTS_KIND;
            t->gen_config.voice = id % 3 == 0 ? "v0" : id % 3 == 1 ? "v1" : "v2";
            task_t * raw = t.get();
            tasks[(size_t) id] = std::move(t);
            q.push(raw);
        }
    });
    for (auto & t : producers) t.join();
    done.wait_for((size_t) P * N);
    running = false; q.terminate();
    for (auto & t : workers) t.join();
    int n_other = 0;
    for (int id = 0; id < P * N; id++) {
        task_t * t = tasks[(size_t) id].get();
        if (t->task == OTHER_KIND) { n_other++; CHECK(t->success); continue; }
        const float want = (float) id + (id % 3 == 1 ? 0.25f : id % 3 == 2 ? 0.5f : 0.0f);
        CHECK(t->success && t->length == 64 && t->sample_rate == 24000.0f);
        for (int k = 0; k < 64; k++) CHECK(((float *) t->response)[k] == want);
        b200::release(t);
    }
    CHECK(others.load() == n_other && (int) done.completed.size() == P * N && max_seen.load() <= 8);
    printf("worker stress OK: %d tasks, %d forwards (largest %d), %d non-TTS\n", P * N, forwards.load(), max_seen.load(), n_other);
    return 0;
}

static void dump(const std::string & path, const float * d, size_t n) {
    FILE * f = fopen(path.c_str(), "wb");
    if (f) { fwrite(d, sizeof(float), n, f); fclose(f); }
}

int main(int argc, char ** argv) {
    if (argc == 2 && std::string(argv[1]) == "selftest") return selftest();
    if (argc == 2 && std::string(argv[1]) == "stress") return stress();
    if (argc < 5) { fprintf(stderr, "usage: %s selftest | model.gguf prompts.txt out_prefix max_batch\n", argv[0]); return 2; }
    std::vector<std::string> lines;
    { std::ifstream in(argv[2]); std::string l; while (std::getline(in, l)) if (!l.empty()) lines.push_back(l); }
    const size_t max_batch = (size_t) atoi(argv[4]);
    const generation_configuration config{ "af_heart", 1, 1.0f, 1.0f, true, "", 0, 1.0f };
    auto a2 = runner_from_file(argv[1], 4, config, true);                              // the worker's runner: a fresh noise stream, like `b` below

    queue_t q; map_t done;
    std::vector<std::unique_ptr<task_t>> tasks;
    for (size_t i = 0; i < lines.size(); i++) {
        auto t = std::make_unique<task_t>();
        t->id = (int) i; t->prompt = lines[i]; t->model = "kokoro"; t->gen_config = config;
        q.push(t.get());
        tasks.push_back(std::move(t));
    }
    std::atomic<bool> running{ true };
    std::vector<size_t> sizes;
    const auto t0 = std::chrono::steady_clock::now();
    std::thread w([&] {
        b200::batch_loop(running, q, done, 300, max_batch, TTS_KIND, [&](task_t *) -> tts_generation_runner & { return *a2; }, [&](task_t *) {},
                         tts_b200_generate_batch, &sizes);
    });
    done.wait_for(lines.size());
    const double ms_worker = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    running = false; q.terminate(); w.join();
    for (size_t i = 0; i < tasks.size(); i++) dump(std::string(argv[3]) + ".worker." + std::to_string(i) + ".f32", (const float *) tasks[i]->response, tasks[i]->length);

    auto b = runner_from_file(argv[1], 4, config, true);
    const auto t1 = std::chrono::steady_clock::now();
    for (size_t i = 0; i < lines.size(); i++) {
        tts_response r{};
        b->generate(lines[i].c_str(), r, config);
        dump(std::string(argv[3]) + ".single." + std::to_string(i) + ".f32", r.data, r.n_outputs);
    }
    const double ms_single = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
    printf("worker_demo: %zu prompts, batches", lines.size());
    for (size_t s : sizes) printf(" %zu", s);
    printf("; worker %.1f ms, one by one %.1f ms (incl. dumps)\n", ms_worker, ms_single);
    (void) a2.release(); (void) b.release();
    return 0;
}

// integration/b200_batch_worker.h -- SURVEY 8(f) row 1: the batch-draining server worker.
//
// The reference's server worker (examples/server/server.cpp:247-271) pops ONE task and calls runner.generate(): on a GPU that leaves the batched forward
// (b2tts_kokoro_run_chunks, 32 utterances in 21 ms) running one utterance at a time.  This header is the replacement for the body of worker::loop(): pop the next task as
// upstream does (blocking), then take every further queued TTS task that can share its forward (same model, same voice / sampling configuration) out of the queue WITHOUT
// blocking, up to `max_batch`, and run them through ONE tts_b200_generate_batch call.  Tasks that cannot join the batch keep their place and order in the queue.
//
// It is written against the SHAPE of upstream's types, not their definitions (they live inside server.cpp, not in a header):
//     Task : task, prompt, gen_config, model, response, length, sample_rate, success, timed_out(int)      (server.cpp:102-125)
//     Queue: rw_mutex, queue (a std::deque<Task*>), get_next()                                            (server.cpp:127-160)
//     Map  : push(Task*)                                                                                  (server.cpp:162-221)
// so that the two-line change a maintainer makes in server.cpp is:
//     void loop() { b200::batch_loop(running, *task_queue, *response_map, task_timeout, max_batch, TTS,
//                                    [&](auto * t) -> tts_generation_runner & { return *runners[t->model]; }, [&](auto * t) { process_task(t); }); }
// Semantics kept from process_task (server.cpp:258-275): a timed-out task is dropped without a response; success = n_outputs != 0; sample_rate from the runner.
// One difference, needed as soon as two tasks share a forward: task->response is a malloc'ed copy owned by the task (upstream hands out a pointer into a runner-owned
// buffer that the next generate() overwrites) -- the HTTP handler frees it after writing the WAV (b200::release).
#pragma once
#include "tts_b200.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

namespace b200 {

// the fields of generation_configuration (include/common.h:40-66) a batched forward shares between its utterances
inline bool same_forward(const generation_configuration & a, const generation_configuration & b) {
    return a.voice == b.voice && a.sample == b.sample && a.top_k == b.top_k && a.temperature == b.temperature && a.repetition_penalty == b.repetition_penalty &&
           a.top_p == b.top_p && a.use_cross_attn == b.use_cross_attn && a.max_tokens == b.max_tokens && a.espeak_voice_id == b.espeak_voice_id;
}

// Takes the queued tasks that can join `head`'s forward out of the queue (front to back, order kept), never blocking and never reordering what stays behind.
template <class Queue, class Task, class TaskKind>
size_t drain_compatible(Queue & q, Task * head, TaskKind tts_kind, size_t max_batch, std::vector<Task *> & batch) {
    size_t taken = 0;
    std::lock_guard<std::mutex> lock(q.rw_mutex);
    for (auto it = q.queue.begin(); it != q.queue.end() && batch.size() < max_batch;) {
        Task * t = *it;
        if (t->task == tts_kind && t->model == head->model && same_forward(t->gen_config, head->gen_config)) {
            batch.push_back(t);
            it = q.queue.erase(it);
            taken++;
        } else {
            ++it;
        }
    }
    return taken;
}

// One batched forward for `batch` (all TTS, one runner).  `generate_batch` has tts_b200_generate_batch's signature; a runner that does not support it (returns false)
// is driven one prompt at a time, as upstream does.
template <class Task, class Map, class GenerateBatch>
void process_batch(std::vector<Task *> & batch, tts_generation_runner & runner, Map & responses, int task_timeout, GenerateBatch && generate_batch) {
    std::vector<Task *> live;
    for (Task * t : batch) if (!t->timed_out(task_timeout)) live.push_back(t);      // upstream drops a timed-out task without a response (server.cpp:259-261)
    if (live.empty()) return;
    std::vector<const char *> prompts;
    for (Task * t : live) prompts.push_back(t->prompt.c_str());
    std::vector<tts_response> outs;
    if (!generate_batch(runner, prompts, outs, live[0]->gen_config)) {
        outs.assign(live.size(), tts_response{});
        for (size_t i = 0; i < live.size(); i++) {
            runner.generate(prompts[i], outs[i], live[i]->gen_config);
            if (outs[i].n_outputs && i + 1 < live.size()) {                         // the runner-owned buffer is about to be overwritten: copy now
                float * own = (float *) std::malloc(outs[i].n_outputs * sizeof(float));
                std::memcpy(own, outs[i].data, outs[i].n_outputs * sizeof(float));
                live[i]->response = own; outs[i].data = nullptr;
            }
        }
    }
    for (size_t i = 0; i < live.size(); i++) {
        Task * t = live[i];
        if (outs[i].data && outs[i].n_outputs) {
            float * own = (float *) std::malloc(outs[i].n_outputs * sizeof(float));
            std::memcpy(own, outs[i].data, outs[i].n_outputs * sizeof(float));
            t->response = own;
        } else if (!outs[i].n_outputs) {
            t->response = nullptr;
        }
        t->length      = outs[i].n_outputs;
        t->sample_rate = runner.sampling_rate;
        t->success     = outs[i].n_outputs != 0;
        responses.push(t);
    }
}

template <class Task> void release(Task * t) { std::free(t->response); t->response = nullptr; }

// The body of worker::loop().  `runner_of(task)` returns the runner of the task's model, `other(task)` handles every non-TTS task (upstream's process_task).
template <class Queue, class Map, class TaskKind, class RunnerOf, class Other, class GenerateBatch>
void batch_loop(std::atomic<bool> & running, Queue & q, Map & responses, int task_timeout, size_t max_batch, TaskKind tts_kind, RunnerOf && runner_of, Other && other,
                GenerateBatch && generate_batch, std::vector<size_t> * batch_sizes = nullptr) {
    while (running) {
        auto * head = q.get_next();                                                 // blocks like upstream; nullptr = queue terminated
        if (!head) continue;
        if (!(head->task == tts_kind)) { other(head); continue; }
        std::vector<decltype(head)> batch{ head };
        if (max_batch > 1) drain_compatible(q, head, tts_kind, max_batch, batch);
        if (batch_sizes) batch_sizes->push_back(batch.size());
        if (std::getenv("B2TTS_WORKER_LOG")) { fprintf(stderr, "b200 worker: forward of %zu task(s)\n", batch.size()); fflush(stderr); }
        process_batch(batch, runner_of(head), responses, task_timeout, generate_batch);
    }
}

template <class Queue, class Map, class TaskKind, class RunnerOf, class Other>
void batch_loop(std::atomic<bool> & running, Queue & q, Map & responses, int task_timeout, size_t max_batch, TaskKind tts_kind, RunnerOf && runner_of, Other && other) {
    batch_loop(running, q, responses, task_timeout, max_batch, tts_kind, runner_of, other, tts_b200_generate_batch);
}

}  // namespace b200

// integration/tts_b200.h -- the one addition to the reference's API surface (SURVEY 8b "Batch"): several prompts through ONE batched forward.
// Declared here, defined next to the B200 runners (integration/kokoro_b200_runner.cpp); everything else is the reference's own include/common.h + src/models/loaders.h.
#pragma once
#include "common.h"

// outputs[i] = what runner.generate(prompts[i], ...) would have returned had the prompts been submitted one after another on this runner (same tokens, same position
// in the reference's process-wide noise stream); outputs[i].data points into a runner-owned buffer valid until the next call on the runner.
// Returns false when `runner` is not a B200 runner that supports it (the caller then loops over generate()).
bool tts_b200_generate_batch(tts_generation_runner & runner, const std::vector<const char *> & prompts, std::vector<tts_response> & outputs, const generation_configuration & config);

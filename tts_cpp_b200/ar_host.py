"""Host-side steps between the autoregressive decode loops and the codec decoders -- the Python mirror of what the C++ shim keeps from the
reference (index shuffles over a few thousand integers; nothing here touches the GPU).

    parler_adjust_output_tokens : parler_tts_runner::adjust_output_tokens (reference src/models/parler/model.cpp:734-760)
    dia_adjust_output_tokens    : dia_runner::adjust_output_tokens        (reference src/models/dia/model.cpp:825-847)

Both undo the delay pattern (frame i takes head h's token from step i + delay[h]) and drop every frame in which some head produced a special id
(>= audio_vocab_size), giving the frame-major [frames][heads] code layout dac_runner::run / b2tts_dac_decode_batch takes."""
from __future__ import annotations

import numpy as np

DIA_DELAY_PATTERN = (0, 8, 9, 10, 11, 12, 13, 14, 15)     # dia_model::delay_pattern (reference src/models/dia/model.h:85)


def parler_adjust_output_tokens(tokens: np.ndarray, audio_vocab: int = 1024) -> np.ndarray:
    """tokens [steps][heads] as sampled -> codes [frames][heads].  Head h is delayed by h steps.  The reference indexes one element past the end when
    i + h == steps (`next_index > size` instead of `>=`, undefined behaviour); here such frames are dropped like the ones further out."""
    tokens = np.asarray(tokens)
    steps, heads = tokens.shape
    out = []
    for i in range(steps):
        if i + heads - 1 >= steps:
            continue
        row = tokens[i + np.arange(heads), np.arange(heads)]
        if (row >= audio_vocab).any():
            continue
        out.append(row)
    return np.asarray(out, np.uint32).reshape(-1, heads)


def dia_adjust_output_tokens(tokens: np.ndarray, audio_vocab: int = 1024, max_delay: int = 15, delay_pattern=DIA_DELAY_PATTERN) -> np.ndarray:
    """tokens [steps][heads] as sampled -> codes [frames][heads]: the first steps - max_delay frames, head h read delay_pattern[h] steps later."""
    tokens = np.asarray(tokens)
    steps, heads = tokens.shape
    d = np.asarray(delay_pattern[:heads])
    out = []
    for i in range(steps - max_delay):
        row = tokens[i + d, np.arange(heads)]
        if (row >= audio_vocab).any():
            continue
        out.append(row)
    return np.asarray(out, np.uint32).reshape(-1, heads)

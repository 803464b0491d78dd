"""Host-side steps between the autoregressive decode loops and the codec decoders -- the Python mirror of what the C++ shim keeps from the
reference (index shuffles over a few thousand integers; nothing here touches the GPU).

    parler_adjust_output_tokens : parler_tts_runner::adjust_output_tokens (reference src/models/parler/model.cpp:734-760)
    dia_adjust_output_tokens    : dia_runner::adjust_output_tokens        (reference src/models/dia/model.cpp:825-847)
    orpheus_n_generated         : the exit test of orpheus_runner::generate_from_batch (reference src/models/orpheus/model.cpp:389-398)
    orpheus_prepare_output_tokens : orpheus_runner::prepare_output_tokens   (reference src/models/orpheus/model.cpp:371-387)

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


ORPHEUS_HEADS = (0, 1, 2, 2, 1, 2, 2)          # orpheus_model::heads (reference src/models/orpheus/model.h:44): SNAC level of each token of a 7-token frame


def orpheus_n_generated(tokens: np.ndarray, stopping_token: int = 128258, max_generation: int = 2100) -> int:
    """How many tokens of a decoded stream the reference's loop would have kept: it stops right after sampling the stopping token (which stays in the
    stream) or once max_generation tokens exist."""
    tokens = np.asarray(tokens).reshape(-1)
    hit = np.nonzero(tokens == stopping_token)[0]
    n = int(hit[0]) + 1 if hit.size else tokens.size
    return min(n, int(max_generation))


def orpheus_prepare_output_tokens(tokens: np.ndarray, heads=ORPHEUS_HEADS) -> list[np.ndarray]:
    """tokens (the kept stream) -> the three SNAC code vectors [coarse, medium, fine]: whole 7-token frames only, token ii of a frame minus
    128266 + (ii % 7) * 4096 goes to level heads[ii] (unsigned 32-bit arithmetic like the reference, so ids below the audio range wrap)."""
    tokens = np.asarray(tokens, np.uint32).reshape(-1)
    chunks = tokens.size // 7
    out = [[] for _ in range(max(heads) + 1)]
    for i in range(chunks):
        for ii in range(7):
            t = (int(tokens[i * 7 + ii]) - 128266 - (ii % 7) * 4096) & 0xFFFFFFFF
            out[heads[ii]].append(t)
    return [np.asarray(o, np.uint32) for o in out]

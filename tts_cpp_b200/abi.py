"""Helpers around include/b2tts.h (no GPU needed): the list of C-ABI symbols the header declares."""
from __future__ import annotations

import os
import re

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "b2tts.h")


def declared_symbols() -> list[str]:
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2tts_[a-z0-9_]+)\s*\(", src)))

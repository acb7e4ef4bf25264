"""The GPT-2 byte <-> unicode table every byte-level tokenizer spells its tokens in (reference zett/utils.py:351-609
``CHARS_TO_BYTES``)."""
from typing import Dict

def _bytes_to_chars() -> Dict[int, str]:
    keep = list(range(33, 127)) + list(range(161, 173)) + list(range(174, 256))
    table, n = {}, 0
    for b in range(256):
        if b in keep:
            table[b] = chr(b)
        else:
            table[b] = chr(256 + n)
            n += 1
    return table


BYTES_TO_CHARS: Dict[int, str] = _bytes_to_chars()
CHARS_TO_BYTES: Dict[str, int] = {c: b for b, c in BYTES_TO_CHARS.items()}


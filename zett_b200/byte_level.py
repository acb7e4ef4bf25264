"""``convert_to_byte_level`` -- rewrite any HF fast tokenizer (Unigram / BPE / WordPiece) as a byte-level tokenizer over
the 256-character GPT-2 alphabet, the step immediately upstream of ``get_surface_form_matrix`` in the reference's
transfer driver (zett/tokenizer_converters.py:78-406, called at scripts/transfer.py:153-159,198-202).

Host-side string surgery on the tokenizer's JSON; runs once per tokenizer.  Behaviour follows the reference:

* every non-special token is mapped to its byte-level spelling (metaspace / continuing-subword-prefix conventions
  folded into a leading ``Ġ``), byte-fallback pieces ``<0xXX>`` become the byte's character when that is free, the
  alphabet characters missing from the vocabulary are appended;
* ``make_whitespace_consistent`` renames tokens with irregular whitespace runs to ``<unused_whitespace__i>`` and
  appends the canonical runs; ``match_special_tokens_to`` re-indexes the special tokens like another tokenizer;
* Unigram keeps its scores (missing alphabet characters get ``-100000``), BPE keeps its merges and gains the merges
  needed to rebuild multi-character atoms no merge produces (a UTF-8 character is several byte-characters);
* normalizer := ``Prepend(" ")``, pre-tokenizer := ``Split(SPLIT_REGEX, removed, inverted) + ByteLevel``, decoder :=
  ``ByteLevel``.

One deliberate difference: the reference walks Python ``set``s when it emits the extra BPE merges, so their relative
order changes with ``PYTHONHASHSEED``; here that walk is sorted, which makes the output reproducible.  The vocabulary,
the set of merges and the position of the original merges are identical (tests/test_byte_level.py checks them against
tokenizers converted by the reference itself).
"""
from __future__ import annotations

import copy
import json
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
from tokenizers import Tokenizer, decoders, models, pre_tokenizers

from .byte_alphabet import BYTES_TO_CHARS, CHARS_TO_BYTES

# zett/utils.py:23,29
NEGATIVE_INF_FILL_VALUE = -100_000
SPLIT_REGEX = r"'s|'t|'re|'ve|'m|'ll|'d| ?[\p{L}\p{M}]+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+"
WHITESPACE_CHARS = ["Ġ", "Ċ", "ĉ"]  # space, newline, tab in the byte alphabet


def _tokenizer_json(tok) -> dict:
    return json.loads(tok._tokenizer.to_str())


def _has_byte_level_pretokenizer(tok, data: dict) -> bool:
    """zett/tokenizer_converters.py:30-37"""
    if isinstance(tok._tokenizer.pre_tokenizer, pre_tokenizers.ByteLevel):
        return True
    pre = data.get("pre_tokenizer") or {}
    return pre.get("type") == "Sequence" and any(p["type"] == "ByteLevel" for p in pre["pretokenizers"])


def _byte_spelling(tok, data: dict) -> Tuple[Callable[[str], str], Optional[str]]:
    """token string -> byte-level spelling, and the continuing-subword prefix that was folded away
    (zett/tokenizer_converters.py:40-75)."""
    if _has_byte_level_pretokenizer(tok, data):
        assert len(data["model"].get("continuing_subword_prefix") or "") == 0
        return (lambda t: t), None
    backend = tok._tokenizer

    def surface(text: str) -> str:
        if backend.normalizer is not None:
            text = backend.normalizer.normalize_str(text)
        if backend.pre_tokenizer is not None:
            text = backend.pre_tokenizer.pre_tokenize_str(text)[0][0]
        return text

    probe = surface(" test")
    meta = probe[0] if (probe[0] != " " and probe != "test") else None  # e.g. the sentencepiece underline
    prefix = data["model"].get("continuing_subword_prefix")

    def spell(token: str) -> str:
        if meta is not None:
            token = token.replace(meta, " ")
        if prefix is not None:
            token = token[len(prefix):] if token.startswith(prefix) else " " + token
        return "".join(BYTES_TO_CHARS[b] for b in token.encode("utf-8"))

    return spell, prefix


def _whitespace_count(s: str) -> int:
    return sum(c in WHITESPACE_CHARS for c in s)


def _canonical_whitespace_tokens() -> List[str]:
    return [c2 + c1 * i for c1 in WHITESPACE_CHARS for i in range(1, 16) for c2 in WHITESPACE_CHARS]


def _fix_postprocessor_ids(post: dict, surface_forms: List[str]) -> None:
    """zett/tokenizer_converters.py:16-27"""
    kind = post["type"]
    if kind == "TemplateProcessing":
        for entry in post["special_tokens"].values():
            entry["ids"] = [surface_forms.index(t) for t in entry["tokens"]]
    elif kind == "RobertaProcessing":
        post["sep"][1] = surface_forms.index(post["sep"][0])
        post["cls"][1] = surface_forms.index(post["cls"][0])
    elif kind == "Sequence":
        for sub in post["processors"]:
            _fix_postprocessor_ids(sub, surface_forms)


# ---- BPE merge repair ------------------------------------------------------------------------------------------------
def _undecomposable_atoms(token: str, producers: Dict[str, List[Tuple[str, str]]]) -> set:
    """Expand ``token`` through every merge that produces it, recursively; return the pieces no merge produces
    (zett/tokenizer_converters.py:284-301 -- the fixed point of that loop does not depend on its iteration order)."""
    leaves, seen, pending = set(), set(), [token]
    while pending:
        piece = pending.pop()
        if piece in seen:
            continue
        seen.add(piece)
        ways = producers.get(piece)
        if ways is None:
            leaves.add(piece)
            continue
        for left, right in ways:
            pending.append(left)
            pending.append(right)
    return leaves


def _merges_to_build(token: str, known: set) -> Tuple[List[str], set]:
    """Greedy pairwise merges that assemble ``token`` from its characters, and the intermediate symbols that are
    missing from the vocabulary (zett/tokenizer_converters.py:303-326; the scan order of that loop is kept)."""
    atoms = list(token)
    merges: List[str] = []
    new_symbols = set()
    while len(atoms) > 1:
        snapshot = list(atoms)
        for left, right in zip(snapshot, snapshot[1:]):
            hit = False
            i = 0
            while i < len(atoms) - 1:
                if atoms[i] == left and atoms[i + 1] == right:
                    atoms[i] = left + right
                    del atoms[i + 1]
                    hit = True
                i += 1
            if hit:
                merges.append(f"{left} {right}")
                if left + right not in known:
                    new_symbols.add(left + right)
    return merges, new_symbols


def convert_to_byte_level(tokenizer, keep_normalizer: bool = False, keep_pretokenizer: bool = False,
                          make_whitespace_consistent: bool = False, match_special_tokens_to=None):
    """Same signature and return value as the reference: ``(tokenizer, n_added_tokens | None)``; the tokenizer object is
    modified in place (its backend is replaced) and returned."""
    match_data = _tokenizer_json(match_special_tokens_to) if match_special_tokens_to is not None else {}
    data = _tokenizer_json(tokenizer)
    data.pop("added_tokens", None)  # they are part of the vocabulary below
    original = copy.deepcopy(data)
    original_length = len(tokenizer)
    keeps_indices = True

    spell, prefix = _byte_spelling(tokenizer, data)
    already_byte_level = _has_byte_level_pretokenizer(tokenizer, data)
    if prefix is not None:
        data["model"]["continuing_subword_prefix"] = ""

    own_specials = set(tokenizer.all_special_tokens)
    surface_forms = [t if t in own_specials else spell(t) for t in tokenizer.convert_ids_to_tokens(range(len(tokenizer)))]

    byte_pieces: Dict[str, str] = {}
    if data["model"].get("byte_fallback"):
        byte_pieces = {f"<0x{i:02X}>": BYTES_TO_CHARS[i] for i in range(255)}  # 0xFF is left out upstream, too
        present = set(surface_forms)
        for i, s in enumerate(surface_forms):
            if s in byte_pieces and byte_pieces[s] not in present:
                surface_forms[i] = byte_pieces[s]

    present = set(surface_forms)
    missing = [c for c in CHARS_TO_BYTES if c not in present]
    if missing:
        print(f"WARNING: {len(missing)} bytes not in surface forms.")
        surface_forms += missing

    if make_whitespace_consistent:
        wanted = _canonical_whitespace_tokens()
        for i, s in enumerate(surface_forms):
            if s in wanted:
                wanted.remove(s)
            elif _whitespace_count(s) > 1 or len(s.strip()) == 0:
                surface_forms[i] = f"<unused_whitespace__{i}>"
        surface_forms += wanted

    if match_special_tokens_to is not None:
        other_specials = set(match_special_tokens_to.all_special_tokens)
        surface_forms = [s for s in surface_forms if s not in own_specials and s not in other_specials]
        ids = match_special_tokens_to.all_special_ids
        toks = match_special_tokens_to.all_special_tokens
        for j in np.argsort(ids):
            surface_forms.insert(ids[j], toks[j])
        special_tokens = list(toks)
        keeps_indices = False
    else:
        special_tokens = list(tokenizer.all_special_tokens)

    prepend = {"type": "Prepend", "prepend": " "}
    byte_pre = {
        "type": "Sequence",
        "pretokenizers": [
            {"type": "Split", "pattern": {"Regex": SPLIT_REGEX}, "behavior": "Removed", "invert": True},
            {"type": "ByteLevel", "add_prefix_space": False, "trim_offsets": True, "use_regex": False},
        ],
    }
    if not keep_normalizer:
        data["normalizer"] = prepend
    else:
        previous = data.get("normalizer")
        data["normalizer"] = {"type": "Sequence", "normalizers": ([previous] if previous is not None else []) + [prepend]}
    if not keep_pretokenizer:
        data["pre_tokenizer"] = byte_pre
    elif not already_byte_level:
        previous = data.get("pre_tokenizer")
        byte_pre["use_regex"] = False
        data["pre_tokenizer"] = {"type": "Sequence", "pretokenizers": ([previous] if previous is not None else []) + [byte_pre]}

    model = tokenizer._tokenizer.model
    if isinstance(model, models.Unigram):
        scores = {spell(piece): score for piece, score in original["model"]["vocab"]}
        for c in CHARS_TO_BYTES:
            scores.setdefault(c, NEGATIVE_INF_FILL_VALUE)  # keeps the filled-in bytes out of every segmentation
        if make_whitespace_consistent:
            for key in [k for k in scores if _whitespace_count(k) > 1]:
                del scores[key]
        data["model"]["vocab"] = [(s, scores.get(s, 0.0)) for s in surface_forms]
    elif isinstance(model, models.BPE):
        known = set(surface_forms)
        producers: Dict[str, List[Tuple[str, str]]] = {}
        merges: List[str] = []
        for merge in data["model"]["merges"]:
            left, right = merge.split(" ") if isinstance(merge, str) else merge
            left, right = spell(left), spell(right)
            joined = left + right
            if make_whitespace_consistent and _whitespace_count(joined) > 1:
                continue
            producers.setdefault(joined, []).append((left, right))
            merges.append(f"{left} {right}")
        to_check = surface_forms[original_length:] if already_byte_level else surface_forms
        specials = set(special_tokens)
        atoms = set()
        for token in to_check:
            if token in specials or token in byte_pieces or token.startswith("<unused_whitespace__"):
                continue
            atoms.update(a for a in _undecomposable_atoms(token, producers) if len(a) > 1)
        seen_merges, before, after, new_symbols = set(), [], [], set()
        for token in sorted(atoms):  # the reference iterates a set here (hash-seed dependent order)
            extra, symbols = _merges_to_build(token, known)
            new_symbols |= symbols
            late = make_whitespace_consistent and _whitespace_count(token) > 1
            for m in extra:
                if m not in seen_merges:
                    seen_merges.add(m)
                    (after if late else before).append(m)
        surface_forms += sorted(new_symbols)
        data["model"]["vocab"] = {s: i for i, s in enumerate(surface_forms)}
        data["model"]["merges"] = before + merges + after
    elif isinstance(model, models.WordPiece):
        data["model"]["vocab"] = {s: i for i, s in enumerate(surface_forms)}
    else:
        raise ValueError(f"Unknown model type: {type(model)}")

    if match_special_tokens_to is not None and match_data.get("post_processor") is not None:
        _fix_postprocessor_ids(match_data["post_processor"], surface_forms)
        data["post_processor"] = match_data["post_processor"]

    tokenizer._tokenizer = Tokenizer.from_str(json.dumps(data))
    tokenizer._tokenizer.decoder = decoders.ByteLevel()
    if match_special_tokens_to is not None:
        for name in ("eos_token", "pad_token", "sep_token", "unk_token", "bos_token", "cls_token", "mask_token"):
            setattr(tokenizer, name, getattr(match_special_tokens_to, name))
    return tokenizer, (len(tokenizer) - original_length if keeps_indices else None)

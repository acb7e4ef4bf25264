"""Synthetic input tokenizers for the convert_to_byte_level parity tests (shared by the golden minting script, which
runs the REFERENCE on them, and by tests/test_byte_level.py, which runs zett_b200.byte_level on the same inputs)."""
from tokenizers import Tokenizer, decoders, models, normalizers, pre_tokenizers, processors
from transformers import PreTrainedTokenizerFast

SAMPLE_TEXTS = ["hello world", " the quick brown fox", "naïve café déjà vu", "a  b\n\tc", "x=1;y=22", "日本語 text", "<s> mixed </s>",
                "The the THE", "tabs\t\tand\n\nnewlines   spaces"]


def _words():
    base = ["the", "quick", "brown", "fox", "hello", "world", "na", "ve", "caf", "text", "mixed", "and", "tabs", "new", "lines",
            "spaces", "th", "he", "ll", "wor", "ld", "qu", "ick", "br", "own", "x", "y", "1", "22", "=", ";", "é", "ï", "à", "d", "j", "v", "u"]
    return base


def unigram_metaspace(byte_fallback=True):
    """sentencepiece-style Unigram: metaspace pre-tokenizer, optional <0xXX> byte-fallback pieces."""
    vocab = [("<unk>", 0.0), ("<s>", 0.0), ("</s>", 0.0)]
    if byte_fallback:
        vocab += [(f"<0x{i:02X}>", -20.0) for i in range(256)]
    score = -3.0
    for w in _words():
        vocab.append(("▁" + w, score))
        vocab.append((w, score - 1.5))
        score -= 0.137
    for ch in "abcdefghijklmnopqrstuvwxyz▁":
        if all(p != ch for p, _ in vocab):
            vocab.append((ch, -9.0))
    tok = Tokenizer(models.Unigram(vocab, unk_id=0, byte_fallback=byte_fallback))
    tok.normalizer = normalizers.Sequence([normalizers.Replace("  ", " ")])
    tok.pre_tokenizer = pre_tokenizers.Metaspace(replacement="▁", prepend_scheme="always")
    tok.decoder = decoders.Metaspace(replacement="▁", prepend_scheme="always")
    return PreTrainedTokenizerFast(tokenizer_object=tok, bos_token="<s>", eos_token="</s>", unk_token="<unk>")


def bpe_bytelevel():
    """GPT-2-style byte-level BPE with an <|endoftext|> special and an added whitespace token."""
    from zett_b200.byte_alphabet import BYTES_TO_CHARS
    alphabet = [BYTES_TO_CHARS[b] for b in range(256)]
    vocab = {c: i for i, c in enumerate(alphabet)}
    merges = []

    def add(a, b):
        merges.append((a, b))
        if a + b not in vocab:
            vocab[a + b] = len(vocab)

    for w in ["the", "quick", "hello", "world", "text", "and", "Ġthe", "Ġworld", "Ġquick", "ĠĠ", "ĊĊ", "Ġand", "ĠĠĠ"]:
        cur = w[0]
        for ch in w[1:]:
            add(cur, ch)
            cur += ch
    vocab["<|endoftext|>"] = len(vocab)
    tok = Tokenizer(models.BPE(vocab=vocab, merges=merges))
    tok.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False)
    tok.decoder = decoders.ByteLevel()
    fast = PreTrainedTokenizerFast(tokenizer_object=tok, bos_token="<|endoftext|>", eos_token="<|endoftext|>",
                                   unk_token="<|endoftext|>")
    fast.add_tokens(["    "])  # an added token made of spaces (GPT-NeoX style)
    return fast


def bpe_metaspace():
    """Llama-style BPE: no pre-tokenizer, Prepend + Replace normalizer, byte fallback, multi-byte characters."""
    vocab = {"<unk>": 0, "<s>": 1, "</s>": 2}
    for i in range(256):
        vocab[f"<0x{i:02X}>"] = len(vocab)
    chars = "▁abcdefghijklmnopqrstuvwxyzéïà=;12"
    for c in chars:
        vocab[c] = len(vocab)
    merges = []

    def add(a, b):
        merges.append((a, b))
        if a + b not in vocab:
            vocab[a + b] = len(vocab)

    for w in ["▁the", "▁quick", "hello", "▁world", "▁caf", "▁café", "▁na", "ïve", "▁text", "▁▁", "▁déjà"]:
        cur = w[0]
        for ch in w[1:]:
            add(cur, ch)
            cur += ch
    tok = Tokenizer(models.BPE(vocab=vocab, merges=merges, unk_token="<unk>", fuse_unk=True, byte_fallback=True))
    tok.normalizer = normalizers.Sequence([normalizers.Prepend("▁"), normalizers.Replace(" ", "▁")])
    tok.decoder = decoders.Sequence([decoders.Replace("▁", " "), decoders.ByteFallback(), decoders.Fuse(), decoders.Strip(" ", 1, 0)])
    tok.post_processor = processors.TemplateProcessing(single="<s> $A", pair="<s> $A <s> $B", special_tokens=[("<s>", 1)])
    return PreTrainedTokenizerFast(tokenizer_object=tok, bos_token="<s>", eos_token="</s>", unk_token="<unk>")


def wordpiece_bert():
    vocab = {"[PAD]": 0, "[UNK]": 1, "[CLS]": 2, "[SEP]": 3, "[MASK]": 4}
    for w in ["the", "quick", "hello", "world", "text", "##s", "##ing", "##ld", "wor", "he", "##llo", "a", "b", "c", "x", "y", "=", ";", "1", "22"]:
        vocab[w] = len(vocab)
    tok = Tokenizer(models.WordPiece(vocab=vocab, unk_token="[UNK]"))
    tok.normalizer = normalizers.BertNormalizer(lowercase=True)
    tok.pre_tokenizer = pre_tokenizers.BertPreTokenizer()
    tok.decoder = decoders.WordPiece()
    return PreTrainedTokenizerFast(tokenizer_object=tok, pad_token="[PAD]", unk_token="[UNK]", cls_token="[CLS]",
                                   sep_token="[SEP]", mask_token="[MASK]")


# case name -> (input builder, kwargs builder)
def cases():
    return {
        "unigram_metaspace": (lambda: unigram_metaspace(True), lambda: {}),
        "unigram_plain": (lambda: unigram_metaspace(False), lambda: {}),
        "bpe_bytelevel": (bpe_bytelevel, lambda: {}),
        "bpe_bytelevel_ws": (bpe_bytelevel, lambda: {"make_whitespace_consistent": True}),
        "bpe_metaspace": (bpe_metaspace, lambda: {}),
        "bpe_metaspace_keep": (bpe_metaspace, lambda: {"keep_normalizer": True, "keep_pretokenizer": True}),
        "wordpiece": (wordpiece_bert, lambda: {}),
        "bpe_bytelevel_match": (bpe_bytelevel, lambda: {"make_whitespace_consistent": True,
                                                         "match_special_tokens_to": bpe_metaspace()}),
        "unigram_match": (lambda: unigram_metaspace(True), lambda: {"match_special_tokens_to": wordpiece_bert()}),
    }

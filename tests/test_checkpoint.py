"""Flax msgpack ingestion: a writer of flax.serialization's format (ext type 1 ndarrays, chunked arrays, bfloat16) feeds
the reader; the name mapping reproduces the reference's PyTorch state_dict (scripts/convert_to_pt.py:35-45)."""
import json
import os

import msgpack
import numpy as np

import zett_synthetic as synthetic
from zett_b200.checkpoint import flax_params_to_state_dict, load_flax_hypernet, read_flax_msgpack


def _pack_array(a, dtype_name=None):
    payload = msgpack.packb((a.shape, dtype_name or a.dtype.name, a.tobytes("C")), use_bin_type=True)
    return msgpack.ExtType(1, payload)


def _to_flax_tree(weights, chunk_name=None):
    """Inverse of the mapping under test: PyTorch names -> Flax tree (kernel = weight.T, scale, embedding, layers_N)."""
    tree = {}
    for name, w in weights.items():
        parts = name.split(".")
        if name == "lang_embeddings.weight":
            parts = ["model", "embeddings", "lang_embedding", "embedding"]
        else:
            is_seq = parts[0] in ("input_projection", "output_projection", "output_projection_out")
            if is_seq:
                parts[1] = "layers_" + parts[1]
            leaf = parts[-1]
            if leaf == "weight":
                if "LayerNorm" in name or ".ln." in name:
                    parts[-1] = "scale"
                elif "embeddings" in name:
                    parts[-1] = "embedding"
                else:
                    parts[-1] = "kernel"
                    w = w.T
        node = tree
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        a = np.ascontiguousarray(w)
        if name == chunk_name:  # exercise the chunked-array encoding
            flat = a.reshape(-1)
            half = flat.size // 2
            node[parts[-1]] = {"__msgpack_chunked_array__": True, "shape": {str(i): s for i, s in enumerate(a.shape)},
                               "chunks": {"0": _pack_array(flat[:half]), "1": _pack_array(flat[half:])}}
        else:
            node[parts[-1]] = _pack_array(a)
    return tree


def test_roundtrip_names_and_values(tmp_path):
    cfg = synthetic.make_config("tiny_lang")
    weights = synthetic.make_weights(cfg, seed=3)
    weights = {k: v for k, v in weights.items() if "word_embeddings" not in k}
    blob = msgpack.packb(_to_flax_tree(weights, chunk_name="input_projection.1.dense1.weight"), use_bin_type=True)
    sd = flax_params_to_state_dict(read_flax_msgpack(blob))
    assert set(sd) == set(weights)
    for k in weights:
        np.testing.assert_array_equal(sd[k], weights[k]), k
    # directory form: config.json + flax_model.msgpack -> ZettHypernet with the reference's state_dict
    cfg.save_pretrained(tmp_path)
    open(os.path.join(tmp_path, "flax_model.msgpack"), "wb").write(blob)
    model = load_flax_hypernet(str(tmp_path))
    got = model.state_dict()
    for k in weights:
        np.testing.assert_array_equal(got[k].numpy(), weights[k])
    assert json.load(open(os.path.join(tmp_path, "config.json")))["model_type"] == "zett_hypernetwork"


def test_bfloat16_leaves_are_widened_exactly():
    x = np.array([1.0, -2.5, 3.140625, 1e-3], dtype=np.float32)
    bf = (x.view(np.uint32) >> 16).astype(np.uint16)  # truncation to bfloat16 bit patterns
    blob = msgpack.packb({"a": {"kernel": _pack_array(bf.reshape(2, 2), "bfloat16")}}, use_bin_type=True)
    sd = flax_params_to_state_dict(read_flax_msgpack(blob))
    want = (bf.astype(np.uint32) << 16).view(np.float32).reshape(2, 2).T
    np.testing.assert_array_equal(sd["a.weight"], want)


def test_checkpoint_with_flax_vocab_1_word_embeddings(tmp_path):
    """The Flax hypernet is built with vocab_size = 1: a checkpoint that carries the [1, H] word-embedding table (never
    read on the inputs_embeds path) must load, as the reference's converter lets it (scripts/convert_to_pt.py:35-45)."""
    cfg = synthetic.make_config("tiny")
    weights = synthetic.make_weights(cfg, seed=3)
    with_we = dict(weights)
    with_we["model.embeddings.word_embeddings.weight"] = np.zeros((1, cfg.hn_hidden_size), dtype=np.float32)
    blob = msgpack.packb(_to_flax_tree(with_we, chunk_name=None), use_bin_type=True)
    cfg.save_pretrained(tmp_path)
    open(os.path.join(tmp_path, "flax_model.msgpack"), "wb").write(blob)
    model = load_flax_hypernet(str(tmp_path))
    got = model.state_dict()
    for k in weights:
        if "word_embeddings" not in k:
            np.testing.assert_array_equal(got[k].numpy(), weights[k])


def test_msgpack_serialize_roundtrip_bias_file():
    """bias.msgpack (scripts/transfer.py:305-310) is a bare ndarray in Flax's msgpack encoding."""
    from zett_b200.checkpoint import msgpack_serialize
    b = np.linspace(-1, 1, 37, dtype=np.float32)
    back = read_flax_msgpack(msgpack_serialize(b))
    np.testing.assert_array_equal(back, b)
    tree = {"a": {"b": b, "c": np.arange(6, dtype=np.int32).reshape(2, 3)}}
    back = read_flax_msgpack(msgpack_serialize(tree))
    np.testing.assert_array_equal(back["a"]["c"], tree["a"]["c"])

"""Host-side logic that needs no GPU: the transfer driver's batching semantics (scripts/transfer.py:54-124) and the
row-sharded multi-process path (world_size 2, gloo) with the oracle standing in for the kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import hypernet_oracle as ho
import zett_synthetic as synthetic
from zett_b200 import parallel
from zett_b200.transfer import batched_inference, default_args


def _fake_predict(sf, priors=None):
    sf = np.asarray(sf, dtype=np.float32)
    return sf[:, :4] * 2.0 + 1.0, sf[:, 1:5] - 3.0, sf.sum(axis=1)


def test_batched_inference_equals_one_shot():
    rng = np.random.default_rng(0)
    sf = rng.integers(0, 1000, size=(1000, 7)).astype(np.int32)
    cfg = synthetic.make_config("tiny")
    cfg.hidden_size = 4
    for bs in (64, 256, 1000, 4096):
        got = batched_inference(sf, None, cfg, default_args(batch_size=bs), _fake_predict, rng=np.random.default_rng(1))
        for g, w in zip(got, _fake_predict(sf)):
            np.testing.assert_array_equal(g, w)
    got = batched_inference(sf, None, cfg, default_args(batch_size=128), _fake_predict, embedding_path_out=None, bias_path=None)
    assert got[1] is None and got[2] is None
    with pytest.raises(NotImplementedError):
        batched_inference(sf, None, cfg, default_args(sample_batches=True), _fake_predict)


def test_shard_bounds_cover_all_rows():
    for n in (0, 1, 7, 50257, 50304, 262144):
        for world in (1, 2, 3, 8):
            seen = 0
            for r in range(world):
                lo, hi, per = parallel.shard_bounds(n, world, r)
                assert 0 <= lo <= hi <= n and hi - lo <= per
                assert lo == min(n, r * per)
                seen += hi - lo
            assert seen == n and per * world >= n
    assert parallel.packed_width(4096, True) == 8196 and parallel.packed_width(768, False) == 772


def test_shard_plan_covers_all_rows_in_order():
    """Super-blocks of world * rows_per_pass rows: every row belongs to exactly one (rank, block), blocks tile the padded
    matrix back to back, padding only past the last row."""
    for n in (0, 1, 7, 1003, 50304, 262144):
        for world in (1, 2, 3, 8):
            for rpp in (1, 16, 4096, 16384, 10 ** 9):
                plan = parallel.shard_plan(n, world, rpp)
                owner = np.full(n, -1)
                nxt = 0
                for base, per in plan:
                    assert base == nxt and 1 <= per <= rpp
                    nxt = base + world * per
                    for r in range(world):
                        lo = min(n, base + r * per)
                        hi = min(n, lo + per)
                        assert (owner[lo:hi] == -1).all()
                        owner[lo:hi] = r
                assert (owner >= 0).all()
                assert parallel.padded_rows(n, world, rpp) == nxt and nxt >= n and nxt - n < max(1, world)
                if plan:
                    assert all(per == min(rpp, per) for _, per in plan[:-1])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_rows, rows_per_pass, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = synthetic.make_config("tiny")
    weights = synthetic.make_weights(cfg, seed=11)
    src = synthetic.make_source_embeddings(cfg, seed=12)
    sf = synthetic.make_random_surface_forms(cfg, n_rows, seed=21)
    D = cfg.n_embd

    def compute(lo, hi, block):  # the oracle stands in for the kernels; the packing / gather logic is what is tested
        a, b, c = ho.hypernet_forward(cfg, weights, sf[lo:hi], src)
        block[: hi - lo, :D] = torch.from_numpy(a)
        block[: hi - lo, D:2 * D] = torch.from_numpy(b)
        block[: hi - lo, 2 * D] = torch.from_numpy(c)

    calls = []

    def counted(lo, hi, block):
        calls.append((lo, hi))
        compute(lo, hi, block)

    pin, pout, pbias = parallel.predict_sharded(n_rows, D, True, counted, torch.device("cpu"), rows_per_pass=rows_per_pass)
    if rows_per_pass:
        assert len(calls) >= n_rows // (world * rows_per_pass)
    ret[rank] = (pin.numpy().copy(), pout.numpy().copy(), pbias.numpy().copy())
    dist.destroy_process_group()


@pytest.mark.parametrize("n_rows,rows_per_pass", [(101, None), (64, None), (101, 16), (67, 5)])
def test_row_sharded_two_ranks_gloo(n_rows, rows_per_pass):
    """world_size 2 over gloo: the sharded result (one super-block, or several with a gather each) equals the
    single-process result on both ranks, and the ranks agree bit for bit."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_rows, rows_per_pass, ret), nprocs=world, join=True)
    cfg = synthetic.make_config("tiny")
    want = ho.hypernet_forward(cfg, synthetic.make_weights(cfg, seed=11), synthetic.make_random_surface_forms(cfg, n_rows, seed=21),
                               synthetic.make_source_embeddings(cfg, seed=12))
    for rank in range(world):
        for g, w in zip(ret[rank], want):
            assert g.shape == w.shape
            # each rank computes its rows in a smaller batch; fp32 GEMM blocking may differ in the last bit
            np.testing.assert_allclose(g, w, rtol=0, atol=5e-6)
    for a, b in zip(ret[0], ret[1]):
        np.testing.assert_array_equal(a, b)


def test_checkpoint_roundtrip_and_automodel(tmp_path):
    """Reference state_dict names in, save_pretrained / from_pretrained / AutoModel out (README.md:93-117 call surface)."""
    import zett_b200
    from transformers import AutoModel
    from zett_b200.modeling_hypernet import ZettHypernet, load_weights_numpy
    cfg = synthetic.make_config("tiny_lang")
    weights = synthetic.make_weights(cfg, seed=11)
    model = load_weights_numpy(ZettHypernet(cfg), weights)
    assert set(model.state_dict()) == set(weights)
    model.save_pretrained(tmp_path)
    zett_b200.register_auto_classes()
    again = AutoModel.from_pretrained(tmp_path)
    assert isinstance(again, ZettHypernet) and again.config.hn_embed_lang_id and again.config.n_langs == 5
    for k, v in again.state_dict().items():
        np.testing.assert_array_equal(v.numpy(), weights[k])


def test_post_step_special_rows_and_splice(tmp_path):
    """scripts/transfer.py:272-304 on a tiny PyTorch GPT-2: special rows come from the source model, the predicted
    matrices become the model's (tied or untied) embeddings, vocab_size follows, the model still runs and saves."""
    from transformers import GPT2Config, GPT2LMHeadModel
    from zett_b200.transfer import overwrite_special_rows, source_embeddings_of, splice_into_model
    rng = np.random.default_rng(0)
    for tied in (True, False):
        cfg = GPT2Config(vocab_size=50, n_positions=16, n_embd=32, n_layer=1, n_head=2, tie_word_embeddings=tied)
        model = GPT2LMHeadModel(cfg)
        src_in, src_out, stacked = source_embeddings_of(model)
        assert stacked.shape == (50, 32 if tied else 64) and (src_out is None) == tied
        pred_in = rng.standard_normal((70, 32)).astype(np.float32)
        pred_out = None if tied else rng.standard_normal((70, 32)).astype(np.float32)
        prev_special, new_special = [3, 49], [60, 0]
        keep_in = src_in.numpy().copy()
        overwrite_special_rows(pred_in, pred_out, None, src_in.numpy(), None if tied else src_out.numpy(), prev_special, new_special)
        np.testing.assert_array_equal(pred_in[60], keep_in[3])
        np.testing.assert_array_equal(pred_in[0], keep_in[49])
        model = splice_into_model(model, pred_in, pred_out)
        assert model.config.vocab_size == 70
        np.testing.assert_array_equal(model.get_input_embeddings().weight.detach().numpy(), pred_in)
        want_out = pred_in if tied else pred_out
        np.testing.assert_array_equal(model.get_output_embeddings().weight.detach().numpy(), want_out)
        logits = model(torch.tensor([[1, 60, 69]])).logits
        assert logits.shape == (1, 3, 70)
        model.save_pretrained(tmp_path / ("tied" if tied else "untied"))
    t = torch.zeros((5, 4))
    overwrite_special_rows(t, None, None, torch.ones((3, 4)), None, [2], [4])
    assert t[4].sum() == 4 and t[:4].sum() == 0


def test_transfer_model_writes_bias_msgpack_when_the_model_has_no_bias(tmp_path, monkeypatch):
    """scripts/transfer.py:305-310: a base model without an output-bias parameter gets the predicted bias as bias.msgpack
    next to the saved model (Flax msgpack encoding of the bare array).  The hypernet is replaced by a stand-in here: the
    post-step is host logic."""
    from transformers import GPT2Config, GPT2LMHeadModel
    from zett_b200 import transfer
    from zett_b200.checkpoint import read_flax_msgpack
    rng = np.random.default_rng(0)
    model = GPT2LMHeadModel(GPT2Config(vocab_size=50, n_positions=16, n_embd=32, n_layer=1, n_head=2))
    hn = synthetic.make_hn_tokenizer("unigram", 600, seed=5)
    target = synthetic.make_hn_tokenizer("unigram", 400, seed=6)   # stands in for a byte-level target tokenizer
    n = len(target)
    pred = (rng.standard_normal((n, 32)).astype(np.float32), None, rng.standard_normal(n).astype(np.float32))
    monkeypatch.setattr(transfer, "make_predict", lambda hypernet, stacked, lang_index=None: (lambda sfm, priors=None: pred))

    class Base:   # the base model's tokenizer: two special tokens that also exist in the target vocabulary
        all_special_tokens = ["<s>", "</s>"]
        all_special_ids = [0, 2]

        def save_pretrained(self, path):
            pass
    hyper = type("H", (), {"config": synthetic.make_config("tiny")})()
    out = tmp_path / "out"
    _, info = transfer.transfer_model(hyper, model, Base(), target, hn, output=str(out))
    assert info["bias_written_to"] == "bias.msgpack" and info["rows"] == n
    np.testing.assert_array_equal(read_flax_msgpack(str(out / "bias.msgpack")), pred[2])
    assert (out / "config.json").exists()

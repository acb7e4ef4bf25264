"""``predict`` / ``batched_inference`` -- the inference driver calls of the reference (scripts/transfer.py:54-124,
221-234) over the B200 hypernet.

The reference closes over module-level state (``hypernet_params``, ``source_embeddings_stacked``, ``lang_index``,
``embedding_path_out``, ``bias_path``); here ``make_predict`` builds the same closure explicitly and
``batched_inference`` takes it as an argument.  Semantics are unchanged: rows are shuffled, split in ``batch_size``
chunks, the last chunk is padded by repeating row 0, every chunk's predictions are scatter-added into fp32 host
arrays and the padding is dropped.  ``predict_whole_vocab`` is the B200-first alternative: one call, the library
pipelines its own passes, one device->host copy at the end.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Callable, Optional

import numpy as np
import torch


def default_args(batch_size: int = 16384, sample_batches: bool = False, min_k: int = 1, n_samples: int = 1):
    """The fields of the reference's ``Args`` that ``batched_inference`` reads (scripts/transfer.py:30-51)."""
    return SimpleNamespace(batch_size=batch_size, sample_batches=sample_batches, min_k=min_k, n_samples=n_samples)


def make_predict(hypernet, source_embeddings_stacked, lang_index=None) -> Callable:
    """``predict(target_surface_form_matrix, target_priors)`` of scripts/transfer.py:221-234.

    ``source_embeddings_stacked`` is placed on the hypernet's CUDA device once (the reference's jit closure constant);
    each call copies the surface forms host->device through pinned memory and the three results device->host."""
    nat = hypernet.native()
    device = nat.device
    src = torch.as_tensor(source_embeddings_stacked).to(device=device, dtype=torch.float32).contiguous()
    lang = None if lang_index is None else int(lang_index)

    def predict(target_surface_form_matrix, target_priors=None):
        sf = torch.as_tensor(np.ascontiguousarray(target_surface_form_matrix, dtype=np.int32))
        sf_dev = sf.pin_memory().to(device, non_blocking=True)
        # target_priors only matter for hn_embed_target_priors hypernets, which the PyTorch reference rejects
        pred_in, pred_out, pred_bias = hypernet(sf_dev, source_embeddings=src, lang_index=lang)
        outs = []
        for t in (pred_in, pred_out, pred_bias):
            if t is None:
                outs.append(None)
                continue
            host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            host.copy_(t, non_blocking=True)
            outs.append(host)
        torch.cuda.current_stream(device).synchronize()
        return tuple(None if h is None else h.numpy() for h in outs)

    return predict


def batched_inference(target_surface_form_matrix, target_priors, config, args, predict: Callable,
                      embedding_path_out: Optional[object] = True, bias_path: Optional[object] = True, rng=None):
    """scripts/transfer.py:54-124.  ``embedding_path_out`` / ``bias_path`` keep the reference's meaning: ``None``
    disables the corresponding output array."""
    original_length = len(target_surface_form_matrix)
    if getattr(args, "sample_batches", False):
        # only meaningful for inter-token-attention hypernets, which the PyTorch reference does not support
        raise NotImplementedError("sample_batches")
    rng = np.random if rng is None else rng
    shuffled_indices = rng.permutation(original_length)
    total_length = math.ceil(original_length / args.batch_size) * args.batch_size
    padded = np.pad(shuffled_indices, (0, total_length - original_length))
    indices = np.array_split(padded, total_length // args.batch_size)
    empty_in_last_batch = total_length - original_length

    hidden = getattr(config, "hidden_size", None) or config.n_embd
    predicted_embeddings_in = np.zeros((original_length, hidden), dtype=np.float32)
    predicted_embeddings_out = np.zeros((original_length, hidden), dtype=np.float32) if embedding_path_out is not None else None
    predicted_bias = np.zeros(original_length, dtype=np.float32) if bias_path is not None else None
    if target_priors is None:
        target_priors = np.zeros(original_length, dtype=np.float32)

    for i, batch_indices in enumerate(indices):
        last_batch = i == len(indices) - 1
        in_b, out_b, bias_b = predict(target_surface_form_matrix[batch_indices], target_priors[batch_indices])
        if last_batch and empty_in_last_batch > 0:
            batch_indices = batch_indices[:-empty_in_last_batch]
            in_b = in_b[:-empty_in_last_batch]
            if predicted_embeddings_out is not None:
                out_b = out_b[:-empty_in_last_batch]
            if predicted_bias is not None:
                bias_b = bias_b[:-empty_in_last_batch]
        predicted_embeddings_in[batch_indices] += in_b
        if predicted_embeddings_out is not None:
            predicted_embeddings_out[batch_indices] += out_b
        if predicted_bias is not None:
            predicted_bias[batch_indices] += bias_b
    return predicted_embeddings_in, predicted_embeddings_out, predicted_bias


def predict_whole_vocab(hypernet, target_surface_form_matrix, source_embeddings_dev, lang_index=None):
    """One call for the whole vocabulary: pinned H2D of the int32 matrix, all passes queued back to back on the
    current stream, pinned D2H of the results.  Returns numpy arrays ``(in, out | None, bias)``."""
    return make_predict(hypernet, source_embeddings_dev, lang_index)(target_surface_form_matrix)


# ----------------------------------------------------------------------------------------------------------------------
# Post-step of the transfer driver (scripts/transfer.py:272-328): special-token rows keep the source model's
# embeddings, the predicted matrices replace the base model's input / output embeddings (and output bias), the model
# is saved.  The reference does this on Flax parameter trees (IN/OUT_EMBEDDING_PATHS, zett/model/__init__.py:15-41) and
# converts to PyTorch afterwards (--save_pt); here it is done on the PyTorch model directly.
# ----------------------------------------------------------------------------------------------------------------------
def overwrite_special_rows(predicted_in, predicted_out, predicted_bias, source_in, source_out, previous_special_ids,
                           new_special_ids):
    """``predicted[new_special_ids] = source[previous_special_ids]`` (scripts/transfer.py:274-302); in place.

    Accepts numpy arrays or torch tensors (host or device).  ``predicted_bias`` is left as predicted, like the reference."""
    prev = torch.as_tensor(np.asarray(previous_special_ids), dtype=torch.long)
    new = torch.as_tensor(np.asarray(new_special_ids), dtype=torch.long)
    if prev.numel() != new.numel():
        raise ValueError("special-token id lists differ in length")

    def put(pred, src):
        if pred is None:
            return None
        if isinstance(pred, np.ndarray):
            pred[new.numpy()] = np.asarray(src)[prev.numpy()]
            return pred
        pred[new.to(pred.device)] = torch.as_tensor(src)[prev].to(pred.device, pred.dtype)
        return pred

    return put(predicted_in, source_in), put(predicted_out, source_out), predicted_bias


def source_embeddings_of(model):
    """``(source_in [V0, D], source_out [V0, D] | None, stacked [V0, E])`` of a PyTorch HF model, the way the reference
    stacks them (scripts/transfer.py:162-191): output embeddings are taken only when the model does not tie them."""
    emb_in = model.get_input_embeddings().weight.detach().float()
    out_layer = model.get_output_embeddings() if hasattr(model, "get_output_embeddings") else None
    tied = bool(getattr(model.config, "tie_word_embeddings", True))
    if out_layer is None or tied:
        return emb_in, None, emb_in
    emb_out = out_layer.weight.detach().float()
    return emb_in, emb_out, torch.cat([emb_in, emb_out], dim=1)


def splice_into_model(model, predicted_in, predicted_out=None, predicted_bias=None):
    """Replace the base model's vocabulary-sized parameters by the predicted ones (scripts/transfer.py:287-304):
    input embeddings <- predicted_in, untied output embeddings <- predicted_out, output bias <- predicted_bias when
    the head has one.  ``config.vocab_size`` follows (``:272``).  Returns the model."""
    pred_in = torch.as_tensor(predicted_in)
    n, d = pred_in.shape
    old_in = model.get_input_embeddings()
    if old_in.weight.shape[1] != d:
        raise ValueError("embedding width %d does not match the model's %d" % (d, old_in.weight.shape[1]))
    dtype, device = old_in.weight.dtype, old_in.weight.device
    new_in = torch.nn.Embedding(n, d, padding_idx=None, dtype=dtype, device=device)
    with torch.no_grad():
        new_in.weight.copy_(pred_in.to(device=device, dtype=dtype))
    model.set_input_embeddings(new_in)
    out_layer = model.get_output_embeddings() if hasattr(model, "get_output_embeddings") else None
    tied = bool(getattr(model.config, "tie_word_embeddings", True))
    if out_layer is not None:
        if tied:
            new_out = torch.nn.Linear(d, n, bias=out_layer.bias is not None, dtype=dtype, device=device)
            new_out.weight = new_in.weight
        else:
            if predicted_out is None:
                raise ValueError("the model has untied output embeddings: predicted_out is required")
            new_out = torch.nn.Linear(d, n, bias=out_layer.bias is not None, dtype=dtype, device=device)
            with torch.no_grad():
                new_out.weight.copy_(torch.as_tensor(predicted_out).to(device=device, dtype=dtype))
        if new_out.bias is not None:
            with torch.no_grad():
                if predicted_bias is not None:
                    new_out.bias.copy_(torch.as_tensor(predicted_bias).to(device=device, dtype=dtype))
                else:
                    new_out.bias.zero_()
        model.set_output_embeddings(new_out)
    model.config.vocab_size = n
    return model


def transfer_model(hypernet, base_model, base_tokenizer, target_tokenizer, hn_tokenizer, lang_index=None, output=None,
                   batch_size: Optional[int] = None):
    """The whole driver for a PyTorch base model: surface forms of the (already byte-level) target tokenizer ->
    prediction -> special rows -> splice -> optional save.  Mirrors ``scripts/transfer.py:204-328``.

    ``target_tokenizer`` must already be byte-level with its special tokens matched to ``base_tokenizer`` (the
    reference's ``convert_to_byte_level(..., match_special_tokens_to=base_tokenizer)``)."""
    from .surface_forms import get_surface_form_matrix

    cfg = hypernet.config
    source_in, source_out, stacked = source_embeddings_of(base_model)
    sfm, n_truncated = get_surface_form_matrix(target_tokenizer, cfg.hn_surface_maxlen, hn_tokenizer)
    predict = make_predict(hypernet, stacked, lang_index)
    if batch_size:
        cfg_like = SimpleNamespace(hidden_size=cfg.n_embd, n_embd=cfg.n_embd)
        pred_in, pred_out, pred_bias = batched_inference(sfm, None, cfg_like, default_args(batch_size=batch_size), predict,
                                                         embedding_path_out=True if source_out is not None else None)
    else:
        pred_in, pred_out, pred_bias = predict(sfm)
    if source_out is None:
        pred_out = None
    vocab = target_tokenizer.get_vocab()
    new_ids = [vocab[t] for t in base_tokenizer.all_special_tokens]
    overwrite_special_rows(pred_in, pred_out, pred_bias, source_in.cpu().numpy(),
                           None if source_out is None else source_out.cpu().numpy(), base_tokenizer.all_special_ids, new_ids)
    model = splice_into_model(base_model, pred_in, pred_out, pred_bias)
    has_bias_param = output_bias_of(model) is not None
    if output is not None:
        import os
        os.makedirs(output, exist_ok=True)
        base_tokenizer.save_pretrained(output)   # tokenizer_config.json and other metadata (transfer.py:281-283)
        target_tokenizer.save_pretrained(output)
        model.save_pretrained(output)
        if not has_bias_param and pred_bias is not None:
            # the base model has no output-bias parameter (BIAS_PATHS has no entry): the predicted bias is kept next to
            # the model in Flax's msgpack encoding, as the reference does (scripts/transfer.py:305-310)
            from .checkpoint import msgpack_serialize
            with open(os.path.join(output, "bias.msgpack"), "wb") as f:
                f.write(msgpack_serialize(np.asarray(pred_bias, dtype=np.float32)))
    return model, dict(n_truncated=n_truncated, rows=len(sfm), bias_written_to="model" if has_bias_param else "bias.msgpack")


def output_bias_of(model):
    """The output-embedding bias parameter of a PyTorch causal LM, or None (the reference's BIAS_PATHS,
    zett/model/__init__.py:35-41, lists the architectures that have one)."""
    out_layer = model.get_output_embeddings() if hasattr(model, "get_output_embeddings") else None
    return None if out_layer is None else getattr(out_layer, "bias", None)


# ----------------------------------------------------------------------------------------------------------------------
# Pipelined end-to-end path: token strings on the host -> predicted matrices in pinned host memory.
# The vocabulary is cut into passes; while the GPU runs pass k the host retokenises pass k + 1, and the device->host
# copy of pass k runs on a side stream under the compute of pass k + 1.  Same results as the one-shot calls (rows are
# independent); only the overlap differs.
# ----------------------------------------------------------------------------------------------------------------------
class TokenPipeline:
    """Reusable buffers + streams for ``predict_from_tokens`` (one instance per hypernet / device)."""

    def __init__(self, hypernet_or_native, hn_tokenizer, source_embeddings_dev, lang_index=None, rows_per_pass: int = 16384,
                 first_chunk_rows=None):
        from .surface_forms import native_model_for
        self.nat = hypernet_or_native.native() if hasattr(hypernet_or_native, "native") else hypernet_or_native
        self.cfg = self.nat.cfg
        self.device = self.nat.device
        self.src = source_embeddings_dev
        self.lang = -1 if (lang_index is None or not self.cfg.hn_embed_lang_id) else int(lang_index)
        self.rows_per_pass = int(rows_per_pass)
        # the GPU idles while the very first pass is retokenised (nothing to overlap it with): that pass is cut in two so
        # that the forward starts after `first_chunk_rows` tokens -- eight ranks sharing one host have two retokenizer
        # threads each, and 16 384 tokens take them ~10 ms
        # (measured on one GPU: with two threads the cut saves 2.4 ms per 50k-token step; with sixteen threads the pass is
        # retokenised in ~1.5 ms anyway and the cut COSTS 4 ms, because de-duplication works per pass -- so it is only made
        # when the retokenizer is short of threads)
        if first_chunk_rows is None:
            from .surface_forms import default_threads
            first_chunk_rows = 4096 if default_threads() <= 4 else 0
        self.first_chunk_rows = int(first_chunk_rows)
        self.tok_model = native_model_for(hn_tokenizer)
        self.pad_id = int(hn_tokenizer.pad_token_id)
        self.special = {t: int(hn_tokenizer.convert_tokens_to_ids(t)) for t in hn_tokenizer.all_special_tokens}
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._cap = 0

    def _ensure(self, n_rows: int, width: int):
        if n_rows <= self._cap:
            return
        L = self.cfg.hn_surface_maxlen
        self.sf_pinned = torch.empty((n_rows, L), dtype=torch.int32, pin_memory=True)
        self.sf_dev = torch.empty((n_rows, L), dtype=torch.int32, device=self.device)
        self.block = torch.zeros((n_rows, width), dtype=torch.float32, device=self.device)
        self.out_pinned = torch.empty((n_rows, width), dtype=torch.float32, pin_memory=True)
        self._cap = n_rows

    def run(self, tokens, after_compute=None, comm=None, full=None, plan=None, rank: int = 0, side=None):
        """Returns ``(out_pinned[:n] as [n, n_out * D + 4] fp32, surface_forms [n, L] int32, n_truncated)``; columns
        are ``pred_in | pred_out | bias`` (``zett_b200.parallel.unpack`` splits them).

        Multi-GPU use: ``plan`` = ``parallel.shard_plan(world * n, world, rows_per_pass)``, ``full`` = the padded full matrix
        on this device, ``comm`` = a ``parallel.NativeComm``, ``side`` = a CUDA stream: pass k is computed straight into this
        rank's slot of super-block k and the in-place all-gather of that super-block runs on ``side`` under the compute of
        pass k + 1.  ``after_compute(block)`` (single-GPU callers) runs on the compute stream after the last pass.

        Like ``ZettHypernet.forward`` the call raises ``IndexError`` for an out-of-range id in ANY pass (the library's flag
        is sticky across passes) and repeats itself with the bf16 operand split when a value left fp16's range
        (single-process use only: a collective must not be repeated by one rank alone)."""
        result = {}

        def enqueue():
            result["v"] = self._run_once(tokens, after_compute, comm, full, plan, rank, side)

        self.nat.run_checked(enqueue, allow_fallback=after_compute is None and comm is None)
        self.copy_stream.synchronize()
        return result["v"]

    def _run_once(self, tokens, after_compute=None, comm=None, full=None, plan=None, rank=0, side=None):
        from .parallel import packed_width
        cfg = self.cfg
        n, D, separate = len(tokens), cfg.n_embd, bool(cfg.separate_out_embeddings)
        width = packed_width(D, separate)
        self._ensure(n, width)
        stream = torch.cuda.current_stream(self.device)
        if plan is None:   # single GPU: a "super-block" is one pass of this rank
            plan = [(lo, min(self.rows_per_pass, n - lo)) for lo in range(0, n, self.rows_per_pass)]
            full, rank = self.block, 0
        world = 1 if comm is None else comm.world
        n_trunc, loc = 0, 0
        for base, per in plan:
            n_here = min(per, n - loc)
            if n_here <= 0:
                break
            lo, hi = loc, loc + n_here
            blk = full[base + rank * per: base + rank * per + n_here]
            c = self.first_chunk_rows
            parts = [(lo, lo + c), (lo + c, hi)] if (loc == 0 and 0 < c and 2 * c <= n_here) else [(lo, hi)]
            for a, b in parts:
                sf, nt = self.tok_model.surface_forms(tokens[a:b], cfg.hn_surface_maxlen, self.pad_id, special_tokens=self.special)  # host
                n_trunc += nt
                self.sf_pinned[a:b].numpy()[...] = sf
                self.sf_dev[a:b].copy_(self.sf_pinned[a:b], non_blocking=True)                              # H2D
                part = blk[a - lo: b - lo]
                self.nat.forward_into(self.sf_dev[a:b], self.src, self.lang, part[:, 0:], part[:, D:] if separate else None,
                                      part[:, (2 if separate else 1) * D:], ld_pred=width, ld_bias=width)
            ev = torch.cuda.Event()
            ev.record(stream)
            if comm is not None and world > 1:                                                              # all-gather
                side.wait_event(ev)
                comm.allgather_rows(full[base: base + world * per], per, stream=side)
            with torch.cuda.stream(self.copy_stream):                                                       # D2H
                self.copy_stream.wait_event(ev)
                self.out_pinned[lo:hi].copy_(blk, non_blocking=True)
            loc = hi
        if comm is not None and world > 1:
            stream.wait_stream(side)
            comm.barrier(stream)   # peer copies complete locally; past this point every rank's rows have landed everywhere
        if after_compute is not None:
            after_compute(self.block[:n])
        return self.out_pinned[:n], self.sf_pinned[:n].numpy(), n_trunc


def predict_from_tokens(hypernet, tokens, hn_tokenizer, source_embeddings_dev, lang_index=None, rows_per_pass: int = 16384):
    """Token strings (byte-level spellings) -> ``(pred_in, pred_out | None, pred_bias)`` numpy arrays, pipelined."""
    from .parallel import unpack
    pipe = TokenPipeline(hypernet, hn_tokenizer, source_embeddings_dev, lang_index, rows_per_pass)
    out, sf, n_trunc = pipe.run(list(tokens))
    cfg = pipe.cfg
    pin, pout, pbias = unpack(out, len(tokens), cfg.n_embd, bool(cfg.separate_out_embeddings))
    return pin.numpy(), (None if pout is None else pout.numpy()), pbias.numpy()

/*
 * zett_b200 -- C ABI of the B200-native ZeTT embedding-prediction hot path (libzett_b200.so).
 *
 * The reference (bminixhofer/zett) has no FFI on this path: its boundary is a Python call surface.  Every entry
 * point below names the reference interface it replaces; `INTEGRATION.md` shows the ctypes stubs a maintainer of the
 * reference would add.  Conventions:
 *   - plain pointers and sizes only (no torch types); every pointer is BORROWED -- the caller keeps the memory
 *     alive until the CUDA stream it passed has been synchronised;
 *   - every function returns 0 on success or a negative zett_status; nothing throws across the ABI; the message of
 *     the last failure on the calling thread is available from zett_last_error();
 *   - a handle is bound to the CUDA device that was current when it was created and is not thread-safe
 *     (one handle per device, one caller at a time); all device work is enqueued on the stream passed in.
 */
#ifndef ZETT_B200_H_
#define ZETT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZETT_B200_ABI_VERSION 2

typedef enum zett_status {
  ZETT_OK = 0,
  ZETT_ERR_INVALID = -1,       /* bad argument / shape / name                                   (ValueError)          */
  ZETT_ERR_UNSUPPORTED = -2,   /* a branch the reference raises NotImplementedError for         (NotImplementedError) */
  ZETT_ERR_CUDA = -3,          /* CUDA runtime / driver failure, incl. "no device"              (RuntimeError)        */
  ZETT_ERR_STATE = -4,         /* call order (forward before finalize, missing weight ...)      (RuntimeError)        */
  ZETT_ERR_INDEX = -5,         /* a surface-form id outside [0, original_vocab_size + n_extra)  (IndexError)          */
  ZETT_ERR_KEY = -6,           /* a token char outside the 256-char byte alphabet               (KeyError)            */
  ZETT_ERR_MISSING_UNK = -7,   /* Unigram needed <unk> but the model has no unk id              (Exception)           */
  ZETT_ERR_RANGE = -8          /* a GEMM operand left fp16's range under split_terms = 2: rerun
                                  after zett_hn_set_split_terms(h, 3)                            (ArithmeticError)     */
} zett_status;

const char* zett_last_error(void);
int zett_abi_version(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Hypernetwork forward.
 * Replaces ZettHypernet.__init__ / __call__   (reference hf_hypernet/modeling_hypernet.py:46-154, 156-267)
 *      ==  Hypernet.setup / __call__          (reference zett/model/__init__.py:217-346, 387-469).
 * ---------------------------------------------------------------------------------------------------------------- */

/* Mirrors ZettHypernetConfig (reference hf_hypernet/configuration_hypernet.py:4-56) plus the fields training writes
 * onto it (train.py:295,314,350,361).  Booleans are 0/1 ints. */
typedef struct zett_hn_config {
  int32_t struct_bytes;                    /* sizeof(zett_hn_config), ABI guard                                    */
  int32_t hn_surface_maxlen;               /* L                                                                    */
  int32_t hn_n_layers;
  int32_t n_embd;                          /* D                                                                    */
  int32_t hn_hidden_size;                  /* H                                                                    */
  int32_t hn_intermediate_size;            /* I                                                                    */
  int32_t hn_num_attention_heads;          /* 0 -> H / 64        (modeling_hypernet.py:73-75)                      */
  int32_t hn_rescale_embeddings;
  int32_t hn_embed_target_priors;          /* must be 0          (modeling_hypernet.py:85-89)                      */
  int32_t hn_add_inter_token_attention;    /* must be 0          (modeling_hypernet.py:85-89)                      */
  int32_t hn_embed_using_source_embeddings;/* must be 1          (modeling_hypernet.py:167-168)                    */
  int32_t hn_concat_last_hidden_state;     /* must be 0          (shape-inconsistent in the reference)             */
  int32_t hn_single_head;
  int32_t hn_predict_bias;
  int32_t hn_embed_lang_id;
  int32_t hn_model_type_is_roberta;        /* must be 1          (modeling_hypernet.py:78-79)                      */
  int32_t n_langs;
  int32_t pad_token_id;
  int32_t original_vocab_size;             /* V0                                                                   */
  int32_t hn_n_extra_tokens;
  int32_t separate_out_embeddings;
  int32_t max_position_embeddings;         /* rows of model.embeddings.position_embeddings (514 for roberta-base)  */
  float encoder_layer_norm_eps;            /* 1e-5 (roberta-base)                                                  */
  /* execution knobs, not model semantics */
  int32_t max_rows_per_pass;               /* rows handled by one pass of the kernels (0 -> 16384, transfer.py:44) */
  int32_t gemm_impl;                       /* 0 = auto (5), 2 = tcgen05 CTA pairs (cta_group::2) on 256 x 256 tiles,
                                              3 = SIMT fp32 debug kernel (checker, never the default),
                                              5 = CTA pairs on 256 x 512 tiles (N a multiple of 512, else as 2)    */
  int32_t split_terms;                     /* operand precision: 2 = fp16 MMA + two e5m2 correction MMAs at fp8 rate
                                              (operands must stay inside fp16's range: zett_hn_check reports
                                              ZETT_ERR_RANGE otherwise), 3 = three bf16 MMA terms (A0W0 + A1W0 + A0W1,
                                              fp32's range), 0 = auto (2; the Python wrapper switches to 3 and reruns
                                              on ZETT_ERR_RANGE), 1 = one bf16 pass (misses the 1e-3 parity budget;
                                              for comparison only, announced on stderr)                            */
} zett_hn_config;

typedef struct zett_hn zett_hn;

/* dtype codes for zett_hn_set_weight */
#define ZETT_F32 0
#define ZETT_F16 1
#define ZETT_BF16 2

int zett_hn_create(const zett_hn_config* cfg, zett_hn** out);

/* `name` is a key of the reference's PyTorch state_dict (e.g. "model.encoder.layer.0.attention.self.query.weight",
 * "input_projection.1.dense1.weight", "in_scaler.w"; full list in SURVEY.md section 8b).  `data` may be a host or a
 * device pointer; it is copied before the call returns.  Unused reference keys ("model.embeddings.word_embeddings.
 * weight", "*.position_ids", "*.token_type_ids") are accepted and ignored. */
int zett_hn_set_weight(zett_hn* h, const char* name, const void* data, int dtype, int ndim, const int64_t* shape);

/* Checks that every weight the config needs is present and writes the Linear weights in the operand format the
 * tensor-core kernels consume (the fp32 originals stay on the device for zett_hn_set_split_terms). */
int zett_hn_finalize(zett_hn* h);

/* Bytes of device workspace a forward of `n_rows` rows uses (allocated lazily by the library, reused across calls). */
size_t zett_hn_workspace_bytes(const zett_hn* h, int64_t n_rows);

/* pred_in[n_rows, D], pred_out[n_rows, D] (NULL unless separate_out_embeddings), pred_bias[n_rows]  <-
 *     ZettHypernet.__call__(target_surface_forms[n_rows, L] int32, source_embeddings[v0_rows, E] fp32, lang_index)
 * All five pointers are DEVICE pointers, row-major.  lang_index < 0 means None.
 * ld_pred = elements between consecutive rows of pred_in / pred_out (0 -> D, i.e. contiguous; otherwise a multiple
 * of 4 and >= D), ld_bias = elements between consecutive pred_bias entries (0 -> 1): a caller that row-shards the
 * vocabulary lets all three land in one [n_rows, 2D + 4] block so that a single all-gather moves them.
 * Work is enqueued on `cuda_stream` (a cudaStream_t; NULL = legacy default stream); the call does not synchronise. */
int zett_hn_forward(zett_hn* h, const int32_t* surface_forms_dev, int64_t n_rows, const float* source_emb_dev,
                    int64_t v0_rows, int32_t lang_index, float* pred_in_dev, float* pred_out_dev,
                    float* pred_bias_dev, int64_t ld_pred, int64_t ld_bias, void* cuda_stream);

/* Synchronises `cuda_stream` and reports what the kernels recorded since the previous check (the flags are sticky over
 * any number of zett_hn_forward calls, and cleared here): ZETT_ERR_INDEX if any surface-form id was outside
 * [0, V0 + max(n_extra, 1)) (the reference would raise an index error), ZETT_ERR_RANGE if a GEMM operand left fp16's
 * range under split_terms = 2, ZETT_ERR_CUDA on a kernel fault. */
int zett_hn_check(zett_hn* h, void* cuda_stream);

/* Switch the operand format of a finalized handle (1, 2 or 3 as in zett_hn_config.split_terms): the Linear weights are
 * rewritten from their fp32 originals and the workspace is re-sized on the next forward.  Synchronises the device. */
int zett_hn_set_split_terms(zett_hn* h, int split_terms);

/* Execution statistics of ALL forwards enqueued between the two most recent zett_hn_check calls (a caller that runs one
 * forward per pass, or several steps, reads the totals after one check): kernels launched, packed (non-pad) positions,
 * rows, GEMM FLOPs and -- with zett_hn_set_timing -- the summed device time of the GEMM launches. */
typedef struct zett_hn_stats {
  int64_t kernel_launches;
  int64_t rows;
  int64_t packed_positions;    /* valid after zett_hn_check (read back from the device)                            */
  int64_t encoder_positions;
  double flops_executed;       /* algorithmic FLOPs of the GEMMs actually issued (one count per product term set)  */
  double gemm_ms;              /* summed CUDA-event time of the GEMM kernel launches (only with zett_hn_set_timing) */
  int64_t gemm_launches;
  int64_t distinct_ids;        /* distinct surface-form ids summed over the passes (input projection runs per id)  */
  int64_t distinct_pairs;      /* distinct (id, position) pairs summed over the passes (first encoder layer's
                                  LayerNorm and query/key/value GEMM run per pair); 0 when that is switched off    */
  int64_t split_terms;         /* the operand format in effect (zett_hn_config.split_terms after "auto")           */
  int64_t gemm_impl;           /* the GEMM implementation in effect (zett_hn_config.gemm_impl after "auto")         */
  int64_t operand_overflows;   /* warps that produced a GEMM operand outside fp16's range since the previous
                                  zett_hn_check (0 unless that check returned ZETT_ERR_RANGE)                       */
} zett_hn_stats;
int zett_hn_get_stats(zett_hn* h, zett_hn_stats* out);

/* enable != 0: bracket every GEMM launch with CUDA events on the caller's stream; zett_hn_check then fills
 * gemm_ms / gemm_launches of the forwards since the previous check (the roofline figure of bench.py). */
int zett_hn_set_timing(zett_hn* h, int enable);

void zett_hn_destroy(zett_hn* h);

/* Stand-alone entry to the GEMM engine (used by the unit tests and the micro-benchmark):
 *   out[m, n] = act(A[m, k] . W[n, k]^T + bias[n])     A, W, out fp32 device pointers, row-major.
 * impl / split_terms as in zett_hn_config; act: 0 none, 1 gelu(tanh), 2 gelu(erf). `elapsed_ms` (nullable) receives
 * the CUDA-event time of `iters` back-to-back launches of the GEMM kernel alone (operands already split). */
int zett_gemm_f32(const float* a_dev, const float* w_dev, const float* bias_dev, float* out_dev, int64_t m, int64_t n,
                  int64_t k, int act, int impl, int split_terms, int iters, float* elapsed_ms, void* cuda_stream);

/* The same with the whole epilogue: y = act(A W^T + bias) + residual[m, n];  y = col_scale[n] * y + col_shift[n]
 * (every pointer after w_dev nullable).  out_operand_dev (nullable, fp32 [m, n]) receives the values DECODED from the
 * operand-line output the kernel writes for the next GEMM (main plane + correction plane), so that a test can hold
 * them against out_dev.  `report` (nullable, report_cap bytes) receives a JSON object with the per-role stall cycles
 * of the last launch when the environment has ZETT_GEMM_PROF=1, else an empty string. */
int zett_gemm_f32_ex(const float* a_dev, const float* w_dev, const float* bias_dev, const float* residual_dev,
                     const float* col_scale_dev, const float* col_shift_dev, float* out_dev, float* out_operand_dev,
                     int64_t m, int64_t n, int64_t k, int act, int impl, int split_terms, int iters, float* elapsed_ms,
                     char* report, int64_t report_cap, void* cuda_stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Multi-GPU: one process per GPU, row blocks of the vocabulary per rank, the predicted rows of all ranks assembled on
 * every rank by ncclAllGather (NVLink / NVSwitch).
 * Replaces the reference's sharding of an inference batch over the local devices
 *      SHARDING = PositionalSharding(jax.local_devices())                 (reference zett/utils.py:26)
 *      jax.device_put(batch, SHARDING.reshape((-1, 1))) ... device_get    (reference scripts/transfer.py:90-91, 105-111)
 * with replicated hypernet parameters and source-embedding table.  NCCL is bound at run time (dlopen libnccl.so.2,
 * preferring the copy already in the process), so single-GPU callers never need it.
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct zett_comm zett_comm;

/* Rank 0 fills 128 bytes (ncclUniqueId) and hands them to the other ranks by whatever means the host application has
 * (MPI, a file, torch.distributed.broadcast_object_list ...). */
int zett_comm_unique_id(void* out_id_128_bytes);

/* Collective over all `world` ranks, each on its own process with its CUDA device current.  world == 1 needs no id
 * (and no NCCL). */
int zett_comm_init(int rank, int world, const void* nccl_unique_id, zett_comm** out);

/* full_dev[world * rows_per_rank, row_elems] <- concatenation over ranks of shard_dev[rows_per_rank, row_elems]
 * (fp32, device pointers, row-major, every rank the same sizes), enqueued on `cuda_stream`; no synchronisation.
 * shard_dev may alias this rank's slot of full_dev (in-place all-gather).  world == 1: shard_dev must be full_dev. */
int zett_allgather_rows(zett_comm* c, const float* shard_dev, int64_t rows_per_rank, int64_t row_elems, float* full_dev,
                        void* cuda_stream);

/* Peer-copy transport.  A full matrix that every rank holds at the same size can be REGISTERED: each rank describes its
 * buffer with zett_comm_ipc_handle (a 64-byte CUDA IPC handle + the buffer's offset inside its allocation), the host
 * application passes everybody's 64-byte handles and offsets to zett_comm_register (collective), and from then on
 * zett_allgather_rows into that buffer does not launch a kernel: every rank pushes its rows into its slot of every peer's
 * buffer with asynchronous peer copies on `cuda_stream` (copy engines over NVLink / NVSwitch, no SM).  The pushes of a
 * rank complete in its own stream order; zett_comm_barrier (a one-element ncclAllReduce on `cuda_stream`) completes when
 * every rank has reached it, i.e. when everything pushed before it on every rank has landed.  Unregister before freeing
 * the buffer. */
int zett_comm_ipc_handle(const void* dev_ptr, void* out_handle_64_bytes, int64_t* out_offset);
int zett_comm_register(zett_comm* c, void* full_dev, int64_t bytes, const void* all_handles, const int64_t* all_offsets);
int zett_comm_unregister(zett_comm* c);
int zett_comm_barrier(zett_comm* c, void* cuda_stream);

/* rank / world of the communicator and the NCCL version in use (0 when world == 1); any pointer may be NULL */
int zett_comm_info(const zett_comm* c, int* rank, int* world, int* nccl_version);

void zett_comm_destroy(zett_comm* c);

/* ------------------------------------------------------------------------------------------------------------------
 * Surface forms.
 * Replaces get_surface_form_matrix                  (reference zett/utils.py:651-689)
 * and the tokenizers.models.{Unigram,BPE}.tokenize call it makes per token (zett/utils.py:681; third-party Rust).
 * Host-side, multi-threaded C++; strings are UTF-8, NUL-terminated.
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct zett_tok zett_tok;

/* Unigram model: pieces[n] with scores[n]; unk_id < 0 = None. */
int zett_tok_create_unigram(const char* const* pieces, const double* scores, int64_t n, int64_t unk_id,
                            int byte_fallback, zett_tok** out);

/* BPE model: vocab[n] (index = id); merges as m pairs of ids (left, right) in rank order.  unk_id < 0 = None.
 * continuing_subword_prefix / end_of_word_suffix may be NULL. */
int zett_tok_create_bpe(const char* const* vocab, int64_t n, const int32_t* merges, int64_t m, int64_t unk_id,
                        const char* continuing_subword_prefix, const char* end_of_word_suffix, int fuse_unk,
                        int byte_fallback, int ignore_merges, zett_tok** out);

/* ids of one token string (Model.tokenize); returns the number of ids (may exceed `cap`; only `cap` are written)
 * or a negative zett_status. */
int64_t zett_tok_tokenize(const zett_tok* t, const char* token, int32_t* out_ids, int64_t cap);

/* out[v + padding, maxlen] int32 (host), pre-filled by the callee with pad_id.
 * special_ids[i] >= 0 marks tokens[i] as one of the hn tokenizer's special tokens with that id (written to column 0);
 * special_ids may be NULL.  Tokens holding a char outside the 256-char byte alphabet fail with ZETT_ERR_KEY
 * (the reference raises KeyError at zett/utils.py:675).  n_threads <= 0 -> hardware concurrency. */
int zett_surface_forms(const zett_tok* t, const char* const* tokens, int64_t v, const int32_t* special_ids,
                       int32_t maxlen, int32_t pad_id, int64_t padding, int32_t* out, int64_t* n_truncated,
                       int n_threads);

/* Same result from ONE buffer: `tokens_blob` holds the v token strings back to back, each terminated by a NUL byte
 * (so `"\0".join(tokens)` plus the terminator every C string carries); `special_blob` / `special_token_ids` list the hn
 * tokenizer's special tokens (hn_tokenizer.all_special_tokens, zett/utils.py:671-673) the same way, n_special of them,
 * and the match against them happens inside.  This is the entry point a binding should prefer for a whole vocabulary:
 * building an array of 50k char* in the host language costs more than the retokenisation itself. */
int zett_surface_forms_blob(const zett_tok* t, const char* tokens_blob, int64_t blob_bytes, int64_t v,
                            const char* special_blob, int64_t special_bytes, const int32_t* special_token_ids,
                            int64_t n_special, int32_t maxlen, int32_t pad_id, int64_t padding, int32_t* out,
                            int64_t* n_truncated, int n_threads);

void zett_tok_destroy(zett_tok* t);

/* ------------------------------------------------------------------------------------------------------------------
 * TokenizerSampler (training side; SURVEY section 8f, last row).
 * Replaces rust_utils.TokenizerSampler.sample_tokenizer          (reference rust_utils/src/lib.rs:69-250),
 * called by the training collator                                  (reference zett/collator.py:341-452).
 * Host-side C++: GPT-2 pre-tokenisation (Unicode-aware), byte-level substring scores, the seed cache of the last
 * batches and the seed list of a Unigram tokenizer.  The reference draws its noise from an unseeded generator and
 * iterates hash maps; here the noise takes a seed (none is drawn at noise_std = 0), the byte alphabet comes in byte order
 * and ties in the score order are broken by the piece's bytes.
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct zett_sampler zett_sampler;

int zett_sampler_create(zett_sampler** out);

/* texts_blob: n_texts UTF-8 strings back to back, NUL-terminated; counts[n_texts] their frequencies (the reference's
 * HashMap<String, u32>).  seed_size, max_length, stride, noise_std, pop_prev, push_current as in the reference
 * (defaults there: stride 1, noise_std 0, pop_prev / push_current true).  On success *out_pieces_blob (out_n strings back to
 * back, NUL-terminated, *out_blob_bytes long) and *out_scores (out_n log-probabilities) are malloc'd by the callee and
 * released with zett_sampler_free. */
int zett_sampler_sample(zett_sampler* s, const char* texts_blob, int64_t blob_bytes, const uint32_t* counts, int64_t n_texts,
                        int64_t seed_size, int64_t max_length, int64_t stride, double noise_std, uint64_t noise_seed,
                        int pop_prev, int push_current, char** out_pieces_blob, int64_t* out_blob_bytes, double** out_scores,
                        int64_t* out_n);

void zett_sampler_free(void* p);
void zett_sampler_destroy(zett_sampler* s);

#ifdef __cplusplus
}
#endif
#endif /* ZETT_B200_H_ */

// The non-GEMM kernels of the hypernetwork forward, plus the SIMT fp32 GEMM used as the on-device checker.
//
//   pack_*_kernel (five)    surface-form rows -> packed (pad-free) position lists, distinct ids / (id, position) pairs
//                           (modeling_hypernet.py:170-177,190)
//   gather_rescale_kernel   ids -> source / fallback embedding rows, in_scaler, 16-bit split (modeling_hypernet.py:179-188)
//   layernorm_kernel        (x [+ residual] [+ type/position embeddings]) -> LayerNorm -> fp32 + split planes
//   attention_kernel        per (row, heads of one warp) softmax(q k^T / sqrt(dh) + mask) v over <= S packed positions
//   split_rows_kernel       fp32 -> operand lines of the GEMM engine (weights at load time, with their per-row scales)
//   gemm_simt_kernel        same contract as gemm_tcgen05_kernel, CUDA cores, for checking
//
// Packing.  A surface-form row holds L ids, most of them pad (mean non-pad length ~2.9 of 7 on the benchmark
// vocabularies).  Pad positions act only as masked keys and their own outputs never reach position 0, so they are
// dropped: every per-position tensor is stored for the kept positions only, rows back to back.  Kept = non-pad
// positions plus position 0 (always the query that is read out).  A row with no valid key at all attends uniformly
// over all S positions in the reference (additive finfo.min mask, eager attention); such rows keep all L positions
// and are flagged so that the attention kernel reproduces the uniform weights.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "epilogue.cuh"
#include "ptx.cuh"

namespace zett {

constexpr int kPackThreads = 1024;   // the single-block scan
constexpr int kPackBlock = 256;      // the per-row / per-position kernels
constexpr int kMaxSurfaceLen = 32;

// counts[] slots
enum : int { kCntSurface = 0, kCntEncoder = 1, kCntBadId = 2, kCntRows = 3, kCntUnique = 4, kCntPairs = 5, kCntSlots = 8 };

struct PackParams {
  const int32_t* ids;    // [n_rows, L]
  int n_rows, L;
  int pad_id, v0, n_fallback;  // n_fallback = max(hn_n_extra_tokens, 1)
  int lang_slot;               // 1 when a lang-id position is appended to every row
  int dedup_ids;               // 0: every position is its own "distinct id" (the all-distinct measurement, ZETT_DEDUP_IDS=0)
  int* counts;                 // [kCntSlots], zeroed before the pass
  unsigned int* sticky_bad;    // handle-wide flag, set when any id of any pass was out of range (cleared by zett_hn_check)
  int* row_cnt;                // [n_rows]      kept positions of each row
  int* row_start1;             // [n_rows + 1]  surface packing (input projection)
  int* row_start2;             // [n_rows + 1]  encoder packing (surface positions + lang slot)
  int* tok_id;                 // [T1]  clamped id of each kept position
  int* tok_pos;                // [T1]  position id
  int* tok_enc;                // [T1]  index of this position in the encoder packing
  int* tok1_row;               // [T1]  row of this position
  int* lang_enc;               // [n_rows] index of the lang slot in the encoder packing (lang_slot only)
  int* tok2_row;               // [T2]
  unsigned char* tok2_valid;   // [T2]  1 = usable as attention key
  // de-duplication of the input projection: it depends on the id only (positions are added after it), so it is
  // evaluated once per DISTINCT id of the pass
  int* id_claim;               // [v0 + n_fallback] scratch, preset to INT_MAX-like: lowest position holding the id
  int* id_slot;                // [v0 + n_fallback] scratch: index of the id among the distinct ids
  int* uniq_src;               // [U]   source row (>= 0) or -1 - fallback row (< 0) of each distinct id
  int* tok_u;                  // [T1]  index into the distinct ids
  // de-duplication of the first encoder layer's input: LayerNorm(projection[id] + type + position[pos]) and the
  // query/key/value rows computed from it depend on the (id, position) pair only, so they are evaluated once per
  // DISTINCT pair of the pass (pair_claim == nullptr switches this off)
  int* pair_claim;             // [(v0 + n_fallback) * L] scratch, preset to INT_MAX-like: lowest position holding the pair
  int* pair_slot;              // [(v0 + n_fallback) * L] scratch: index of the pair among the distinct pairs
  int* pair_u;                 // [P]   distinct-id index of each distinct pair   (slot 0 = the lang-id position when lang_slot)
  int* pair_pos;               // [P]   position id of each distinct pair
  int* enc_pair;               // [T2]  distinct-pair index of every encoder position
};

__device__ __forceinline__ uint32_t kept_mask(const int32_t* row, int L, int pad_id, int lang_slot, uint32_t& nonpad) {
  nonpad = 0;
  for (int p = 0; p < L; ++p) nonpad |= (row[p] != pad_id ? 1u : 0u) << p;
  uint32_t kept = nonpad | 1u;
  if (!lang_slot && nonpad == 0) kept = (L >= 32) ? 0xFFFFFFFFu : ((1u << L) - 1u);
  return kept;
}

// The packing runs as five small grid-wide kernels on the caller's stream (a single block took 0.7 ms per 16 384 rows,
// 11 % of a whole XLM-R-shape step).  Slots of distinct ids / pairs are handed out by atomic counters, so their ORDER
// varies from run to run; the outputs do not (every GEMM / LayerNorm row is computed independently of its index).

// (1) one thread per row: kept positions, id range check
__global__ void __launch_bounds__(kPackBlock) pack_count_kernel(const PackParams p) {
  const int r = blockIdx.x * kPackBlock + threadIdx.x;
  if (r >= p.n_rows) return;
  const int32_t* row = p.ids + static_cast<long long>(r) * p.L;
  const int id_limit = p.v0 + p.n_fallback;
  uint32_t nonpad;
  p.row_cnt[r] = __popc(kept_mask(row, p.L, p.pad_id, p.lang_slot, nonpad));
  int bad = 0;
  for (int q = 0; q < p.L; ++q) bad |= (row[q] < 0 || row[q] >= id_limit) ? 1 : 0;
  if (bad) {
    atomicOr(&p.counts[kCntBadId], 1);
    atomicOr(p.sticky_bad, 1u);
  }
}

// (2) one block: exclusive scan of the per-row counts -> row starts of both packings, totals
__global__ void __launch_bounds__(kPackThreads, 1) pack_scan_kernel(const PackParams p) {
  __shared__ int warp_sums[32];
  __shared__ int warp_excl[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (p.n_rows + kPackThreads - 1) / kPackThreads;
  const int r0 = min(p.n_rows, tid * per), r1 = min(p.n_rows, r0 + per);
  int local = 0;
  for (int r = r0; r < r1; ++r) local += p.row_cnt[r];
  int incl = local;
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int w = warp_sums[lane];
    int wi = w;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xFFFFFFFFu, wi, o);
      if (lane >= o) wi += t;
    }
    warp_excl[lane] = wi - w;
    if (lane == 31) {
      p.counts[kCntSurface] = wi;
      if (!p.dedup_ids) p.counts[kCntUnique] = wi;
      p.counts[kCntEncoder] = wi + (p.lang_slot ? p.n_rows : 0);
      p.counts[kCntRows] = p.n_rows;
      p.counts[kCntPairs] = (p.pair_claim && p.lang_slot) ? 1 : 0;  // pair 0 = the lang-id position
      p.row_start1[p.n_rows] = wi;
      p.row_start2[p.n_rows] = wi + (p.lang_slot ? p.n_rows : 0);
    }
  }
  __syncthreads();
  int t1 = warp_excl[warp] + incl - local;
  for (int r = r0; r < r1; ++r) {
    p.row_start1[r] = t1;
    p.row_start2[r] = t1 + (p.lang_slot ? r : 0);
    t1 += p.row_cnt[r];
  }
}

// (3) one thread per row: position lists of both packings; every position bids for its id and its (id, position) pair
__global__ void __launch_bounds__(kPackBlock) pack_emit_kernel(const PackParams p) {
  const int r = blockIdx.x * kPackBlock + threadIdx.x;
  if (r >= p.n_rows) return;
  const int32_t* row = p.ids + static_cast<long long>(r) * p.L;
  const int id_limit = p.v0 + p.n_fallback;
  uint32_t nonpad;
  const uint32_t kept = kept_mask(row, p.L, p.pad_id, p.lang_slot, nonpad);
  const int t1 = p.row_start1[r], t2 = p.row_start2[r];
  int j = 0;
  for (int q = 0; q < p.L; ++q) {
    if (!((kept >> q) & 1u)) continue;
    const int id = max(0, min(row[q], id_limit - 1));  // never read out of bounds; kCntBadId reports the violation
    p.tok_id[t1 + j] = id;
    p.tok_pos[t1 + j] = q;
    p.tok_enc[t1 + j] = t2 + j;
    p.tok1_row[t1 + j] = r;
    p.tok2_row[t2 + j] = r;
    p.tok2_valid[t2 + j] = static_cast<unsigned char>((nonpad >> q) & 1u);
    if (p.dedup_ids) atomicMin(&p.id_claim[id], t1 + j);
    if (p.pair_claim) atomicMin(&p.pair_claim[id * p.L + q], t1 + j);
    ++j;
  }
  if (p.lang_slot) {
    p.lang_enc[r] = t2 + j;
    p.tok2_row[t2 + j] = r;
    p.tok2_valid[t2 + j] = 1;
    if (p.pair_claim) p.enc_pair[t2 + j] = 0;
  }
}

// (4) one thread per surface position: the lowest position holding an id / a pair owns it and takes the next free slot
__global__ void __launch_bounds__(kPackBlock) pack_owner_kernel(const PackParams p) {
  const int t = blockIdx.x * kPackBlock + threadIdx.x;
  if (t == 0 && p.pair_claim && p.lang_slot) { p.pair_u[0] = 0; p.pair_pos[0] = 0; }  // placeholder, overwritten by the lang-id LayerNorm
  if (t >= p.counts[kCntSurface]) return;
  const int id = p.tok_id[t];
  if (!p.dedup_ids) {
    p.uniq_src[t] = (id >= p.v0) ? (-1 - (id - p.v0)) : id;
  } else if (p.id_claim[id] == t) {
    const int u = atomicAdd(&p.counts[kCntUnique], 1);
    p.id_slot[id] = u;
    p.uniq_src[u] = (id >= p.v0) ? (-1 - (id - p.v0)) : id;
  }
  if (p.pair_claim) {
    const int pos = p.tok_pos[t];
    if (p.pair_claim[id * p.L + pos] == t) {
      const int pu = atomicAdd(&p.counts[kCntPairs], 1);
      p.pair_slot[id * p.L + pos] = pu;
      p.pair_u[pu] = id;  // the id for now; pack_index_kernel turns it into the distinct-id index
      p.pair_pos[pu] = pos;
    }
  }
}

// (5) one thread per surface position (and per distinct pair): indices into the distinct ids / pairs
__global__ void __launch_bounds__(kPackBlock) pack_index_kernel(const PackParams p) {
  const int t = blockIdx.x * kPackBlock + threadIdx.x;
  if (p.pair_claim && t >= (p.lang_slot ? 1 : 0) && t < p.counts[kCntPairs]) p.pair_u[t] = p.id_slot[p.pair_u[t]];
  if (t >= p.counts[kCntSurface]) return;
  const int id = p.tok_id[t];
  p.tok_u[t] = p.dedup_ids ? p.id_slot[id] : t;
  if (p.pair_claim) p.enc_pair[p.tok_enc[t]] = p.pair_slot[id * p.L + p.tok_pos[t]];
}

// -------------------------------------------------------------------------------------------------------------------
// Gather + rescale + split.  One CTA streams whole embedding rows: each row is fetched with ONE bulk async copy
// (cp.async.bulk, the TMA engine's linear mode) into a double-buffered shared-memory stage, so the HBM reads are
// full-line and independent of the thread mapping; the threads then apply  w * x + b  (in_scaler, not for fallback
// rows) and write the operand lines the input-projection GEMM consumes (operand.cuh).
// Algorithmic bytes per position: 4 E read + 4 E written.
// -------------------------------------------------------------------------------------------------------------------
constexpr int kGatherThreads = 256;

struct GatherParams {
  const float* source;     // [v0_rows, E]
  long long v0_rows;
  const float* fallback;   // [n_fallback, E]
  const float* scale_w;    // [E] or nullptr
  const float* scale_b;
  const int* tok_src;
  const int* n_tok;        // device count
  int E;
  OperandOut out;          // [cap, E] operand lines
};

__global__ void __launch_bounds__(kGatherThreads) gather_rescale_kernel(const GatherParams p) {
  extern __shared__ __align__(128) uint8_t gsmem[];
  __shared__ __align__(8) unsigned long long bars[2];
  float* stage[2] = {reinterpret_cast<float*>(gsmem), reinterpret_cast<float*>(gsmem) + p.E};
  const uint32_t bar[2] = {smem_u32(&bars[0]), smem_u32(&bars[1])};
  const int n_tok = *p.n_tok;
  const uint32_t row_bytes = static_cast<uint32_t>(p.E) * 4u;
  const int tid = threadIdx.x;

  if (tid == 0) {
    mbar_init(bar[0], 1);
    mbar_init(bar[1], 1);
    fence_barrier_init();
  }
  __syncthreads();

  auto row_ptr = [&](int t) -> const float* {
    const int s = __ldg(p.tok_src + t);
    if (s >= 0) return p.source + static_cast<long long>(min(static_cast<long long>(s), p.v0_rows - 1)) * p.E;
    return p.fallback + static_cast<long long>(-1 - s) * p.E;
  };
  int t = blockIdx.x;
  if (tid == 0 && t < n_tok) {
    mbar_expect_tx(bar[0], row_bytes);
    bulk_load_1d(smem_u32(stage[0]), row_ptr(t), row_bytes, bar[0]);
  }
  uint32_t phase[2] = {0u, 0u};
  uint32_t bad = 0;
  int buf = 0;
  for (; t < n_tok; t += gridDim.x, buf ^= 1) {
    const int t_next = t + gridDim.x;
    if (tid == 0 && t_next < n_tok) {  // the other stage was released by the __syncthreads of the last iteration
      mbar_expect_tx(bar[buf ^ 1], row_bytes);
      bulk_load_1d(smem_u32(stage[buf ^ 1]), row_ptr(t_next), row_bytes, bar[buf ^ 1]);
    }
    mbar_wait(bar[buf], phase[buf], 16);
    phase[buf] ^= 1u;
    const bool rescale = p.scale_w != nullptr && __ldg(p.tok_src + t) >= 0;
    const float* x = stage[buf];
    for (int e0 = tid * 8; e0 < p.E; e0 += kGatherThreads * 8) {
      float v[8];
      *reinterpret_cast<float4*>(v) = *reinterpret_cast<const float4*>(x + e0);
      *reinterpret_cast<float4*>(v + 4) = *reinterpret_cast<const float4*>(x + e0 + 4);
      if (rescale) {
        float w[8], b[8];
        *reinterpret_cast<float4*>(w) = __ldg(reinterpret_cast<const float4*>(p.scale_w + e0));
        *reinterpret_cast<float4*>(w + 4) = __ldg(reinterpret_cast<const float4*>(p.scale_w + e0 + 4));
        *reinterpret_cast<float4*>(b) = __ldg(reinterpret_cast<const float4*>(p.scale_b + e0));
        *reinterpret_cast<float4*>(b + 4) = __ldg(reinterpret_cast<const float4*>(p.scale_b + e0 + 4));
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __fadd_rn(__fmul_rn(w[i], v[i]), b[i]);  // Rescaler: w * x + b, two roundings
      }
      store_operand8(p.out, t, e0, v, false, bad);
    }
    __syncthreads();  // everyone is done reading stage[buf] before it is refilled two iterations from now
  }
  report_saturation(p.out.sat, bad);
}

// -------------------------------------------------------------------------------------------------------------------
// LayerNorm over H with the additions that precede it in the reference:
//     v = a[t] (+ res[t]) (+ vec0) (+ table[idx])        y = (v - mean) / sqrt(var + eps) * gamma + beta
// ProjectorBlock.ln (eps 1e-6, modeling_hypernet.py:33,40), RobertaEmbeddings (type + position embeddings, eps 1e-5),
// RobertaSelfOutput / RobertaOutput (residual, eps 1e-5).  Two-pass statistics in fp32 like torch.  One CTA per row.
// Outputs: fp32 (the residual stream) and the split planes the next GEMM reads; optionally scattered through
// out_index (surface packing -> encoder packing), optionally duplicated for the first position of every row into
// compact [n_rows, H] buffers (the pruned last layer reads only those), optionally a dot product with a vector
// (bias_projection, modeling_hypernet.py:260-265).
// -------------------------------------------------------------------------------------------------------------------
constexpr int kLnBlockVec = 8;  // block per row: float4 per thread; 128 threads up to H = 4096, 256 up to H = 8192
constexpr int kLnWarpVec = 16;  // warp per row: float4 per lane, H <= 4 * 16 * 32 = 2048

struct LnParams {
  const float* a;
  long long lda;             // 0 broadcasts one vector to every row
  const int* in_index;       // nullable: row of `a` to read for position t (distinct-id table -> positions)
  const float* res;          // nullable, row stride H
  const float* vec0;         // nullable [H]
  const float* table;        // nullable [*, H]
  const int* table_idx;      // nullable -> table_const
  int table_const;
  const float* gamma;
  const float* beta;
  float eps;
  int H;
  const int* n_dev;          // nullable -> n_host
  int n_host;
  const int* out_index;      // nullable
  float* out_f32;            // nullable [*, H]
  OperandOut out_op;         // base nullable: [*, H] operand lines
  const int* tok_row;        // nullable: enables the compact copy for positions with row_start[tok_row[t]] == t
  const int* row_start;
  float* c_f32;
  OperandOut c_op;           // base nullable
  const float* dot_w;        // nullable [H]
  const float* dot_b;        // [1]
  float* dot_out;            // [n * dot_ld]
  long long dot_ld;
};

__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  __syncthreads();  // protect `red` from the previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < nwarp) ? red[lane] : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, o);
  return t;
}

__device__ __forceinline__ void store_split4(const OperandOut& o, long long row, int col, const float4 y, bool is_weight, uint32_t& bad) {
  const float v[4] = {y.x, y.y, y.z, y.w};
  store_operand4(o, row, col, v, is_weight, bad);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

// WARP = false: one block per row (H up to 4 * VEC * blockDim), block-wide reductions.
// WARP = true : one warp per row (H up to 128 * VEC), shuffle reductions only -- rows of a few KB (H <= 2048) are
//               latency-bound on the two block barriers otherwise (43 % of the HBM roofline at H = 768).
// FMT is the operand format of out_op / c_op as a compile-time constant, and VEC is sized to the row by the launcher:
// with the format read from the struct and VEC = 8 for every block-per-row shape the kernel executed ~200 instructions
// per float4 (three pack variants behind uniform branches, twice; half of the unrolled loop dead) and ran at 58 % of
// the SM's issue rate and 35 % of HBM -- issue-bound, not memory-bound (profiles/layernorm_r2_before.csv).
// (minimum blocks per SM, MINB: without a bound ptxas hoists every gamma / beta / table load of the unrolled store loop to
// the top -- 196 registers, ONE resident block per SM, 13 % of the HBM rate.  Eight float4 per lane run best with three
// blocks of <= 80 registers (at 64 they spill 16-24 bytes; XLM-R's LayerNorms 625 -> 568 us per pass), four float4 with
// four blocks, the 16-float4-per-lane warp variant keeps its row in 64 registers and gets two.)
// T = threads per row (32 for the warp form, the block size otherwise) is a template constant as well: the quads of a lane
// are T apart, so every address of the row is "per-row base + lane offset" (computed once per row) plus a compile-time
// multiple of 16 T bytes (fp32 rows and operand lines alike) -- no per-quad address arithmetic.
template <bool WARP, int VEC, int FMT, int T, int MINB = 4>
__global__ void __launch_bounds__(256, (WARP && VEC > 8) ? 2 : MINB) layernorm_kernel(const LnParams p) {
  __shared__ float red[32];
  const int n = p.n_dev ? *p.n_dev : p.n_host;
  const int H4 = p.H >> 2;
  const int lane_in_row = WARP ? (threadIdx.x & 31) : threadIdx.x;
  constexpr int rows_per_block = WARP ? 8 : 1;
  auto row_sum = [&](float v) { return WARP ? warp_sum(v) : block_sum(v, red); };
  const float fh = static_cast<float>(p.H);
  const float4* g4 = reinterpret_cast<const float4*>(p.gamma) + lane_in_row;
  const float4* b4 = reinterpret_cast<const float4*>(p.beta) + lane_in_row;
  const float4* w4 = p.dot_w ? reinterpret_cast<const float4*>(p.dot_w) + lane_in_row : nullptr;
  const float4* z4 = p.vec0 ? reinterpret_cast<const float4*>(p.vec0) + lane_in_row : nullptr;
  const bool has_op = p.out_op.base != nullptr;
  // quads this lane owns: c < n_c  (lane_in_row + c * T < H4)
  const int n_c = lane_in_row < H4 ? min(VEC, (H4 - lane_in_row + T - 1) / T) : 0;
  uint32_t bad = 0;
  for (int t = blockIdx.x * rows_per_block + (WARP ? (threadIdx.x >> 5) : 0); t < n; t += gridDim.x * rows_per_block) {
    float4 v[VEC];
    const float4* a4 = reinterpret_cast<const float4*>(p.a + static_cast<long long>(p.in_index ? __ldg(p.in_index + t) : t) * p.lda) + lane_in_row;
    const float4* r4 = p.res ? reinterpret_cast<const float4*>(p.res + static_cast<long long>(t) * p.H) + lane_in_row : nullptr;
    const float4* e4 = nullptr;
    if (p.table) {
      const int idx = p.table_idx ? __ldg(p.table_idx + t) : p.table_const;
      e4 = reinterpret_cast<const float4*>(p.table + static_cast<long long>(idx) * p.H) + lane_in_row;
    }
    // everything the store phase needs that does not depend on the statistics, fetched while the row is in flight
    const long long orow = p.out_index ? __ldg(p.out_index + t) : t;
    long long crow = -1;
    if (p.tok_row) {
      const int r = __ldg(p.tok_row + t);
      if (__ldg(p.row_start + r) == t) crow = r;
    }
#pragma unroll
    for (int c = 0; c < VEC; ++c)
      if (c < n_c) v[c] = a4[c * T];
    if (r4) {
#pragma unroll
      for (int c = 0; c < VEC; ++c)
        if (c < n_c) { const float4 y = r4[c * T]; v[c].x += y.x; v[c].y += y.y; v[c].z += y.z; v[c].w += y.w; }
    }
    if (z4) {
#pragma unroll
      for (int c = 0; c < VEC; ++c)
        if (c < n_c) { const float4 y = __ldg(z4 + c * T); v[c].x += y.x; v[c].y += y.y; v[c].z += y.z; v[c].w += y.w; }
    }
    if (e4) {
#pragma unroll
      for (int c = 0; c < VEC; ++c)
        if (c < n_c) { const float4 y = __ldg(e4 + c * T); v[c].x += y.x; v[c].y += y.y; v[c].z += y.z; v[c].w += y.w; }
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < VEC; ++c)
      if (c < n_c) s += (v[c].x + v[c].y) + (v[c].z + v[c].w);
    const float mean = row_sum(s) / fh;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
      if (c < n_c) {
        v[c].x -= mean; v[c].y -= mean; v[c].z -= mean; v[c].w -= mean;
        q += (v[c].x * v[c].x + v[c].y * v[c].y) + (v[c].z * v[c].z + v[c].w * v[c].w);
      }
    }
    const float var = row_sum(q) / fh;
    const float rstd = 1.0f / sqrtf(var + p.eps);

    float4* of = p.out_f32 ? reinterpret_cast<float4*>(p.out_f32 + orow * p.H) + lane_in_row : nullptr;
    float4* cf = (crow >= 0 && p.c_f32) ? reinterpret_cast<float4*>(p.c_f32 + crow * p.H) + lane_in_row : nullptr;
    const bool has_cop = crow >= 0 && p.c_op.base != nullptr;
    OperandCursor oc{nullptr, nullptr}, cc{nullptr, nullptr};
    if (has_op) oc = operand_cursor<FMT>(p.out_op, orow, lane_in_row);
    if (has_cop) cc = operand_cursor<FMT>(p.c_op, crow, lane_in_row);
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
      if (c < n_c) {
        const float4 g = __ldg(g4 + c * T), b = __ldg(b4 + c * T);
        float y[4];
        y[0] = v[c].x * rstd * g.x + b.x;
        y[1] = v[c].y * rstd * g.y + b.y;
        y[2] = v[c].z * rstd * g.z + b.z;
        y[3] = v[c].w * rstd * g.w + b.w;
        const float4 y4 = make_float4(y[0], y[1], y[2], y[3]);
        if (of) of[c * T] = y4;
        if (cf) cf[c * T] = y4;
        if (has_op || has_cop) {
          const Packed4 pk = pack_operand4(y, FMT, false, bad);
          if (has_op) store_packed4_cursor<FMT>(oc, c * operand_quad_stride<FMT>(T), pk);
          if (has_cop) store_packed4_cursor<FMT>(cc, c * operand_quad_stride<FMT>(T), pk);
        }
        if (w4) {
          const float4 w = __ldg(w4 + c * T);
          dot += (y[0] * w.x + y[1] * w.y) + (y[2] * w.z + y[3] * w.w);
        }
      }
    }
    if (w4) {
      const float d = row_sum(dot);
      if (lane_in_row == 0) p.dot_out[static_cast<long long>(t) * p.dot_ld] = d + __ldg(p.dot_b);
    }
  }
  if (p.out_op.base || p.c_op.base) report_saturation(p.out_op.base ? p.out_op.sat : p.c_op.sat, bad);
}

// -------------------------------------------------------------------------------------------------------------------
// Attention over the <= S packed positions of one row.  LPH lanes hold one head (4 * VPL consecutive elements each,
// 16-byte loads), so a warp serves 32 / LPH heads of a row at once; dot products reduce with shuffles inside the LPH
// lanes, softmax is an online running max / sum in fp32 (HF eager attention: scores * dh^-0.5 + additive mask,
// softmax, @ V).  A row without any valid key gets uniform weights over all of its positions (what finfo.min +
// softmax yields).  row0_only: only the first position of each row is a query (pruned last layer); q and out are
// then indexed by row.  qkv_index: q / k / v rows are read through an index (first layer on distinct pairs).
// -------------------------------------------------------------------------------------------------------------------
struct AttnParams {
  const float* q; long long ldq;
  const float* k; long long ldk;
  const float* v; long long ldv;
  const int* row_start;          // encoder packing [n_rows + 1]
  const unsigned char* valid;    // [T2]
  const int* qkv_index;          // nullable [T2]: row of q / k / v holding encoder position t (distinct-pair rows)
  int n_rows, n_heads, dh;
  float scale;
  int row0_only;
  OperandOut out;                // [*, H] operand lines
};

// Lane mapping: with one float4 per lane (LPH = dh / 4) a key cost five shuffle + add steps and a full set of softmax
// scalars for four FMAs of dot product, and the kernel ran at 77 % of the SM's issue rate (profiles/attention_r2_mistral.csv)
// -- issue-bound.  Eight lanes per head with VPL float4 each (dh = 32 * VPL) need three shuffle steps per 4 * VPL FMAs and
// share the scalars among four heads: 2.3x fewer instructions per (query, key, head) at dh = 128, 1.5x at dh = 64.
template <int LPH, int VPL, int FMT>
__global__ void __launch_bounds__(256) attention_kernel(const AttnParams p) {
  constexpr int HPW = 32 / LPH;  // heads per warp
  const int lane = threadIdx.x & 31;
  const int groups = (p.n_heads + HPW - 1) / HPW;
  const long long wid = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= static_cast<long long>(p.n_rows) * groups) return;
  const int r = static_cast<int>(wid / groups);
  const int h = static_cast<int>(wid % groups) * HPW + lane / LPH;
  const bool head_ok = h < p.n_heads;               // lanes of a missing head still take part in the shuffles
  const int hoff = (head_ok ? h : 0) * p.dh + (lane % LPH) * 4;
  const int t0 = __ldg(p.row_start + r), t1 = __ldg(p.row_start + r + 1);
  const int n = t1 - t0;
  uint32_t valid_mask = 0;
  for (int j = 0; j < n; ++j) valid_mask |= (__ldg(p.valid + t0 + j) != 0 ? 1u : 0u) << j;
  const bool any_valid = valid_mask != 0;
  const int n_q = p.row0_only ? 1 : n;
  uint32_t bad = 0;
  // (Measured and not kept, round 2: prefetch.global.L1 of all of a row's K / V lines ahead of the query loop -- no change
  // on any of the three shapes; 64 registers for four resident blocks at VPL = 4 -- inside the box-to-box noise.)
  // K and V are streamed per query (L1-resident re-reads).  Keeping a row's keys and values in registers (all loads issued
  // up front, 128 registers, two resident blocks per SM) was measured slower twice: round 1, and again in round 2 as a
  // separate kernel -- 15.4 vs 14.5 ms per XLM-R-shape step.
  for (int i = 0; i < n_q; ++i) {
    const long long qi = p.row0_only ? r : (p.qkv_index ? __ldg(p.qkv_index + t0 + i) : (t0 + i));
    float4 qv[VPL];
#pragma unroll
    for (int d = 0; d < VPL; ++d) qv[d] = __ldg(reinterpret_cast<const float4*>(p.q + qi * p.ldq + hoff + 4 * LPH * d));
    float m = -INFINITY, l = 0.f;
    float4 acc[VPL];
#pragma unroll
    for (int d = 0; d < VPL; ++d) acc[d] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < n; ++j) {
      if (any_valid && !((valid_mask >> j) & 1u)) continue;  // masked key: weight exactly 0
      const long long kvj = p.qkv_index ? __ldg(p.qkv_index + t0 + j) : (t0 + j);
      float s = 0.f;
      if (any_valid) {
        const float* kj = p.k + kvj * p.ldk + hoff;
#pragma unroll
        for (int d = 0; d < VPL; ++d) {
          const float4 kk = __ldg(reinterpret_cast<const float4*>(kj + 4 * LPH * d));
          s += (qv[d].x * kk.x + qv[d].y * kk.y) + (qv[d].z * kk.z + qv[d].w * kk.w);
        }
#pragma unroll
        for (int o = LPH / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
        s *= p.scale;
      }
      const float m_new = fmaxf(m, s);
      const float corr = __expf(m - m_new);  // exp(-inf) = 0 on the first key
      const float w = __expf(s - m_new);
      l = l * corr + w;
      const float* vj = p.v + kvj * p.ldv + hoff;
#pragma unroll
      for (int d = 0; d < VPL; ++d) {
        const float4 vv = __ldg(reinterpret_cast<const float4*>(vj + 4 * LPH * d));
        acc[d].x = acc[d].x * corr + w * vv.x; acc[d].y = acc[d].y * corr + w * vv.y;
        acc[d].z = acc[d].z * corr + w * vv.z; acc[d].w = acc[d].w * corr + w * vv.w;
      }
      m = m_new;
    }
    const float inv = 1.0f / l;
    const long long orow = p.row0_only ? r : (t0 + i);
    if (head_ok) {
#pragma unroll
      for (int d = 0; d < VPL; ++d) {
        const float y[4] = {acc[d].x * inv, acc[d].y * inv, acc[d].z * inv, acc[d].w * inv};
        store_packed4_as<FMT>(p.out, orow, hoff + 4 * LPH * d, pack_operand4(y, FMT, false, bad));
      }
    }
  }
  report_saturation(p.out.sat, bad);
}

// fp32 [rows, k] -> operand lines, one block per row (k a multiple of 4).  Weights of kFmtF16F8 are stored scaled by a
// power of two per row that brings the row maximum into [16, 32) (operand.cuh); `inv_scale[row]` receives the inverse,
// which the GEMM epilogue applies to output column `row`.  Values outside fp16's range after that are counted in out.sat.
__global__ void __launch_bounds__(256) split_rows_kernel(const float* __restrict__ x, long long rows, int k, OperandOut out,
                                                         long long row_off, bool is_weight, float* inv_scale) {
  __shared__ float red[32];
  uint32_t bad = 0;
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const float4* x4 = reinterpret_cast<const float4*>(x + r * k);
    float scale = 1.0f;
    if (inv_scale) {
      float m = 0.f;
      for (int i = threadIdx.x; i < k / 4; i += blockDim.x) {
        const float4 y = __ldg(x4 + i);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(y.x), fabsf(y.y))), fmaxf(fabsf(y.z), fabsf(y.w)));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
      __syncthreads();
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
      __syncthreads();
      m = 0.f;
      for (int w = 0; w < (blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
      int e = 0;
      if (m > 0.f && isfinite(m)) {
        frexpf(m, &e);            // m = f * 2^e, f in [0.5, 1)  ->  m * 2^(5 - e) in [16, 32)
        e = max(-100, min(100, 5 - e));
        scale = ldexpf(1.0f, e);
      }
      if (threadIdx.x == 0) inv_scale[row_off + r] = 1.0f / scale;   // a power of two: exact
    }
    for (int i = threadIdx.x; i < k / 4; i += blockDim.x) {
      float4 y = __ldg(x4 + i);
      y.x *= scale; y.y *= scale; y.z *= scale; y.w *= scale;
      store_split4(out, row_off + r, 4 * i, y, is_weight, bad);
    }
  }
  report_saturation(out.sat, bad);
}

// -------------------------------------------------------------------------------------------------------------------
// SIMT checker GEMM: identical contract and identical product terms (fp32 accumulation) as the tcgen05 kernel, on CUDA
// cores, reading the same operand lines.  128 x 32 output tile per CTA, one thread per row.  Used by the tests to
// validate the tensor-core path on the device and selectable (gemm_impl = 3) to bisect a failure; never the default.
// -------------------------------------------------------------------------------------------------------------------
struct SimtGemmParams {
  const uint8_t* a; long long a_ld;   // A operand lines [M, a_ld bytes]
  const uint8_t* w; long long w_ld;   // W operand lines [N, w_ld bytes]
  int m_host; const int* m_dev;
  int n, k;
  int fmt;
};

__global__ void __launch_bounds__(128) gemm_simt_kernel(const SimtGemmParams s, const EpilogueParams ep) {
  // term t of the product list is As[t] . Ws[t]:  bf16x3: (hi,hi) (lo,hi) (hi,lo);  f16+fp8: (p0,p0) (q0,q0) (q1,q1);  bf16x1: (hi,hi)
  __shared__ float As[3][128][17], Ws[3][32][17];
  const int M = s.m_dev ? *s.m_dev : s.m_host;
  const int row0 = blockIdx.y * 128, col0 = blockIdx.x * 32;
  if (row0 >= M) return;
  const int tid = threadIdx.x;
  const bool f8 = s.fmt == kFmtF16F8;
  float acc[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) acc[c] = 0.f;
  for (int k0 = 0; k0 < s.k; k0 += 16) {
    for (int i = tid; i < 128 * 16; i += 128) {
      const int r = i >> 4, c = i & 15;
      const long long gr = row0 + r;
      float x0 = 0.f, x1 = 0.f, x2 = 0.f;
      if (gr < M && k0 + c < s.k) {
        x0 = operand_plane(s.a, s.a_ld, s.fmt, gr, k0 + c, 0);
        x1 = operand_plane(s.a, s.a_ld, s.fmt, gr, k0 + c, 1);
        x2 = f8 ? operand_plane(s.a, s.a_ld, s.fmt, gr, k0 + c, 2) : x0;
      }
      As[0][r][c] = x0; As[1][r][c] = x1; As[2][r][c] = x2;
    }
    for (int i = tid; i < 32 * 16; i += 128) {
      const int r = i >> 4, c = i & 15;
      const long long gn = col0 + r;
      float y0 = 0.f, y1 = 0.f, y2 = 0.f;
      if (gn < s.n && k0 + c < s.k) {
        y0 = operand_plane(s.w, s.w_ld, s.fmt, gn, k0 + c, 0);
        if (f8) {
          y1 = operand_plane(s.w, s.w_ld, s.fmt, gn, k0 + c, 1);
          y2 = operand_plane(s.w, s.w_ld, s.fmt, gn, k0 + c, 2);
        } else {
          y1 = y0;                                                    // A_lo . W_hi (A_lo is zero in the one-term format)
          y2 = operand_plane(s.w, s.w_ld, s.fmt, gn, k0 + c, 1);      // A_hi . W_lo
        }
      }
      Ws[0][r][c] = y0; Ws[1][r][c] = y1; Ws[2][r][c] = y2;
    }
    __syncthreads();
#pragma unroll 4
    for (int kk = 0; kk < 16; ++kk) {
      const float x0 = As[0][tid][kk], x1 = As[1][tid][kk], x2 = As[2][tid][kk];
#pragma unroll
      for (int c = 0; c < 32; ++c) acc[c] += x0 * Ws[0][c][kk] + x1 * Ws[1][c][kk] + x2 * Ws[2][c][kk];
    }
    __syncthreads();
  }
  const int row = row0 + tid;
  uint32_t bad = 0;
  if (row < M) epilogue_row32(ep, row, col0, s.n, acc, bad);
  report_saturation(ep.out_op.sat, bad);
}

}  // namespace zett

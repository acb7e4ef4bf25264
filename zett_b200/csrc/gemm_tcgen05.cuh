// Persistent, warp-specialised tcgen05 GEMM for the hypernetwork's Linear layers:
//     out[m, n] = epilogue( sum_k A[m, k] * W[n, k] )          (nn.Linear: y = x W^T + b, both operands K-major)
//
//  * Operands are rows of 128-byte lines (operand.cuh); ONE line per row is a pipeline stage: the A box is {128 B, 128
//    rows}, the W box {128 B, this CTA's share of the tile's W rows}, both fetched with the 128-byte swizzle, one TMA
//    each.  All product terms of a format (fp16 main term + two e5m2 correction terms; three bf16 terms; one bf16 term)
//    are tcgen05.mma instructions on 32-byte K-slices of the same swizzled rows, accumulating into the SAME TMEM
//    accumulator -- the list is a compile-time table (MmaList), so the issue loop is straight-line code.
//  * A CTA pair (cta_group::2, UMMA M = 256) owns a tile: each CTA loads its 128 rows of A and its half of the W rows,
//    only the leader arrives on the full barrier (expecting both CTAs' bytes) and issues the MMAs; commits are multicast.
//  * Roles: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2..9 = epilogue.  Role dispatch is on a
//    warp-uniform warp index and the producer / issuer loops run with all 32 lanes converged, electing one lane only
//    around the asynchronous instructions: their operands (descriptors, barrier addresses, coordinates) then live in
//    uniform registers.  (Round 1 branched on `lane == 0`; ptxas wrapped every tcgen05.mma in a five-R2UR waterfall
//    loop, ~25 instructions per issue, and the single issuing thread could not keep the tensor pipe busy: 51 % active on
//    the one-term probe.)
//  * HALVES == 1: 256 x block_n tiles, two TMEM accumulators so the epilogue of tile i overlaps the main loop of tile
//    i + 1.  HALVES == 2: 256 x 512 tiles, the two accumulators hold the two N halves of ONE tile: every line of A is
//    fetched once for 512 columns (25 % fewer operand bytes per FLOP), the epilogue does not overlap the next main loop.
//  * Epilogue: eight warps, two per TMEM lane quarter.  A warp reads a 32-lane x 32-column accumulator chunk
//    (tcgen05.ld 32x32b.x32: one ROW per thread), transposes it through a private 4 KB shared-memory tile (16-byte
//    chunks XOR-swizzled by row, conflict-free both ways) and continues with one QUAD per lane (epilogue.cuh): eight
//    lanes cover 128 contiguous bytes of an output row, so residual loads, fp32 stores and operand-line stores are whole
//    lines.  (Row-per-thread stores touch 32 lines per instruction; the LSU retires one line per cycle, which made the
//    epilogue of a 256 x 512 tile 15 000 cycles long.)
//  * M may live in device memory (packed-position counts are data dependent); tiles beyond it are skipped.
//  * Every barrier wait is bounded by a watchdog (ptx.cuh).  With GemmShape::prof set, every role accumulates the cycles
//    it spent waiting on each kind of barrier (the stall picture of one launch, read by the probes in tests/).
#pragma once
#include <cuda.h>
#include "ptx.cuh"
#include "epilogue.cuh"

namespace zett {

constexpr int kBlockM = 128;         // rows of A per CTA (UMMA M = 256 per pair)
constexpr int kMaxStages = 8;
constexpr int kEpilogueWarps = 8;    // two per TMEM lane quarter
constexpr int kGemmThreads = 64 + 32 * kEpilogueWarps;   // TMA producer warp, MMA warp, epilogue warps
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kLineBytes = 128;
constexpr uint32_t kABytes = kBlockM * kLineBytes;        // A part of a stage: 16 KB
constexpr uint32_t kStagingBytes = kEpilogueWarps * 4096; // per-warp transpose tiles
constexpr int kProfSlots = 12;       // per CTA: 0 total, 1 producer empty-wait, 2 MMA full-wait, 3 MMA accumulator-wait,
                                     //          4 epilogue accumulator-wait (warp 2), 5 epilogue busy (warp 2), 6 tiles,
                                     //          7 epilogue tcgen05.ld wait, 8 staging stores, 9 quads (warp 2)

struct GemmShape {
  int m_host;          // rows of A / out when m_dev == nullptr
  const int* m_dev;    // optional device-resident row count
  int n;
  int k_lines;         // lines per operand row
  int block_n;         // UMMA N: a multiple of 32 up to 256
  int group_m;         // rasterisation: m-tiles per group
  int chunk_n;         // rasterisation: n-tiles per L2-resident W chunk
  int num_stages;
  uint32_t stage_bytes;
  uint64_t hint_a, hint_b;   // L2 eviction policies of the A / W loads (ptx.cuh)
  uint32_t idesc16;    // tcgen05 instruction descriptor, kind::f16
  uint32_t idesc8;     // tcgen05 instruction descriptor, kind::f8f6f4 (e5m2 x e5m2)
  unsigned long long* prof;  // nullable [gridDim.x, kProfSlots]
  unsigned int* wave_sync;   // nullable [3], zero between launches: {arrivals, exits, gave up} of the producers' wave barrier
  int sync_every;            // the producers of all CTAs meet every `sync_every` full waves of tiles (0: never)
};

struct TileCoord { int m_blk, n_blk; };

// Rasterisation for L2 residency: N is cut into chunks whose W panel stays in L2 while all of M sweeps past it in small
// m-groups; inside a group m varies fastest, so the CTA pairs of a wave share W tiles and each A tile is fetched once per
// chunk.
__device__ __forceinline__ TileCoord tile_coord(int tile, int m_tiles, int n_tiles, int group_m, int chunk_n) {
  const int chunk_tiles = m_tiles * chunk_n;            // tiles of a full chunk
  const int c = tile / chunk_tiles;
  const int first_n = c * chunk_n;
  const int cn = min(chunk_n, n_tiles - first_n);       // n-tiles of this chunk
  const int in_chunk = tile - c * chunk_tiles;          // (full chunks precede: their size is m_tiles * chunk_n)
  const int group_tiles = group_m * cn;
  const int g = in_chunk / group_tiles;
  const int first_m = g * group_m;
  const int gm = min(group_m, m_tiles - first_m);
  const int r = in_chunk - g * group_tiles;
  TileCoord t;
  t.m_blk = first_m + r % gm;
  t.n_blk = first_n + r / gm;
  return t;
}

// The tcgen05.mma instructions issued per line of K, as (kind is fp8?, byte offset of the A slice, of the W slice).
template <int FMT> struct MmaList;
template <> struct MmaList<kFmtF16F8> {
  static constexpr int kCount = 4;
  __device__ static constexpr bool f8(int i) { return i >= 2; }
  __device__ static constexpr uint32_t a_off(int i) { return 32u * i; }   // p0[0:16] p0[16:32] | q0 | q1
  __device__ static constexpr uint32_t b_off(int i) { return 32u * i; }
};
template <> struct MmaList<kFmtBf16x3> {
  static constexpr int kCount = 6;
  __device__ static constexpr bool f8(int) { return false; }
  // hi.hi (two K slices), lo.hi, hi.lo
  __device__ static constexpr uint32_t a_off(int i) { return i < 2 ? 32u * i : (i < 4 ? 64u + 32u * (i - 2) : 32u * (i - 4)); }
  __device__ static constexpr uint32_t b_off(int i) { return i < 2 ? 32u * i : (i < 4 ? 32u * (i - 2) : 64u + 32u * (i - 4)); }
};
template <> struct MmaList<kFmtBf16x1> {
  static constexpr int kCount = 4;
  __device__ static constexpr bool f8(int) { return false; }
  __device__ static constexpr uint32_t a_off(int i) { return 32u * i; }
  __device__ static constexpr uint32_t b_off(int i) { return 32u * i; }
};

// hi word of the shared-memory matrix descriptor of a K-major, 128-byte-swizzled tile: stride between 8-row atoms 1024 B,
// descriptor version 1 (bit 46), SWIZZLE_128B (bits 61-63 = 2); the lo word is (address >> 4) | leading-offset field 1
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  return (static_cast<uint64_t>(kDescHi) << 32) | static_cast<uint64_t>(((smem_addr & 0x3FFFFu) >> 4) | (1u << 16));
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 ld_shared_v4(uint32_t addr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr) : "memory");
  return r;
}

// One 32-row x 32-column accumulator chunk, second half of its trip: the rows the warp's threads hold (tcgen05.ld layout:
// one row per thread) have been written to the warp's staging tile; lane l now takes the quads (row 4 i + l / 8, columns
// 4 (l % 8) ..) for i = 0..7.  The epilogue shares the SM's four issue ports with nothing else, and with K = 768 (XLM-R
// shape) a tile's main loop is only ~12 000 cycles long, so this loop is written for instruction count: the outputs a
// GEMM produces (fp32 and / or operand lines) and its activation are template parameters, row pointers advance by
// constants, and the loop is rolled in two halves of four quads (fully unrolled with both GELUs inlined the epilogue was
// ~60 KB of SASS per activation -- instruction-cache misses on top).
// FLAGS: what the GEMM's epilogue does, as compile-time bits (kEpiGeneric = decide at run time, the catch-all)
enum : int { kEpiF32 = 1, kEpiOp = 2, kEpiRes = 4, kEpiAffine = 8, kEpiGeneric = 16 };

// The residual of a chunk as one lane sees it: the quads (row 4 i + lane / 8, 4 columns), i = 0..7, fetched ONE CHUNK AHEAD
// (load_residual): a residual row was written by a LayerNorm kernel hundreds of megabytes ago, so each of these loads is a
// DRAM round trip, and issued inside the quad loop their latency was exposed eight times per tile -- longer than the whole
// main loop of a K = 768 tile.
struct ResidualRegs {
  float4 r[8];
};

__device__ __forceinline__ void load_residual(const EpilogueParams& ep, long long row0, int M, int col, int n, int lane, ResidualRegs& rr) {
  const long long row = row0 + (lane >> 3);
  const float* pr = ep.residual + row * ep.ld_res + col;
  const int n_rows = static_cast<int>(min(static_cast<long long>(32), static_cast<long long>(M) - row));
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    rr.r[i] = (col < n && 4 * i < n_rows) ? __ldg(reinterpret_cast<const float4*>(pr)) : make_float4(0.f, 0.f, 0.f, 0.f);
    pr += 4 * ep.ld_res;
  }
}

// pull the residual lines of a whole tile share (32 rows x `cols` columns from col0) towards L2 before they are needed
__device__ __forceinline__ void prefetch_residual_l2(const EpilogueParams& ep, long long row0, int M, int col0, int cols, int n, int lane) {
  const long long row = row0 + lane;   // one row per lane, one prefetch per 128-byte line
  if (row >= M) return;
  const float* pr = ep.residual + row * ep.ld_res;
  for (int c = col0; c < col0 + cols && c < n; c += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(pr + c));
}

template <int FMT, int ACT, int FLAGS>
__device__ __forceinline__ void epilogue_quads(const EpilogueParams& ep, uint32_t stg, int lane, long long row0, int M, int col,
                                               const QuadConsts& q, const ResidualRegs& rres, uint32_t& bad) {
  constexpr bool generic = (FLAGS & kEpiGeneric) != 0;
  const bool f32 = generic ? ep.out_f32 != nullptr : (FLAGS & kEpiF32) != 0;
  const bool op = generic ? ep.out_op.base != nullptr : (FLAGS & kEpiOp) != 0;
  const bool res = generic ? ep.residual != nullptr : (FLAGS & kEpiRes) != 0;
  const bool affine = generic ? ep.col_scale != nullptr : (FLAGS & kEpiAffine) != 0;
  const bool scaled = generic ? (ep.w_scale != nullptr || ep.bias != nullptr) : true;   // (identity constants otherwise)
  const int jj = lane & 7, rsub = lane >> 3;
  const long long row = row0 + rsub;                    // this lane's rows are row, row + 4, ..., row + 28
  const int n_rows = static_cast<int>(min(static_cast<long long>(32), static_cast<long long>(M) - row));   // valid while 4 i < n_rows
  float* po = f32 ? ep.out_f32 + row * ep.ld_out + col : nullptr;
  // operand lines of the output: FMT is the engine's format, the one the next GEMM reads
  constexpr int kShift = FMT == kFmtBf16x1 ? 6 : 5;
  const int e = col & ((1 << kShift) - 1);
  uint8_t* pl = op ? ep.out_op.base + row * ep.out_op.ld_bytes + static_cast<long long>(col >> kShift) * 128 : nullptr;
  const long long so = 4 * ep.ld_out, sl = 4 * ep.out_op.ld_bytes;
  uint32_t rd = stg + static_cast<uint32_t>(rsub) * 128u;
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    float4 x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rr = 4 * i + rsub;                       // (16 * half does not change rr & 7)
      x[i] = ld_shared_v4(rd + static_cast<uint32_t>(4 * i) * 128u + (static_cast<uint32_t>(jj ^ (rr & 7)) << 4));
    }
    rd += 16u * 128u;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      // (arithmetic is unconditional, only the memory accesses are guarded: straight-line code the scheduler can interleave)
      const bool live = 16 * half + 4 * i < n_rows;
      float v[4] = {x[i].x, x[i].y, x[i].z, x[i].w};
      if (scaled) {
        v[0] = fmaf(v[0], q.ws.x, q.b.x); v[1] = fmaf(v[1], q.ws.y, q.b.y);
        v[2] = fmaf(v[2], q.ws.z, q.b.z); v[3] = fmaf(v[3], q.ws.w, q.b.w);
      }
      if (ACT == kActGeluTanh) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = gelu_tanh_f(v[j]);
      } else if (ACT == kActGeluErf) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = gelu_erf_f(v[j]);
      }
      if (res) {
        const float4 t = half == 0 ? rres.r[i] : rres.r[4 + i];
        v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
      }
      if (affine) {  // Rescaler: w * y + b with two roundings, as torch evaluates it (modeling_hypernet.py:18-19)
        v[0] = __fadd_rn(__fmul_rn(q.cs.x, v[0]), q.ct.x); v[1] = __fadd_rn(__fmul_rn(q.cs.y, v[1]), q.ct.y);
        v[2] = __fadd_rn(__fmul_rn(q.cs.z, v[2]), q.ct.z); v[3] = __fadd_rn(__fmul_rn(q.cs.w, v[3]), q.ct.w);
      }
      if (f32) {
        if (live) {
          const float4 y = make_float4(v[0], v[1], v[2], v[3]);
          if (ep.stream_f32) __stcs(reinterpret_cast<float4*>(po), y); else *reinterpret_cast<float4*>(po) = y;
        }
        po += so;
      }
      if (op) {
        uint32_t bad_i = 0;
        const Packed4 p = pack_operand4(v, FMT, false, bad_i);
        if (live) {
          bad |= bad_i;
          *reinterpret_cast<uint2*>(pl + 2 * e) = p.m;
          if (FMT == kFmtF16F8) {
            *reinterpret_cast<uint32_t*>(pl + 64 + e) = p.s.x;
            *reinterpret_cast<uint32_t*>(pl + 96 + e) = p.t;
          } else if (FMT == kFmtBf16x3) {
            *reinterpret_cast<uint2*>(pl + 64 + 2 * e) = p.s;
          }
        }
        pl += sl;
      }
    }
  }
}

// The epilogues the hypernetwork forward actually issues get straight-line instantiations; anything else (the unit tests'
// combinations) takes the run-time-flag version.  `code` = activation * 16 + flags, uniform over the launch.
template <int FMT>
__device__ __forceinline__ void epilogue_dispatch(int code, const EpilogueParams& ep, uint32_t stg, int lane, long long row0, int M,
                                                  int col, const QuadConsts& q, const ResidualRegs& rres, uint32_t& bad) {
  switch (code) {
    case kActNone * 16 + kEpiF32:                          // QKV, last-layer K/V and Q
      epilogue_quads<FMT, kActNone, kEpiF32>(ep, stg, lane, row0, M, col, q, rres, bad); break;
    case kActNone * 16 + (kEpiF32 | kEpiRes):              // attention output and MLP down projections (+ residual)
      epilogue_quads<FMT, kActNone, kEpiF32 | kEpiRes>(ep, stg, lane, row0, M, col, q, rres, bad); break;
    case kActGeluErf * 16 + kEpiOp:                        // encoder MLP up projection
      epilogue_quads<FMT, kActGeluErf, kEpiOp>(ep, stg, lane, row0, M, col, q, rres, bad); break;
    case kActGeluTanh * 16 + kEpiOp:                       // ProjectorBlock dense1
      epilogue_quads<FMT, kActGeluTanh, kEpiOp>(ep, stg, lane, row0, M, col, q, rres, bad); break;
    case kActGeluTanh * 16 + (kEpiF32 | kEpiRes):          // ProjectorBlock dense2 (+ residual)
      epilogue_quads<FMT, kActGeluTanh, kEpiF32 | kEpiRes>(ep, stg, lane, row0, M, col, q, rres, bad); break;
    case kActNone * 16 + (kEpiF32 | kEpiOp):               // input projection Linear
      epilogue_quads<FMT, kActNone, kEpiF32 | kEpiOp>(ep, stg, lane, row0, M, col, q, rres, bad); break;
    case kActNone * 16 + (kEpiF32 | kEpiAffine):           // output heads with the Rescaler
      epilogue_quads<FMT, kActNone, kEpiF32 | kEpiAffine>(ep, stg, lane, row0, M, col, q, rres, bad); break;
    default:
      if (ep.act == kActGeluErf) epilogue_quads<FMT, kActGeluErf, kEpiGeneric>(ep, stg, lane, row0, M, col, q, rres, bad);
      else if (ep.act == kActGeluTanh) epilogue_quads<FMT, kActGeluTanh, kEpiGeneric>(ep, stg, lane, row0, M, col, q, rres, bad);
      else epilogue_quads<FMT, kActNone, kEpiGeneric>(ep, stg, lane, row0, M, col, q, rres, bad);
  }
}

// tmap_a: box {128 B, 128 rows} over A's lines;  tmap_b: box {128 B, HALVES * block_n / 2 rows} over W's lines
template <int FMT, int HALVES>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmShape s,
                    const EpilogueParams ep) {
  using L = MmaList<FMT>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // 1024-byte alignment for the 128B swizzle atoms
  const uint32_t stg_base = smem_base + static_cast<uint32_t>(s.num_stages) * s.stage_bytes;
  const uint32_t bar_base = stg_base + kStagingBytes;
  auto full_bar = [&](int i) { return bar_base + 8u * i; };
  auto empty_bar = [&](int i) { return bar_base + 8u * (kMaxStages + i); };
  auto tmem_full_bar = [&](int i) { return bar_base + 8u * (2 * kMaxStages + i); };
  auto tmem_empty_bar = [&](int i) { return bar_base + 8u * (2 * kMaxStages + 2 + i); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 4);

  const int warp = warp_index_uniform();
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();   // rank inside the CTA pair
  const bool leader = cta_rank == 0;

  // (shuffles from lane 0 tell the compiler these loaded values are warp-uniform: loop bounds and MMA operands derive from them)
  const int M = __shfl_sync(0xFFFFFFFFu, s.m_dev ? *s.m_dev : s.m_host, 0);
  constexpr int tile_m = 2 * kBlockM;
  const int m_tiles = (M + tile_m - 1) / tile_m;
  const int tile_n = s.block_n * HALVES;
  const int n_tiles = (s.n + tile_n - 1) / tile_n;
  const int total_tiles = m_tiles * n_tiles;
  const int num_kb = s.k_lines;
  const int first_tile = blockIdx.x >> 1;
  const int tile_step = gridDim.x >> 1;
  const int load_n = s.block_n >> 1;  // rows of the W tile this CTA's smem holds, per N half
  const bool prof = s.prof != nullptr;
  const long long t_start = prof ? clock64() : 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int i = 0; i < s.num_stages; ++i) {
      mbar_init(full_bar(i), 1);    // the leader's arrive.expect_tx covers the bytes landing in both CTAs of the pair
      mbar_init(empty_bar(i), 1);   // one tcgen05.commit (multicast to both CTAs)
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tmem_full_bar(i), 1);
      // one arrival per epilogue warp that reads this accumulator, in both CTAs of the pair: all eight warps
      // (HALVES == 1) or the four that own this N half (HALVES == 2)
      mbar_init(tmem_empty_bar(i), (HALVES == 2 ? 4 : 8) * 2);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<2>(tmem_slot, kTmemCols);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xFFFFFFFFu, tmem_base, 0);

  if (warp == 0) {
    // ===================================== TMA producer ==========================================
    int stage = 0;
    uint32_t phase = 0;
    long long waited = 0;
    // Wave barrier.  The CTA pairs that share an A or W tile read it from L2 within microseconds of each other only while
    // they run in lockstep; nothing keeps them there, and over the ~70 waves of a large GEMM they drift apart until a
    // tile's second reader finds it evicted (DRAM reads per wave grow with the length of the launch:
    // profiles/gemm_traffic_vs_n_r2.txt).  Every `sync_every` waves in which all pairs still have a tile, the producers
    // meet before loading the next one.
    const bool syncing = s.sync_every > 0 && s.wave_sync != nullptr;
    const int full_iters = total_tiles / tile_step;
    int iter = 0;
    for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++iter) {
      if (syncing && iter > 0 && iter < full_iters && iter % s.sync_every == 0) {
        if (lane == 0) {
          const unsigned int target = gridDim.x * static_cast<unsigned int>(iter / s.sync_every);
          volatile unsigned int* ws = s.wave_sync;
          atomicAdd(s.wave_sync, 1u);
          // Best effort: the barrier only shapes the timing.  A wait is normally microseconds; a producer that has waited
          // 2 ms (the CTAs are not all resident: SMs held by another kernel or process) switches the barrier off for the
          // rest of this launch, for everybody.
          if (ws[2] == 0u) {
            const unsigned long long t0 = globaltimer_ns();
            while (ws[0] < target && ws[2] == 0u) {
              __nanosleep(32);
              if (globaltimer_ns() - t0 > 2000000ull) ws[2] = 1u;
            }
          }
        }
        __syncwarp();
      }
      const TileCoord tc = tile_coord(tile, m_tiles, n_tiles, s.group_m, s.chunk_n);
      const int row_a = tc.m_blk * tile_m + static_cast<int>(cta_rank) * kBlockM;
      const int row_b = tc.n_blk * tile_n + static_cast<int>(cta_rank) * load_n * HALVES;
      for (int kb = 0; kb < num_kb; ++kb) {
        const long long t0 = prof ? clock64() : 0;
        mbar_wait(empty_bar(stage), phase ^ 1u, 1);
        if (prof) waited += clock64() - t0;
        if (elect_one()) {
          // only the leader arrives (expecting the bytes landing in both CTAs of the pair); the copies of both CTAs credit
          // that barrier.  Bytes may land before the expect_tx: the phase cannot complete until that arrive.
          if (leader) mbar_expect_tx(full_bar(stage), s.stage_bytes * 2u);
          const uint32_t a_dst = smem_base + static_cast<uint32_t>(stage) * s.stage_bytes;
          tma_load_2d_2sm(&tmap_a, full_bar(stage), a_dst, kb * static_cast<int>(kLineBytes), row_a, s.hint_a);
          tma_load_2d_2sm(&tmap_b, full_bar(stage), a_dst + kABytes, kb * static_cast<int>(kLineBytes), row_b, s.hint_b);
        }
        __syncwarp();
        if (++stage == s.num_stages) { stage = 0; phase ^= 1u; }
      }
    }
    if (prof && lane == 0) s.prof[static_cast<size_t>(blockIdx.x) * kProfSlots + 1] = static_cast<unsigned long long>(waited);
    if (syncing && lane == 0) {
      // the last producer to leave zeroes the counters for the next launch (everybody is past every barrier by then)
      if (atomicAdd(s.wave_sync + 1, 1u) == gridDim.x - 1) {
        s.wave_sync[0] = 0u;
        s.wave_sync[1] = 0u;
        s.wave_sync[2] = 0u;
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ============================================
    if (leader) {
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      long long waited_full = 0, waited_acc = 0;
      const uint32_t half_bytes = static_cast<uint32_t>(load_n) * kLineBytes;   // W rows of one N half in this CTA's stage
      for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++iter) {
        for (int kb = 0; kb < num_kb; ++kb) {
          if (kb == 0) {
            const long long t0 = prof ? clock64() : 0;
            if (HALVES == 2) {
              mbar_wait(tmem_empty_bar(0), (iter & 1u) ^ 1u, 2);
              mbar_wait(tmem_empty_bar(1), (iter & 1u) ^ 1u, 2);
            } else {
              mbar_wait(tmem_empty_bar(iter & 1), ((iter >> 1) & 1u) ^ 1u, 2);
            }
            if (prof) waited_acc += clock64() - t0;
          }
          const long long t1 = prof ? clock64() : 0;
          mbar_wait(full_bar(stage), phase, 3);
          if (prof) waited_full += clock64() - t1;
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a0 = smem_base + static_cast<uint32_t>(stage) * s.stage_bytes;
            const uint64_t da = make_desc(a0);
#pragma unroll
            for (int hf = 0; hf < HALVES; ++hf) {
              // accumulator: the N half of the tile (HALVES == 2) or the buffer this tile alternates to (HALVES == 1)
              const int acc = HALVES == 2 ? hf : (iter & 1);
              const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * s.block_n);
              const uint64_t db = make_desc(a0 + kABytes + static_cast<uint32_t>(hf) * half_bytes);
#pragma unroll
              for (int i = 0; i < L::kCount; ++i) {
                const uint32_t accumulate = (i > 0 || kb > 0) ? 1u : 0u;
                if (L::f8(i)) umma_f8<2>(tmem_d, da + (L::a_off(i) >> 4), db + (L::b_off(i) >> 4), s.idesc8, accumulate);
                else umma_f16<2>(tmem_d, da + (L::a_off(i) >> 4), db + (L::b_off(i) >> 4), s.idesc16, accumulate);
              }
              if (kb == num_kb - 1) umma_commit<2>(tmem_full_bar(acc), 3);   // accumulator complete -> both CTAs' epilogues
            }
            umma_commit<2>(empty_bar(stage), 3);                              // frees the stage in both CTAs
          }
          __syncwarp();
          if (++stage == s.num_stages) { stage = 0; phase ^= 1u; }
        }
      }
      if (prof && lane == 0) {
        s.prof[static_cast<size_t>(blockIdx.x) * kProfSlots + 2] = static_cast<unsigned long long>(waited_full);
        s.prof[static_cast<size_t>(blockIdx.x) * kProfSlots + 3] = static_cast<unsigned long long>(waited_acc);
      }
    }
  } else {
    // ===================================== epilogue warps ========================================
    const int quarter = warp & 3;        // TMEM lanes [32 * quarter, 32 * quarter + 32) are the ones this warp may read
    const int group = (warp - 2) >> 2;   // 0 or 1
    const uint32_t stg = stg_base + static_cast<uint32_t>(warp - 2) * 4096u;
    uint32_t bad = 0;
    long long waited = 0, busy = 0, t_ldwait = 0, t_stage = 0, t_quads = 0;
    int iter = 0;
    // a straight-line instantiation exists only for epilogues with a bias / weight scale (every Linear of the forward)
    const int epi_flags = (ep.out_f32 ? kEpiF32 : 0) | (ep.out_op.base ? kEpiOp : 0) | (ep.residual ? kEpiRes : 0) | (ep.col_scale ? kEpiAffine : 0);
    const int epi_code = (ep.bias != nullptr || ep.w_scale != nullptr) ? ep.act * 16 + epi_flags : -1;
    for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++iter) {
      const TileCoord tc = tile_coord(tile, m_tiles, n_tiles, s.group_m, s.chunk_n);
      const long long row0 = static_cast<long long>(tc.m_blk) * tile_m + static_cast<int>(cta_rank) * kBlockM + quarter * 32;
      // HALVES == 2: group g drains N half g.  HALVES == 1: both groups drain the tile's accumulator, group g taking
      // the 32-column chunks g, g + 2, ...
      const int hf = HALVES == 2 ? group : 0;
      const int acc = HALVES == 2 ? hf : (iter & 1);
      const uint32_t acc_phase = HALVES == 2 ? (iter & 1u) : ((iter >> 1) & 1u);
      const int c_first = HALVES == 2 ? 0 : 32 * group;
      const int c_step = HALVES == 2 ? 32 : 64;
      // accumulator column c holds W row (c < load_n ? CTA 0's : CTA 1's) share of this half:
      //   output column = tile origin + (c / load_n) * load_n * HALVES + hf * load_n + c % load_n   (= origin + c when HALVES == 1)
      const int col_tile = tc.n_blk * tile_n + hf * load_n;
      auto gcol = [&](int c) { return col_tile + (c < load_n ? c : c + load_n * (HALVES - 1)); };
      // the residual does not depend on the accumulator: start pulling it in while the MMAs of this tile are still running
      ResidualRegs rres;
      const bool has_res = ep.residual != nullptr;
      if (has_res) {
        for (int c = c_first; c < s.block_n; c += c_step) prefetch_residual_l2(ep, row0, M, gcol(c), 32, s.n, lane);
        if (c_first < s.block_n) load_residual(ep, row0, M, gcol(c_first) + 4 * (lane & 7), s.n, lane, rres);
      }
      const long long t0 = prof ? clock64() : 0;
      mbar_wait(tmem_full_bar(acc), acc_phase, 4);
      const long long t1 = prof ? clock64() : 0;
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * s.block_n);
      // one register chunk: its rows go to the staging tile, then the tcgen05.ld of the NEXT chunk is issued into the same
      // registers and lands while this chunk's quads are processed out of shared memory
      float v[32];
      const uint32_t wr = stg + static_cast<uint32_t>(lane) * 128u;
      const uint32_t sw = static_cast<uint32_t>(lane & 7);
      // the per-column constants of a chunk (bias, weight scale, affine) are fetched one chunk ahead: with the shared-memory
      // carve-out at its maximum there is next to no L1, so every one of these loads is an L2 round trip
      QuadConsts q_next;
      int col_next = gcol(c_first) + 4 * (lane & 7);
      if (c_first < s.block_n) {
        tmem_ld_32x32(taddr + static_cast<uint32_t>(c_first), v);
        if (col_next < s.n) q_next = load_quad_consts(ep, col_next);
      }
#pragma unroll 1
      for (int c = c_first; c < s.block_n; c += c_step) {
        const long long p0 = prof ? clock64() : 0;
        tmem_ld_wait();
        const long long p1 = prof ? clock64() : 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) st_shared_v4(wr + ((static_cast<uint32_t>(j) ^ sw) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        const QuadConsts q = q_next;
        const int col = col_next;
        if (c + c_step < s.block_n) {
          tmem_ld_32x32(taddr + static_cast<uint32_t>(c + c_step), v);
          col_next = gcol(c + c_step) + 4 * (lane & 7);
          if (col_next < s.n) q_next = load_quad_consts(ep, col_next);
        }
        __syncwarp();
        const long long p2 = prof ? clock64() : 0;
        if (col < s.n) epilogue_dispatch<FMT>(epi_code, ep, stg, lane, row0, M, col, q, rres, bad);
        if (has_res && c + c_step < s.block_n) load_residual(ep, row0, M, col_next, s.n, lane, rres);   // for the next chunk
        __syncwarp();  // the staging tile is rewritten by the next chunk
        if (prof) { t_ldwait += p1 - p0; t_stage += p2 - p1; t_quads += clock64() - p2; }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(tmem_empty_bar(acc), 0);   // the leader's barrier
      if (prof) { waited += t1 - t0; busy += clock64() - t1; }
    }
    report_saturation(ep.out_op.sat, bad);
    if (prof && warp == 2 && lane == 0) {
      s.prof[static_cast<size_t>(blockIdx.x) * kProfSlots + 4] = static_cast<unsigned long long>(waited);
      s.prof[static_cast<size_t>(blockIdx.x) * kProfSlots + 5] = static_cast<unsigned long long>(busy);
      s.prof[static_cast<size_t>(blockIdx.x) * kProfSlots + 6] = static_cast<unsigned long long>(iter);
      s.prof[static_cast<size_t>(blockIdx.x) * kProfSlots + 7] = static_cast<unsigned long long>(t_ldwait);
      s.prof[static_cast<size_t>(blockIdx.x) * kProfSlots + 8] = static_cast<unsigned long long>(t_stage);
      s.prof[static_cast<size_t>(blockIdx.x) * kProfSlots + 9] = static_cast<unsigned long long>(t_quads);
    }
  }

  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc<2>(tmem_base, kTmemCols);
  if (prof && threadIdx.x == 0) s.prof[static_cast<size_t>(blockIdx.x) * kProfSlots] = static_cast<unsigned long long>(clock64() - t_start);
}

}  // namespace zett

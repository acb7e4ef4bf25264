// Persistent, warp-specialised tcgen05 GEMM for the hypernetwork's Linear layers:
//     out[m, n] = epilogue( sum_k A[m, k] * W[n, k] )          (nn.Linear: y = x W^T + b, both operands K-major)
//
//  * operand formats (epilogue.cuh): 16-bit planes  plane 0 = round(x), plane 1 = round(x - plane 0)  issued as
//    A0*B0 + A1*B0 + A0*B1 into the SAME TMEM accumulator (n_terms == 3; n_terms == 1 issues A0*B0 only), which restores
//    ~fp32 operand precision (the 1e-3 parity budget rules out single-pass bf16); or fp16 + two e5m2
//    correction planes (f8): one kind::f16 MMA + two kind::f8f6f4 MMAs per k-step.
//  * warp 0 = TMA producer (3-D tensor maps {K, rows, plane}; 128-byte swizzle: 16-bit rows of BLOCK_K = 64, or the two
//    fp8 planes of a k-block interleaved in one line), warp 1 = MMA issuer + TMEM owner, warps 2..9 = epilogue
//    (tcgen05.ld -> bias / GELU / residual / column affine -> fp32 and/or operand planes for the next GEMM).
//  * n_halves == 1: accumulators are double-buffered in TMEM (2 x BLOCK_N columns) so the epilogue of tile i overlaps the
//    main loop of tile i+1; smem stages form an mbarrier ring; every wait is bounded by a watchdog (ptx.cuh).
//  * CG == 2 pairs two CTAs (cta_group::2, UMMA M = 256): each CTA loads its 128 rows of A and half of the B tile; the
//    leader alone arrives on the full barrier (expecting both CTAs' bytes), issues the MMAs and multicasts the commits.
//  * eight epilogue warps, two per TMEM lane quarter: the two groups of four take alternate 32-column chunks of one
//    accumulator, or one N half each when n_halves == 2 (the epilogue is instruction-issue bound -- GELU, format
//    conversion, 16-byte stores -- and with K = 768 it is as long as the main loop of a tile).
//  * CP == 2 puts two such pairs in one cluster of four CTAs working on vertically adjacent 256-row tiles of the same
//    N tile: every CTA fetches only HALF of its pair's share of the W tile and TMA-multicasts it to the CTA holding the
//    same share in the other pair (.multicast::cluster), so a cluster reads 4 A blocks + 2 W blocks from the L2 slices
//    instead of 4 + 4.  Measured: -19 % L2 slice reads, -16 % DRAM reads, but the bytes ARRIVING at
//    each SM are unchanged and only 33 clusters of four fit the 148 SMs, so it ends within 2 % of plain pairs; kept as
//    gemm_impl 4, not the default.  A smem stage of a CTA is then written by two CTAs, so every empty barrier counts one
//    tcgen05.commit from EACH pair leader (multicast to all four CTAs); the full barriers stay per pair.
//  * n_halves == 2 (gemm_impl 5) widens the tile of a CTA pair to 256 x 512: both accumulators of TMEM hold the two N halves
//    of ONE tile, every k-block of A is fetched once for 512 columns (25 % fewer operand bytes per FLOP, the largest item of
//    the kernel's energy budget after the MMAs), at the price of two 96 KB smem stages instead of
//    three 64 KB ones and an epilogue that no longer overlaps the next tile's main loop.
//  * M may live in device memory (packed-position counts are data dependent); tiles beyond it are skipped.
#pragma once
#include <cuda.h>
#include "ptx.cuh"
#include "epilogue.cuh"

namespace zett {

constexpr int kBlockM = 128;
constexpr int kMaxBlockK = 64;      // 64 x 2 B = one 128-byte swizzle row; block_k = 32 halves the rows (more, finer stages)
constexpr int kUmmaK = 16;
constexpr int kMaxStages = 8;
constexpr int kEpilogueWarps = 8;    // two per TMEM lane quarter
constexpr int kGemmThreads = 64 + 32 * kEpilogueWarps;   // TMA producer warp, MMA warp, epilogue warps
constexpr uint32_t kTmemCols = 512;

struct GemmShape {
  int m_host;          // rows of A / out when m_dev == nullptr
  const int* m_dev;    // optional device-resident row count
  int n, k;
  int block_n;         // UMMA N: 32, 64, 128 or 256
  int n_halves;        // 1: tile N = block_n, accumulators double-buffered; 2: tile N = 2 * block_n (block_n == 256 only)
  int group_m;         // rasterisation: m-tiles per group
  int chunk_n;         // rasterisation: n-tiles per L2-resident W chunk
  int block_k;         // K elements per pipeline stage: 64 (128-byte rows of 16-bit operands) or 32
  int n_terms;         // 1 or 3
  int n_planes;        // planes held by the 16-bit tensor maps' boxes (1 or 2)
  int f8;              // 1: operand format kFmtF16F8 -- one fp16 plane + two e5m2 correction planes (epilogue.cuh)
  int mma_mask;        // diagnostic (ZETT_MMA_MASK, default 7): bit 0 main term, bit 1 16-bit correction terms, bit 2 fp8 terms
  uint64_t hint_a, hint_b;   // L2 eviction policies of the A / W loads (ptx.cuh)
  uint32_t idesc;      // tcgen05 instruction descriptor (kind::f16)
  uint32_t idesc8;     // tcgen05 instruction descriptor (kind::f8f6f4), f8 only
  int num_stages;
  uint32_t stage_bytes, a_plane_bytes, b_plane_bytes;   // 16-bit planes: 128-byte rows
  uint32_t a8_bytes, b8_bytes;                          // interleaved fp8 planes: 128-byte rows (q0[64] | q1[64])
};

struct TileCoord { int m_blk, n_blk; };

// Rasterisation for L2 residency (the main loop is bound by operand-fetch latency x limited smem buffering, so L2 hits
// matter more than anything else): N is cut into chunks whose W panel fits comfortably in L2 and stays there while
// all of M sweeps past it in small m-groups; inside a group m varies fastest, so the CTAs of a wave share W tiles
// (one DRAM fetch, the rest L2 hits) and each A tile is fetched once per chunk.
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

// tmap_a / tmap_a8: box {block_k, 128 rows, all planes};  tmap_b / tmap_b8: box {block_k, block_n / (CG * CP) rows,
// all planes (CP == 1) or ONE plane (CP == 2: a multicast share lands inside each plane of the [plane][row] smem layout)}
template <int CG, int CP>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_a8, const __grid_constant__ CUtensorMap tmap_b8,
                    const GemmShape s, const EpilogueParams ep) {
  static_assert(CP == 1 || CG == 2, "pairs of pairs need cta_group::2");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment for the 128B swizzle atoms
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + s.num_stages * s.stage_bytes;
  auto full_bar = [&](int i) { return bar_base + 8u * i; };
  auto empty_bar = [&](int i) { return bar_base + 8u * (kMaxStages + i); };
  auto tmem_full_bar = [&](int i) { return bar_base + 8u * (2 * kMaxStages + i); };
  auto tmem_empty_bar = [&](int i) { return bar_base + 8u * (2 * kMaxStages + 2 + i); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cluster_rank = (CG * CP > 1) ? cluster_ctarank() : 0u;
  const uint32_t cta_rank = cluster_rank & (CG - 1);   // rank inside the CTA pair (cta_group::2 pairs ranks 2p, 2p + 1)
  const uint32_t pair = cluster_rank / CG;             // pair inside the cluster
  const bool leader = cta_rank == 0;

  const int M = s.m_dev ? *s.m_dev : s.m_host;
  const int tile_m = kBlockM * CG * CP;                // rows of one cluster tile
  const int m_tiles = (M + tile_m - 1) / tile_m;
  const int tile_n = s.block_n * s.n_halves;
  const int n_tiles = (s.n + tile_n - 1) / tile_n;
  const int total_tiles = m_tiles * n_tiles;
  const int num_kb = (s.k + s.block_k - 1) / s.block_k;
  const int first_tile = blockIdx.x / (CG * CP);
  const int tile_step = gridDim.x / (CG * CP);
  const int load_n = s.block_n / CG;  // rows of the W tile this CTA's smem holds, per N half
  const int cta_row0 = static_cast<int>(pair) * kBlockM * CG + static_cast<int>(cta_rank) * kBlockM;  // inside the cluster tile

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    if (s.f8) {
      prefetch_tmap(&tmap_a8);
      prefetch_tmap(&tmap_b8);
    }
    for (int i = 0; i < s.num_stages; ++i) {
      mbar_init(full_bar(i), 1);    // the leader's arrive.expect_tx covers the bytes landing in both CTAs of the pair
      mbar_init(empty_bar(i), CP);  // one tcgen05.commit per pair that reads or multicasts into this stage
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tmem_full_bar(i), 1);
      // every epilogue thread that reads this accumulator, in every CTA of the pair: both warp groups (n_halves == 1)
      // or the one group that owns this N half (n_halves == 2)
      mbar_init(tmem_empty_bar(i), (s.n_halves == 2 ? 128 : 256) * CG);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<CG>(tmem_slot, kTmemCols);
  tc_fence_before();
  if constexpr (CG * CP > 1) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===================================== TMA producer ==========================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        const TileCoord tc = tile_coord(tile, m_tiles, n_tiles, s.group_m, s.chunk_n);
        const int row_a = tc.m_blk * tile_m + cta_row0;
        const int row_b = tc.n_blk * tile_n + static_cast<int>(cta_rank) * load_n * s.n_halves;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u, 1);
          const uint32_t a_dst = smem_base + stage * s.stage_bytes;
          const uint32_t b_dst = a_dst + s.n_planes * s.a_plane_bytes;
          const uint32_t a8_dst = b_dst + s.n_planes * s.b_plane_bytes;
          const uint32_t b8_dst = a8_dst + s.a8_bytes;
          if constexpr (CG == 1) {
            mbar_expect_tx(full_bar(stage), s.stage_bytes);
            tma_load_3d(&tmap_a, full_bar(stage), a_dst, kb * s.block_k, row_a, 0, s.hint_a);
            tma_load_3d(&tmap_b, full_bar(stage), b_dst, kb * s.block_k, row_b, 0, s.hint_b);
            if (s.f8) {
              tma_load_3d(&tmap_a8, full_bar(stage), a8_dst, kb * 128, row_a, 0, s.hint_a);
              tma_load_3d(&tmap_b8, full_bar(stage), b8_dst, kb * 128, row_b, 0, s.hint_b);
            }
          } else {
            // only the leader arrives (expecting the bytes landing in both CTAs of its pair); the other copies credit
            // that barrier directly, so no other loop carries a cluster-scope operation.  Bytes may land before the
            // leader's expect_tx: the phase cannot complete until that arrive, and the transaction count is signed.
            if (leader) mbar_expect_tx(full_bar(stage), s.stage_bytes * 2u);
            tma_load_3d_2sm(&tmap_a, full_bar(stage), a_dst, kb * s.block_k, row_a, 0, s.hint_a);
            if (s.f8) tma_load_3d_2sm(&tmap_a8, full_bar(stage), a8_dst, kb * 128, row_a, 0, s.hint_a);
            if constexpr (CP == 1) {
              tma_load_3d_2sm(&tmap_b, full_bar(stage), b_dst, kb * s.block_k, row_b, 0, s.hint_b);
              if (s.f8) tma_load_3d_2sm(&tmap_b8, full_bar(stage), b8_dst, kb * 128, row_b, 0, s.hint_b);
            } else {
              // this CTA fetches rows [pair * load_n / 2, +load_n / 2) of its share, plane by plane, for itself and for the
              // CTA of equal pair rank in the other pair; the other half arrives from there.  Each copy credits the full
              // barrier of the destination's own pair leader.
              const int share = load_n / CP;
              const int row_s = row_b + static_cast<int>(pair) * share;
              const uint16_t mask = static_cast<uint16_t>(0x5u << cta_rank);
              const uint32_t off16 = pair * static_cast<uint32_t>(share) * static_cast<uint32_t>(s.block_k) * 2u;
              for (int pl = 0; pl < s.n_planes; ++pl)
                tma_load_3d_2sm_mc(&tmap_b, full_bar(stage), b_dst + pl * s.b_plane_bytes + off16, kb * s.block_k, row_s, pl, mask, s.hint_b);
              if (s.f8)
                tma_load_3d_2sm_mc(&tmap_b8, full_bar(stage), b8_dst + pair * static_cast<uint32_t>(share) * 128u, kb * 128, row_s, 0, mask, s.hint_b);
            }
          }
          if (++stage == s.num_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ============================================
    if (lane == 0 && leader) {
      const uint16_t kAllMask = static_cast<uint16_t>((1u << (CG * CP)) - 1u);
      const uint16_t kPairMask = static_cast<uint16_t>(((1u << CG) - 1u) << (pair * CG));
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++iter) {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase, 3);
          tc_fence_after();
          const uint32_t a0 = smem_base + stage * s.stage_bytes;
          const uint32_t b0 = a0 + s.n_planes * s.a_plane_bytes;
          const uint32_t row16 = static_cast<uint32_t>(s.block_k) * 2u;  // bytes per row of a 16-bit plane
          const uint64_t da0 = umma_desc_kmajor(a0, row16), da1 = umma_desc_kmajor(a0 + s.a_plane_bytes, row16);
          const uint32_t a8 = b0 + s.n_planes * s.b_plane_bytes;
          const uint64_t dqa = umma_desc_kmajor(a8, 128u);
          const int ksteps = s.block_k / kUmmaK;
          for (int hf = 0; hf < s.n_halves; ++hf) {
            // accumulator: the N half of the tile (n_halves == 2) or the buffer this tile alternates to (n_halves == 1)
            const int acc = s.n_halves == 2 ? hf : (iter & 1);
            if (kb == 0) {
              const uint32_t acc_phase = s.n_halves == 2 ? (iter & 1u) : ((iter >> 1) & 1u);
              mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1u, 2);
              tc_fence_after();
            }
            const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * s.block_n);
            const uint32_t bh = b0 + static_cast<uint32_t>(hf * load_n) * row16;  // this half's rows of the W planes
            const uint64_t db0 = umma_desc_kmajor(bh, row16), db1 = umma_desc_kmajor(bh + s.b_plane_bytes, row16);
#pragma unroll 4
            for (int kk = 0; kk < ksteps; ++kk) {
              const uint64_t koff = static_cast<uint64_t>((kk * kUmmaK * 2) >> 4);  // 32 B per K step inside the atom
              if (s.mma_mask & 1) umma_f16<CG>(tmem_d, da0 + koff, db0 + koff, s.idesc, (kb | kk) != 0);
              if (s.n_terms == 3 && !s.f8 && (s.mma_mask & 2)) {
                umma_f16<CG>(tmem_d, da1 + koff, db0 + koff, s.idesc, 1u);
                umma_f16<CG>(tmem_d, da0 + koff, db1 + koff, s.idesc, 1u);
              }
            }
            if (s.f8 && (s.mma_mask & 4)) {  // first-order corrections at fp8 rate: Aq0 . Wq0 + Aq1 . Wq1, K = 32 per instruction
              const uint64_t dqb = umma_desc_kmajor(a8 + s.a8_bytes + static_cast<uint32_t>(hf * load_n) * 128u, 128u);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)  // bytes [0, 64) of a row are q0 of the k-block, [64, 128) are q1
                umma_f8<CG>(tmem_d, dqa + 2u * kk, dqb + 2u * kk, s.idesc8, (s.mma_mask & 1) | kb | kk);
            }
            if (kb == num_kb - 1) umma_commit<CG>(tmem_full_bar(acc), kPairMask);  // accumulator complete (this pair)
          }
          umma_commit<CG>(empty_bar(stage), kAllMask);                            // frees the stage in every CTA of the cluster
          if (++stage == s.num_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ===================================== epilogue warps ========================================
    const int quarter = warp & 3;        // TMEM lanes [32 * quarter, 32 * quarter + 32) are the ones this warp may read
    const int group = (warp - 2) >> 2;   // 0 or 1
    int iter = 0;
    for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++iter) {
      const TileCoord tc = tile_coord(tile, m_tiles, n_tiles, s.group_m, s.chunk_n);
      const int row = tc.m_blk * tile_m + cta_row0 + quarter * 32 + lane;
      const bool row_ok = row < M;
      // n_halves == 2: group g drains N half g.  n_halves == 1: both groups drain the tile's accumulator, group g taking
      // the 32-column chunks g, g + 2, ...
      const int hf = s.n_halves == 2 ? group : 0;
      const int acc = s.n_halves == 2 ? hf : (iter & 1);
      const uint32_t acc_phase = s.n_halves == 2 ? (iter & 1u) : ((iter >> 1) & 1u);
      const int c_first = s.n_halves == 2 ? 0 : 32 * group;
      const int c_step = s.n_halves == 2 ? 32 : 64;
      mbar_wait(tmem_full_bar(acc), acc_phase, 4);
      tc_fence_after();
      // accumulator column c holds W row (c < load_n ? CTA 0's : CTA 1's) share of this half:
      //   output column = tile origin + (c / load_n) * load_n * n_halves + hf * load_n + c % load_n   (= origin + c when n_halves == 1)
      const int col_tile = tc.n_blk * tile_n + hf * load_n;
      auto gcol = [&](int c) { return col_tile + (c < load_n ? c : c + load_n * (s.n_halves - 1)); };
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * s.block_n);
      // two register chunks: the tcgen05.ld of the next chunk is in flight while the current one is processed
      float va[32], vb[32];
      if (c_first < s.block_n) tmem_ld_32x32(taddr + static_cast<uint32_t>(c_first), va);
      for (int c = c_first; c < s.block_n; c += 2 * c_step) {
        const int c2 = c + c_step, c3 = c2 + c_step;
        tmem_ld_wait();
        if (c2 < s.block_n) tmem_ld_32x32(taddr + static_cast<uint32_t>(c2), vb);
        if (row_ok && gcol(c) < s.n) epilogue_store32(ep, row, gcol(c), min(32, s.n - gcol(c)), va);
        if (c2 < s.block_n) {
          tmem_ld_wait();
          if (c3 < s.block_n) tmem_ld_32x32(taddr + static_cast<uint32_t>(c3), va);
          if (row_ok && gcol(c2) < s.n) epilogue_store32(ep, row, gcol(c2), min(32, s.n - gcol(c2)), vb);
        }
      }
      tc_fence_before();
      if constexpr (CG == 1) mbar_arrive(tmem_empty_bar(acc));
      else mbar_arrive_cluster(tmem_empty_bar(acc), pair * CG);
    }
  }

  __syncwarp();
  tc_fence_before();
  if constexpr (CG * CP > 1) cluster_sync_all(); else __syncthreads();
  if (warp == 1) tmem_dealloc<CG>(tmem_base, kTmemCols);
}

}  // namespace zett

// Operand formats of the GEMM engine.
//
// Every GEMM operand (activations [rows, K] and nn.Linear weights [N, K], both K-major) lives in HBM as rows of
// 128-byte LINES; one line holds everything the tensor cores need for a group of consecutive K-elements of one row, so
// a TMA box {128 bytes, rows} with the 128-byte swizzle is one pipeline stage made of whole L2 lines, and every
// tcgen05.mma of the stage addresses its operand as a 32-byte K-slice inside the swizzled row:
//
//   kFmtF16F8  (default, 4 B/element, 32 elements per line)
//       bytes [ 0, 64)  p0[32] = fp16(x)
//       bytes [64, 96)  q0[32] e5m2        activations: 2^6 (x - p0)     weights: 2^-6 w
//       bytes [96,128)  q1[32] e5m2        activations: 2^-6 x           weights: 2^6 (w - p0)
//     A.W ~= A_p0 W_p0 (two kind::f16 MMAs, K = 16 each) + A_q0 W_q0 + A_q1 W_q1 (two kind::f8f6f4 MMAs, K = 32 each, at
//     twice the fp16 rate): the first-order rounding errors of both fp16 operands, to e5m2 accuracy.  The 2^+-6 factors
//     cancel inside each product and keep both fp8 operands in e5m2's normal range.
//   kFmtBf16x3 (4 B/element, 32 per line)    bytes [0,64) hi[32] = bf16(x), bytes [64,128) lo[32] = bf16(x - hi)
//     A.W ~= A_hi W_hi + A_lo W_hi + A_hi W_lo (six kind::f16 MMAs per line); fp32's exponent range, ~16 mantissa bits.
//   kFmtBf16x1 (2 B/element, 64 per line)    plain bf16; one pass, misses the 1e-3 parity budget: probes only.
//
// Rows are padded to whole lines (ld_bytes = ceil(K / elements-per-line) * 128) and the padding stays zero (buffers are
// zeroed once when they are allocated; writers touch valid elements only), so K only has to be a multiple of 8.
//
// Weights of kFmtF16F8 are stored pre-scaled by a power of two per output row (row maximum brought into [16, 32)): fp16
// and e5m2 have 5 exponent bits, and a trained checkpoint's rows span more dynamic range than N(0, 1/fan_in) does.  The
// GEMM epilogue multiplies column n by the inverse scale (exact).  Activations are not rescaled; a value beyond fp16's
// range is COUNTED (EpilogueParams::sat and friends) and the host falls back to kFmtBf16x3 (zett_hn_check).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cstdint>

namespace zett {

enum SplitFmt : int { kFmtBf16x3 = 0, kFmtBf16x1 = 1, kFmtF16F8 = 2 };
constexpr float kF8Up = 64.0f, kF8Down = 0.015625f;
constexpr float kF16Max = 65504.0f;

__host__ __device__ inline int fmt_elems_per_line(int fmt) { return fmt == kFmtBf16x1 ? 64 : 32; }
__host__ __device__ inline long long operand_lines(int fmt, long long k) {
  const int e = fmt_elems_per_line(fmt);
  return (k + e - 1) / e;
}
__host__ __device__ inline long long operand_ld_bytes(int fmt, long long k) { return operand_lines(fmt, k) * 128; }

// where a producer writes an operand: rows of `ld_bytes`, format `fmt`; `sat` (nullable) counts values outside fp16's range
struct OperandOut {
  uint8_t* base;
  long long ld_bytes;
  int fmt;
  unsigned int* sat;
};

__device__ __forceinline__ float e5m2_to_float(uint8_t v) {
  const __half_raw hr = __nv_cvt_fp8_to_halfraw(v, __NV_E5M2);
  return __half2float(__half(hr));
}

// The 16 bytes four consecutive values occupy in a line, as {main-plane 8 bytes, second 4 or 8, third 4}:
//   kFmtF16F8 : m = p0[4], s.x = q0[4], t = q1[4]        kFmtBf16x3: m = hi[4], s = lo[4]        kFmtBf16x1: m only
struct Packed4 {
  uint2 m;
  uint2 s;
  uint32_t t;
};

__device__ __forceinline__ uint32_t half2_bits(const __half2 h) { return *reinterpret_cast<const uint32_t*>(&h); }

__device__ __forceinline__ Packed4 pack_operand4(const float* y, int fmt, bool is_weight, uint32_t& bad) {
  Packed4 r;
  r.s = make_uint2(0u, 0u);
  r.t = 0u;
  if (fmt == kFmtF16F8) {
    // paired conversions (cvt.rn.f16x2.f32, cvt.rn.satfinite.e5m2x2.f32).  No clamp: a value beyond fp16's range is
    // counted, and a forward that counted any is repeated in the bf16 format (zett_hn_check), so what is stored for it
    // does not matter.
    const float m = fmaxf(fmaxf(fabsf(y[0]), fabsf(y[1])), fmaxf(fabsf(y[2]), fabsf(y[3])));
    bad |= (m > kF16Max) ? 1u : 0u;
    const __half2 h01 = __floats2half2_rn(y[0], y[1]), h23 = __floats2half2_rn(y[2], y[3]);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const float2 lo01 = make_float2((y[0] - f01.x) * kF8Up, (y[1] - f01.y) * kF8Up);
    const float2 lo23 = make_float2((y[2] - f23.x) * kF8Up, (y[3] - f23.y) * kF8Up);
    const float2 dn01 = make_float2(y[0] * kF8Down, y[1] * kF8Down), dn23 = make_float2(y[2] * kF8Down, y[3] * kF8Down);
    const uint32_t lo = uint32_t(__nv_cvt_float2_to_fp8x2(lo01, __NV_SATFINITE, __NV_E5M2)) |
                        (uint32_t(__nv_cvt_float2_to_fp8x2(lo23, __NV_SATFINITE, __NV_E5M2)) << 16);
    const uint32_t dn = uint32_t(__nv_cvt_float2_to_fp8x2(dn01, __NV_SATFINITE, __NV_E5M2)) |
                        (uint32_t(__nv_cvt_float2_to_fp8x2(dn23, __NV_SATFINITE, __NV_E5M2)) << 16);
    r.m.x = half2_bits(h01); r.m.y = half2_bits(h23);
    r.s.x = is_weight ? dn : lo;   // q0: weights 2^-6 w        activations 2^6 (x - p0)
    r.t = is_weight ? lo : dn;     // q1: weights 2^6 (w - p0)   activations 2^-6 x
  } else {
    const __nv_bfloat162 h01 = __floats2bfloat162_rn(y[0], y[1]), h23 = __floats2bfloat162_rn(y[2], y[3]);
    r.m.x = *reinterpret_cast<const uint32_t*>(&h01); r.m.y = *reinterpret_cast<const uint32_t*>(&h23);
    if (fmt == kFmtBf16x3) {
      const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
      const __nv_bfloat162 l01 = __floats2bfloat162_rn(y[0] - f01.x, y[1] - f01.y), l23 = __floats2bfloat162_rn(y[2] - f23.x, y[3] - f23.y);
      r.s.x = *reinterpret_cast<const uint32_t*>(&l01); r.s.y = *reinterpret_cast<const uint32_t*>(&l23);
    }
  }
  return r;
}

// address of the line holding element (row, col) and the element's index inside it
__device__ __forceinline__ uint8_t* operand_line(const OperandOut& o, long long row, int col, int& e) {
  if (o.fmt == kFmtBf16x1) {
    e = col & 63;
    return o.base + row * o.ld_bytes + static_cast<long long>(col >> 6) * 128;
  }
  e = col & 31;
  return o.base + row * o.ld_bytes + static_cast<long long>(col >> 5) * 128;
}

__device__ __forceinline__ void store_packed4(const OperandOut& o, long long row, int col, const Packed4& p) {
  int e;
  uint8_t* line = operand_line(o, row, col, e);
  *reinterpret_cast<uint2*>(line + 2 * e) = p.m;
  if (o.fmt == kFmtF16F8) {
    *reinterpret_cast<uint32_t*>(line + 64 + e) = p.s.x;
    *reinterpret_cast<uint32_t*>(line + 96 + e) = p.t;
  } else if (o.fmt == kFmtBf16x3) {
    *reinterpret_cast<uint2*>(line + 64 + 2 * e) = p.s;
  }
}

// the same with the format as a compile-time constant (producers that are instantiated per format)
template <int FMT>
__device__ __forceinline__ void store_packed4_as(const OperandOut& o, long long row, int col, const Packed4& p) {
  constexpr int kShift = FMT == kFmtBf16x1 ? 6 : 5;
  const int e = col & ((1 << kShift) - 1);
  uint8_t* line = o.base + row * o.ld_bytes + static_cast<long long>(col >> kShift) * 128;
  *reinterpret_cast<uint2*>(line + 2 * e) = p.m;
  if (FMT == kFmtF16F8) {
    *reinterpret_cast<uint32_t*>(line + 64 + e) = p.s.x;
    *reinterpret_cast<uint32_t*>(line + 96 + e) = p.t;
  } else if (FMT == kFmtBf16x3) {
    *reinterpret_cast<uint2*>(line + 64 + 2 * e) = p.s;
  }
}

// A producer that walks a row in quads a fixed distance apart (LayerNorm: quad q, q + T, q + 2T ... with T a multiple of
// 16) resolves the line arithmetic once: `m` points at the quad's main-plane bytes, `s` at its second-plane bytes, and
// the quad `step` places further on sits operand_quad_stride<FMT>(T) * step bytes behind both.
struct OperandCursor {
  uint8_t* m;
  uint8_t* s;
};
template <int FMT>
__device__ __forceinline__ OperandCursor operand_cursor(const OperandOut& o, long long row, int quad) {
  constexpr int kQuadsPerLine = FMT == kFmtBf16x1 ? 16 : 8;
  uint8_t* line = o.base + row * o.ld_bytes + static_cast<long long>(quad / kQuadsPerLine) * 128;
  const int e = quad % kQuadsPerLine;
  OperandCursor c;
  c.m = line + 8 * e;
  c.s = line + 64 + (FMT == kFmtF16F8 ? 4 : 8) * e;
  return c;
}
template <int FMT>
__host__ __device__ constexpr int operand_quad_stride(int quads) { return quads * (FMT == kFmtBf16x1 ? 8 : 16); }
template <int FMT>
__device__ __forceinline__ void store_packed4_cursor(const OperandCursor& c, int byte_offset, const Packed4& p) {
  *reinterpret_cast<uint2*>(c.m + byte_offset) = p.m;
  if (FMT == kFmtF16F8) {
    *reinterpret_cast<uint32_t*>(c.s + byte_offset) = p.s.x;
    *reinterpret_cast<uint32_t*>(c.s + 32 + byte_offset) = p.t;
  } else if (FMT == kFmtBf16x3) {
    *reinterpret_cast<uint2*>(c.s + byte_offset) = p.s;
  }
}

// 4 consecutive elements of a row (col a multiple of 4)
__device__ __forceinline__ void store_operand4(const OperandOut& o, long long row, int col, const float* y, bool is_weight,
                                               uint32_t& bad) {
  store_packed4(o, row, col, pack_operand4(y, o.fmt, is_weight, bad));
}

// 8 consecutive elements (col a multiple of 8): 16-byte stores for the main plane
__device__ __forceinline__ void store_operand8(const OperandOut& o, long long row, int col, const float* y, bool is_weight,
                                               uint32_t& bad) {
  const Packed4 a = pack_operand4(y, o.fmt, is_weight, bad), b = pack_operand4(y + 4, o.fmt, is_weight, bad);
  int e;
  uint8_t* line = operand_line(o, row, col, e);
  *reinterpret_cast<uint4*>(line + 2 * e) = make_uint4(a.m.x, a.m.y, b.m.x, b.m.y);
  if (o.fmt == kFmtF16F8) {
    *reinterpret_cast<uint2*>(line + 64 + e) = make_uint2(a.s.x, b.s.x);
    *reinterpret_cast<uint2*>(line + 96 + e) = make_uint2(a.t, b.t);
  } else if (o.fmt == kFmtBf16x3) {
    *reinterpret_cast<uint4*>(line + 64 + 2 * e) = make_uint4(a.s.x, a.s.y, b.s.x, b.s.y);
  }
}

// a warp reports the values it found outside fp16's range (one atomic per warp that saw any)
__device__ __forceinline__ void report_saturation(unsigned int* sat, uint32_t bad) {
  if (sat == nullptr) return;
  const uint32_t any = __ballot_sync(__activemask(), bad != 0);
  if (any != 0 && (threadIdx.x & 31) == (__ffs(any) - 1)) atomicAdd(sat, static_cast<unsigned int>(__popc(any)));
}

// value of plane `pl` (0 main; 1, 2 correction planes) of element (row, k) of an operand (SIMT checker GEMM)
__device__ __forceinline__ float operand_plane(const uint8_t* base, long long ld_bytes, int fmt, long long row, int k, int pl) {
  if (fmt == kFmtBf16x1) {
    if (pl != 0) return 0.f;
    const uint8_t* line = base + row * ld_bytes + static_cast<long long>(k >> 6) * 128;
    return __bfloat162float(__ushort_as_bfloat16(*reinterpret_cast<const uint16_t*>(line + 2 * (k & 63))));
  }
  const uint8_t* line = base + row * ld_bytes + static_cast<long long>(k >> 5) * 128;
  const int e = k & 31;
  if (fmt == kFmtF16F8) {
    if (pl == 0) return __half2float(__ushort_as_half(*reinterpret_cast<const uint16_t*>(line + 2 * e)));
    return e5m2_to_float(line[(pl == 1 ? 64 : 96) + e]);
  }
  if (pl == 2) return 0.f;
  return __bfloat162float(__ushort_as_bfloat16(*reinterpret_cast<const uint16_t*>(line + (pl == 1 ? 64 : 0) + 2 * e)));
}

}  // namespace zett

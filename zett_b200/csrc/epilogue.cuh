// Fused GEMM epilogue shared by the tcgen05 kernel and the SIMT debug kernel:
//     y = act(acc + bias[n]);  y += residual[m, n];  y = col_scale[n] * y + col_shift[n]
// then store fp32 and/or the 16-bit split planes (plane 0 = round(y), plane 1 = round(y - plane0)) that the
// next GEMM consumes as its A operand.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cstdint>

namespace zett {

enum Act : int { kActNone = 0, kActGeluTanh = 1, kActGeluErf = 2 };
// Operand formats of the GEMM engine.  An operand x is stored as planes whose products restore ~fp32 accuracy:
//   kFmtBf16 / kFmtFp16 : p0 = round16(x), p1 = round16(x - p0);        A.W ~= A0 W0 + A1 W0 + A0 W1      (3 f16-kind MMAs)
//   kFmtF16F8           : p0 = fp16(x) and two e5m2 planes that carry the first-order corrections at fp8 rate
//                           activations: q0 = e5m2(2^6 (x - p0)),  q1 = e5m2(2^-6 x)
//                           weights    : q0 = e5m2(2^-6 w),        q1 = e5m2(2^6 (w - p0))
//                         A.W ~= A0 W0 (f16 MMA) + Aq0 Wq0 + Aq1 Wq1 (two f8f6f4 MMAs, half the cost each);
//                         the 2^+-6 factors cancel inside each product and keep both fp8 operands in e5m2's normal range.
//   The two fp8 planes live where plane 1 lives (same 2 bytes per element), INTERLEAVED per block of 64 K-elements so
//   that one 128-byte line holds q0[64] | q1[64] of a (row, k-block): the TMA boxes of the fp8 planes then fetch whole
//   lines with the 128-byte swizzle, like the 16-bit planes (two 64-byte half-line planes cost 1.5x the L2 requests).
//   Byte address of element offset `off` (row * K + k, K a multiple of 64):  q0 at 2 * off - (off & 63), q1 64 further.
enum SplitFmt : int { kFmtBf16 = 0, kFmtFp16 = 1, kFmtF16F8 = 2 };
constexpr float kF8Up = 64.0f, kF8Down = 0.015625f;

struct EpilogueParams {
  const float* bias;        // [n] or nullptr
  int act;
  const float* residual;    // fp32 [m, ld_res] or nullptr (added after the activation)
  long long ld_res;
  const float* col_scale;   // [n] or nullptr: Rescaler  y = w * y + b  (hf_hypernet/modeling_hypernet.py:9-19)
  const float* col_shift;
  float* out_f32;           // nullable, [m, ld_out]
  long long ld_out;
  uint16_t* out_p0;         // nullable split planes, [m, ld_split] each
  uint16_t* out_p1;
  long long ld_split;
  int split_fmt;
  int stream_f32;           // 1: fp32 output stored with the streaming (evict-first) hint, it is larger than L2
};

// F.gelu(x, approximate="tanh")  (ProjectorBlock, hf_hypernet/modeling_hypernet.py:36-39)
__device__ __forceinline__ float gelu_tanh_f(float x) {
  const float c = 0.7978845608028654f;  // sqrt(2/pi)
  return 0.5f * x * (1.0f + tanhf(c * (x + 0.044715f * x * x * x)));
}
// exact GELU, hidden_act="gelu" of the RoBERTa encoder
__device__ __forceinline__ float gelu_erf_f(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.7071067811865476f));
}

// kFmtF16F8: returns p0 (fp16 bits) and the two e5m2 bytes
__device__ __forceinline__ void split_f16f8(float x, bool is_weight, uint16_t& p0, uint8_t& q0, uint8_t& q1) {
  const float xc = fminf(fmaxf(x, -65504.f), 65504.f);
  const __half h = __float2half_rn(xc);
  const float lo = x - __half2float(h);
  p0 = __half_as_ushort(h);
  if (is_weight) {
    q0 = __nv_cvt_float_to_fp8(x * kF8Down, __NV_SATFINITE, __NV_E5M2);
    q1 = __nv_cvt_float_to_fp8(lo * kF8Up, __NV_SATFINITE, __NV_E5M2);
  } else {
    q0 = __nv_cvt_float_to_fp8(lo * kF8Up, __NV_SATFINITE, __NV_E5M2);
    q1 = __nv_cvt_float_to_fp8(x * kF8Down, __NV_SATFINITE, __NV_E5M2);
  }
}

__device__ __forceinline__ float e5m2_to_float(uint8_t v) {
  const __half_raw hr = __nv_cvt_fp8_to_halfraw(v, __NV_E5M2);
  return __half2float(__half(hr));
}

// byte offset of q0 of the element at offset `off` inside the interleaved fp8 region (q1 = +64)
__device__ __forceinline__ long long f8_offset(long long off) { return 2 * off - (off & 63); }

__device__ __forceinline__ uint32_t pack4_u8(const uint8_t* b) {
  return uint32_t(b[0]) | (uint32_t(b[1]) << 8) | (uint32_t(b[2]) << 16) | (uint32_t(b[3]) << 24);
}

// Store 4 consecutive elements (offset `off`, a multiple of 4) of an operand in format `fmt`.
__device__ __forceinline__ void store_operand4(uint16_t* p0, uint16_t* p1, long long off, const float* y, int fmt, bool is_weight) {
  if (fmt == kFmtF16F8) {
    uint16_t a[4];
    uint8_t b[4], c[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_f16f8(y[i], is_weight, a[i], b[i], c[i]);
    uint2 pa;
    pa.x = a[0] | (uint32_t(a[1]) << 16); pa.y = a[2] | (uint32_t(a[3]) << 16);
    *reinterpret_cast<uint2*>(p0 + off) = pa;
    uint8_t* q0 = reinterpret_cast<uint8_t*>(p1) + f8_offset(off);
    *reinterpret_cast<uint32_t*>(q0) = pack4_u8(b);
    *reinterpret_cast<uint32_t*>(q0 + 64) = pack4_u8(c);
  } else {
    uint16_t a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (fmt == kFmtBf16) {
        const __nv_bfloat16 h = __float2bfloat16_rn(y[i]);
        a[i] = __bfloat16_as_ushort(h);
        b[i] = __bfloat16_as_ushort(__float2bfloat16_rn(y[i] - __bfloat162float(h)));
      } else {
        const __half h = __float2half_rn(y[i]);
        a[i] = __half_as_ushort(h);
        b[i] = __half_as_ushort(__float2half_rn(y[i] - __half2float(h)));
      }
    }
    uint2 pa, pb;
    pa.x = a[0] | (uint32_t(a[1]) << 16); pa.y = a[2] | (uint32_t(a[3]) << 16);
    pb.x = b[0] | (uint32_t(b[1]) << 16); pb.y = b[2] | (uint32_t(b[3]) << 16);
    *reinterpret_cast<uint2*>(p0 + off) = pa;
    if (p1) *reinterpret_cast<uint2*>(p1 + off) = pb;
  }
}

// 8 consecutive elements (offset a multiple of 8): 16-byte stores for the 16-bit planes, 8-byte for the fp8 planes
__device__ __forceinline__ void store_operand8(uint16_t* p0, uint16_t* p1, long long off, const float* y, int fmt, bool is_weight) {
  if (fmt == kFmtF16F8) {
    uint16_t a[8];
    uint8_t b[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) split_f16f8(y[i], is_weight, a[i], b[i], c[i]);
    uint4 pa;
    pa.x = a[0] | (uint32_t(a[1]) << 16); pa.y = a[2] | (uint32_t(a[3]) << 16);
    pa.z = a[4] | (uint32_t(a[5]) << 16); pa.w = a[6] | (uint32_t(a[7]) << 16);
    *reinterpret_cast<uint4*>(p0 + off) = pa;
    uint8_t* q0 = reinterpret_cast<uint8_t*>(p1) + f8_offset(off);
    *reinterpret_cast<uint2*>(q0) = make_uint2(pack4_u8(b), pack4_u8(b + 4));
    *reinterpret_cast<uint2*>(q0 + 64) = make_uint2(pack4_u8(c), pack4_u8(c + 4));
  } else {
    store_operand4(p0, p1, off, y, fmt, is_weight);
    store_operand4(p0, p1, off + 4, y + 4, fmt, is_weight);
  }
}

// one element (attention output, ragged tails)
__device__ __forceinline__ void store_operand1(uint16_t* p0, uint16_t* p1, long long off, float y, int fmt) {
  if (fmt == kFmtF16F8) {
    uint16_t a;
    uint8_t b, c;
    split_f16f8(y, false, a, b, c);
    p0[off] = a;
    uint8_t* q0 = reinterpret_cast<uint8_t*>(p1) + f8_offset(off);
    q0[0] = b;
    q0[64] = c;
  } else if (fmt == kFmtBf16) {
    const __nv_bfloat16 h = __float2bfloat16_rn(y);
    p0[off] = __bfloat16_as_ushort(h);
    if (p1) p1[off] = __bfloat16_as_ushort(__float2bfloat16_rn(y - __bfloat162float(h)));
  } else {
    const __half h = __float2half_rn(y);
    p0[off] = __half_as_ushort(h);
    if (p1) p1[off] = __half_as_ushort(__float2half_rn(y - __half2float(h)));
  }
}

__device__ __forceinline__ void split16(float x, int fmt, uint16_t& p0, uint16_t& p1) {
  if (fmt == kFmtBf16) {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
    p0 = __bfloat16_as_ushort(h);
    p1 = __bfloat16_as_ushort(l);
  } else {
    const __half h = __float2half_rn(x);
    const __half l = __float2half_rn(x - __half2float(h));
    p0 = __half_as_ushort(h);
    p1 = __half_as_ushort(l);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// One thread owns output row `row`, columns [col0, col0 + 32) (col0 a multiple of 32; leading dimensions multiples of 8,
// so every vector access below is aligned).  The full-width path keeps the 32 values in registers: every loop is
// compile-time unrolled and the activation is a template parameter, so nothing is indexed dynamically.
// ---------------------------------------------------------------------------------------------------------------
template <int ACT>
__device__ __forceinline__ void epilogue_full32(const EpilogueParams& ep, int row, int col0, float (&v)[32]) {
  if (ep.bias) {
    const float4* b4 = reinterpret_cast<const float4*>(ep.bias + col0);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 b = __ldg(b4 + q);
      v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
    }
  }
  if (ACT == kActGeluTanh) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_tanh_f(v[j]);
  } else if (ACT == kActGeluErf) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf_f(v[j]);
  }
  if (ep.residual) {
    const float4* r4 = reinterpret_cast<const float4*>(ep.residual + static_cast<long long>(row) * ep.ld_res + col0);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 t = __ldg(r4 + q);
      v[4 * q] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
    }
  }
  if (ep.col_scale) {
    const float4* s4 = reinterpret_cast<const float4*>(ep.col_scale + col0);
    const float4* t4 = reinterpret_cast<const float4*>(ep.col_shift + col0);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 w = __ldg(s4 + q), b = __ldg(t4 + q);
      // Rescaler: w * y + b with two roundings, as torch evaluates it (modeling_hypernet.py:18-19)
      v[4 * q] = __fadd_rn(__fmul_rn(w.x, v[4 * q]), b.x);
      v[4 * q + 1] = __fadd_rn(__fmul_rn(w.y, v[4 * q + 1]), b.y);
      v[4 * q + 2] = __fadd_rn(__fmul_rn(w.z, v[4 * q + 2]), b.z);
      v[4 * q + 3] = __fadd_rn(__fmul_rn(w.w, v[4 * q + 3]), b.w);
    }
  }
  if (ep.out_f32) {
    float4* o = reinterpret_cast<float4*>(ep.out_f32 + static_cast<long long>(row) * ep.ld_out + col0);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 y = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      if (ep.stream_f32) __stcs(o + q, y); else o[q] = y;
    }
  }
  if (ep.out_p0) {
    const long long off = static_cast<long long>(row) * ep.ld_split + col0;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      const float y[8] = {v[j], v[j + 1], v[j + 2], v[j + 3], v[j + 4], v[j + 5], v[j + 6], v[j + 7]};
      store_operand8(ep.out_p0, ep.out_p1, off + j, y, ep.split_fmt, false);
    }
  }
}

// ragged tail (N not a multiple of 32): element-wise, rarely taken
__device__ __noinline__ void epilogue_ragged(const EpilogueParams& ep, int row, int col0, int ncols, const float* v) {
  for (int j = 0; j < ncols; ++j) {
    float y = v[j];
    if (ep.bias) y += __ldg(ep.bias + col0 + j);
    if (ep.act == kActGeluTanh) y = gelu_tanh_f(y);
    else if (ep.act == kActGeluErf) y = gelu_erf_f(y);
    if (ep.residual) y += __ldg(ep.residual + static_cast<long long>(row) * ep.ld_res + col0 + j);
    if (ep.col_scale) y = __fadd_rn(__fmul_rn(__ldg(ep.col_scale + col0 + j), y), __ldg(ep.col_shift + col0 + j));
    if (ep.out_f32) ep.out_f32[static_cast<long long>(row) * ep.ld_out + col0 + j] = y;
    if (ep.out_p0) store_operand1(ep.out_p0, ep.out_p1, static_cast<long long>(row) * ep.ld_split + col0 + j, y, ep.split_fmt);
  }
}

__device__ __forceinline__ void epilogue_store32(const EpilogueParams& ep, int row, int col0, int ncols, float (&v)[32]) {
  if (ncols == 32) {
    if (ep.act == kActGeluTanh) epilogue_full32<kActGeluTanh>(ep, row, col0, v);
    else if (ep.act == kActGeluErf) epilogue_full32<kActGeluErf>(ep, row, col0, v);
    else epilogue_full32<kActNone>(ep, row, col0, v);
  } else {
    float t[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) t[j] = v[j];
    epilogue_ragged(ep, row, col0, ncols, t);
  }
}

}  // namespace zett

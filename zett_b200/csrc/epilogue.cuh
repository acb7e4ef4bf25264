// Fused GEMM epilogue shared by the tcgen05 kernel and the SIMT debug kernel:
//     y = act(acc + bias[n]);  y += residual[m, n];  y = col_scale[n] * y + col_shift[n]
// then store fp32 and/or the 16-bit split planes (plane 0 = round(y), plane 1 = round(y - plane0)) that the
// next GEMM consumes as its A operand.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>

namespace zett {

enum Act : int { kActNone = 0, kActGeluTanh = 1, kActGeluErf = 2 };
enum SplitFmt : int { kFmtBf16 = 0, kFmtFp16 = 1 };

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

// One thread owns output row `row`, columns [col0, col0 + ncols), ncols <= 32.  col0 is a multiple of 32 and the
// leading dimensions are multiples of 8, so 16-byte vector stores are aligned whenever a full group is in range.
__device__ __forceinline__ void epilogue_store32(const EpilogueParams& ep, int row, int col0, int ncols, float* v) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j < ncols) {
      float y = v[j];
      if (ep.bias) y += __ldg(ep.bias + col0 + j);
      if (ep.act == kActGeluTanh) y = gelu_tanh_f(y);
      else if (ep.act == kActGeluErf) y = gelu_erf_f(y);
      v[j] = y;
    }
  }
  if (ep.residual) {
    const float* r = ep.residual + static_cast<long long>(row) * ep.ld_res + col0;
    if (ncols == 32) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(r + j));
        v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
      }
    } else {
      for (int j = 0; j < ncols; ++j) v[j] += __ldg(r + j);
    }
  }
  if (ep.col_scale) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < ncols) v[j] = __ldg(ep.col_scale + col0 + j) * v[j] + __ldg(ep.col_shift + col0 + j);
  }
  if (ep.out_f32) {
    float* o = ep.out_f32 + static_cast<long long>(row) * ep.ld_out + col0;
    if (ncols == 32) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
      for (int j = 0; j < ncols; ++j) o[j] = v[j];
    }
  }
  if (ep.out_p0) {
    uint16_t* o0 = ep.out_p0 + static_cast<long long>(row) * ep.ld_split + col0;
    uint16_t* o1 = ep.out_p1 ? ep.out_p1 + static_cast<long long>(row) * ep.ld_split + col0 : nullptr;
    if (ncols == 32) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint16_t a[8], b[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) split16(v[j + t], ep.split_fmt, a[t], b[t]);
        uint4 pa, pb;
        pa.x = a[0] | (uint32_t(a[1]) << 16); pa.y = a[2] | (uint32_t(a[3]) << 16);
        pa.z = a[4] | (uint32_t(a[5]) << 16); pa.w = a[6] | (uint32_t(a[7]) << 16);
        pb.x = b[0] | (uint32_t(b[1]) << 16); pb.y = b[2] | (uint32_t(b[3]) << 16);
        pb.z = b[4] | (uint32_t(b[5]) << 16); pb.w = b[6] | (uint32_t(b[7]) << 16);
        *reinterpret_cast<uint4*>(o0 + j) = pa;
        if (o1) *reinterpret_cast<uint4*>(o1 + j) = pb;
      }
    } else {
      for (int j = 0; j < ncols; ++j) {
        uint16_t a, b;
        split16(v[j], ep.split_fmt, a, b);
        o0[j] = a;
        if (o1) o1[j] = b;
      }
    }
  }
}

}  // namespace zett

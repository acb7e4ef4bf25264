// Fused GEMM epilogue shared by the tcgen05 kernel and the SIMT checker kernel:
//     y = act(acc * w_scale[n] + bias[n]);  y += residual[m, n];  y = col_scale[n] * y + col_shift[n]
// then store fp32 and/or the operand lines (operand.cuh) the next GEMM consumes as its A operand.
// The unit of work is a QUAD: four consecutive columns of one row (16 bytes of fp32), which is what one lane holds after
// the tcgen05 kernel has transposed an accumulator chunk through shared memory (gemm_tcgen05.cuh) -- eight lanes then
// cover 128 contiguous bytes of an output row, so every global access of the epilogue is made of whole lines.
#pragma once
#include "operand.cuh"

namespace zett {

enum Act : int { kActNone = 0, kActGeluTanh = 1, kActGeluErf = 2 };

struct EpilogueParams {
  const float* bias;        // [n] or nullptr
  const float* w_scale;     // [n] or nullptr: inverse of the power-of-two row scale the weights were stored with
  int act;
  const float* residual;    // fp32 [m, ld_res] or nullptr (added after the activation)
  long long ld_res;
  const float* col_scale;   // [n] or nullptr: Rescaler  y = w * y + b  (hf_hypernet/modeling_hypernet.py:9-19)
  const float* col_shift;
  float* out_f32;           // nullable, [m, ld_out]
  long long ld_out;
  OperandOut out_op;        // base nullable: operand lines of the next GEMM, rows of out_op.ld_bytes
  int stream_f32;           // 1: fp32 output stored with the streaming (evict-first) hint
};

// F.gelu(x, approximate="tanh")  (ProjectorBlock, hf_hypernet/modeling_hypernet.py:36-39):
//     0.5 x (1 + tanh(u)) = x / (1 + exp(-2 u)),  u = sqrt(2/pi) (x + 0.044715 x^3)
// one MUFU.EX2 and one MUFU.RCP instead of tanhf's branchy ~25 instructions; relative error ~2e-7.  (The epilogue warps
// share the SM's four issue ports with nothing else, and at K = 768 a tile's main loop is only ~12 000 cycles.)
// (rcp.approx / ex2.approx: one MUFU each, ~1 ulp; __frcp_rn is the IEEE-rounded sequence and cost more than tanhf did)
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float gelu_tanh_f(float x) {
  const float c2 = -2.0f * 0.7978845608028654f;  // -2 sqrt(2/pi)
  const float u2 = c2 * fmaf(0.044715f * x * x, x, x);
  return x * rcp_approx(1.0f + __expf(u2));
}
// exact GELU, hidden_act="gelu" of the RoBERTa encoder: 0.5 x (1 + erf(x / sqrt 2)), erf by Abramowitz & Stegun 7.1.26
// (|error| <= 1.5e-7 ABSOLUTE, which is what matters next to the 1; branch-free: erff's two ranges diverge inside a warp)
__device__ __forceinline__ float gelu_erf_f(float x) {
  const float z = x * 0.7071067811865476f;
  const float az = fabsf(z);
  const float t = rcp_approx(fmaf(0.3275911f, az, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = p * t * __expf(-az * az);            // 1 - erf(|z|)
  const float one_plus_erf = z >= 0.f ? 2.0f - e : e;  // 1 + erf(z)
  return 0.5f * x * one_plus_erf;
}

// per-column constants of a quad (loaded once per lane and accumulator chunk: the lane's columns do not change with the row)
struct QuadConsts {
  float4 ws, b, cs, ct;
};

__device__ __forceinline__ QuadConsts load_quad_consts(const EpilogueParams& ep, int col) {
  QuadConsts q;
  q.ws = ep.w_scale ? __ldg(reinterpret_cast<const float4*>(ep.w_scale + col)) : make_float4(1.f, 1.f, 1.f, 1.f);
  q.b = ep.bias ? __ldg(reinterpret_cast<const float4*>(ep.bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
  if (ep.col_scale) {
    q.cs = __ldg(reinterpret_cast<const float4*>(ep.col_scale + col));
    q.ct = __ldg(reinterpret_cast<const float4*>(ep.col_shift + col));
  } else {
    q.cs = make_float4(1.f, 1.f, 1.f, 1.f);
    q.ct = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  return q;
}

template <int ACT>
__device__ __forceinline__ void epilogue_quad(const EpilogueParams& ep, long long row, int col, float4 acc, const QuadConsts& q,
                                              uint32_t& bad) {
  float v[4] = {acc.x, acc.y, acc.z, acc.w};
  if (ep.w_scale) {
    v[0] = fmaf(v[0], q.ws.x, q.b.x); v[1] = fmaf(v[1], q.ws.y, q.b.y);
    v[2] = fmaf(v[2], q.ws.z, q.b.z); v[3] = fmaf(v[3], q.ws.w, q.b.w);
  } else if (ep.bias) {
    v[0] += q.b.x; v[1] += q.b.y; v[2] += q.b.z; v[3] += q.b.w;
  }
  if (ACT == kActGeluTanh) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = gelu_tanh_f(v[j]);
  } else if (ACT == kActGeluErf) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = gelu_erf_f(v[j]);
  }
  if (ep.residual) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(ep.residual + row * ep.ld_res + col));
    v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
  }
  if (ep.col_scale) {
    // Rescaler: w * y + b with two roundings, as torch evaluates it (modeling_hypernet.py:18-19)
    v[0] = __fadd_rn(__fmul_rn(q.cs.x, v[0]), q.ct.x); v[1] = __fadd_rn(__fmul_rn(q.cs.y, v[1]), q.ct.y);
    v[2] = __fadd_rn(__fmul_rn(q.cs.z, v[2]), q.ct.z); v[3] = __fadd_rn(__fmul_rn(q.cs.w, v[3]), q.ct.w);
  }
  if (ep.out_f32) {
    float4* o = reinterpret_cast<float4*>(ep.out_f32 + row * ep.ld_out + col);
    const float4 y = make_float4(v[0], v[1], v[2], v[3]);
    if (ep.stream_f32) __stcs(o, y); else *o = y;
  }
  if (ep.out_op.base) store_operand4(ep.out_op, row, col, v, false, bad);
}

__device__ __forceinline__ void epilogue_quad_dyn(const EpilogueParams& ep, long long row, int col, float4 acc, const QuadConsts& q,
                                                  uint32_t& bad) {
  if (ep.act == kActGeluTanh) epilogue_quad<kActGeluTanh>(ep, row, col, acc, q, bad);
  else if (ep.act == kActGeluErf) epilogue_quad<kActGeluErf>(ep, row, col, acc, q, bad);
  else epilogue_quad<kActNone>(ep, row, col, acc, q, bad);
}

// row-per-thread form (SIMT checker): `v` holds columns [col0, col0 + 32) of `row`; n is a multiple of 4
__device__ __forceinline__ void epilogue_row32(const EpilogueParams& ep, long long row, int col0, int n, const float* v, uint32_t& bad) {
  for (int j = 0; j < 32 && col0 + j < n; j += 4) {
    const QuadConsts q = load_quad_consts(ep, col0 + j);
    epilogue_quad_dyn(ep, row, col0 + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]), q, bad);
  }
}

}  // namespace zett

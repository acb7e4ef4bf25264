// Stand-alone probe (built and run on the GPU box by scripts/gpu_r2l.sh; not part of the library):
//   1. which die every SM sits on -- latency of an atomic (executed at the line's home L2 slice) from every SM to a set
//      of lines: for one line the SMs fall into a near and a far group, and the partition is the same for every line
//      up to which side is "near";
//   2. whether a buffer that SMs of BOTH dies read is fetched from DRAM once or once per die -- the same 32 MB read by
//      every SM in lockstep, then by the SMs of one die only (dram__bytes_read.sum per launch under ncu).
// nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o /tmp/l2_die_probe scripts/l2_die_probe.cu
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kLines = 48;
constexpr int kReps = 24;
constexpr long long kLineStride = 1 << 20;  // bytes between probed lines

__device__ __forceinline__ unsigned smid() {
  unsigned r;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
  return r;
}

// one CTA per SM, thread 0 only: min-of-reps latency of atomicAdd(line, 0) for every probed line
__global__ void latency_kernel(unsigned* buf, int* out_lat, int* out_smid, volatile int* turn) {
  if (threadIdx.x != 0) return;
  const unsigned sm = smid();
  out_smid[blockIdx.x] = static_cast<int>(sm);
  while (*turn != static_cast<int>(blockIdx.x)) __nanosleep(200);  // one SM at a time: no queueing at the line's slice
  for (int l = 0; l < kLines; ++l) {
    unsigned* p = buf + l * (kLineStride / 4);
    int best = 1 << 30;
    for (int r = 0; r < kReps; ++r) {
      const long long t0 = clock64();
      const unsigned v = atomicAdd(p, 0u);
      // make the timestamp depend on the result
      long long t1;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %1, 0xFFFFFFFF;\n\t@p trap;\n\tmov.u64 %0, %%clock64;\n\t}" : "=l"(t1) : "r"(v) : "memory");
      best = min(best, static_cast<int>(t1 - t0));
    }
    out_lat[blockIdx.x * kLines + l] = best;
  }
  __threadfence();
  *turn = static_cast<int>(blockIdx.x) + 1;
}

// every selected CTA reads the whole buffer (16-byte loads that bypass L1); which = -1: all SMs, 0 / 1: that die only
__global__ void shared_read_kernel(const uint4* buf, long long n16, const int* die_of_sm, int which, unsigned* sink) {
  const int die = die_of_sm[smid()];
  if (which >= 0 && die != which) return;
  unsigned acc = 0;
  for (long long i = threadIdx.x; i < n16; i += blockDim.x) {
    const uint4 v = __ldcg(buf + i);
    acc ^= v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x12345678u) *sink = acc;
}

int main() {
  int dev = 0, n_sm = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  unsigned* lines;
  CK(cudaMalloc(&lines, kLines * kLineStride));
  CK(cudaMemset(lines, 0, kLines * kLineStride));
  int *d_lat, *d_smid;
  CK(cudaMalloc(&d_lat, n_sm * kLines * sizeof(int)));
  CK(cudaMalloc(&d_smid, n_sm * sizeof(int)));
  CK(cudaMemset(d_smid, 0xFF, n_sm * sizeof(int)));
  // 1 CTA per SM: a block that needs most of the SM's shared memory cannot share it
  CK(cudaFuncSetAttribute(latency_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  int* d_turn;
  CK(cudaMalloc(&d_turn, sizeof(int)));
  for (int it = 0; it < 2; ++it) {
    CK(cudaMemset(d_turn, 0, sizeof(int)));
    latency_kernel<<<n_sm, 32, 200 * 1024>>>(lines, d_lat, d_smid, d_turn);
    CK(cudaDeviceSynchronize());
  }
  CK(cudaDeviceSynchronize());
  std::vector<int> lat(n_sm * kLines), sm_of_block(n_sm);
  CK(cudaMemcpy(lat.data(), d_lat, lat.size() * sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(sm_of_block.data(), d_smid, n_sm * sizeof(int), cudaMemcpyDeviceToHost));
  // per line: threshold halfway between the 10th and 90th percentile; reference partition = line 0's; other lines vote
  std::vector<int> votes(n_sm, 0);
  std::vector<int> ref(n_sm, 0);
  int used_lines = 0;
  for (int l = 0; l < kLines; ++l) {
    std::vector<int> v(n_sm);
    for (int b = 0; b < n_sm; ++b) v[b] = lat[b * kLines + l];
    std::vector<int> s = v;
    std::sort(s.begin(), s.end());
    const int lo = s[n_sm / 10], hi = s[n_sm - 1 - n_sm / 10];
    if (l < 4) printf("line %d: latency p10 %d  p50 %d  p90 %d  (min %d max %d)\n", l, lo, s[n_sm / 2], hi, s[0], s[n_sm - 1]);
    if (hi - lo < 12) continue;  // not bimodal
    const int thr = (lo + hi) / 2;
    std::vector<int> side(n_sm);
    for (int b = 0; b < n_sm; ++b) side[b] = v[b] > thr ? 1 : 0;
    if (used_lines == 0) ref = side;
    int agree = 0;
    for (int b = 0; b < n_sm; ++b) agree += side[b] == ref[b];
    const bool flip = agree < n_sm / 2;
    for (int b = 0; b < n_sm; ++b) votes[b] += (side[b] ^ (flip ? 1 : 0)) ? 1 : -1;
    ++used_lines;
  }
  std::vector<int> die_of_sm(256, 0);
  int n1 = 0, weak = 0;
  for (int b = 0; b < n_sm; ++b) {
    const int d = votes[b] > 0 ? 1 : 0;
    if (sm_of_block[b] >= 0) die_of_sm[sm_of_block[b]] = d;
    n1 += d;
    if (abs(votes[b]) < used_lines / 2) ++weak;
  }
  printf("lines used %d of %d; SMs on die 1: %d of %d; SMs with a weak vote: %d\n", used_lines, kLines, n1, n_sm, weak);
  printf("die of smid 0..%d: ", n_sm - 1);
  for (int s = 0; s < n_sm; ++s) printf("%d", die_of_sm[s]);
  printf("\n");
  // do CTA pairs (smid 2i, 2i+1) share a die?
  int split_pairs = 0;
  for (int s = 0; s + 1 < n_sm; s += 2) split_pairs += die_of_sm[s] != die_of_sm[s + 1];
  printf("TPCs (smid 2i, 2i+1) split over dies: %d\n", split_pairs);

  int* d_die;
  CK(cudaMalloc(&d_die, 256 * sizeof(int)));
  CK(cudaMemcpy(d_die, die_of_sm.data(), 256 * sizeof(int), cudaMemcpyHostToDevice));
  const long long bytes = 32ll << 20;
  uint4* big;
  unsigned* sink;
  CK(cudaMalloc(&big, bytes));
  CK(cudaMalloc(&sink, 4));
  // something larger than L2 to flush it between the launches
  char* flush;
  const long long flush_bytes = 512ll << 20;
  CK(cudaMalloc(&flush, flush_bytes));
  CK(cudaFuncSetAttribute(shared_read_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int which_list[4] = {-1, 0, 1, -1};
  for (int i = 0; i < 4; ++i) {
    CK(cudaMemset(big, 1, bytes));
    CK(cudaMemset(flush, i, flush_bytes));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    shared_read_kernel<<<n_sm, 512, 200 * 1024>>>(big, bytes / 16, d_die, which_list[i], sink);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("shared read of %lld MB by %s: %.3f ms\n", bytes >> 20, which_list[i] < 0 ? "all SMs" : (which_list[i] == 0 ? "die 0 only" : "die 1 only"), ms);
  }
  return 0;
}

// Microbenchmark: issue rate of VABSDIFF.U32 (|a-b|+c, one instruction) vs the FADD pair of the weighted
// kernel, alone and mixed, on one B200.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 vabsdiff.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>  // 0: FADD pair, 1: VABSDIFF, 2: mixed 2 VABSDIFF : 1 FADD pair (x = 2/3), 3: mixed 1:1
__global__ void k(const float* in, float* out, int iters) {
  float fa[24]; unsigned ua[24];
  float fb = in[threadIdx.x]; unsigned ub = __float_as_uint(in[threadIdx.x + 32]);
#pragma unroll
  for (int j = 0; j < 24; ++j) { fa[j] = in[j]; ua[j] = j; }
  float facc[24]; unsigned uacc[24];
#pragma unroll
  for (int j = 0; j < 24; ++j) { facc[j] = 0.f; uacc[j] = 0u; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 24; ++j) {
      const bool usad = MODE == 1 || (MODE == 2 && (j % 3) != 0) || (MODE == 3 && (j & 1));
      if (usad) uacc[j] = __usad(ua[j], ub, uacc[j]);
      else facc[j] += fabsf(fa[j] - fb);
    }
    fb += 1.0f; ub += 3u;
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 24; ++j) s += facc[j] + __uint_as_float(uacc[j]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, float* in, float* out) {
  const int iters = 20000, blocks = 148 * 4, threads = 256;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(in, out, 100); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(in, out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double pn = double(blocks) * threads * iters * 24;  // pair-node updates
  printf("%-28s %.3f ms  %.1f pair-node updates / clk / SM (at 1.965 GHz)\n", name, ms, pn / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
  float *in, *out; cudaMalloc(&in, 4096); cudaMalloc(&out, 148 * 4 * 256 * 4); cudaMemset(in, 0, 4096);
  run<0>("FADD + FADD|.| (now)", in, out);
  run<1>("VABSDIFF.U32", in, out);
  run<2>("mixed 2 VABSDIFF : 1 FADD pair", in, out);
  run<3>("mixed 1 : 1", in, out);
  return 0;
}

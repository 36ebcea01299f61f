// Micro-benchmark (diagnostics): latency of the warp primitives the wedge kernel leans on, one warp alone on an SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_warp tools/ubench_warp.cu && /tmp/ubench_warp
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int distinct, long long* out, unsigned* sink) {
  const int lane = threadIdx.x & 31;
  unsigned v = (unsigned)(lane % distinct) * 7u + 3u;
  unsigned acc = 0;
  const int N = 2000;
  long long t0 = clock64();
  for (int i = 0; i < N; i++) { unsigned m = __match_any_sync(0xffffffffu, v); v += (m & 1u); acc += m; }
  long long t1 = clock64();
  for (int i = 0; i < N; i++) { unsigned m = __ballot_sync(0xffffffffu, v & 1u); v += (m & 2u); acc += m; }
  long long t2 = clock64();
  for (int i = 0; i < N; i++) { unsigned m = __shfl_sync(0xffffffffu, v, (lane + 1) & 31); v += (m & 1u); acc += m; }
  long long t3 = clock64();
  for (int i = 0; i < N; i++) { unsigned m = __reduce_max_sync(0xffffffffu, v); v += (m & 1u); acc += m; }
  long long t4 = clock64();
  unsigned long long vv = v;
  for (int i = 0; i < N; i++) { unsigned m = __match_any_sync(0xffffffffu, vv); vv += (m & 1u); acc += m; }
  long long t5 = clock64();
  for (int i = 0; i < N; i++) { v = v * 3u + 1u; }
  long long t6 = clock64();
  if (lane == 0) { out[0] = (t1 - t0) / N; out[1] = (t2 - t1) / N; out[2] = (t3 - t2) / N; out[3] = (t4 - t3) / N; out[4] = (t5 - t4) / N; out[5] = (t6 - t5) / N; }
  sink[threadIdx.x] = acc + v + (unsigned)vv;
}
int main() {
  long long* out; unsigned* sink;
  cudaMallocManaged(&out, 64); cudaMalloc(&sink, 4096);
  for (int d : {1, 2, 4, 8, 16, 26, 32}) {
    k<<<1, 32>>>(d, out, sink);
    cudaDeviceSynchronize();
    printf("distinct %2d: match.any b32 %lld  ballot %lld  shfl %lld  redux.max %lld  match.any b64 %lld  imad %lld cycles (dependent chain)\n", d, out[0], out[1], out[2], out[3], out[4], out[5]);
  }
  return 0;
}

// handshake_probe.cu — cost of passing a token between two warps of a CTA (diagnostic, not product code):
// named barriers (bar.arrive / bar.sync) versus a polled shared-memory word.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256, 1) bar_pingpong(int rounds, long long* out) {
  const int warp = threadIdx.x >> 5, pair = warp & 3;
  const bool leader = warp < 4;
  const int P = 1 + 2 * pair, Q = 2 + 2 * pair;
  if (!leader) asm volatile("bar.arrive %0, 64;" ::"r"(Q) : "memory");
  const long long t0 = clock64();
  for (int i = 0; i < rounds; ++i) {
    if (leader) {
      asm volatile("bar.sync %0, 64;" ::"r"(Q) : "memory");
      asm volatile("bar.arrive %0, 64;" ::"r"(P) : "memory");
    } else {
      asm volatile("bar.sync %0, 64;" ::"r"(P) : "memory");
      asm volatile("bar.arrive %0, 64;" ::"r"(Q) : "memory");
    }
  }
  if (leader) asm volatile("bar.sync %0, 64;" ::"r"(Q) : "memory");
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (t1 - t0);
}
__global__ void __launch_bounds__(256, 1) flag_pingpong(int rounds, long long* out) {
  __shared__ volatile int turn[4];
  const int warp = threadIdx.x >> 5, pair = warp & 3, lane = threadIdx.x & 31;
  const bool leader = warp < 4;
  if (threadIdx.x < 4) turn[threadIdx.x] = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < rounds; ++i) {
    const int want = leader ? 2 * i : 2 * i + 1;       // leader runs on even values
    while (turn[pair] != want) { }
    __syncwarp();
    if (lane == 0) turn[pair] = want + 1;
    __syncwarp();
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = (t1 - t0);
}
int main() {
  long long* d; cudaMalloc(&d, 16); cudaMemset(d, 0, 16);
  const int rounds = 100000;
  bar_pingpong<<<148, 256>>>(rounds, d);
  flag_pingpong<<<148, 256>>>(rounds, d);
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("named barriers: %.1f cycles per round (two hand-overs)\n", (double)h[0] / rounds);
  printf("polled flag   : %.1f cycles per round (two hand-overs)\n", (double)h[1] / rounds);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}

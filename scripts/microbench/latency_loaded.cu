// Mailbox hand-off latency between two SMs with and without a background streaming kernel. Diagnostics only.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long ldr(const unsigned long long* p){unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];":"=l"(v):"l"(p):"memory"); return v;}
__device__ __forceinline__ void str(unsigned long long* p, unsigned long long v){asm volatile("st.relaxed.gpu.global.u64 [%0], %1;"::"l"(p),"l"(v):"memory");}
__device__ __forceinline__ unsigned long long gt(){unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;":"=l"(t)::"memory"); return t;}
// block 0 <-> block 1 ping-pong; other blocks stream `bytes` of memory with cp.async-like plain loads
__global__ void k(unsigned long long* box, long long* ns, int rounds, const float4* src, size_t n4, float* sink, int nlanes){
  if (blockIdx.x >= 2) {   // background streamers
    float acc = 0.f;
    for (int rep = 0; rep < 64; ++rep) {
      if (ldr(box + 64) != 0) break;    // stop flag
      for (size_t i = (size_t)(blockIdx.x - 2) * blockDim.x + threadIdx.x; i < n4; i += (size_t)(gridDim.x - 2) * blockDim.x) {
        float4 v = __ldcs(src + i); acc += v.x + v.y + v.z + v.w;
      }
    }
    if (acc == 1.2345f) sink[0] = acc;
    return;
  }
  if (threadIdx.x >= nlanes) return;
  int me = blockIdx.x; unsigned long long* mine = box + (me ? 1024 : 0) + threadIdx.x * 88;  // strided like the kernel's rows
  unsigned long long* other = box + (me ? 0 : 1024) + threadIdx.x * 88;
  unsigned long long t0 = gt();
  for (int r = 1; r <= rounds; r++) {
    if (me == 0) { str(mine, (unsigned long long)r); while (ldr(other) != (unsigned long long)r) {} }
    else         { while (ldr(other) != (unsigned long long)r) {} str(mine, (unsigned long long)r); }
    __syncwarp();
  }
  unsigned long long t1 = gt();
  if (me == 0 && threadIdx.x == 0) { ns[0] = (long long)(t1 - t0); str(box + 64, 1ull); }
}
int main(){
  unsigned long long* box; cudaMalloc(&box, 1<<20); long long* ns; cudaMalloc(&ns, 64); float* sink; cudaMalloc(&sink, 64);
  size_t bytes = (size_t)2 << 30; float4* src; cudaMalloc(&src, bytes); cudaMemset(src, 0, bytes);
  for (int nl : {1, 8, 32}) for (int bg : {0, 146}) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(box, 0, 1<<20);
      k<<<2 + bg, 512>>>(box, ns, 300, src, bytes / 16, sink, nl); cudaDeviceSynchronize();
    }
    long long h; cudaMemcpy(&h, ns, 8, cudaMemcpyDeviceToHost);
    printf("lanes %2d background CTAs %3d: round trip %.0f ns -> one-way hand-off %.0f ns  (%s)\n", nl, bg, h / 300.0, h / 600.0, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}

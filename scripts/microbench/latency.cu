// Dependent-chain latencies on B200 (cycles) + cross-SM mailbox round trip (ns).  Diagnostics only.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2f(float x){float y; asm volatile("ex2.approx.ftz.f32 %0, %1;":"=f"(y):"f"(x)); return y;}
__device__ __forceinline__ float lg2f(float x){float y; asm volatile("lg2.approx.ftz.f32 %0, %1;":"=f"(y):"f"(x)); return y;}
template<int OP> __global__ void chain(float* out, long long* cyc, float seed, int iters){
  float x = seed + threadIdx.x*1e-3f; float y = seed*0.5f;
  __shared__ float sm[64];
  sm[threadIdx.x & 63] = 0.f; __syncthreads();
  long long t0 = clock64();
  #pragma unroll 1
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int u=0;u<16;u++){
      if(OP==0) x = x + y;                                  // FADD
      if(OP==1) x = fmaxf(x, y) + 0.0f*x;                   // FMNMX (+fma to keep dep) -> measure pair
      if(OP==2) x = __shfl_sync(0xffffffffu, x, (u*7+3)&31); // SHFL.IDX
      if(OP==3) x = ex2f(x)*1e-3f;                           // MUFU.EX2 + FMUL
      if(OP==4) x = lg2f(fabsf(x)+2.0f);                     // MUFU.LG2 + FADD
      if(OP==5) { int idx = __float_as_int(x) & 63; x = sm[idx] ; } // LDS dependent
      if(OP==6) x = fmaf(x, 1.0001f, y);                     // FFMA
      if(OP==7) x = fmaxf(fmaxf(x, y), seed);                // 2x FMNMX or FMNMX3
    }
  }
  long long t1 = clock64();
  if(threadIdx.x==0) cyc[0] = (t1-t0);
  out[threadIdx.x] = x;
}
// ping-pong between two CTAs through global memory words {value, tag}
__device__ __forceinline__ unsigned long long ldr(const unsigned long long* p){unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];":"=l"(v):"l"(p):"memory"); return v;}
__device__ __forceinline__ void str(unsigned long long* p, unsigned long long v){asm volatile("st.relaxed.gpu.global.u64 [%0], %1;"::"l"(p),"l"(v):"memory");}
__global__ void pingpong(unsigned long long* box, long long* ns, int rounds){
  if(threadIdx.x!=0) return;
  unsigned long long t0; asm volatile("mov.u64 %0, %%globaltimer;":"=l"(t0));
  int me = blockIdx.x; // 0 or 1
  for(int r=1;r<=rounds;r++){
    if(me==0){ str(box+0, (unsigned long long)r); while(ldr(box+16)!=(unsigned long long)r){} }
    else     { while(ldr(box+0)!=(unsigned long long)r){} str(box+16,(unsigned long long)r); }
  }
  unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;":"=l"(t1));
  if(me==0) ns[0]=(long long)(t1-t0);
}
int main(){
  float* out; long long* cyc; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64);
  const char* names[]={"FADD","FMNMX+FFMA","SHFL.IDX","EX2+FMUL","LG2+FADD(abs)","LDS dep","FFMA","FMNMX x2"};
  int iters=2000;
  for(int op=0;op<8;op++){
    for(int rep=0;rep<2;rep++){
      switch(op){case 0:chain<0><<<1,32>>>(out,cyc,1.0f,iters);break;case 1:chain<1><<<1,32>>>(out,cyc,1.0f,iters);break;
        case 2:chain<2><<<1,32>>>(out,cyc,1.0f,iters);break;case 3:chain<3><<<1,32>>>(out,cyc,1.0f,iters);break;
        case 4:chain<4><<<1,32>>>(out,cyc,1.0f,iters);break;case 5:chain<5><<<1,32>>>(out,cyc,1.0f,iters);break;
        case 6:chain<6><<<1,32>>>(out,cyc,1.0f,iters);break;case 7:chain<7><<<1,32>>>(out,cyc,1.0f,iters);break;}
      cudaDeviceSynchronize();
    }
    long long h; cudaMemcpy(&h,cyc,8,cudaMemcpyDeviceToHost);
    printf("%-16s %.2f cycles per dependent op-group\n", names[op], (double)h/(iters*16.0));
  }
  unsigned long long* box; cudaMalloc(&box, 4096); cudaMemset(box,0,4096); long long* ns; cudaMalloc(&ns,64);
  for(int rep=0;rep<2;rep++){ cudaMemset(box,0,4096); pingpong<<<2,32>>>(box,ns,2000); cudaDeviceSynchronize(); }
  long long h; cudaMemcpy(&h,ns,8,cudaMemcpyDeviceToHost);
  printf("cross-SM mailbox round trip (st.relaxed.gpu -> ld.relaxed.gpu spin, both ways): %.1f ns  => one-way ~%.1f ns\n", (double)h/2000.0, (double)h/4000.0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0); printf("clockRate attr %d kHz\n", clk);
  return 0;
}

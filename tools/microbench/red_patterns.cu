// Dev microbenchmark: throughput of warp-wide RED.ADD.F64 (and gathers) as a function of how the 32 lane addresses
// spread over sectors/lines -- decides whether lane->entry mapping and locus numbering matter for the fused kernel.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <algorithm>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA %s @%d: %s\n",#x,__LINE__,cudaGetErrorString(e)); exit(1);} }while(0)
__device__ __forceinline__ uint32_t hash32(uint32_t x){ x^=x>>16; x*=0x7feb352dU; x^=x>>15; x*=0x846ca68bU; x^=x>>16; return x; }

// each warp, per iteration: base = random column (aligned to `align` doubles); lane address = base + lane*stride (mod K)
template<int MODE> // 0 = RED, 1 = gather (ldg)
__global__ void k_pat(double* acc, const double* __restrict__ tab, int K, int stride, int align, int iters, int rnd_lanes, double* sink){
  const int lane = threadIdx.x & 31;
  uint32_t w = (blockIdx.x*blockDim.x + threadIdx.x) >> 5;
  double s = 0;
  for(int it=0; it<iters; ++it){
    uint32_t h = hash32(w*7919u + it*104729u + 17u);
    int base = (int)(h % (uint32_t)(K - 32*stride - align)); base -= base % align;
    int c = rnd_lanes ? (int)(hash32(h + lane*2654435761u) % (uint32_t)K) : base + lane*stride;
    if (MODE==1) s += __ldg(tab + c);
    if (MODE==0) atomicAdd(acc + c, 1.0);
  }
  if (s == -1.0) sink[0] = s;
}
template<class F> float timeit(F f){ cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); f(); CK(cudaDeviceSynchronize()); float best=1e30f;
  for(int r=0;r<3;r++){ cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms,a,b); best=std::min(best,ms);} return best; }
int main(){
  const int K=30000, iters=2000; int nsm; cudaDeviceGetAttribute(&nsm,cudaDevAttrMultiProcessorCount,0);
  double *acc,*tab,*sink; CK(cudaMalloc(&acc,K*8*2)); CK(cudaMalloc(&tab,K*8)); CK(cudaMalloc(&sink,64)); CK(cudaMemset(acc,0,K*8*2)); CK(cudaMemset(tab,0,K*8));
  const int blocks=nsm*4, threads=512; double ops=(double)blocks*threads*iters;
  struct P{const char* name; int stride, align, rnd;} pats[]={{"random 32 addresses",1,1,1},{"contiguous 32 doubles, 128B-aligned",1,16,0},{"contiguous 32 doubles, unaligned",1,1,0},
     {"stride 2 (16 sectors)",2,1,0},{"stride 4 (1 per sector, 8 lines)",4,1,0},{"stride 16 (1 per line)",16,1,0}};
  for(auto&p:pats){
    float r=timeit([&]{k_pat<0><<<blocks,threads>>>(acc,tab,K,p.stride,p.align,iters,p.rnd,sink);});
    float g=timeit([&]{k_pat<1><<<blocks,threads>>>(acc,tab,K,p.stride,p.align,iters,p.rnd,sink);});
    printf("%-40s RED %8.3f ms %7.1f Gop/s | LDG %8.3f ms %7.1f Gop/s\n",p.name,r,ops/r/1e6,g,ops/g/1e6);
  }
  return 0;
}

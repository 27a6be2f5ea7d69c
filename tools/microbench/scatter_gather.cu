// Design microbenchmark (dev tooling, not product): measures on a B200 the building blocks the fused
// E+M kernel is made of -- 12 B/nnz streaming, K-vector gathers from smem vs L1/L2, and fp64 scatter-adds
// into smem (CAS loop) vs global (REDG.F64) -- plus three row-processing schemes end to end.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o scatter_gather scatter_gather.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA %s @%d: %s\n",#x,__LINE__,cudaGetErrorString(e)); exit(1);} }while(0)

__device__ __forceinline__ uint32_t hash32(uint32_t x){ x^=x>>16; x*=0x7feb352dU; x^=x>>15; x*=0x846ca68bU; x^=x>>16; return x; }

__global__ void k_gen(double* q, int* col, long long nnz, int K, int skew){
  long long i = blockIdx.x*(long long)blockDim.x + threadIdx.x;
  for(; i<nnz; i+=(long long)gridDim.x*blockDim.x){
    uint32_t h = hash32((uint32_t)i*2654435761U + 12345U);
    double u = (h>>8) * (1.0/16777216.0);
    int c;
    if(skew) { double v=u*u*u; c = (int)(v*K); } else c = (int)(u*K);
    if(c>=K) c=K-1;
    col[i]=c;
    q[i] = 1.0 + (hash32(h)>>9)*(1.0/8388608.0);
  }
}

// A: flat stream only
__global__ void __launch_bounds__(512) k_stream(const double2* __restrict__ q2, const int4* __restrict__ c4, long long n4, double* sink){
  double acc=0; long long i = blockIdx.x*(long long)blockDim.x + threadIdx.x;
  for(; i<n4; i+=(long long)gridDim.x*blockDim.x){
    double2 a=q2[2*i], b=q2[2*i+1]; int4 c=c4[i];
    acc += a.x+a.y+b.x+b.y + (double)(c.x^c.y^c.z^c.w);
  }
  if(acc==-1.0) sink[0]=acc;
}
// B/C: stream + gather (smem table of S cols; rest from global)
template<bool SMEM>
__global__ void __launch_bounds__(1024) k_gather(const double2* __restrict__ q2, const int4* __restrict__ c4, long long n4,
                         const double* __restrict__ tab, int K, int S, double* sink){
  extern __shared__ double s_tab[];
  if(SMEM){ for(int i=threadIdx.x;i<S;i+=blockDim.x) s_tab[i]=tab[i]; __syncthreads(); }
  double acc=0; long long i = blockIdx.x*(long long)blockDim.x + threadIdx.x;
  for(; i<n4; i+=(long long)gridDim.x*blockDim.x){
    double2 a=q2[2*i], b=q2[2*i+1]; int4 c=c4[i];
    double t0,t1,t2,t3;
    if(SMEM){ t0 = c.x<S? s_tab[c.x]:__ldg(tab+c.x); t1 = c.y<S? s_tab[c.y]:__ldg(tab+c.y); t2 = c.z<S? s_tab[c.z]:__ldg(tab+c.z); t3 = c.w<S? s_tab[c.w]:__ldg(tab+c.w); }
    else { t0=__ldg(tab+c.x); t1=__ldg(tab+c.y); t2=__ldg(tab+c.z); t3=__ldg(tab+c.w); }
    acc += a.x*t0 + a.y*t1 + b.x*t2 + b.y*t3;
  }
  if(acc==-1.0) sink[0]=acc;
}
// D: stream + REDG scatter with R replicas
__global__ void __launch_bounds__(512) k_red_global(const double2* __restrict__ q2, const int4* __restrict__ c4, long long n4,
                         double* acc, int K, int R){
  double* my = acc + (size_t)(blockIdx.x % R)*K;
  long long i = blockIdx.x*(long long)blockDim.x + threadIdx.x;
  for(; i<n4; i+=(long long)gridDim.x*blockDim.x){
    double2 a=q2[2*i], b=q2[2*i+1]; int4 c=c4[i];
    atomicAdd(my+c.x, a.x); atomicAdd(my+c.y, a.y); atomicAdd(my+c.z, b.x); atomicAdd(my+c.w, b.y);
  }
}
// E: stream + smem CAS scatter (S cols in smem, rest REDG)
__global__ void __launch_bounds__(1024) k_red_smem(const double2* __restrict__ q2, const int4* __restrict__ c4, long long n4,
                         double* acc, int K, int S){
  extern __shared__ double s_acc[];
  for(int i=threadIdx.x;i<S;i+=blockDim.x) s_acc[i]=0; __syncthreads();
  long long i = blockIdx.x*(long long)blockDim.x + threadIdx.x;
  for(; i<n4; i+=(long long)gridDim.x*blockDim.x){
    double2 a=q2[2*i], b=q2[2*i+1]; int4 c=c4[i];
    if(c.x<S) atomicAdd(s_acc+c.x, a.x); else atomicAdd(acc+c.x, a.x);
    if(c.y<S) atomicAdd(s_acc+c.y, a.y); else atomicAdd(acc+c.y, a.y);
    if(c.z<S) atomicAdd(s_acc+c.z, b.x); else atomicAdd(acc+c.z, b.x);
    if(c.w<S) atomicAdd(s_acc+c.w, b.y); else atomicAdd(acc+c.w, b.y);
  }
  __syncthreads();
  for(int i=threadIdx.x;i<S;i+=blockDim.x) atomicAdd(acc+i, s_acc[i]);
}
// F: gather smem + REDG
__global__ void __launch_bounds__(1024) k_gather_red(const double2* __restrict__ q2, const int4* __restrict__ c4, long long n4,
                         const double* __restrict__ tab, double* acc, int K, int S, int R){
  extern __shared__ double s_tab[];
  for(int i=threadIdx.x;i<S;i+=blockDim.x) s_tab[i]=tab[i]; __syncthreads();
  double* my = acc + (size_t)(blockIdx.x % R)*K;
  long long i = blockIdx.x*(long long)blockDim.x + threadIdx.x;
  for(; i<n4; i+=(long long)gridDim.x*blockDim.x){
    double2 a=q2[2*i], b=q2[2*i+1]; int4 c=c4[i];
    double t0 = c.x<S? s_tab[c.x]:__ldg(tab+c.x), t1 = c.y<S? s_tab[c.y]:__ldg(tab+c.y), t2 = c.z<S? s_tab[c.z]:__ldg(tab+c.z), t3 = c.w<S? s_tab[c.w]:__ldg(tab+c.w);
    atomicAdd(my+c.x, a.x*t0); atomicAdd(my+c.y, a.y*t1); atomicAdd(my+c.z, b.x*t2); atomicAdd(my+c.w, b.y*t3);
  }
}
// G: warp-per-row fused E+M (G lanes per row), gather smem, shuffle reduce, REDG
template<int G>
__global__ void __launch_bounds__(1024) k_rows_vec(const double* __restrict__ q, const int* __restrict__ col, const int* __restrict__ indptr, int N,
                         const double* __restrict__ tab, const double* __restrict__ wy, double* acc, int K, int S, int R, int do_red){
  extern __shared__ double s_tab[];
  for(int i=threadIdx.x;i<S;i+=blockDim.x) s_tab[i]=tab[i]; __syncthreads();
  double* my = acc + (size_t)(blockIdx.x % R)*K;
  const int lane = threadIdx.x & (G-1);
  long long grp = (blockIdx.x*(long long)blockDim.x + threadIdx.x)/G;
  long long ngrp = (long long)gridDim.x*blockDim.x/G;
  double sink=0;
  for(long long r=grp; r<N; r+=ngrp){
    int s=indptr[r], e=indptr[r+1];
    double sum=0;
    for(int k=s+lane;k<e;k+=G){ int c=col[k]; double t = c<S? s_tab[c]:__ldg(tab+c); sum += q[k]*t; }
    #pragma unroll
    for(int o=G/2;o>0;o>>=1) sum += __shfl_xor_sync(0xffffffffu, sum, o, G);
    double u = wy[r]/sum;
    for(int k=s+lane;k<e;k+=G){ int c=col[k]; double t = c<S? s_tab[c]:__ldg(tab+c); double v=q[k]*t*u; if(do_red) atomicAdd(my+c, v); else sink+=v; }
  }
  if(sink==-1.0) acc[0]=sink;
}
// H: thread-per-row straight from global
__global__ void __launch_bounds__(1024) k_rows_thread(const double* __restrict__ q, const int* __restrict__ col, const int* __restrict__ indptr, int N,
                         const double* __restrict__ tab, const double* __restrict__ wy, double* acc, int K, int S, int R, int do_red){
  extern __shared__ double s_tab[];
  for(int i=threadIdx.x;i<S;i+=blockDim.x) s_tab[i]=tab[i]; __syncthreads();
  double* my = acc + (size_t)(blockIdx.x % R)*K;
  double sink=0;
  for(long long r=blockIdx.x*(long long)blockDim.x+threadIdx.x; r<N; r+=(long long)gridDim.x*blockDim.x){
    int s=indptr[r], e=indptr[r+1];
    double sum=0;
    for(int k=s;k<e;k++){ int c=col[k]; double t = c<S? s_tab[c]:__ldg(tab+c); sum += q[k]*t; }
    double u = wy[r]/sum;
    for(int k=s;k<e;k++){ int c=col[k]; double t = c<S? s_tab[c]:__ldg(tab+c); double v=q[k]*t*u; if(do_red) atomicAdd(my+c, v); else sink+=v; }
  }
  if(sink==-1.0) acc[0]=sink;
}

template<class F> float timeit(F f, int reps=5){
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); CK(cudaDeviceSynchronize());
  float best=1e30f;
  for(int r=0;r<reps;r++){ cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms,a,b); best=std::min(best,ms);} 
  return best;
}

int main(int argc,char**argv){
  int N = argc>1? atoi(argv[1]) : 5000000; int K = argc>2? atoi(argv[2]) : 30000;
  std::vector<int> indptr(N+1); indptr[0]=0; uint32_t st=12345;
  for(int i=0;i<N;i++){ st = st*1664525u+1013904223u; int len = ((st>>10)%5==0)? 1 : 8 + (st>>12)%33; indptr[i+1]=indptr[i]+len; }
  long long nnz = indptr[N]; nnz -= nnz%4; // flat kernels use nnz/4 quads
  printf("N=%d K=%d nnz=%lld avg=%.2f\n",N,K,(long long)indptr[N],(double)indptr[N]/N);
  double *q,*tab,*acc,*wy,*sink; int *col,*ip; const int RMAX=32;
  CK(cudaMalloc(&q,(size_t)indptr[N]*8)); CK(cudaMalloc(&col,(size_t)indptr[N]*4)); CK(cudaMalloc(&ip,(size_t)(N+1)*4));
  CK(cudaMalloc(&tab,(size_t)K*8)); CK(cudaMalloc(&acc,(size_t)K*8*RMAX)); CK(cudaMalloc(&wy,(size_t)N*8)); CK(cudaMalloc(&sink,64));
  CK(cudaMemcpy(ip,indptr.data(),(size_t)(N+1)*4,cudaMemcpyHostToDevice));
  std::vector<double> ones(std::max(N,K),1.0); CK(cudaMemcpy(tab,ones.data(),(size_t)K*8,cudaMemcpyHostToDevice)); CK(cudaMemcpy(wy,ones.data(),(size_t)N*8,cudaMemcpyHostToDevice));
  CK(cudaMemset(acc,0,(size_t)K*8*RMAX));
  int nsm; cudaDeviceGetAttribute(&nsm,cudaDevAttrMultiProcessorCount,0);
  const int SM_S = std::min(K, 28000); size_t smem = (size_t)SM_S*8;
  CK(cudaFuncSetAttribute(k_gather<true>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem));
  CK(cudaFuncSetAttribute(k_red_smem,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem));
  CK(cudaFuncSetAttribute(k_gather_red,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem));
  CK(cudaFuncSetAttribute(k_rows_vec<32>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem));
  CK(cudaFuncSetAttribute(k_rows_vec<16>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem));
  CK(cudaFuncSetAttribute(k_rows_vec<8>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem));
  CK(cudaFuncSetAttribute(k_rows_thread,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem));
  long long n4=nnz/4; double GB = nnz*12.0/1e9;
  for(int skew=0;skew<2;skew++){
    k_gen<<<nsm*8,256>>>(q,col,indptr[N],K,skew); CK(cudaDeviceSynchronize());
    printf("---- cols %s ----\n", skew?"skewed(u^3)":"uniform");
    auto rep=[&](const char* name,float ms){ printf("%-44s %8.3f ms  %7.1f GB/s(12B/nnz)  %6.2f Gnnz/s\n",name,ms,GB/ms*1e3,nnz/ms/1e6); fflush(stdout); };
    if(!skew){
      rep("A stream only (512x4/SM)", timeit([&]{k_stream<<<nsm*4,512>>>((double2*)q,(int4*)col,n4,sink);}));
      rep("A stream only (512x16/SM grid)", timeit([&]{k_stream<<<nsm*16,512>>>((double2*)q,(int4*)col,n4,sink);}));
    }
    rep("B stream+gather smem(28k)+ldg rest", timeit([&]{k_gather<true><<<nsm,1024,smem>>>((double2*)q,(int4*)col,n4,tab,K,SM_S,sink);}));
    rep("C stream+gather ldg (L1/L2)", timeit([&]{k_gather<false><<<nsm*2,1024>>>((double2*)q,(int4*)col,n4,tab,K,0,sink);}));
    for(int R: {1,8,32}){ char nm[64]; snprintf(nm,64,"D stream+REDG.F64 global R=%d",R);
      rep(nm, timeit([&]{k_red_global<<<nsm*4,512>>>((double2*)q,(int4*)col,n4,acc,K,R);})); }
    rep("E stream+smem CAS f64 (28k) + REDG rest", timeit([&]{k_red_smem<<<nsm,1024,smem>>>((double2*)q,(int4*)col,n4,acc,K,SM_S);}));
    { int S2=std::min(K,8000); rep("E' stream+smem CAS f64 (8k hot) + REDG rest", timeit([&]{k_red_smem<<<nsm,1024,smem>>>((double2*)q,(int4*)col,n4,acc,K,S2);})); }
    for(int R: {1,8}){ char nm[64]; snprintf(nm,64,"F gather smem + REDG R=%d",R);
      rep(nm, timeit([&]{k_gather_red<<<nsm,1024,smem>>>((double2*)q,(int4*)col,n4,tab,acc,K,SM_S,R);})); }
    for(int dr=0; dr<2; dr++){
      char nm[64];
      snprintf(nm,64,"G rows warp(32)/row red=%d",dr); rep(nm, timeit([&]{k_rows_vec<32><<<nsm,1024,smem>>>(q,col,ip,N,tab,wy,acc,K,SM_S,8,dr);}));
      snprintf(nm,64,"G rows 16 lanes/row red=%d",dr); rep(nm, timeit([&]{k_rows_vec<16><<<nsm,1024,smem>>>(q,col,ip,N,tab,wy,acc,K,SM_S,8,dr);}));
      snprintf(nm,64,"G rows 8 lanes/row red=%d",dr); rep(nm, timeit([&]{k_rows_vec<8><<<nsm,1024,smem>>>(q,col,ip,N,tab,wy,acc,K,SM_S,8,dr);}));
      snprintf(nm,64,"H rows thread/row red=%d",dr); rep(nm, timeit([&]{k_rows_thread<<<nsm,1024,smem>>>(q,col,ip,N,tab,wy,acc,K,SM_S,8,dr);}));
    }
  }
  return 0;
}

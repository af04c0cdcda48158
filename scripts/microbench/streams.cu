// Memory-system ceiling probes for the D3Q19 access pattern (no LBM math).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o streams streams.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

struct P { const float* in[19]; float* out[19]; uint32_t n; int sx, sy; };
__constant__ int c_off[19];

// (a) plain copy
__global__ void k_copy4(const float4* __restrict__ a, float4* __restrict__ b, size_t n4){
  size_t i=(size_t)blockIdx.x*blockDim.x+threadIdx.x; if(i<n4) b[i]=a[i];
}
// (b) 19 planes, aligned
template<int CS> __global__ void __launch_bounds__(256) k_soa(const P p){
  uint32_t i=blockIdx.x*blockDim.x+threadIdx.x; if(i>=p.n) return;
  float f[19];
  #pragma unroll
  for(int s=0;s<19;s++) f[s]=__ldg(p.in[s]+i);
  #pragma unroll
  for(int s=0;s<19;s++){ if(CS) __stcs(p.out[s]+i,f[s]); else p.out[s][i]=f[s]; }
}
// (c) 19 planes, pull-shifted reads (in[] pre-shifted by host)
// same kernel as (b) with shifted pointers
// (e) 2 nodes per thread, float2 everywhere (aligned variant)
__global__ void __launch_bounds__(256) k_soa2(const P p){
  uint32_t i=(blockIdx.x*blockDim.x+threadIdx.x); if(2*i>=p.n) return;
  float2 f[19];
  #pragma unroll
  for(int s=0;s<19;s++) f[s]=__ldg((const float2*)p.in[s]+i);
  #pragma unroll
  for(int s=0;s<19;s++) ((float2*)p.out[s])[i]=f[s];
}
__global__ void __launch_bounds__(256) k_soa4(const P p){
  uint32_t i=(blockIdx.x*blockDim.x+threadIdx.x); if(4*i>=p.n) return;
  float4 f[19];
  #pragma unroll
  for(int s=0;s<19;s++) f[s]=__ldg((const float4*)p.in[s]+i);
  #pragma unroll
  for(int s=0;s<19;s++) ((float4*)p.out[s])[i]=f[s];
}
// (d) blocked layout [tile][19][B]: node i -> tile=i/B, w=i%B
template<int B> __global__ void __launch_bounds__(256) k_blocked(const float* __restrict__ in, float* __restrict__ out, uint32_t n){
  uint32_t i=blockIdx.x*blockDim.x+threadIdx.x; if(i>=n) return;
  size_t base=(size_t)(i/B)*19*B + (i%B);
  float f[19];
  #pragma unroll
  for(int s=0;s<19;s++) f[s]=__ldg(in+base+s*B);
  #pragma unroll
  for(int s=0;s<19;s++) out[base+s*B]=f[s];
}
// persistent grid-stride variant of (b)
__global__ void __launch_bounds__(256) k_soa_persist(const P p){
  for(uint32_t i=blockIdx.x*blockDim.x+threadIdx.x; i<p.n; i+=gridDim.x*blockDim.x){
    float f[19];
    #pragma unroll
    for(int s=0;s<19;s++) f[s]=__ldg(p.in[s]+i);
    #pragma unroll
    for(int s=0;s<19;s++) p.out[s][i]=f[s];
  }
}
// fake compute: ~200 FMAs between load and store
__global__ void __launch_bounds__(256) k_soa_compute(const P p, int iters){
  uint32_t i=blockIdx.x*blockDim.x+threadIdx.x; if(i>=p.n) return;
  float f[19];
  #pragma unroll
  for(int s=0;s<19;s++) f[s]=__ldg(p.in[s]+i);
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int s=0;s<19;s++) f[s]=fmaf(f[s],1.0001f,f[(s+1)%19]*0.0001f);
  }
  #pragma unroll
  for(int s=0;s<19;s++) p.out[s][i]=f[s];
}

template<class F> float timeit(F f, int rep=20){
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  for(int i=0;i<3;i++) f();
  cudaEventRecord(a); for(int i=0;i<rep;i++) f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms,a,b); CK(cudaGetLastError()); return ms/rep;
}
int main(int argc,char**argv){
  int nx=256,ny=256,nz=256; if(argc>1) nx=ny=nz=atoi(argv[1]);
  size_t N=(size_t)nx*ny*nz, pad=(size_t)ny*nz+nz+32;
  size_t tot=19*N+2*pad;
  float *A,*B; CK(cudaMalloc(&A,tot*4)); CK(cudaMalloc(&B,tot*4)); CK(cudaMemset(A,0,tot*4)); CK(cudaMemset(B,0,tot*4));
  double bytes=2.0*19*N*4;
  const int e[19][3]={{0,0,0},{1,0,0},{-1,0,0},{0,1,0},{0,-1,0},{0,0,1},{0,0,-1},{1,1,0},{-1,-1,0},{1,-1,0},{-1,1,0},{1,0,1},{-1,0,-1},{1,0,-1},{-1,0,1},{0,1,1},{0,-1,-1},{0,1,-1},{0,-1,1}};
  P al, sh; al.n=sh.n=(uint32_t)N; al.sx=sh.sx=ny*nz; al.sy=sh.sy=nz;
  for(int s=0;s<19;s++){ al.in[s]=A+pad+s*N; al.out[s]=B+pad+s*N; sh.out[s]=B+pad+s*N; sh.in[s]=A+pad+s*N-((long long)e[s][0]*ny*nz+e[s][1]*nz+e[s][2]); }
  float ms;
  ms=timeit([&]{ k_copy4<<<(19*N/4+255)/256,256>>>((const float4*)(A+pad),(float4*)(B+pad),19*N/4); });
  printf("copy float4 2-stream           %.3f ms  %.0f GB/s\n",ms,bytes/ms/1e6);
  for(int blk: {128,256}){
    ms=timeit([&]{ k_soa<0><<<(N+blk-1)/blk,blk>>>(al); });
    printf("soa19 aligned blk%-4d           %.3f ms  %.0f GB/s\n",blk,ms,bytes/ms/1e6);
  }
  ms=timeit([&]{ k_soa<1><<<(N+255)/256,256>>>(al); });
  printf("soa19 aligned st.cs            %.3f ms  %.0f GB/s\n",ms,bytes/ms/1e6);
  ms=timeit([&]{ k_soa<0><<<(N+255)/256,256>>>(sh); });
  printf("soa19 pull-shifted             %.3f ms  %.0f GB/s\n",ms,bytes/ms/1e6);
  ms=timeit([&]{ k_soa<1><<<(N+255)/256,256>>>(sh); });
  printf("soa19 pull-shifted st.cs       %.3f ms  %.0f GB/s\n",ms,bytes/ms/1e6);
  ms=timeit([&]{ k_soa2<<<(N/2+255)/256,256>>>(al); });
  printf("soa19 aligned float2           %.3f ms  %.0f GB/s\n",ms,bytes/ms/1e6);
  ms=timeit([&]{ k_soa4<<<(N/4+255)/256,256>>>(al); });
  printf("soa19 aligned float4           %.3f ms  %.0f GB/s\n",ms,bytes/ms/1e6);
  ms=timeit([&]{ k_blocked<256><<<(N+255)/256,256>>>(A+pad,B+pad,(uint32_t)N); });
  printf("blocked [tile][19][256]        %.3f ms  %.0f GB/s\n",ms,bytes/ms/1e6);
  ms=timeit([&]{ k_blocked<32><<<(N+255)/256,256>>>(A+pad,B+pad,(uint32_t)N); });
  printf("blocked [tile][19][32]         %.3f ms  %.0f GB/s\n",ms,bytes/ms/1e6);
  ms=timeit([&]{ k_blocked<1024><<<(N+255)/256,256>>>(A+pad,B+pad,(uint32_t)N); });
  printf("blocked [tile][19][1024]       %.3f ms  %.0f GB/s\n",ms,bytes/ms/1e6);
  for(int g: {148*4,148*6,148*8}){
    ms=timeit([&]{ k_soa_persist<<<g,256>>>(al); });
    printf("soa19 persistent grid %-5d     %.3f ms  %.0f GB/s\n",g,ms,bytes/ms/1e6);
  }
  for(int it: {0,2,5,10}){
    ms=timeit([&]{ k_soa_compute<<<(N+255)/256,256>>>(sh,it); });
    printf("soa19 shifted + %2d x19 FMA x2   %.3f ms  %.0f GB/s\n",it,ms,bytes/ms/1e6);
  }
  return 0;
}

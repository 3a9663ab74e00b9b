// Micro-benchmark: peak FP64 rate of DMMA (mma.sync m8n8k4 f64) vs DFMA on this GPU, plus a
// stream-copy HBM number, used to choose the Legendre kernel's inner product and to state
// the fp64 roofline denominator (there is no fp64 entry in MEASURED_PEAKS.json).
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

template<int NACC>
__global__ void dmma_peak(double* out, int iters){
  double c[NACC][2];
  #pragma unroll
  for(int i=0;i<NACC;i++){c[i][0]=0; c[i][1]=0;}
  double a = 1.0 + threadIdx.x*1e-9, b = 1.0 - threadIdx.x*1e-9;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<NACC;i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]),"+d"(c[i][1]) : "d"(a),"d"(b));
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i][0]+c[i][1];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int NACC>
__global__ void dfma_peak(double* out, int iters){
  double c[NACC];
  #pragma unroll
  for(int i=0;i<NACC;i++) c[i]=i;
  double a = 1.0 + threadIdx.x*1e-9, b = 1e-9*threadIdx.x;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<NACC;i++) c[i] = fma(c[i], a, b);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void copyk(double2* __restrict__ d, const double2* __restrict__ s, size_t n){
  size_t i = blockIdx.x*(size_t)blockDim.x+threadIdx.x; size_t st=(size_t)gridDim.x*blockDim.x;
  for(; i<n; i+=st) d[i]=s[i];
}
int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  printf("device %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  double* out; CK(cudaMalloc(&out, sizeof(double)*148*8*1024));
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters=20000; float ms;
  for(int warps=4; warps<=32; warps*=2){
    int blocks=p.multiProcessorCount*2, threads=warps*32/2; if(threads>1024) threads=1024;
    dmma_peak<8><<<blocks,threads>>>(out, 100); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); dmma_peak<8><<<blocks,threads>>>(out, iters); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms,e0,e1);
    double fl = (double)blocks*(threads/32)*iters*8*(2.0*8*8*4);
    printf("DMMA m8n8k4  blocks %d threads %d : %.2f TFLOP/s (%.3f ms)\n", blocks, threads, fl/ms*1e-9, ms);
    dfma_peak<8><<<blocks,threads>>>(out, 100); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); dfma_peak<8><<<blocks,threads>>>(out, iters*8); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms,e0,e1);
    fl = (double)blocks*threads*iters*8.0*8*2.0;
    printf("DFMA         blocks %d threads %d : %.2f TFLOP/s (%.3f ms)\n", blocks, threads, fl/ms*1e-9, ms);
  }
  size_t n = (size_t)1<<27; // 2 GiB per buffer of double2
  double2 *s,*d; CK(cudaMalloc(&s,n*16)); CK(cudaMalloc(&d,n*16)); CK(cudaMemset(s,1,n*16));
  for(int r=0;r<3;r++){
    cudaEventRecord(e0); copyk<<<148*16,512>>>(d,s,n); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms,e0,e1);
    printf("copy 2x%.1f GiB: %.1f GB/s\n", n*16/1073741824.0, 2.0*n*16/ms*1e-6);
  }
  return 0;
}

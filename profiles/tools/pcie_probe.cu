// pcie_probe.cu — how fast do the copies of the e2e path run on this box?  (nvcc -O2 -o pcie_probe pcie_probe.cu)
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)
int main(){
  const int W=1241,H=376,N=512; const size_t pitch=1280, rows=H+6;
  unsigned char *h,*d,*d2; CK(cudaHostAlloc((void**)&h,(size_t)N*W*H,cudaHostAllocDefault)); CK(cudaMalloc((void**)&d,(size_t)N*pitch*rows)); CK(cudaMalloc((void**)&d2,(size_t)N*W*H));
  cudaStream_t s; CK(cudaStreamCreate(&s)); cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); float ms;
  for(int rep=0;rep<2;++rep){
    cudaEventRecord(a,s); CK(cudaMemcpyAsync(d2,h,(size_t)N*W*H,cudaMemcpyHostToDevice,s)); cudaEventRecord(b,s); cudaEventSynchronize(b); cudaEventElapsedTime(&ms,a,b);
    printf("H2D contiguous   %7.2f ms  %6.1f GB/s\n",ms,(double)N*W*H/ms/1e6);
    cudaMemcpy3DParms p={}; p.srcPtr=make_cudaPitchedPtr(h,W,W,H); p.dstPtr=make_cudaPitchedPtr(d+3*pitch+4,pitch,W,rows); p.extent=make_cudaExtent(W,H,N); p.kind=cudaMemcpyHostToDevice;
    cudaEventRecord(a,s); CK(cudaMemcpy3DAsync(&p,s)); cudaEventRecord(b,s); cudaEventSynchronize(b); cudaEventElapsedTime(&ms,a,b);
    printf("H2D 3D pitched   %7.2f ms  %6.1f GB/s\n",ms,(double)N*W*H/ms/1e6);
    cudaEventRecord(a,s); for(int i=0;i<N;++i) CK(cudaMemcpy2DAsync(d+(size_t)i*pitch*rows+3*pitch+4,pitch,h+(size_t)i*W*H,W,W,H,cudaMemcpyHostToDevice,s)); cudaEventRecord(b,s); cudaEventSynchronize(b); cudaEventElapsedTime(&ms,a,b);
    printf("H2D 2D per image %7.2f ms  %6.1f GB/s\n",ms,(double)N*W*H/ms/1e6);
    p.srcPtr=make_cudaPitchedPtr(d2,W,W,H); p.kind=cudaMemcpyDeviceToDevice;
    cudaEventRecord(a,s); CK(cudaMemcpy3DAsync(&p,s)); cudaEventRecord(b,s); cudaEventSynchronize(b); cudaEventElapsedTime(&ms,a,b);
    printf("D2D 3D repack    %7.2f ms  %6.1f GB/s\n",ms,(double)N*W*H/ms/1e6);
    cudaEventRecord(a,s); CK(cudaMemcpyAsync(h,d2,(size_t)N*W*H,cudaMemcpyDeviceToHost,s)); cudaEventRecord(b,s); cudaEventSynchronize(b); cudaEventElapsedTime(&ms,a,b);
    printf("D2H contiguous   %7.2f ms  %6.1f GB/s\n",ms,(double)N*W*H/ms/1e6);
  }
  return 0;
}

// Calibration only (not linked into the product): time cub::DeviceRadixSort::SortPairs on the same
// record volume (u64 key + u64 value, 42 key bits) to see what NVIDIA's own onesweep reaches on this GPU.
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdlib>
typedef unsigned long long u64;
__global__ void fill(u64* k, u64* v, u64 n, int nbits){
  u64 stride=(u64)gridDim.x*blockDim.x; u64 mask=(1ull<<nbits)-1;
  for(u64 i=(u64)blockIdx.x*blockDim.x+threadIdx.x;i<n;i+=stride){u64 x=i+12345*0x9e3779b97f4a7c15ull; x^=x>>30; x*=0xbf58476d1ce4e5b9ull; x^=x>>27; x*=0x94d049bb133111ebull; x^=x>>31; k[i]=x&mask; v[i]=i;}
}
int main(int argc,char**argv){
  u64 n = argc>1? (u64)atof(argv[1]) : 100000000ull; int nbits = argc>2? atoi(argv[2]):42;
  u64 *k0,*k1,*v0,*v1; cudaMalloc(&k0,n*8);cudaMalloc(&k1,n*8);cudaMalloc(&v0,n*8);cudaMalloc(&v1,n*8);
  size_t tb=0; cub::DeviceRadixSort::SortPairs(nullptr,tb,k0,k1,v0,v1,n,0,nbits);
  void* tmp; cudaMalloc(&tmp,tb);
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  for(int r=0;r<4;r++){
    fill<<<1184,256>>>(k0,v0,n,nbits);
    cudaEventRecord(a);
    cub::DeviceRadixSort::SortPairs(tmp,tb,k0,k1,v0,v1,n,0,nbits);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms,a,b);
    int passes=(nbits+7)/8;
    printf("cub SortPairs n=%llu bits=%d: %.3f ms total, %.3f ms/pass (%d passes), %.1f GB/s per pass-equivalent\n",n,nbits,ms,ms/passes,passes, 32.0*n/(ms/passes*1e-3)/1e9);
  }
  return 0;
}

#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s){ uint32_t d; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s)); return d; }
__global__ void k(uint32_t* o, uint32_t x, uint32_t y){
  o[0] = prmt(x, 0, 0xB391);   // x=0x80FF7F01 -> expect 0xFF80007F
  o[1] = prmt(y, 0, 0xAA88);   // y=0x00800000 -> expect 0xFFFF0000
  o[2] = prmt(y, 0, 0xBB99);   // bytes 1,3 sign: y -> b1=0x00,b3=0x00 -> 0
  o[3] = prmt(x, 0, 0x4341);   // zero-extend hi bytes
}
int main(){ uint32_t* d; cudaMalloc(&d, 64); k<<<1,1>>>(d, 0x80FF7F01u, 0x00800000u); uint32_t h[4]; cudaMemcpy(h,d,16,cudaMemcpyDeviceToHost);
  for(int i=0;i<4;i++) printf("o[%d]=0x%08X\n", i, h[i]); return 0; }

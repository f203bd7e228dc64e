// Checks, on the device, the per-halfword semantics the DPX fill kernel relies on.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(uint32_t* o){
  // 0x7F00 + 0x0300 = 0x8200 wraps negative in 16 bit: does viaddmax wrap before the max?
  o[0] = __viaddmax_s16x2(0x7F007F00u, 0x03000100u, 0x00050005u);   // expect wrap: lo: 0x7F00+0x0100=0x8000 (neg) -> max(.,5)=5 ; hi: 0x8200 -> 5
  o[1] = __viaddmin_s16x2(0x7F007F00u, 0x03000100u, 0x00050005u);   // wrap: negative values stay
  o[2] = __vimax3_s16x2(0x80008000u, 0x7FFF0001u, 0xFFFF0000u);     // hi: max(-32768,32767,-1)=7FFF ; lo: max(-32768,1,0)=1
  o[3] = __vminu2(0x80000001u, 0x7FFF8000u);                   // unsigned: hi 7FFF lo 0001
  o[4] = __vmaxu2(0x80000001u, 0x7FFF8000u);                   // hi 8000 lo 8000
  o[5] = __byte_perm(0x80FF7F01u, 0, 0xB391);                       // sign-extend hi bytes of each half: b1=0x7F -> 0x007F ; b3=0x80 -> 0xFF80
  o[6] = __byte_perm(0x80FF7F01u, 0, 0x4341);                       // zero-extend: 0x0080007F
  o[7] = __byte_perm(0x00800000u << 0, 0, 0xAA88);                  // sign replicate byte0 (0x00) and byte2 (0x80): expect 0xFFFF0000
  o[8] = __vadd2(0xFFFF0001u, 0x0001FFFFu);                         // 0x00000000 per-half wrap, no carry across
  o[9] = __vsub2(0x00000000u, 0x00010001u);                         // 0xFFFFFFFF
  o[10] = __vmaxs2(0x01030103u, 0x01040102u);                  // 0x01040103
  bool ph, pl; o[11] = __vibmax_s16x2(0x00050003u, 0x00040003u, &ph, &pl); o[12] = (ph?2:0)|(pl?1:0);
}
int main(){ uint32_t* d; cudaMalloc(&d, 64*4); k<<<1,1>>>(d); uint32_t h[16]; cudaMemcpy(h,d,64,cudaMemcpyDeviceToHost);
  for(int i=0;i<13;i++) printf("o[%d]=0x%08X\n", i, h[i]); return 0; }

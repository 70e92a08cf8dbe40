// pipe_probe.cu -- issue-rate probe for the integer instructions the DP kernels are built from (sm_100a).
// Each kernel runs a long dependent-free stream of one instruction kind (8 independent chains per thread) and
// reports warp-instructions per clock per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096
enum { OP_VIMNMX3 = 0, OP_VIADDMNMX, OP_VIMNMX_P, OP_VIMNMX, OP_LOP3, OP_IADD3, OP_IMAD, OP_PRMT, OP_SEL, OP_VIADD16,
       OP_MIX_ALU_FMA, OP_MIX_DPX_IMAD, OP_CELL, OP_N };
static const char* NAMES[] = {"VIMNMX3.U16x2", "VIADDMNMX.U16x2", "VIMNMX.S16x2+preds(+2 SEL)", "VIMNMX.U16x2", "LOP3", "IADD3",
                              "IMAD", "PRMT", "SEL", "VIADD.16x2", "LOP3+IMAD 1:1", "VIADDMNMX+IMAD 1:1", "cell mix (22 ALU + 6 FMA)"};

template <int OP> __global__ void probe(uint32_t* out, uint32_t seed, long long* clk)
{
  uint32_t r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = seed * (i + 1) + threadIdx.x;
  uint32_t c1 = seed ^ 0x00070007u, c2 = seed + 0x00080008u;
  long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint32_t x = r[i];
      if (OP == OP_VIMNMX3) x = __vimax3_u16x2(x, c1, c2);
      if (OP == OP_VIADDMNMX) x = __viaddmax_u16x2(x, c1, c2);
      if (OP == OP_VIMNMX_P) { bool ph, pl; x = __vibmax_s16x2(x, c1, &ph, &pl); c2 = pl ? c2 : 0x11u; c1 = ph ? c1 : 0x22u; }
      if (OP == OP_VIMNMX) x = __vmaxu2(x, c1);
      if (OP == OP_LOP3) x = (x & c1) ^ c2;
      if (OP == OP_IADD3) x = x + c1 + c2;
      if (OP == OP_IMAD) x = x * c1 + c2;
      if (OP == OP_PRMT) x = __byte_perm(x, c1, 0x5410 + (c2 & 0x1000));
      if (OP == OP_SEL) x = ((int)c1 > it) ? x : c2;
      if (OP == OP_VIADD16) x = __vadd2(x, c1);
      if (OP == OP_MIX_ALU_FMA) { x = (x & c1) ^ c2; x = x * c1 + c2; }
      if (OP == OP_MIX_DPX_IMAD) { x = __viaddmax_u16x2(x, c1, c2); x = x * c1 + c2; }
      if (OP == OP_CELL) {
        // instruction mix of one packed 5-state cell: 12 DPX, 5 LOP3, 5 IADD3, 1 PRMT, 6 IMAD
        uint32_t a = __vimax3_u16x2(x, c1, c2), b = __vimax3_u16x2(a, c2, x);
        uint32_t p0 = b & 0x00070007u; uint32_t m = b - p0 + c1;
        uint32_t d = __viaddmax_u16x2(m, c1, x); d = __vimax3_u16x2(d, a, c2);
        uint32_t p1 = d & 0x00070007u; d = d - p1 + c2;
        uint32_t e = __viaddmax_u16x2(d, c1, m); uint32_t p2 = e & 0x00070007u; e = e - p2 + c1;
        uint32_t f = __vmaxu2(m, e); f = __viaddmax_u16x2(f, c1, d); f = __viaddmax_u16x2(a, c2, f);
        uint32_t p3 = f & 0x00070007u; f = f - p3 + c1;
        uint32_t g = __viaddmax_u16x2(f, c2, e); uint32_t p4 = g & 0x00070007u; g = g - p4 + c2;
        uint32_t q = x ^ c1; q = __vminu2(q, 0x00010001u); q = q * c2 + c1;
        uint32_t cell = p4; cell = cell * 8 + p3; cell = cell * 8 + p2; cell = cell * 8 + p1; cell = cell * 8 + p0;
        x = __byte_perm(cell, g + q, 0x5410);
      }
      r[i] = x;
    }
  }
  long long t1 = clock64();
  uint32_t acc = c1 ^ c2;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc ^= r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int OP> void run(int instrPerIter, uint32_t* dOut, long long* dClk, int warpsPerSm)
{
  const int threads = warpsPerSm * 32;
  probe<OP><<<148, threads>>>(dOut, 12345u, dClk);
  cudaDeviceSynchronize();
  probe<OP><<<148, threads>>>(dOut, 12345u, dClk);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, dClk, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += (double)h[i];
  avg /= 148;
  const double winstr = (double)ITER * 8 * instrPerIter * warpsPerSm;
  printf("%-32s warps/SM %2d  %.3f warp-instr/clk/SM (%.2f per SMSP)\n", NAMES[OP], warpsPerSm, winstr / avg, winstr / avg / 4);
}

int main()
{
  uint32_t* dOut; long long* dClk;
  cudaMalloc(&dOut, 148 * 1024 * 4);
  cudaMalloc(&dClk, 148 * 8);
  for (int w : {8, 16, 32}) {
    run<OP_VIMNMX3>(1, dOut, dClk, w);   run<OP_VIADDMNMX>(1, dOut, dClk, w); run<OP_VIMNMX_P>(3, dOut, dClk, w);
    run<OP_VIMNMX>(1, dOut, dClk, w);    run<OP_LOP3>(1, dOut, dClk, w);      run<OP_IADD3>(1, dOut, dClk, w);
    run<OP_IMAD>(1, dOut, dClk, w);      run<OP_PRMT>(1, dOut, dClk, w);      run<OP_SEL>(1, dOut, dClk, w);
    run<OP_VIADD16>(1, dOut, dClk, w);   run<OP_MIX_ALU_FMA>(2, dOut, dClk, w); run<OP_MIX_DPX_IMAD>(2, dOut, dClk, w);
    run<OP_CELL>(29, dOut, dClk, w);
  }
  printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

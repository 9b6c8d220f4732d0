// pipe_mix.cu -- how do the FP64 pipe (DFMA), the heavy FMA pipe (IMAD.WIDE) and the ALU pipe (IADD3 carry pairs) of a
// B200 SM share issue slots?  Synthetic instruction mixes, F DFMA : I IMAD.WIDE : A IADD3 per block, each type on 8
// independent dependency chains.  Decides whether a hybrid field multiplier (limb products on DFMA, Montgomery
// reduction on IMAD.WIDE) can beat the 276 IMAD.WIDE product.  Development tool, not product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/pipe_mix tools/pipe_mix.cu && build/pipe_mix
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

typedef uint32_t u32;
typedef uint64_t u64;

template <int F, int I, int A, int REP>
__global__ void __launch_bounds__(256) mix_kernel(u32* sink, u32 iters, u32 seed) {
  const u32 tid = blockIdx.x * blockDim.x + threadIdx.x;
  double d[8];
  u64 w[8];
  u32 lo[8], hi[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    d[k] = (double)(seed + tid * 8 + k);
    w[k] = seed + tid * 8 + k;
    lo[k] = seed ^ (tid + k);
    hi[k] = seed + k;
  }
  const double x = (double)(seed | 1u), y = 1.0 / (double)(seed | 3u);
  const u32 m = seed | 1u;
  for (u32 it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < REP; r++) {
      // interleave the three types as evenly as integer counts allow
      constexpr int T = (F > I ? (F > A ? F : A) : (I > A ? I : A));
      int f = 0, i = 0, a = 0;
#pragma unroll
      for (int s = 1; s <= T; s++) {
        if (f * T < s * F) {
          asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d[f & 7]) : "d"((f & 1) ? x : y), "d"(y));
          f++;
        }
        if (i * T < s * I) {
          // the second multiplicand changes every iteration: a loop-invariant product would be hoisted into an addition
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i & 7]) : "r"(m), "r"(it + (u32)(i + 3)));
          i++;
        }
        if (a * T < s * A) {
          asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(lo[(a >> 1) & 7]), "+r"(hi[(a >> 1) & 7]) : "r"(m), "r"(seed));
          a += 2;
        }
      }
    }
  }
  u64 s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s ^= (u64)__double_as_longlong(d[k]) ^ w[k] ^ lo[k] ^ ((u64)hi[k] << 32);
  if (s == 0x12345678u) sink[0] = (u32)s;
}

template <int F, int I, int A, int REP>
static void run(const char* what, int sms, double clk_ghz) {
  u32* sink;
  cudaMalloc(&sink, 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const u32 grid = sms * 8, block = 256, iters = 2048;
  float ms = 0.f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    mix_kernel<F, I, A, REP><<<grid, block>>>(sink, iters, 12345u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  // clocks per SM per thread-level block (F + I + A instructions): time * f / (threads per SM * iters * REP)
  const double thr_per_sm = 8.0 * 256.0;
  const double clk_per_block = ms * 1e-3 * clk_ghz * 1e9 / (thr_per_sm * iters * REP);
  printf("%-44s F=%3d I=%3d A=%3d  %8.3f ms  %7.3f clk/SM per thread-block  (F/clk %.1f  I/clk %.1f  A/clk %.1f  total/clk %.1f)\n", what, F,
         I, A, ms, clk_per_block, F / clk_per_block, I / clk_per_block, A / clk_per_block, (F + I + A) / clk_per_block);
  cudaFree(sink);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz / 1e6;
  printf("%s, %d SMs, %.3f GHz nominal (clk/SM figures assume this clock)\n", p.name, p.multiProcessorCount, ghz);
  const int sms = p.multiProcessorCount;
  run<8, 0, 0, 8>("DFMA alone", sms, ghz);
  run<0, 8, 0, 8>("IMAD.WIDE alone", sms, ghz);
  run<0, 0, 8, 8>("IADD3 carry pairs alone", sms, ghz);
  run<8, 8, 0, 8>("DFMA + IMAD.WIDE 1:1", sms, ghz);
  run<16, 8, 0, 4>("DFMA + IMAD.WIDE 2:1", sms, ghz);
  run<8, 0, 8, 8>("DFMA + IADD3 1:1", sms, ghz);
  run<0, 8, 8, 8>("IMAD.WIDE + IADD3 1:1", sms, ghz);
  run<0, 8, 16, 4>("IMAD.WIDE + IADD3 1:2", sms, ghz);
  run<8, 8, 8, 8>("all three 1:1:1", sms, ghz);
  run<16, 8, 16, 4>("all three 2:1:2", sms, ghz);
  run<0, 276, 60, 1>("today's Fq product: 276 IMAD.WIDE + 60 IADD3", sms, ghz);
  run<208, 132, 300, 1>("hybrid 48-bit limbs: 208 F + 132 I + 300 A", sms, ghz);
  run<208, 132, 200, 1>("hybrid 48-bit limbs, leaner ALU: 208/132/200", sms, ghz);
  run<319, 132, 311, 1>("hybrid 24-bit limbs: 319 F + 132 I + 311 A", sms, ghz);
  run<288, 132, 160, 1>("hybrid 24-bit, lean conversions: 288/132/160", sms, ghz);
  run<400, 0, 300, 1>("all-DFMA 48-bit limbs: 400 F + 300 A", sms, ghz);
  run<120, 210, 150, 1>("half hybrid: 120 F + 210 I + 150 A", sms, ghz);
  run<408, 0, 420, 1>("all-DFMA 48-bit limbs as counted in DESIGN.md: 408 F + 420 A", sms, ghz);
  run<384, 0, 256, 1>("all-DFMA floor: 384 F + 256 A (no conversions)", sms, ghz);
  return 0;
}

// line_atomicity.cu -- litmus test for the assumption the tagged ("flag in data") rows of csrc/tagged_kernels.cuh rest on:
//
//   a 128-byte line written by ONE warp-wide store instruction (8 lanes x 16 bytes, st.global.cg.v2.f64 -- the pattern of
//   NCCL's LL128 protocol) is observed by a 128-byte load (8 lanes x ld.global.cg.v2.f64, one instruction) either entirely
//   old or entirely new: never a mix of sectors from two different stores.
//
// Writers (half of the CTAs) keep rewriting every line of a small array with the value k in all 16 doubles, k = 1, 2, ...
// (so every store of a line carries a different, self-describing payload); readers (the other half) keep loading whole
// lines and count loads in which the 16 values are not all equal ("torn").  Both 8 x 16-byte and 4 x 32-byte
// (st.global.cg.v4.f64) store shapes are tested, lines hot in L2 (4 K lines) and streaming through DRAM (64 M lines).
// The PTX memory model does not promise this atomicity; the test documents what sm_100 does.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o line_atomicity line_atomicity.cu && ./line_atomicity
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ void st2(double* p, double a, double b) {
  asm volatile("st.global.cg.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void ld2(const double* p, double& a, double& b) {
  asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void st4(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.cg.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void ld4(const double* p, double& a, double& b, double& c, double& d) {
  asm volatile("ld.global.cg.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p) : "memory");
}

// WIDE = false: a group = 8 lanes x 16 B;  WIDE = true: a group = 4 lanes x 32 B
template <bool WIDE>
__global__ void litmus(double* lines, long long num_lines, int rounds, unsigned long long* torn, unsigned long long* loads,
                       unsigned long long* changed) {
  constexpr int LPR = WIDE ? 4 : 8;
  const int lane = threadIdx.x & 31, gl = lane % LPR, gw = lane / LPR;
  const unsigned gmask = (((1u << LPR) - 1u) << (gw * LPR));
  const long long groups_per_side = (long long)(gridDim.x / 2) * (blockDim.x / LPR);
  const long long g = (long long)(blockIdx.x / 2) * (blockDim.x / LPR) + threadIdx.x / LPR;
  const bool writer = (blockIdx.x & 1) == 0;
  unsigned long long my_torn = 0, my_loads = 0, my_changed = 0;
  for (int r = 1; r <= rounds; r++) {
    for (long long ln = g; ln < num_lines; ln += groups_per_side) {
      double* p = lines + ln * 16;
      if (writer) {
        const double v = (double)r;
        if (WIDE) st4(p + 4 * gl, v, v, v, v);
        else st2(p + 2 * gl, v, v);
      } else {
        double a, b, c = 0, d = 0;
        if (WIDE) ld4(p + 4 * gl, a, b, c, d);
        else { ld2(p + 2 * gl, a, b); c = a; d = a; }
        const double first = __shfl_sync(gmask, a, gw * LPR);
        const bool same = (a == first) && (b == first) && (c == first) && (d == first);
        const unsigned ok = __ballot_sync(gmask, same);
        if (gl == 0) {
          my_loads++;
          if ((ok & gmask) != gmask) my_torn++;
          if (first != 0.0) my_changed++;
        }
      }
    }
  }
  if (!writer && gl == 0) {
    atomicAdd(torn, my_torn);
    atomicAdd(loads, my_loads);
    atomicAdd(changed, my_changed);
  }
}

int main() {
  unsigned long long* ctr;
  cudaMallocManaged(&ctr, 3 * sizeof(unsigned long long));
  int sm = 0;
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
  struct Case { long long lines; int rounds; const char* what; } cases[] = {
      {4096, 20000, "4 K lines (L2-resident, heavy write/read overlap)"},
      {1 << 20, 200, "1 M lines (128 MB: spills L2)"},
      {1 << 26, 4, "64 M lines (8 GB: DRAM streaming)"}};
  int bad = 0;
  for (const Case& c : cases) {
    double* lines = nullptr;
    if (cudaMalloc(&lines, (size_t)c.lines * 128) != cudaSuccess) { printf("skip %s: alloc failed\n", c.what); continue; }
    for (int wide = 0; wide < 2; wide++) {
      cudaMemset(lines, 0, (size_t)c.lines * 128);
      ctr[0] = ctr[1] = ctr[2] = 0;
      if (wide) litmus<true><<<sm * 4, 256>>>(lines, c.lines, c.rounds, ctr, ctr + 1, ctr + 2);
      else litmus<false><<<sm * 4, 256>>>(lines, c.lines, c.rounds, ctr, ctr + 1, ctr + 2);
      cudaError_t e = cudaDeviceSynchronize();
      printf("%-52s %s: loads %llu, saw a writer's value %llu, torn %llu%s\n", c.what, wide ? "4 x 32 B" : "8 x 16 B", ctr[1], ctr[2], ctr[0],
             e == cudaSuccess ? "" : cudaGetErrorString(e));
      if (ctr[0]) bad = 1;
    }
    cudaFree(lines);
  }
  printf(bad ? "TORN LINES OBSERVED\n" : "no torn line observed\n");
  return bad;
}

// latency micro-benchmark: dependent chains of DMMA, DFMA, sqrt, div, log, LDS->DFMA on one warp (clock64)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double a, double b) {
  __shared__ double sm[64];
  if (threadIdx.x < 64) sm[threadIdx.x] = 1.0 + threadIdx.x * 1e-9;
  __syncthreads();
  double c0 = threadIdx.x, c1 = -1.0;
  long long t0 = clock64();
  for (int i = 0; i < 1024; ++i)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
  long long t1 = clock64();
  double f = c0;
  for (int i = 0; i < 1024; ++i) f = fma(f, a, b);
  long long t2 = clock64();
  double s = fabs(f) + 2.0;
  for (int i = 0; i < 256; ++i) s = sqrt(s) + 1.5;
  long long t3 = clock64();
  double d = s;
  for (int i = 0; i < 256; ++i) d = 1.0 / d + 0.5;
  long long t4 = clock64();
  double l = d + 2.0;
  for (int i = 0; i < 256; ++i) l = log(l) + 3.0;
  long long t5 = clock64();
  double m = l;
  for (int i = 0; i < 256; ++i) m = fma(m, 1e-9, sm[(i + (int)m) & 63]);
  long long t6 = clock64();
  double r = m;
  for (int i = 0; i < 256; ++i) r = rsqrt(r) + 1.5;
  long long t7 = clock64();
  // two interleaved independent DMMA chains
  double e0 = 1, e1 = 2, g0 = 3, g1 = 4;
  for (int i = 0; i < 1024; ++i) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(e0), "+d"(e1) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(g0), "+d"(g1) : "d"(a), "d"(b));
  }
  long long t8 = clock64();
  if (threadIdx.x == 0) {
    cyc[0] = (t1 - t0); cyc[1] = (t2 - t1); cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; cyc[5] = t6 - t5; cyc[6] = t7 - t6; cyc[7] = t8 - t7;
  }
  out[threadIdx.x] = c0 + c1 + f + s + d + l + m + r + e0 + e1 + g0 + g1;
}
int main() {
  double* o; long long* c; cudaMalloc(&o, 8 * 32); cudaMalloc(&c, 64);
  k<<<1, 32>>>(o, c, 1.0000001, 1e-9); k<<<1, 32>>>(o, c, 1.0000001, 1e-9);
  long long h[8]; cudaMemcpy(h, c, 64, cudaMemcpyDeviceToHost);
  printf("{\"dmma_dep_cyc\": %.1f, \"dfma_dep_cyc\": %.1f, \"sqrt_add_cyc\": %.1f, \"div_add_cyc\": %.1f, \"log_add_cyc\": %.1f, \"lds_fma_cyc\": %.1f, \"rsqrt_add_cyc\": %.1f, \"dmma_2chain_pair_cyc\": %.1f}\n",
         h[0] / 1024.0, h[1] / 1024.0, h[2] / 256.0, h[3] / 256.0, h[4] / 256.0, h[5] / 256.0, h[6] / 256.0, h[7] / 1024.0);
  return 0;
}

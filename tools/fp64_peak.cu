// FP64 peak micro-benchmark for B200 (sm_100a): DFMA (vector) and DMMA (mma.sync f64) issue rates.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
// Output: one JSON object on stdout. Used as the FP64 roofline denominator (SURVEY.md section 6:
// "FP64 peak is not in MEASURED_PEAKS.json - it must be measured on the box").
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; } } while (0)

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == 12345.678) out[0] = s;
}

// m8n8k4: A 1 reg, B 1 reg, C 2 regs per thread; 256 FMA per warp instruction.
template <int ILP>
__global__ void dmma884_kernel(double* out, int iters, double a, double b) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c0[i] = i; c1[i] = -i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
  if (s == 12345.678) out[0] = s;
}

// m16n8k8: A 4 regs, B 2 regs, C 4 regs; 1024 FMA per warp instruction.
template <int ILP>
__global__ void dmma1688_kernel(double* out, int iters, double a, double b) {
  double c[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 12345.678) out[0] = s;
}

// m16n8k16: A 8 regs, B 4 regs, C 4 regs; 2048 FMA per warp instruction.
template <int ILP>
__global__ void dmma16816_kernel(double* out, int iters, double a, double b) {
  double c[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a), "d"(b), "d"(a));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 12345.678) out[0] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); launch(); cudaDeviceSynchronize();
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, 8));
  const int iters = 4096;
  printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
  for (int warps = 4; warps <= 32; warps *= 2) {
    int threads = warps * 32, blocks = sms * 2;
    {
      constexpr int ILP = 8;
      double ms = time_ms([&] { dfma_kernel<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
      double fl = 2.0 * blocks * threads * (double)iters * ILP;
      printf(", \"dfma_tflops_w%d\": %.3f", warps * 2, fl / ms * 1e-9);
    }
    {
      constexpr int ILP = 8;
      double ms = time_ms([&] { dmma884_kernel<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
      double fl = 2.0 * 256 * blocks * warps * (double)iters * ILP;
      printf(", \"dmma_m8n8k4_tflops_w%d\": %.3f", warps * 2, fl / ms * 1e-9);
    }
    {
      constexpr int ILP = 4;
      double ms = time_ms([&] { dmma1688_kernel<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
      double fl = 2.0 * 1024 * blocks * warps * (double)iters * ILP;
      printf(", \"dmma_m16n8k8_tflops_w%d\": %.3f", warps * 2, fl / ms * 1e-9);
    }
    {
      constexpr int ILP = 4;
      double ms = time_ms([&] { dmma16816_kernel<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
      double fl = 2.0 * 2048 * blocks * warps * (double)iters * ILP;
      printf(", \"dmma_m16n8k16_tflops_w%d\": %.3f", warps * 2, fl / ms * 1e-9);
    }
  }
  CK(cudaGetLastError());
  printf("}\n");
  return 0;
}

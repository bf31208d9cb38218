// What does one FP64 tensor-core instruction (mma.sync.aligned.m8n8k4.f64) compute, bit for bit?
// Compares D = A(8x4) B(4x8) + C against candidate scalar formulas on random data with a wide exponent spread:
//   (1) ascending-k chain of fma:  d = fma(a3,b3, fma(a2,b2, fma(a1,b1, fma(a0,b0,c))))
//   (2) descending-k chain, (3) pairwise, (4) products rounded then added.
// Prints the fraction of bit-identical outputs per candidate.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_semantics tools/dmma_semantics.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

__global__ void k(const double* A, const double* B, const double* C, double* D, int n) {
  // one warp per problem; lane = 4*r + q: A[r][q], B[q][n=r]... fragments: a = A[row=lane>>2][k=lane&3], b = B[k=lane&3][col=lane>>2]
  const int p = blockIdx.x, lane = threadIdx.x;
  const double* a = A + p * 32; const double* b = B + p * 32; const double* c = C + p * 64;
  const int g = lane >> 2, t = lane & 3;
  double av = a[g * 4 + t];          // A[g][t]
  double bv = b[t * 8 + g];          // B[t][g]
  double d0 = c[g * 8 + 2 * t], d1 = c[g * 8 + 2 * t + 1];
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(av), "d"(bv));
  D[p * 64 + g * 8 + 2 * t] = d0; D[p * 64 + g * 8 + 2 * t + 1] = d1;
}

int main() {
  const int n = 20000;
  std::vector<double> A(n * 32), B(n * 32), C(n * 64), D(n * 64);
  srand(7);
  auto rnd = [&](int spread) { double m = (double)rand() / RAND_MAX * 2 - 1; int e = rand() % (2 * spread + 1) - spread; return ldexp(m, e); };
  for (auto& v : A) v = rnd(12);
  for (auto& v : B) v = rnd(12);
  for (auto& v : C) v = rnd(20);
  double *dA, *dB, *dC, *dD;
  cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dB, B.size() * 8); cudaMalloc(&dC, C.size() * 8); cudaMalloc(&dD, D.size() * 8);
  cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dC, C.data(), C.size() * 8, cudaMemcpyHostToDevice);
  k<<<n, 32>>>(dA, dB, dC, dD, n);
  cudaMemcpy(D.data(), dD, D.size() * 8, cudaMemcpyDeviceToHost);
  long m1 = 0, m2 = 0, m3 = 0, m4 = 0, tot = 0;
  for (int p = 0; p < n; ++p)
    for (int r = 0; r < 8; ++r)
      for (int c = 0; c < 8; ++c) {
        const double* a = &A[p * 32 + r * 4]; const double* b = &B[p * 32 + c];
        const double cc = C[p * 64 + r * 8 + c], d = D[p * 64 + r * 8 + c];
        double s1 = cc; for (int q = 0; q < 4; ++q) s1 = fma(a[q], b[q * 8], s1);
        double s2 = cc; for (int q = 3; q >= 0; --q) s2 = fma(a[q], b[q * 8], s2);
        double s3 = fma(a[1], b[8], a[0] * b[0]) + fma(a[3], b[24], a[2] * b[16]) + cc;
        double s4 = cc; for (int q = 0; q < 4; ++q) s4 = s4 + a[q] * b[q * 8];
        m1 += (s1 == d); m2 += (s2 == d); m3 += (s3 == d); m4 += (s4 == d); ++tot;
      }
  printf("DMMA m8n8k4 vs scalar candidates over %ld outputs: ascending fma chain %.6f, descending %.6f, pairwise %.6f, mul+add %.6f\n",
         tot, (double)m1 / tot, (double)m2 / tot, (double)m3 / tot, (double)m4 / tot);
  return 0;
}

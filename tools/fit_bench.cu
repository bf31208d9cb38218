// Standalone micro-benchmark of gp_fit_kernel (phase A of the batched GP): worker/slot shapes side by side, and - with
// -DCNGP_FIT_TIMING - the cycles block 0's warps spend in each phase of a tile column.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo [-DCNGP_FIT_TIMING] tools/fit_bench.cu -o tools/fit_bench
//   tools/fit_bench [windows=4096] [N=256] [lag_table=1]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../corenav_gp_b200/csrc/gp_fit.cuh"

using namespace cngp;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int NW, int T>
static void run(const char* name, FitArgs fa, int B, double* d_lml, long long* d_dbg) {
  const size_t smem = fit_smem_bytes(fa.nt, NW * T);
  auto kfn = gp_fit_kernel<KID_RBF_PER, NW, T>;
  if (NW * T < fa.nt) { printf("%-9s skipped (nt > NW*T)\n", name); return; }
  cudaFuncAttributes at;
  CK(cudaFuncGetAttributes(&at, kfn));
  if (smem + at.sharedSizeBytes > 232448) { printf("%-9s skipped (smem %zu + %zu)\n", name, smem, at.sharedSizeBytes); return; }
  CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#ifdef CNGP_FIT_TIMING
  CK(cudaMemset(d_dbg, 0, 64 * 8 * sizeof(long long)));
  fa.dbg = d_dbg;
#endif
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  kfn<<<B, (NW + 1) * 32, smem>>>(fa);
  CK(cudaDeviceSynchronize());
  float best = 1e9f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    kfn<<<B, (NW + 1) * 32, smem>>>(fa);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    best = fminf(best, ms);
  }
  std::vector<double> lml(B);
  CK(cudaMemcpy(lml.data(), d_lml, B * sizeof(double), cudaMemcpyDeviceToHost));
  double cs = 0; for (double v : lml) cs += v;
  printf("%-9s regs %3d  %.3f ms   lml[0] %.12f  sum %.9f\n", name, at.numRegs, best, lml[0], cs);
#ifdef CNGP_FIT_TIMING
  std::vector<long long> dbg(64 * 8);
  CK(cudaMemcpy(dbg.data(), d_dbg, dbg.size() * sizeof(long long), cudaMemcpyDeviceToHost));
  const int reps = 6, nt = fa.nt;
  printf("   cycles per column (block 0):  worker: pass | eval | barrier 1 | (d) | barrier 2 | (a)      diag: chol8 | store + barrier 2 | update   (a blocked barrier shows in the phase after it)\n");
  for (int w = 0; w <= NW; ++w) {
    printf("   warp %2d:", w);
    for (int k = 0; k < 6; ++k) printf(" %7.0f", (double)dbg[w * 8 + k] / reps / nt);
    printf("\n");
  }
#endif
}

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 4096, N = argc > 2 ? atoi(argv[2]) : 256;
  const int nt = (N + 7) / 8;
  std::vector<double> x((size_t)B * N), y((size_t)B * N);
  unsigned long long s = 88172645463325252ull;
  for (int b = 0; b < B; ++b)
    for (int j = 0; j < N; ++j) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      const double u = (double)(s >> 11) / 9007199254740992.0 - 0.5;
      x[(size_t)b * N + j] = 20.0 + j;
      y[(size_t)b * N + j] = 0.02 + 0.05 * sin(2 * M_PI * j / 37.0) + 0.03 * j / N + 0.1 * u;
    }
  const double th[6] = {0.01, 10.0, 0.0025, 37.0, 1.0, 1e-3};
  double *dx, *dy, *dth, *dL, *dz, *dfeat, *dlml; int* dst; long long* ddbg;
  CK(cudaMalloc(&dx, x.size() * 8)); CK(cudaMalloc(&dy, y.size() * 8)); CK(cudaMalloc(&dth, sizeof th));
  CK(cudaMalloc(&dL, (size_t)B * tiles_in_lower(nt) * 512)); CK(cudaMalloc(&dz, (size_t)B * nt * 64));
  CK(cudaMalloc(&dfeat, (size_t)B * 4 * nt * 64)); CK(cudaMalloc(&dlml, B * 8)); CK(cudaMalloc(&dst, B * 4));
  CK(cudaMalloc(&ddbg, 64 * 8 * sizeof(long long)));
  CK(cudaMemcpy(dx, x.data(), x.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dy, y.data(), y.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dth, th, sizeof th, cudaMemcpyHostToDevice));
  FitArgs fa;
  memset(&fa, 0, sizeof fa);
  fa.kp.n_params = 5; fa.kp.fast_id = KID_RBF_PER;
  fa.theta = dth; fa.theta_stride = 0; fa.theta_mode = 0; fa.x = dx; fa.y = dy; fa.N = N; fa.nt = nt; fa.n_windows = B;
  fa.L = dL; fa.z = dz; fa.feat = dfeat; fa.lml = dlml; fa.status = dst;
  fa.lag_ok = argc > 3 ? atoi(argv[3]) : 1;      // K(X,X) from the integer-lag table (stamps are 20 + j)
  fa.kp.n_terms = 2; fa.kp.n_leaves = 2; fa.kp.term_start[0] = 0; fa.kp.term_start[1] = 1; fa.kp.term_start[2] = 2;
  fa.kp.leaf_type[0] = CNGP_K_RBF; fa.kp.leaf_type[1] = CNGP_K_STDPERIODIC; fa.kp.leaf_param[0] = 0; fa.kp.leaf_param[1] = 2;
  printf("gp_fit_kernel<rbf+stdperiodic>  windows %d  N %d\n", B, N);
  run<11, 3>("11x3", fa, B, dlml, ddbg);
  return 0;
}

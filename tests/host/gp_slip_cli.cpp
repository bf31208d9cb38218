// Test driver for gp_slip_predict (include/gp_predictor_b200.hpp): reads one GP_Input window, runs the callback and
// writes mean / sigma.   usage: gp_slip_cli <in.bin: n, time[n], slip[n] as doubles> <out.bin: m, mean[m], sigma[m]> [fit]
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/gp_predictor_b200.hpp"

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 3;
  double nd = 0;
  if (fread(&nd, sizeof(double), 1, f) != 1) return 4;
  const int n = (int)nd;
  core_nav::GP_Input in;
  in.header.seq = 7;
  in.time_array.resize(n);
  in.slip_array.resize(n);
  if (fread(in.time_array.data(), sizeof(double), n, f) != (size_t)n) return 4;
  if (fread(in.slip_array.data(), sizeof(double), n, f) != (size_t)n) return 4;
  fclose(f);
  cngp_ctx* ctx = nullptr;
  if (cngp_create(nullptr, &ctx) != CNGP_OK) { fprintf(stderr, "cngp_create failed: %s\n", cngp_last_error(nullptr)); return 5; }
  const bool fit = argc > 3 && !strcmp(argv[3], "fit");
  const double theta[4] = {0.01, 10.0, 0.05, 1e-3};       // rbf*brownian: variance, lengthscale, brownian variance, noise
  try {
    const core_nav::GP_Output out = gp_slip_predict(ctx, in, "rbf*brownian", fit ? nullptr : theta);
    FILE* o = fopen(argv[2], "wb");
    const double md = (double)out.mean.size();
    fwrite(&md, sizeof(double), 1, o);
    fwrite(out.mean.data(), sizeof(double), out.mean.size(), o);
    fwrite(out.sigma.data(), sizeof(double), out.sigma.size(), o);
    fclose(o);
    printf("{\"seq\": %u, \"m\": %zu}\n", out.header.seq, out.mean.size());
  } catch (const std::exception& e) {
    fprintf(stderr, "%s\n", e.what());
    cngp_destroy(ctx);
    return 6;
  }
  cngp_destroy(ctx);
  return 0;
}

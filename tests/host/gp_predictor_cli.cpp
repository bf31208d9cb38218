// Test driver for the ROS-free GpPredictor (include/gp_predictor_b200.hpp): plays one GP_Output message and one
// SetStopping response read from a binary file through GpPredictor::GPCallBack and prints what the node would have
// published.  Linked against the real libcngp.so for the GPU test and against tests/host/fake_cngp.c (the C oracle
// behind the same ABI) for the CPU test of the host logic.
//   usage: gp_predictor_cli <input.bin> <seconds the clock advances during the callback> <service_ok 0|1>
//   input.bin: doubles  P[225] Q[225] STM[225] Hvec[60] pos[3] M mean[M] sigma[M]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/gp_predictor_b200.hpp"

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 3;
  std::vector<double> head(225 * 3 + 60 + 3 + 1);
  if (fread(head.data(), sizeof(double), head.size(), f) != head.size()) return 4;
  const int M = (int)head.back();
  core_nav::GP_Output msg;
  msg.mean.resize(M);
  msg.sigma.resize(M);
  if (fread(msg.mean.data(), sizeof(double), M, f) != (size_t)M) return 4;
  if (fread(msg.sigma.data(), sizeof(double), M, f) != (size_t)M) return 4;
  fclose(f);
  const double advance = atof(argv[2]);
  const bool service_ok = atoi(argv[3]) != 0;

  cngp_ctx* ctx = nullptr;
  if (cngp_create(nullptr, &ctx) != CNGP_OK) {
    fprintf(stderr, "cngp_create failed: %s\n", cngp_last_error(nullptr));
    return 5;
  }
  int clock_calls = 0, published = 0;
  double stop_cmd = -1.0;
  GpPredictor node(
      ctx,
      [&](core_nav::SetStopping& srv) {
        if (!service_ok) return false;
        for (int k = 0; k < 225; ++k) {
          srv.response.PvecData[k] = head[k];
          srv.response.QvecData[k] = head[225 + k];
          srv.response.STMvecData[k] = head[450 + k];
        }
        for (int k = 0; k < 60; ++k) srv.response.HvecData[k] = head[675 + k];
        srv.response.PosData.x = head[735];
        srv.response.PosData.y = head[736];
        srv.response.PosData.z = head[737];
        return srv.request.stopping;
      },
      [&](const std_msgs::Float64& m) { ++published; stop_cmd = m.data; },
      [&]() { return clock_calls++ == 0 ? 100.0 : 100.0 + advance; });
  const bool ret = node.GPCallBack(msg);
  const GpPredictor::Vector3 enu = node.llh_to_enu(node.savePos[0], node.savePos[1], node.savePos[2] + 1.0);
  printf("{\"returned\": %d, \"published\": %d, \"stop_cmd\": %.17g, \"i\": %d, \"slip_i\": %d, \"xy_errSlip\": %.17g, "
         "\"flag\": %d, \"enu_up\": [%.17g, %.17g, %.17g]",
         (int)ret, published, stop_cmd, node.i, node.slip_i, node.xy_errSlip, (int)node.new_gp_data_arrived_, enu[0], enu[1],
         enu[2]);
  // the public matrix members of gp_predictor.h:36-46 after the callback
  auto dump = [](const char* name, const double* v, int n) {
    printf(", \"%s\": [", name);
    for (int k = 0; k < n; ++k) printf("%s%.17g", k ? ", " : "", v[k]);
    printf("]");
  };
  dump("P_pred", node.P_pred.data(), 225);
  dump("K_pred", node.K_pred.data(), 60);
  dump("R_IP", node.R_IP.data(), 16);
  dump("R_IP_1", node.R_IP_1.data(), 16);
  dump("R_IP_2", node.R_IP_2.data(), 16);
  dump("ins_enu_slip", node.ins_enu_slip.data(), 3);
  dump("ins_enu_slip3p", node.ins_enu_slip3p.data(), 3);
  dump("ins_enu_slip_3p", node.ins_enu_slip_3p.data(), 3);
  printf("}\n");
  cngp_destroy(ctx);
  return 0;
}

// Test driver for include/cngp_wire.hpp: reads a serialised message, decodes it into the C++ message struct, changes
// nothing, encodes it again and writes the bytes back - tests/test_wire.py compares them with corenav_gp_b200/wire.py.
//   wire_cli <gp_input|gp_output|set_stopping_response|set_stopping_request|float64> <in> <out>
// For gp_input the decoded window is also pushed through core_nav::GP_Input -> (time, slip) -> summary on stdout.
#include <cstdio>
#include <fstream>
#include <iterator>
#include <string>

#include "../../include/cngp_wire.hpp"

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  const std::string kind = argv[1];
  std::ifstream in(argv[2], std::ios::binary);
  cngp_wire::Bytes data((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  cngp_wire::Bytes out;
  try {
    if (kind == "gp_input") {
      core_nav::GP_Input m;
      cngp_wire::deserialize(data, m);
      std::printf("{\"seq\": %u, \"stamp\": %.9f, \"frame_id\": \"%s\", \"n\": %zu, \"m\": %zu}\n", m.header.seq, m.header.stamp,
                  m.header.frame_id.c_str(), m.time_array.size(), m.slip_array.size());
      out = cngp_wire::serialize(m);
    } else if (kind == "gp_output") {
      core_nav::GP_Output m;
      cngp_wire::deserialize(data, m);
      out = cngp_wire::serialize(m);
    } else if (kind == "set_stopping_response") {
      core_nav::SetStopping::Response m;
      cngp_wire::deserialize(data, m);
      std::printf("{\"P00\": %.17g, \"H59\": %.17g, \"z\": %.17g}\n", m.PvecData[0], m.HvecData[59], m.PosData.z);
      out = cngp_wire::serialize(m);
    } else if (kind == "set_stopping_request") {
      core_nav::SetStopping::Request m;
      cngp_wire::deserialize(data, m);
      out = cngp_wire::serialize(m);
    } else if (kind == "float64") {
      std_msgs::Float64 m;
      cngp_wire::deserialize(data, m);
      out = cngp_wire::serialize(m);
    } else if (kind == "framed_gp_output") {   // a TCPROS stream: two frames back to back, re-emit them
      size_t at = 0;
      while (at < data.size()) {
        cngp_wire::Bytes body;
        const size_t used = cngp_wire::unframe(data.data() + at, data.size() - at, body);
        if (!used) return 3;
        core_nav::GP_Output m;
        cngp_wire::deserialize(body, m);
        const cngp_wire::Bytes f = cngp_wire::frame(cngp_wire::serialize(m));
        out.insert(out.end(), f.begin(), f.end());
        at += used;
      }
    } else {
      return 2;
    }
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 4;
  }
  std::ofstream o(argv[3], std::ios::binary);
  o.write(reinterpret_cast<const char*>(out.data()), (std::streamsize)out.size());
  return 0;
}

// cpp_host_demo.cpp — the C++11 host class without ReaK: the CRS-A465-style 6-DOF arm is assembled with
// reak_b200::chain_builder exactly as examples/robot_airship/old/CRS_A465_models.cpp:348-640 assembles it
// (actuator, rotor inertia_gen, revolute_joint_3D, rigid_link_3D, inertia_3D per joint; axes z,-y,-y,z,-y,z),
// evaluated and integrated on GPU 0 through kte_batch_propagator, and the results are printed so that
// tests/test_gpu_parity.py can compare them with the same chain built through the Python mirror.
//
//   g++ -std=c++11 -O2 -Iinclude examples/cpp_host_demo.cpp -Lreak_b200/lib -lreak_b200 -o cpp_host_demo
#include <cstdio>
#include <vector>

#include "reak_b200/kte_batch_propagator.hpp"

int main() {
  using namespace reak_b200;
  const double axes[6][3] = {{0, 0, 1}, {0, -1, 0}, {0, -1, 0}, {0, 0, 1}, {0, -1, 0}, {0, 0, 1}};
  const double link_z[6] = {0.3302, 0.3048, 0.1500, 0.1802, 0.0762, 0.0};
  chain_builder b(3);
  int cur = b.add_frame();
  rkb_base_frame base = rkb_base_frame();
  base.position[1] = -3.3; base.position[2] = 0.3;
  base.quat[0] = 0.70710678118654757; base.quat[3] = 0.70710678118654746;  // rot_z(pi/2) as axis_angle gives it
  base.acceleration[2] = 9.81;
  b.set_base(cur, base);
  std::uint64_t upstream = 0;
  for (int k = 0; k < 6; ++k) {
    const int coord = b.add_coord(), end = b.add_frame(), nxt = b.add_frame();
    const int act = b.driving_actuator_gen(coord, -1, k);
    b.inertia_gen(coord, 1.0);
    const int joint = b.revolute_joint_3D(coord, axes[k][0], axes[k][1], axes[k][2], cur, end);
    b.set_actuator_joint(act, joint);
    const double p[3] = {0.0, 0.0, link_z[k]}, q[4] = {1.0, 0.0, 0.0, 0.0};
    b.rigid_link_3D(end, nxt, p, q);
    upstream |= std::uint64_t(1) << coord;
    const double tensor[6] = {1.0, 0.0, 0.0, 1.0, 0.0, 1.0};
    b.inertia_3D(nxt, 1.0, tensor, upstream);
    cur = nxt;
  }
  try {
    kte_batch_propagator prop(b, 0, 1e-3);
    if (!prop.is_serial() || prop.get_state_dimensions() != 12 || prop.get_input_dimensions() != 6) return 2;
    const std::size_t N = 64;
    std::vector<double> x(N * 12), u(N * 6), xd(N * 12), xo(N * 12), M(N * 36);
    unsigned long long s = 88172645463325252ull;  // xorshift64: the test regenerates the same inputs
    for (std::size_t i = 0; i < x.size() + u.size(); ++i) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      const double r = (double)(s >> 11) / 9007199254740992.0 * 2.0 - 1.0;
      if (i < x.size()) x[i] = r; else u[i - x.size()] = r;
    }
    std::vector<int32_t> st(N);
    prop.get_state_derivatives(N, &x[0], &u[0], &xd[0], &st[0]);
    prop.get_mass_matrices(N, &x[0], &M[0]);
    prop.get_next_states(N, &x[0], &u[0], 25, 1e-3, &xo[0], &st[0]);
    for (std::size_t i = 0; i < N; ++i) if (st[i]) return 3;
    // single-sample concept call must agree with the batched one
    std::vector<double> p0(x.begin(), x.begin() + 12), u0(u.begin(), u.begin() + 6);
    std::vector<double> d0 = prop.get_state_derivative(0, p0, u0, 0.0);
    for (int k = 0; k < 12; ++k) if (d0[k] != xd[k]) return 4;
    bool threw = false;
    try { p0.pop_back(); prop.get_state_derivative(0, p0, u0, 0.0); } catch (std::range_error&) { threw = true; }
    if (!threw) return 5;
    for (std::size_t i = 0; i < N * 12; ++i) std::printf("%.17g %.17g\n", xd[i], xo[i]);
    for (std::size_t i = 0; i < 36; ++i) std::printf("M %.17g\n", M[i]);
  } catch (std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  return 0;
}

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
  int joint_end[6];
  for (int k = 0; k < 6; ++k) {
    const int coord = b.add_coord(), end = b.add_frame(), nxt = b.add_frame();
    joint_end[k] = end;
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
    // the arm's proximity model against the lab (CRS_A465_geom_model.cpp:88-148, build_MD148_lab.cpp:103-149):
    // findMinimumDistance at the propagated states, and the planner's is_free on one of them
    const double h = 0.70710678118654757;  // axis_angle(pi/2, x).getQuaternion() = (cos pi/4, sin pi/4, 0, 0)
    struct local {
      static rkb_shape shape(int kind, int anchor, double px, double py, double pz, double qw, double qx, double qy, double qz,
                             double d0, double d1, double d2) {
        rkb_shape s = rkb_shape();
        s.kind = kind; s.anchor = anchor;
        s.position[0] = px; s.position[1] = py; s.position[2] = pz;
        s.quat[0] = qw; s.quat[1] = qx; s.quat[2] = qy; s.quat[3] = qz;
        s.dims[0] = d0; s.dims[1] = d1; s.dims[2] = d2;
        return s;
      }
    };
    std::vector<rkb_shape> robot, lab;
    robot.push_back(local::shape(RKB_SHAPE_CCYLINDER, joint_end[0], 0, 0, 0.3302, h, h, 0, 0, 0.34, 0.09, 0));
    robot.push_back(local::shape(RKB_SHAPE_CCYLINDER, joint_end[1], 0, 0, 0.15, 1, 0, 0, 0, 0.3, 0.07, 0));
    robot.push_back(local::shape(RKB_SHAPE_CCYLINDER, joint_end[2], 0, 0, 0.165, 1, 0, 0, 0, 0.33, 0.07, 0));
    robot.push_back(local::shape(RKB_SHAPE_CCYLINDER, joint_end[4], 0, 0, 0.0381, 1, 0, 0, 0, 0.0762, 0.05, 0));
    robot.push_back(local::shape(RKB_SHAPE_SPHERE, joint_end[5], -0.04, 0, 0.05, 1, 0, 0, 0, 0.11, 0, 0));
    lab.push_back(local::shape(RKB_SHAPE_PLANE, -1, -0.8, -1.0, 0.0, 1, 0, 0, 0, 4.0, 6.0, 0));
    lab.push_back(local::shape(RKB_SHAPE_PLANE, -1, 1.2, -1.0, 1.5, h, 0, -h, 0, 3.0, 6.0, 0));
    lab.push_back(local::shape(RKB_SHAPE_PLANE, -1, -0.8, 2.0, 1.5, h, h, 0, 0, 4.0, 3.0, 0));
    lab.push_back(local::shape(RKB_SHAPE_CCYLINDER, -1, 0.1, -1.71, 0.15, h, h, 0, 0, 3.42, 0.18, 0));
    lab.push_back(local::shape(RKB_SHAPE_CCYLINDER, -1, -0.1, -1.71, 0.15, h, h, 0, 0, 3.42, 0.18, 0));
    rkb_proxy* pair = prop.make_proxy_pair(robot, lab);
    if (rkb_proxy_finder_count(pair) != 25) return 6;
    std::vector<double> dist(N);
    std::vector<int32_t> finder(N);
    for (std::size_t i = 0; i < x.size(); ++i) x[i] *= 3.0;
    prop.get_min_distances(pair, N, &x[0], &dist[0], &finder[0]);
    std::vector<const rkb_proxy*> pairs(1, pair);
    std::vector<double> p1(x.begin(), x.begin() + 12);
    if (prop.is_free(pairs, p1) != !(dist[0] < 0.0)) return 7;
    for (std::size_t i = 0; i < N; ++i) std::printf("D %.17g %d\n", dist[i], (int)finder[i]);
    rkb_proxy_destroy(pair);
  } catch (std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  return 0;
}

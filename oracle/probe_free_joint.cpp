// probe_free_joint.cpp — TEST INFRASTRUCTURE (a recorded experiment, not part of any checker).
// Does the reference evaluate an arm on a free base (BASELINE config 4's "free" variant)?
// free_joint_3D + platform inertia_3D + one revolute joint + link + inertia_3D, through
// kte_nl_system::get_state_derivative of the UNMODIFIED sources:
//   ./probe_free_joint 0   no rotor inertia on the joint      -> evaluates (15 state derivatives printed)
//   ./probe_free_joint 1   joint carries an inertia_gen rotor -> SEGFAULT in mass_matrix_calc::get_TMT_TdMT:
//                          the 3D-frame x gen-inertia branch tests mUpStreamJoints.find(mCoords[i]) and then
//                          dereferences mUpStream3DJoints[mFrames3D[i]] (mass_matrix_calculator.cpp:226-233),
//                          which operator[] has just created as a null pointer.
// Every joint of the CRS A465 chains carries such a rotor (CRS_A465_models.cpp:304-640), so the reference has no
// behaviour to match for a CRS arm on a floating base; the prismatic-track variant (:304-347) is what is built.
//   make -C oracle probe-free-joint
#include <ReaK/core/base/defs.hpp>
#include <ReaK/core/kinetostatics/kinetostatics.hpp>
#include <ReaK/core/kinetostatics/motion_jacobians.hpp>
#include <ReaK/ctrl/mbd_kte/kte_map_chain.hpp>
#include <ReaK/ctrl/mbd_kte/revolute_joint.hpp>
#include <ReaK/ctrl/mbd_kte/free_joints.hpp>
#include <ReaK/ctrl/mbd_kte/rigid_link.hpp>
#include <ReaK/ctrl/mbd_kte/inertia.hpp>
#include <ReaK/ctrl/mbd_kte/jacobian_joint_map.hpp>
#include <ReaK/ctrl/mbd_kte/mass_matrix_calculator.hpp>
#include <ReaK/ctrl/ctrl_sys/kte_nl_system.hpp>
#include <cstdio>
#include <cstdlib>
using namespace ReaK;
int main(int argc, char** argv) {
  const bool with_rotor = argc > 1 && atoi(argv[1]) != 0;
  typedef frame_3D<double> F;
  shared_ptr<F> base(new F()), fcoord(new F()), f1(new F()), f2(new F()), f3(new F());
  base->Acceleration = vect<double,3>(0, 0, 9.81);
  shared_ptr<jacobian_3D_3D<double> > J0(new jacobian_3D_3D<double>());
  shared_ptr<gen_coord<double> > q1(new gen_coord<double>());
  shared_ptr<jacobian_gen_3D<double> > J1(new jacobian_gen_3D<double>());
  shared_ptr<kte::free_joint_3D> fj(new kte::free_joint_3D("free", fcoord, base, f1, J0));
  shared_ptr<kte::revolute_joint_3D> rj(new kte::revolute_joint_3D("rev", q1, vect<double,3>(0, 1, 0), f1, f2, J1));
  shared_ptr<kte::rigid_link_3D> lk(new kte::rigid_link_3D("link", f2, f3, pose_3D<double>(weak_ptr<pose_3D<double> >(), vect<double,3>(0, 0, 0.3), quaternion<double>())));
  shared_ptr<kte::joint_dependent_frame_3D> dep(new kte::joint_dependent_frame_3D(f3));
  dep->add_joint(q1, J1);
  dep->add_joint(fcoord, J0);
  shared_ptr<kte::inertia_3D> in3(new kte::inertia_3D("body", dep, 2.0, mat<double,mat_structure::symmetric>(mat<double,mat_structure::identity>(3))));
  shared_ptr<kte::joint_dependent_frame_3D> dep0(new kte::joint_dependent_frame_3D(f1));
  dep0->add_joint(fcoord, J0);
  shared_ptr<kte::inertia_3D> in0(new kte::inertia_3D("platform", dep0, 5.0, mat<double,mat_structure::symmetric>(mat<double,mat_structure::identity>(3))));
  shared_ptr<kte::kte_map_chain> chain(new kte::kte_map_chain("chain"));
  shared_ptr<kte::mass_matrix_calc> mc(new kte::mass_matrix_calc("mc"));
  shared_ptr<kte::inertia_gen> rotor;
  if (with_rotor) {
    shared_ptr<kte::joint_dependent_gen_coord> depg(new kte::joint_dependent_gen_coord(q1));
    depg->add_joint(q1, shared_ptr<jacobian_gen_gen<double> >(new jacobian_gen_gen<double>(1.0, 0.0)));
    rotor = shared_ptr<kte::inertia_gen>(new kte::inertia_gen("rotor", depg, 1.0));
  }
  *chain << fj << in0;
  if (rotor) *chain << rotor;
  *chain << rj << lk << in3;
  *mc << in0 << in3;
  if (rotor) *mc << rotor;
  *mc << q1 << fcoord;
  ctrl::kte_nl_system sys("sys");
  sys.chain = chain; sys.mass_calc = mc;
  sys.dofs_gen.push_back(q1); sys.dofs_3D.push_back(fcoord);
  vect_n<double> x(2 + 13, 0.0), u(0);
  x[0] = 0.3; x[1] = 0.1; x[2] = 0.1; x[3] = 0.2; x[4] = 0.3; x[5] = 1.0; x[6] = 0.1; x[7] = -0.2; x[8] = 0.05;
  x[9] = 0.3; x[10] = -0.1; x[11] = 0.2; x[12] = 0.1; x[13] = 0.2; x[14] = -0.3;
  std::printf("evaluating (rotor inertia_gen: %d)...\n", (int)with_rotor); std::fflush(stdout);
  vect_n<double> xd = sys.get_state_derivative(sys, x, u, 0.0);
  for (std::size_t i = 0; i < xd.size(); ++i) std::printf("%.17g ", xd[i]);
  std::printf("\n");
  return 0;
}

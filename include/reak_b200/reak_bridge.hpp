// reak_bridge.hpp — the reference-side binding: turns LIVE ReaK objects into the flat chain
// descriptor of reak_b200.h and wraps the GPU propagator in a class that models ReaK's
// state-space-system concepts with ReaK's own types.  Needs the ReaK headers on the include path
// (this is the one header of the repo that does); it uses public accessors only, so it compiles
// against an unmodified ReaK tree.
//
//   reak_b200::compile_kte_system(sys)   walks kte_nl_system::chain->getKTEs()
//                                        (ctrl/mbd_kte/kte_map_chain.hpp:59) the way
//                                        ctrl/mbd_kte/kte_chain_visitation.hpp:39-84 does, reading
//                                        joints, links, inertias (+ their mUpStreamJoints maps,
//                                        jacobian_joint_map.hpp:76-331), springs, dampers and
//                                        driving actuators through their accessors.
//   ReaK::pp::kte_steer_space            a steerable space (steer_position_toward of SteerableSpaceConcept,
//                                        ctrl/topologies/steerable_space_concept.hpp:60-85) whose steering runs on
//                                        the batched propagator, single pairs and whole planner iterations.
//   ReaK::ctrl::kte_batch_system         SSSystemConcept + DiscreteSSSConcept on vect_n<double>,
//                                        drop-in where kte_nl_system / num_int_dtnl_sys are used
//                                        (e.g. runge_kutta4_integrate_impl, sys_integrators/
//                                        runge_kutta4_integrator_sys.hpp:50-97), plus the batched calls.
#ifndef REAK_B200_REAK_BRIDGE_HPP
#define REAK_B200_REAK_BRIDGE_HPP

#include <cmath>
#include <limits>
#include <iterator>
#include <map>
#include <random>
#include <stdexcept>
#include <utility>
#include <vector>

#include <ReaK/core/lin_alg/vect_alg.hpp>
#include <ReaK/core/lin_alg/mat_num_exceptions.hpp>
#include <ReaK/core/kinetostatics/kinetostatics.hpp>
#include <ReaK/ctrl/mbd_kte/kte_map_chain.hpp>
#include <ReaK/ctrl/mbd_kte/revolute_joint.hpp>
#include <ReaK/ctrl/mbd_kte/prismatic_joint.hpp>
#include <ReaK/ctrl/mbd_kte/free_joints.hpp>
#include <ReaK/ctrl/mbd_kte/rigid_link.hpp>
#include <ReaK/ctrl/mbd_kte/inertia.hpp>
#include <ReaK/ctrl/mbd_kte/spring.hpp>
#include <ReaK/ctrl/mbd_kte/damper.hpp>
#include <ReaK/ctrl/mbd_kte/torsion_spring.hpp>
#include <ReaK/ctrl/mbd_kte/torsion_damper.hpp>
#include <ReaK/ctrl/mbd_kte/driving_actuator.hpp>
#include <ReaK/ctrl/mbd_kte/mass_matrix_calculator.hpp>
#include <ReaK/ctrl/ctrl_sys/kte_nl_system.hpp>
#include <ReaK/geometry/shapes/plane.hpp>
#include <ReaK/geometry/shapes/sphere.hpp>
#include <ReaK/geometry/shapes/capped_cylinder.hpp>
#include <ReaK/geometry/shapes/cylinder.hpp>
#include <ReaK/geometry/shapes/box.hpp>
#include <ReaK/geometry/shapes/circle.hpp>
#include <ReaK/geometry/shapes/capped_rectangle.hpp>
#include <ReaK/geometry/shapes/rectangle.hpp>
#include <ReaK/geometry/proximity/proxy_query_model.hpp>
#include <ReaK/ctrl/topologies/metric_space_concept.hpp>
#include <ReaK/ctrl/topologies/subspace_concept.hpp>
#include <ReaK/ctrl/topologies/random_sampler_concept.hpp>
#include <ReaK/ctrl/topologies/default_random_sampler.hpp>
#include <ReaK/ctrl/topologies/steerable_space_concept.hpp>

#include "kte_batch_propagator.hpp"

namespace reak_b200 {

class unsupported_chain : public std::runtime_error {
 public:
  explicit unsupported_chain(const std::string& what) : std::runtime_error("reak_b200: " + what) {}
};

namespace detail {

struct id_map {
  std::map<const void*, int> ids;
  int lookup(const void* p) const {
    std::map<const void*, int>::const_iterator it = ids.find(p);
    return it == ids.end() ? -1 : it->second;
  }
};

template <typename JointMap>
std::uint64_t upstream_mask(const JointMap& m, const id_map& coords) {
  std::uint64_t mask = 0;
  for (typename JointMap::const_iterator it = m.begin(); it != m.end(); ++it) {
    int c = coords.lookup(it->first.get());
    if (c < 0) throw unsupported_chain("an inertia depends on a coordinate that is not a system dof");
    mask |= std::uint64_t(1) << c;
  }
  return mask;
}

}  // namespace detail

/// Flatten (chain, mass_calc, dofs_gen, inputs) — the public members of kte_nl_system
/// (kte_nl_system.hpp:70-78) — into a chain_builder.  State slot j <-> dofs_gen[j]
/// (kte_nl_system.hpp:189-193); input k <-> inputs[k] (kte_nl_system.hpp:221-224).
inline chain_builder compile_kte_system(const ReaK::ctrl::kte_nl_system& sys) {
  using namespace ReaK;
  using namespace ReaK::kte;
  using ReaK::rtti::rk_dynamic_ptr_cast;
  if (!sys.chain || !sys.mass_calc) throw unsupported_chain("kte_nl_system without chain or mass_calc");
  if (!sys.dofs_2D.empty() && !sys.dofs_3D.empty()) throw unsupported_chain("2D and 3D free-frame dofs in one system");
  if (sys.mass_calc->Coords().size() != sys.dofs_gen.size()) throw unsupported_chain("mass_matrix_calc coordinates differ from the system dofs");
  if (sys.mass_calc->Frames3D().size() != sys.dofs_3D.size() || sys.mass_calc->Frames2D().size() != sys.dofs_2D.size())
    throw unsupported_chain("mass_matrix_calc free frames differ from the system dofs");
  detail::id_map coords, inputs, frames, elems, free3;  // free3: the coordinate frames of the free joints, 2D or 3D
  for (std::size_t i = 0; i < sys.dofs_3D.size(); ++i) {  // state block 2 n + 13 i <-> dofs_3D[i] (kte_nl_system.hpp:205-219)
    if (sys.mass_calc->Frames3D()[i] != sys.dofs_3D[i]) throw unsupported_chain("mass_matrix_calc free frames differ from the system dofs");
    free3.ids[sys.dofs_3D[i].get()] = static_cast<int>(i);
  }
  for (std::size_t i = 0; i < sys.dofs_2D.size(); ++i) {  // state block 2 n + 7 i <-> dofs_2D[i] (kte_nl_system.hpp:194-204)
    if (sys.mass_calc->Frames2D()[i] != sys.dofs_2D[i]) throw unsupported_chain("mass_matrix_calc free frames differ from the system dofs");
    free3.ids[sys.dofs_2D[i].get()] = static_cast<int>(i);
  }
  for (std::size_t i = 0; i < sys.dofs_gen.size(); ++i) {
    if (sys.mass_calc->Coords()[i] != sys.dofs_gen[i]) throw unsupported_chain("mass_matrix_calc coordinates differ from the system dofs");
    coords.ids[sys.dofs_gen[i].get()] = static_cast<int>(i);
  }
  int n_inputs = 0;
  for (std::size_t i = 0; i < sys.inputs.size(); ++i) {
    if (sys.inputs[i]->getInputCount() != 1) throw unsupported_chain("only single-input driving_actuator_gen inputs are compiled");
    inputs.ids[sys.inputs[i].get()] = n_inputs++;
  }
  const std::vector<shared_ptr<kte_map> >& ktes = sys.chain->getKTEs();
  int dim = 0;
  for (std::size_t e = 0; e < ktes.size() && !dim; ++e) {
    if (rk_dynamic_ptr_cast<revolute_joint_3D>(ktes[e]) || rk_dynamic_ptr_cast<prismatic_joint_3D>(ktes[e]) ||
        rk_dynamic_ptr_cast<rigid_link_3D>(ktes[e]) || rk_dynamic_ptr_cast<inertia_3D>(ktes[e]) ||
        rk_dynamic_ptr_cast<free_joint_3D>(ktes[e])) dim = 3;
    else if (rk_dynamic_ptr_cast<revolute_joint_2D>(ktes[e]) || rk_dynamic_ptr_cast<prismatic_joint_2D>(ktes[e]) ||
             rk_dynamic_ptr_cast<rigid_link_2D>(ktes[e]) || rk_dynamic_ptr_cast<inertia_2D>(ktes[e]) ||
             rk_dynamic_ptr_cast<free_joint_2D>(ktes[e])) dim = 2;
  }
  if (!dim) throw unsupported_chain("chain has no 2D or 3D element");
  chain_builder b(dim);
  for (std::size_t i = 0; i < sys.dofs_gen.size(); ++i) b.add_coord();
  std::map<int, int> written;  // frame id -> 1 when some element's doMotion writes it
  struct local {
    static int fid(chain_builder& bb, detail::id_map& fr, const void* p) {
      if (!p) throw unsupported_chain("element with a null frame");
      int id = fr.lookup(p);
      if (id < 0) { id = bb.add_frame(); fr.ids[p] = id; }
      return id;
    }
    static int cid(const detail::id_map& co, const void* p) {
      int id = co.lookup(p);
      if (id < 0) throw unsupported_chain("element refers to a coordinate that is not a system dof");
      return id;
    }
  };
  // a generalized coordinate of a _gen element: a system dof, or an auxiliary gen_coord declared (with the values it
  // holds) right before the first element that uses it
  detail::id_map aux;
  struct gen {
    static int id(chain_builder& bb, const detail::id_map& co, detail::id_map& ax, const shared_ptr<gen_coord<double> >& c) {
      if (!c) throw unsupported_chain("_gen element with a null coordinate");
      int i = co.lookup(c.get());
      if (i >= 0) return i;
      i = ax.lookup(c.get());
      if (i < 0) { i = bb.add_aux_coord(c->q, c->q_dot, c->q_ddot); ax.ids[c.get()] = i; }
      return i;
    }
  };
  std::vector<std::pair<int, const void*> > pending_actuators;  // (element index, joint object)
  const void* base3 = NULL;
  std::map<int, const void*> frame_obj;
  for (std::size_t e = 0; e < ktes.size(); ++e) {
    const shared_ptr<kte_map>& k = ktes[e];
    int idx = -1;
    if (shared_ptr<revolute_joint_3D> j = rk_dynamic_ptr_cast<revolute_joint_3D>(k)) {
      int fa = local::fid(b, frames, j->BaseFrame().get()), fb = local::fid(b, frames, j->EndFrame().get());
      frame_obj[fa] = j->BaseFrame().get(); written[fb] = 1;
      vect<double, 3> a = j->Axis();
      idx = b.revolute_joint_3D(local::cid(coords, j->Angle().get()), a[0], a[1], a[2], fa, fb);
    } else if (shared_ptr<prismatic_joint_3D> j = rk_dynamic_ptr_cast<prismatic_joint_3D>(k)) {
      int fa = local::fid(b, frames, j->BaseFrame().get()), fb = local::fid(b, frames, j->EndFrame().get());
      frame_obj[fa] = j->BaseFrame().get(); written[fb] = 1;
      vect<double, 3> a = j->Axis();
      idx = b.prismatic_joint_3D(local::cid(coords, j->Coord().get()), a[0], a[1], a[2], fa, fb);
    } else if (shared_ptr<free_joint_3D> j = rk_dynamic_ptr_cast<free_joint_3D>(k)) {
      int fa = local::fid(b, frames, j->BaseFrame().get()), fb = local::fid(b, frames, j->EndFrame().get());
      frame_obj[fa] = j->BaseFrame().get(); written[fb] = 1;
      const int fi = free3.lookup(j->Coord().get());
      if (fi < 0) throw unsupported_chain("a free joint's coordinate frame is not listed in the system's dofs_3D");
      idx = b.free_joint_3D(fi, fa, fb);
    } else if (shared_ptr<free_joint_2D> j = rk_dynamic_ptr_cast<free_joint_2D>(k)) {
      int fa = local::fid(b, frames, j->BaseFrame().get()), fb = local::fid(b, frames, j->EndFrame().get());
      frame_obj[fa] = j->BaseFrame().get(); written[fb] = 1;
      const int fi = free3.lookup(j->Coord().get());
      if (fi < 0) throw unsupported_chain("a free joint's coordinate frame is not listed in the system's dofs_2D");
      idx = b.free_joint_2D(fi, fa, fb);
    } else if (shared_ptr<revolute_joint_2D> j = rk_dynamic_ptr_cast<revolute_joint_2D>(k)) {
      int fa = local::fid(b, frames, j->BaseFrame().get()), fb = local::fid(b, frames, j->EndFrame().get());
      frame_obj[fa] = j->BaseFrame().get(); written[fb] = 1;
      idx = b.revolute_joint_2D(local::cid(coords, j->Angle().get()), fa, fb);
    } else if (shared_ptr<prismatic_joint_2D> j = rk_dynamic_ptr_cast<prismatic_joint_2D>(k)) {
      int fa = local::fid(b, frames, j->BaseFrame().get()), fb = local::fid(b, frames, j->EndFrame().get());
      frame_obj[fa] = j->BaseFrame().get(); written[fb] = 1;
      vect<double, 2> a = j->Axis();
      idx = b.prismatic_joint_2D(local::cid(coords, j->Coord().get()), a[0], a[1], fa, fb);
    } else if (shared_ptr<rigid_link_3D> l = rk_dynamic_ptr_cast<rigid_link_3D>(k)) {
      int fa = local::fid(b, frames, l->BaseFrame().get()), fb = local::fid(b, frames, l->EndFrame().get());
      frame_obj[fa] = l->BaseFrame().get(); written[fb] = 1;
      pose_3D<double> o = l->PoseOffset();
      double p[3] = {o.Position[0], o.Position[1], o.Position[2]};
      double q[4] = {o.Quat[0], o.Quat[1], o.Quat[2], o.Quat[3]};
      idx = b.rigid_link_3D(fa, fb, p, q);
    } else if (shared_ptr<rigid_link_2D> l = rk_dynamic_ptr_cast<rigid_link_2D>(k)) {
      int fa = local::fid(b, frames, l->BaseFrame().get()), fb = local::fid(b, frames, l->EndFrame().get());
      frame_obj[fa] = l->BaseFrame().get(); written[fb] = 1;
      pose_2D<double> o = l->PoseOffset();
      idx = b.rigid_link_2D(fa, fb, o.Position[0], o.Position[1], o.Rotation.getAngle());
    } else if (shared_ptr<inertia_3D> in = rk_dynamic_ptr_cast<inertia_3D>(k)) {
      int f = local::fid(b, frames, in->CenterOfMass()->mFrame.get());
      frame_obj[f] = in->CenterOfMass()->mFrame.get();
      mat<double, mat_structure::symmetric> I = in->InertiaTensor();
      double t[6] = {I(0, 0), I(0, 1), I(0, 2), I(1, 1), I(1, 2), I(2, 2)};
      std::uint64_t up = detail::upstream_mask(in->CenterOfMass()->mUpStreamJoints, coords);
      if (!in->CenterOfMass()->mUpStream2DJoints.empty()) throw unsupported_chain("an inertia depends on a 2D free joint");
      up |= detail::upstream_mask(in->CenterOfMass()->mUpStream3DJoints, free3) << 32;  // free-joint frames: bits 32 and up
      idx = b.inertia_3D(f, in->Mass(), t, up);
    } else if (shared_ptr<inertia_2D> in = rk_dynamic_ptr_cast<inertia_2D>(k)) {
      int f = local::fid(b, frames, in->CenterOfMass()->mFrame.get());
      frame_obj[f] = in->CenterOfMass()->mFrame.get();
      std::uint64_t up = detail::upstream_mask(in->CenterOfMass()->mUpStreamJoints, coords);
      if (!in->CenterOfMass()->mUpStream3DJoints.empty()) throw unsupported_chain("a 2D inertia depends on a 3D free joint");
      up |= detail::upstream_mask(in->CenterOfMass()->mUpStream2DJoints, free3) << 32;  // free-joint frames: bits 32 and up
      idx = b.inertia_2D(f, in->Mass(), in->MomentOfInertia(), up);
    } else if (shared_ptr<inertia_gen> in = rk_dynamic_ptr_cast<inertia_gen>(k)) {
      int c = local::cid(coords, in->CenterOfMass()->mFrame.get());
      if (detail::upstream_mask(in->CenterOfMass()->mUpStreamJoints, coords) != (std::uint64_t(1) << c))
        throw unsupported_chain("inertia_gen must depend on its own coordinate only");
      idx = b.inertia_gen(c, in->Mass());
    } else if (shared_ptr<driving_actuator_gen> a = rk_dynamic_ptr_cast<driving_actuator_gen>(k)) {
      int in = inputs.lookup(static_cast<system_input*>(a.get()));
      if (in < 0) throw unsupported_chain("a driving actuator of the chain is not listed in the system inputs");
      idx = b.driving_actuator_gen(local::cid(coords, a->Frame().get()), -1, in);
      pending_actuators.push_back(std::make_pair(idx, static_cast<const void*>(a->Joint().get())));
    } else if (shared_ptr<torsion_spring_3D> s = rk_dynamic_ptr_cast<torsion_spring_3D>(k)) {
      idx = b.torsion_spring(local::fid(b, frames, s->Anchor1().get()), local::fid(b, frames, s->Anchor2().get()), s->Stiffness(), s->Saturation());
    } else if (shared_ptr<torsion_spring_2D> s = rk_dynamic_ptr_cast<torsion_spring_2D>(k)) {
      idx = b.torsion_spring(local::fid(b, frames, s->Anchor1().get()), local::fid(b, frames, s->Anchor2().get()), s->Stiffness(), s->Saturation());
    } else if (shared_ptr<torsion_damper_3D> s = rk_dynamic_ptr_cast<torsion_damper_3D>(k)) {
      idx = b.torsion_damper(local::fid(b, frames, s->Anchor1().get()), local::fid(b, frames, s->Anchor2().get()), s->Damping());
    } else if (shared_ptr<torsion_damper_2D> s = rk_dynamic_ptr_cast<torsion_damper_2D>(k)) {
      idx = b.torsion_damper(local::fid(b, frames, s->Anchor1().get()), local::fid(b, frames, s->Anchor2().get()), s->Damping());
    } else if (shared_ptr<rigid_link_gen> l = rk_dynamic_ptr_cast<rigid_link_gen>(k)) {
      const int ca = gen::id(b, coords, aux, l->BaseFrame()), cb = gen::id(b, coords, aux, l->EndFrame());
      if (coords.lookup(l->EndFrame().get()) >= 0) throw unsupported_chain("rigid_link_gen ends on a system dof");
      idx = b.rigid_link_gen(ca, cb, l->Offset());
    } else if (shared_ptr<spring_gen> s = rk_dynamic_ptr_cast<spring_gen>(k)) {
      const int ca = gen::id(b, coords, aux, s->Anchor1()), cb = gen::id(b, coords, aux, s->Anchor2());
      idx = b.spring_gen(ca, cb, s->RestLength(), s->Stiffness(), s->Saturation());
    } else if (shared_ptr<damper_gen> s = rk_dynamic_ptr_cast<damper_gen>(k)) {
      const int ca = gen::id(b, coords, aux, s->Anchor1()), cb = gen::id(b, coords, aux, s->Anchor2());
      idx = b.damper_gen(ca, cb, s->Damping());
    } else if (shared_ptr<spring_3D> s = rk_dynamic_ptr_cast<spring_3D>(k)) {
      idx = b.spring(local::fid(b, frames, s->Anchor1().get()), local::fid(b, frames, s->Anchor2().get()), s->RestLength(), s->Stiffness(), s->Saturation());
    } else if (shared_ptr<spring_2D> s = rk_dynamic_ptr_cast<spring_2D>(k)) {
      idx = b.spring(local::fid(b, frames, s->Anchor1().get()), local::fid(b, frames, s->Anchor2().get()), s->RestLength(), s->Stiffness(), s->Saturation());
    } else if (shared_ptr<damper_3D> s = rk_dynamic_ptr_cast<damper_3D>(k)) {
      idx = b.damper(local::fid(b, frames, s->Anchor1().get()), local::fid(b, frames, s->Anchor2().get()), s->Damping());
    } else if (shared_ptr<damper_2D> s = rk_dynamic_ptr_cast<damper_2D>(k)) {
      idx = b.damper(local::fid(b, frames, s->Anchor1().get()), local::fid(b, frames, s->Anchor2().get()), s->Damping());
    } else {
      throw unsupported_chain("KTE '" + k->getName() + "' is outside the compiled element set");
    }
    elems.ids[k.get()] = idx;
  }
  for (std::size_t i = 0; i < pending_actuators.size(); ++i) {
    // the joint is addressed through its reacting_kte_gen base: find the chain element that is the same object
    int j = -1;
    for (std::size_t e = 0; e < ktes.size(); ++e)
      if (static_cast<const void*>(rk_dynamic_ptr_cast<reacting_kte_gen>(ktes[e]).get()) == pending_actuators[i].second &&
          pending_actuators[i].second)
        j = elems.lookup(ktes[e].get());
    if (j < 0) throw unsupported_chain("an actuator drives a joint that is not in the chain");
    b.set_actuator_joint(pending_actuators[i].first, j);
  }
  // the root: the one frame no element writes
  int root = -1, n_roots = 0;
  for (std::map<const void*, int>::const_iterator it = frames.ids.begin(); it != frames.ids.end(); ++it)
    if (!written.count(it->second)) { root = it->second; base3 = it->first; ++n_roots; }
  if (n_roots != 1) throw unsupported_chain("the chain must have exactly one un-driven base frame");
  rkb_base_frame bf = rkb_base_frame();
  if (dim == 3) {
    const frame_3D<double>* B = static_cast<const frame_3D<double>*>(base3);
    for (int i = 0; i < 3; ++i) {
      bf.position[i] = B->Position[i]; bf.velocity[i] = B->Velocity[i]; bf.ang_velocity[i] = B->AngVelocity[i];
      bf.acceleration[i] = B->Acceleration[i]; bf.ang_acceleration[i] = B->AngAcceleration[i];
    }
    for (int i = 0; i < 4; ++i) bf.quat[i] = B->Quat[i];
  } else {
    const frame_2D<double>* B = static_cast<const frame_2D<double>*>(base3);
    for (int i = 0; i < 2; ++i) { bf.position[i] = B->Position[i]; bf.velocity[i] = B->Velocity[i]; bf.acceleration[i] = B->Acceleration[i]; }
    bf.quat[0] = B->Rotation.getAngle();
    bf.ang_velocity[0] = B->AngVelocity;
    bf.ang_acceleration[0] = B->AngAcceleration;
  }
  b.set_base(root, bf);
  b.frame_ids = frames.ids;
  return b;
}

/// Flatten a proximity model (geometry/proximity/proxy_query_model.hpp:153-182) into the shape list of
/// rkb_proxy_create.  Public accessors only: shape_3D::getAnchor() / getPose(), the shapes' getDimensions() /
/// getRadius() / getLength().  A shape must ride on a frame of the compiled chain (looked up in
/// builder.frame_ids) or have no anchor; a shape whose anchor is some other pose, or whose pose chain goes
/// deeper, is refused.
inline std::vector<rkb_shape> compile_proxy_model(const ReaK::geom::proxy_query_model_3D& mdl, const chain_builder& builder) {
  using namespace ReaK;
  using ReaK::rtti::rk_dynamic_ptr_cast;
  std::vector<rkb_shape> out;
  for (std::size_t i = 0; i < mdl.mShapeList.size(); ++i) {
    const shared_ptr<geom::shape_3D>& sh = mdl.mShapeList[i];
    if (!sh) continue;  // createProxFinderList skips null shapes (proxy_query_model.cpp:219-220)
    rkb_shape s = rkb_shape();
    if (shared_ptr<geom::plane> p = rk_dynamic_ptr_cast<geom::plane>(sh)) {
      s.kind = RKB_SHAPE_PLANE; s.dims[0] = p->getDimensions()[0]; s.dims[1] = p->getDimensions()[1];
    } else if (shared_ptr<geom::sphere> p = rk_dynamic_ptr_cast<geom::sphere>(sh)) {
      s.kind = RKB_SHAPE_SPHERE; s.dims[0] = p->getRadius();
    } else if (shared_ptr<geom::capped_cylinder> p = rk_dynamic_ptr_cast<geom::capped_cylinder>(sh)) {
      s.kind = RKB_SHAPE_CCYLINDER; s.dims[0] = p->getLength(); s.dims[1] = p->getRadius();
    } else if (shared_ptr<geom::cylinder> p = rk_dynamic_ptr_cast<geom::cylinder>(sh)) {
      s.kind = RKB_SHAPE_CYLINDER; s.dims[0] = p->getLength(); s.dims[1] = p->getRadius();
    } else if (shared_ptr<geom::box> p = rk_dynamic_ptr_cast<geom::box>(sh)) {
      s.kind = RKB_SHAPE_BOX; for (int k = 0; k < 3; ++k) s.dims[k] = p->getDimensions()[k];
    } else {
      throw unsupported_chain("proximity shape outside plane / sphere / capped_cylinder / cylinder / box");
    }
    const shared_ptr<pose_3D<double> >& anchor = sh->getAnchor();
    if (!anchor) s.anchor = -1;
    else {
      std::map<const void*, int>::const_iterator it = builder.frame_ids.find(static_cast<const void*>(anchor.get()));
      if (it == builder.frame_ids.end()) {
        // frame_3D derives from pose_3D: the same object may be reached through a different base sub-object
        for (it = builder.frame_ids.begin(); it != builder.frame_ids.end(); ++it)
          if (static_cast<const pose_3D<double>*>(static_cast<const frame_3D<double>*>(it->first)) == anchor.get()) break;
      }
      if (it == builder.frame_ids.end()) throw unsupported_chain("proximity shape anchored to a pose that is not a frame of the chain");
      s.anchor = it->second;
    }
    const pose_3D<double>& P = sh->getPose();
    for (int k = 0; k < 3; ++k) s.position[k] = P.Position[k];
    for (int k = 0; k < 4; ++k) s.quat[k] = P.Quat[k];
    out.push_back(s);
  }
  return out;
}

/// The planar counterpart (proxy_query_model_2D: circle / capped_rectangle / rectangle riding on frame_2D's of a planar
/// chain): rkb_shape carries position[0..1] and quat[0..1] = the (cos, sin) rot_mat_2D holds.
inline std::vector<rkb_shape> compile_proxy_model(const ReaK::geom::proxy_query_model_2D& mdl, const chain_builder& builder) {
  using namespace ReaK;
  using ReaK::rtti::rk_dynamic_ptr_cast;
  std::vector<rkb_shape> out;
  for (std::size_t i = 0; i < mdl.mShapeList.size(); ++i) {
    const shared_ptr<geom::shape_2D>& sh = mdl.mShapeList[i];
    if (!sh) continue;  // createProxFinderList skips null shapes (proxy_query_model.cpp:81-82)
    rkb_shape s = rkb_shape();
    if (shared_ptr<geom::circle> p = rk_dynamic_ptr_cast<geom::circle>(sh)) {
      s.kind = RKB_SHAPE_CIRCLE; s.dims[0] = p->getRadius();
    } else if (shared_ptr<geom::capped_rectangle> p = rk_dynamic_ptr_cast<geom::capped_rectangle>(sh)) {
      s.kind = RKB_SHAPE_CRECT; s.dims[0] = p->getDimensions()[0]; s.dims[1] = p->getDimensions()[1];
    } else if (shared_ptr<geom::rectangle> p = rk_dynamic_ptr_cast<geom::rectangle>(sh)) {
      s.kind = RKB_SHAPE_RECTANGLE; s.dims[0] = p->getDimensions()[0]; s.dims[1] = p->getDimensions()[1];
    } else {
      throw unsupported_chain("planar proximity shape outside circle / capped_rectangle / rectangle");
    }
    const shared_ptr<pose_2D<double> >& anchor = sh->getAnchor();
    if (!anchor) s.anchor = -1;
    else {
      std::map<const void*, int>::const_iterator it = builder.frame_ids.find(static_cast<const void*>(anchor.get()));
      if (it == builder.frame_ids.end()) {
        for (it = builder.frame_ids.begin(); it != builder.frame_ids.end(); ++it)
          if (static_cast<const pose_2D<double>*>(static_cast<const frame_2D<double>*>(it->first)) == anchor.get()) break;
      }
      if (it == builder.frame_ids.end()) throw unsupported_chain("proximity shape anchored to a pose that is not a frame of the chain");
      s.anchor = it->second;
    }
    const pose_2D<double>& P = sh->getPose();
    s.position[0] = P.Position[0]; s.position[1] = P.Position[1];
    s.quat[0] = P.Rotation(0, 0); s.quat[1] = P.Rotation(1, 0);  // rot_mat_2D: [[c, -s], [s, c]]
    out.push_back(s);
  }
  return out;
}

}  // namespace reak_b200

namespace ReaK {
namespace ctrl {

/// Batched, GPU-backed counterpart of kte_nl_system (+ num_int_dtnl_sys with RK4).  Models
/// SSSystemConcept and DiscreteSSSConcept on the same point/input types as kte_nl_system.
class kte_batch_system {
 public:
  typedef vect_n<double> point_type;
  typedef vect_n<double> point_difference_type;
  typedef vect_n<double> point_derivative_type;
  typedef double time_type;
  typedef double time_difference_type;
  typedef vect_n<double> input_type;
  typedef vect_n<double> output_type;
  typedef std::size_t size_type;
  BOOST_STATIC_CONSTANT(std::size_t, dimensions = 0);
  BOOST_STATIC_CONSTANT(std::size_t, input_dimensions = 0);
  BOOST_STATIC_CONSTANT(std::size_t, output_dimensions = 0);

  explicit kte_batch_system(const kte_nl_system& sys, int device = 0, double time_step = 1e-3)
      : mBuilder(reak_b200::compile_kte_system(sys)), mProp(mBuilder, device, time_step) {}

  size_type get_state_dimensions() const { return mProp.get_state_dimensions(); }
  size_type get_input_dimensions() const { return mProp.get_input_dimensions(); }
  size_type get_output_dimensions() const { return 0; }
  time_difference_type get_time_step() const { return mProp.get_time_step(); }

  template <typename StateSpaceType>
  point_derivative_type get_state_derivative(const StateSpaceType& space, const point_type& p, const input_type& u, const time_type& t = 0) const {
    return to_vect(translate_errors_derivative(space, p, u, t));
  }
  template <typename StateSpaceType>
  point_type get_next_state(const StateSpaceType& space, const point_type& p, const input_type& u, const time_type& t = 0) const {
    try {
      return to_vect(mProp.get_next_state(space, p, u, t));
    } catch (reak_b200::singularity_error&) { throw singularity_error("A"); }
  }
  template <typename StateSpaceType>
  output_type get_output(const StateSpaceType&, const point_type&, const input_type&, const time_type& = 0) const { return output_type(); }

  const reak_b200::kte_batch_propagator& batch() const { return mProp; }

  /// The batched counterpart of a proxy_query_pair_3D(aModel1, aModel2) whose shapes ride on this system's frames:
  /// batch().get_min_distances(pair, ...) is findMinimumDistance at every state, batch().is_free(...) the test of
  /// manip_dk_proxy_env_impl::is_free (ctrl/topologies/manip_free_workspace.hpp:77-99).  Owned by the caller
  /// (rkb_proxy_destroy).
  rkb_proxy* make_proxy_pair(const geom::proxy_query_model_3D& aModel1, const geom::proxy_query_model_3D& aModel2) const {
    return mProp.make_proxy_pair(reak_b200::compile_proxy_model(aModel1, mBuilder), reak_b200::compile_proxy_model(aModel2, mBuilder));
  }
  rkb_proxy* make_proxy_pair(const geom::proxy_query_model_2D& aModel1, const geom::proxy_query_model_2D& aModel2) const {
    return mProp.make_proxy_pair(reak_b200::compile_proxy_model(aModel1, mBuilder), reak_b200::compile_proxy_model(aModel2, mBuilder));
  }
  const reak_b200::chain_builder& descriptor() const { return mBuilder; }

 private:
  template <typename StateSpaceType>
  std::vector<double> translate_errors_derivative(const StateSpaceType& space, const point_type& p, const input_type& u, const time_type& t) const {
    try {
      return mProp.get_state_derivative(space, p, u, t);
    } catch (reak_b200::singularity_error&) { throw singularity_error("A"); }
  }
  static vect_n<double> to_vect(const std::vector<double>& v) {
    vect_n<double> r(v.size());
    for (std::size_t i = 0; i < v.size(); ++i) r[i] = v[i];
    return r;
  }
  reak_b200::chain_builder mBuilder;
  reak_b200::kte_batch_propagator mProp;
};

}  // namespace ctrl

namespace pp {

/// A steerable state space over a KTE chain whose steering is served by the batched GPU propagator
/// (SURVEY 8(f) rank 1).  It offers the valid expression of SteerableSpaceConcept
/// (ctrl/topologies/steerable_space_concept.hpp:60-85),
///     tie(p, st_rec) = space.steer_position_toward(p1, d, p3);
/// next to the metric-space members a planner uses on vect_n states (distance, difference, adjust,
/// origin), and a batched form that steers all candidate extensions of a planner iteration in one call.
///
/// Steering rule — the sampling one of kinodynamic RRT, with the rollouts of steer_with_constant_control
/// (examples/misc/MEAQR_topology.hpp:539-547: one RK4 control interval after the other, input held):
/// n_controls constant inputs are drawn uniformly from the input box, each is rolled out from p1 for
/// n_intervals intervals of steps_per_interval RK4 steps, and the rollout whose end state is nearest
/// (Euclidean) to the target p1 + d (p3 - p1) wins; its state after every interval is the steer record.
class kte_steer_space {
 public:
  typedef vect_n<double> point_type;
  typedef vect_n<double> point_difference_type;
  typedef std::vector<point_type> steer_record_type;
  // what the planners' traits read: metric_space_traits (metric_space_concept.hpp:188-191), point_distribution_traits
  // (random_sampler_concept.hpp:75-78), subspace_traits (subspace_concept.hpp:50-54).  The space is its own super-space:
  // planning_visitor measures edges with get(distance_metric, space.get_super_space()) (planning_visitors.hpp:259).
  typedef default_distance_metric distance_metric_type;
  typedef default_random_sampler random_sampler_type;
  typedef kte_steer_space super_space_type;
  BOOST_STATIC_CONSTANT(std::size_t, dimensions = 0);

  kte_steer_space(const ctrl::kte_nl_system& sys, const vect_n<double>& u_lower, const vect_n<double>& u_upper,
                  std::size_t n_controls, int n_intervals, int steps_per_interval, double dt, unsigned long long seed = 0,
                  int device = 0)
      : mSys(sys, device, dt), mLo(u_lower), mHi(u_upper), mControls(n_controls), mIntervals(n_intervals),
        mSteps(steps_per_interval), mDt(dt), mRng(seed) {
    if (mLo.size() != mSys.get_input_dimensions() || mHi.size() != mLo.size()) throw std::range_error("Input vector dimension mismatch!");
    if (n_controls < 1 || n_intervals < 1 || steps_per_interval < 1) throw std::range_error("steering needs at least one control, interval and step");
  }

  // ---- the metric-space side, on plain vectors ---------------------------------------------------
  double distance(const point_type& a, const point_type& b) const { return norm_2(difference(b, a)); }
  double norm(const point_difference_type& d) const { return norm_2(d); }
  point_difference_type difference(const point_type& a, const point_type& b) const { return a - b; }
  point_type adjust(const point_type& a, const point_difference_type& d) const { return a + d; }
  point_type origin() const { return point_type(mSys.get_state_dimensions(), 0.0); }
  point_type move_position_toward(const point_type& a, double fraction, const point_type& b) const { return a + fraction * (b - a); }

  // ---- what planning_visitor asks of its space besides steering ----------------------------------------
  const super_space_type& get_super_space() const { return *this; }  // SubSpaceConcept (subspace_concept.hpp:63-73)
  /// The box random_point() samples (and nothing else: steering may leave it, like the reference's dynamic spaces).
  void set_state_bounds(const point_type& lower, const point_type& upper) {
    if (lower.size() != mSys.get_state_dimensions() || upper.size() != lower.size()) throw std::range_error("State vector dimension mismatch!");
    mXLo = lower; mXHi = upper;
  }
  /// default_random_sampler (default_random_sampler.hpp:63-66) calls this: uniform in the state box.
  point_type random_point() const {
    const std::size_t nx = mSys.get_state_dimensions();
    if (mXLo.size() != nx) throw std::logic_error("kte_steer_space::random_point needs set_state_bounds first");
    std::uniform_real_distribution<double> uni(0.0, 1.0);
    point_type p(nx);
    for (std::size_t k = 0; k < nx; ++k) p[k] = mXLo[k] + (mXHi[k] - mXLo[k]) * uni(mRng);
    return p;
  }
  /// Collision environment: the proxy pairs manip_dk_proxy_env_impl keeps (manip_free_workspace.hpp:60-99), lowered once.
  void add_proxy_pair(const geom::proxy_query_model_3D& aModel1, const geom::proxy_query_model_3D& aModel2) {
    mPairs.push_back(mSys.make_proxy_pair(aModel1, aModel2));
  }
  /// is_free of the planners (planning_visitors.hpp:242; manip_free_workspace.hpp:77-99): no proxy pair reports a
  /// negative minimum distance at p (always free without pairs).  Served by rkb_is_free.
  bool is_free(const point_type& p) const {
    if (mPairs.empty()) return true;
    std::vector<point_type> one(1, p);
    std::vector<char> out;
    are_free(one, out);
    return out[0] != 0;
  }
  /// ... for all candidate points of a planner iteration in one launch
  void are_free(const std::vector<point_type>& pts, std::vector<char>& out) const {
    const std::size_t n = pts.size(), nx = mSys.get_state_dimensions();
    out.assign(n, 1);
    if (mPairs.empty() || n == 0) return;
    std::vector<double> x(n * nx);
    for (std::size_t i = 0; i < n; ++i) {
      if (pts[i].size() != nx) throw std::range_error("State vector dimension mismatch!");
      for (std::size_t k = 0; k < nx; ++k) x[i * nx + k] = pts[i][k];
    }
    std::vector<int32_t> fr(n);
    std::vector<const rkb_proxy*> pairs(mPairs.begin(), mPairs.end());
    mSys.batch().get_is_free(pairs, n, &x[0], &fr[0]);
    for (std::size_t i = 0; i < n; ++i) out[i] = fr[i] ? 1 : 0;
  }
  ~kte_steer_space() { for (std::size_t i = 0; i < mPairs.size(); ++i) rkb_proxy_destroy(mPairs[i]); }

  void reseed(unsigned long long seed) { mRng.seed(seed); }
  /// The inputs tried by the last steering call, [pairs][n_controls][n_inputs], and the winners' indices.
  const std::vector<double>& last_controls() const { return mU; }
  const std::vector<int32_t>& last_choice() const { return mBest; }
  const ctrl::kte_batch_system& system() const { return mSys; }

  /// SteerableSpaceConcept: steer a fraction d of the way from p1 towards p3.
  std::pair<point_type, steer_record_type> steer_position_toward(const point_type& p1, double d, const point_type& p3) const {
    std::vector<point_type> a(1, p1), b(1, p3), res;
    std::vector<steer_record_type> rec;
    steer_positions_toward(a, d, b, res, &rec);
    return std::make_pair(res[0], rec[0]);
  }

  /// All candidate extensions of one planner iteration: pair i steers from a[i] towards b[i].
  void steer_positions_toward(const std::vector<point_type>& a, double fraction, const std::vector<point_type>& b,
                              std::vector<point_type>& result, std::vector<steer_record_type>* records = NULL) const {
    const std::size_t P = a.size(), nx = mSys.get_state_dimensions(), nu = mSys.get_input_dimensions(), R = mControls;
    if (b.size() != P) throw std::range_error("as many goals as starts are needed");
    std::vector<double> x0(P * nx), goal(P * nx), bx(P * nx), cost(P);
    for (std::size_t i = 0; i < P; ++i) {
      if (a[i].size() != nx || b[i].size() != nx) throw std::range_error("State vector dimension mismatch!");
      for (std::size_t k = 0; k < nx; ++k) { x0[i * nx + k] = a[i][k]; goal[i * nx + k] = a[i][k] + fraction * (b[i][k] - a[i][k]); }
    }
    mU.assign(P * R * (nu ? nu : 1), 0.0);
    std::uniform_real_distribution<double> uni(0.0, 1.0);
    for (std::size_t i = 0; i < P * R; ++i)
      for (std::size_t k = 0; k < nu; ++k) mU[i * nu + k] = mLo[k] + (mHi[k] - mLo[k]) * uni(mRng);
    mBest.assign(P, 0);
    try {
      mSys.batch().steer_batch(P, R, &x0[0], &goal[0], &mU[0], mIntervals * mSteps, mDt, &mBest[0], &bx[0], &cost[0]);
      result.assign(P, point_type(nx));
      for (std::size_t i = 0; i < P; ++i)
        for (std::size_t k = 0; k < nx; ++k) result[i][k] = bx[i * nx + k];
      if (records) {
        // the winners once more, this time keeping the state after every control interval
        std::vector<double> useq(P * mIntervals * (nu ? nu : 1)), xe(P * nx), traj(P * mIntervals * nx);
        for (std::size_t i = 0; i < P; ++i)
          for (int j = 0; j < mIntervals; ++j)
            for (std::size_t k = 0; k < nu; ++k) useq[(i * mIntervals + j) * nu + k] = mU[(i * R + mBest[i]) * nu + k];
        mSys.batch().rollout(P, &x0[0], &useq[0], RKB_SCHEME_RK4, mIntervals, mSteps, mDt, &xe[0], &traj[0]);
        records->assign(P, steer_record_type());
        for (std::size_t i = 0; i < P; ++i) {
          (*records)[i].push_back(a[i]);  // the record starts at the start point (MEAQR_topology.hpp:607-608)
          for (int j = 0; j < mIntervals; ++j) {
            point_type w(nx);
            for (std::size_t k = 0; k < nx; ++k) w[k] = traj[(i * mIntervals + j) * nx + k];
            (*records)[i].push_back(w);
          }
        }
      }
    } catch (reak_b200::singularity_error&) { throw singularity_error("A"); }
  }

 private:
  ctrl::kte_batch_system mSys;
  vect_n<double> mLo, mHi;
  std::size_t mControls;
  int mIntervals, mSteps;
  double mDt;
  mutable std::mt19937_64 mRng;
  mutable std::vector<double> mU;
  mutable std::vector<int32_t> mBest;
  vect_n<double> mXLo, mXHi;
  std::vector<rkb_proxy*> mPairs;
  kte_steer_space(const kte_steer_space&);             // owns device handles: not copyable (planners hold spaces by shared_ptr)
  kte_steer_space& operator=(const kte_steer_space&);
};

// The planners dispatch on these traits: planning_visitor::dispatched_steer_towards_position takes its
// `space.steer_position_toward` branch only for is_steerable_space (planning_visitors.hpp:251-296; the primary
// template is false_, steerable_space_concept.hpp:95), and get(distance_metric, space) / get(random_sampler, space)
// exist only for is_metric_space / is_point_distribution (metric_space_concept.hpp:288-293,
// default_random_sampler.hpp:82-88).
template <> struct is_steerable_space<kte_steer_space> : boost::mpl::true_ {};
template <> struct is_metric_space<kte_steer_space> : boost::mpl::true_ {};
template <> struct is_point_distribution<kte_steer_space> : boost::mpl::true_ {};

/// One planner iteration's steering, batched.  A sampling planner pulls a new node by steering from each of the
/// nearest neighbours `Nc` of the sample towards it, one at a time, until one succeeds
/// (rrg_node_puller::expand_to_nearest, ctrl/graph_alg/node_generators.hpp:59-75, calling
/// planning_visitor::steer_towards_position, planning_visitors.hpp:349-360).  This visitor answers the same
/// `steer_towards_position(p, u, g)` question, but from a table filled by ONE batched steering call over all the
/// candidates (prepare), so the reference's own expand_to_nearest can run on top of it unchanged.  The answers are
/// those of the one-at-a-time visitor (steer_towards_position of a fresh visitor with the same seed) as long as
/// the candidates are asked in the order they were prepared — which is what expand_to_nearest does.
///   Graph: g[u].position is the vertex' point; Graph::edge_bundled has `steer_record` (mg_edge_data,
///   any_motion_graphs.hpp) and, for the optimal motion graphs, `weight`.
template <typename Graph>
class batched_steer_visitor {
 public:
  typedef kte_steer_space::point_type point_type;
  typedef typename Graph::vertex_descriptor Vertex;
  typedef typename Graph::edge_bundled EdgeProp;
  typedef boost::tuple<point_type, bool, EdgeProp> ResultType;

  batched_steer_visitor(const kte_steer_space& aSpace, double aProgressTolerance) : mSpace(&aSpace), mTol(aProgressTolerance) {}

  /// Steer from every vertex of Nc towards p in one call (fraction 1, like steer_towards_position) and, when the
  /// space has a collision environment, test all the end points in one more.
  void prepare(const point_type& p, const std::vector<Vertex>& Nc, const Graph& g) {
    std::vector<point_type> a, b(Nc.size(), p), res;
    for (std::size_t i = 0; i < Nc.size(); ++i) a.push_back(g[Nc[i]].position);
    std::vector<kte_steer_space::steer_record_type> rec;
    mSpace->steer_positions_toward(a, 1.0, b, res, &rec);
    std::vector<char> fr;
    mSpace->are_free(res, fr);
    mTable.clear();
    const kte_steer_space& super = mSpace->get_super_space();
    for (std::size_t i = 0; i < Nc.size(); ++i) {
      ResultType r;
      boost::get<0>(r) = res[i];
      boost::get<2>(r).steer_record = rec[i];
      // planning_visitors.hpp:353-357
      const double traveled = get(distance_metric, super)(a[i], res[i], super);
      const double best_case = get(distance_metric, super)(a[i], p, super);
      set_weight(boost::get<2>(r), traveled, 0);
      boost::get<1>(r) = (!std::isinf(traveled)) && (traveled < 2.0 * best_case) && (traveled > mTol * best_case) && fr[i] != 0;
      mTable.push_back(std::make_pair(Nc[i], r));
    }
    mTarget = p;
  }
  /// NodePullingVisitorConcept: the prepared answer for vertex u.
  ResultType steer_towards_position(const point_type& p, Vertex u, Graph&) const {
    for (std::size_t i = 0; i < mTable.size(); ++i)
      if (mTable[i].first == u) return mTable[i].second;
    throw std::logic_error("batched_steer_visitor: vertex was not prepared");
  }
  std::size_t prepared() const { return mTable.size(); }

 private:
  template <typename E> static auto set_weight(E& e, double w, int) -> decltype(e.weight = w, void()) { e.weight = w; }
  template <typename E> static void set_weight(E&, double, long) {}
  const kte_steer_space* mSpace;
  double mTol;
  point_type mTarget;
  std::vector<std::pair<Vertex, ResultType> > mTable;
};

/// Batched counterpart of linear_neighbor_search's iterator forms (ctrl/path_planning/topological_search.hpp:647-692)
/// and of what a dvp_tree answers (metric_space_search.hpp): the nearest candidates of [first, last) to EVERY point of a
/// batch — all the samples of one planner iteration — in one rkb_nearest call.  Points are vect_n-like (size(),
/// operator[]); the metric is the one the vect_n topologies use, norm_2(difference) (vect_alg.hpp:2314-2333), and the
/// distances are bit-identical to it.  Among candidates at equal distance the first of the range wins.
class batched_neighbor_search {
 public:
  explicit batched_neighbor_search(int device = 0) : mDevice(device) {}

  /// nearest candidate to each point; `last` for every point when the range is empty
  template <typename Point, typename ForwardIter, typename Topology, typename PositionMap>
  std::vector<ForwardIter> operator()(const std::vector<Point>& ps, ForwardIter first, ForwardIter last, const Topology&,
                                      PositionMap position) const {
    std::vector<ForwardIter> cand;
    std::vector<int32_t> idx;
    std::vector<double> dist;
    search(ps, first, last, position, 1, std::numeric_limits<double>::infinity(), cand, idx, dist);
    std::vector<ForwardIter> out(ps.size(), last);
    for (std::size_t i = 0; i < ps.size(); ++i) if (idx[i] >= 0) out[i] = cand[idx[i]];
    return out;
  }

  /// out[i]: up to max_neighbors candidates strictly closer than `radius` to ps[i], by ascending distance
  template <typename Point, typename ForwardIter, typename Topology, typename PositionMap>
  void operator()(const std::vector<Point>& ps, ForwardIter first, ForwardIter last,
                  std::vector<std::vector<typename std::iterator_traits<ForwardIter>::value_type> >& out, const Topology&,
                  PositionMap position, std::size_t max_neighbors = 1, double radius = std::numeric_limits<double>::infinity()) const {
    std::vector<ForwardIter> cand;
    std::vector<int32_t> idx;
    std::vector<double> dist;
    if (max_neighbors > RKB_NEAREST_MAX_K) throw std::range_error("batched_neighbor_search: at most RKB_NEAREST_MAX_K neighbours per query");
    search(ps, first, last, position, static_cast<int>(max_neighbors), radius, cand, idx, dist);
    out.assign(ps.size(), std::vector<typename std::iterator_traits<ForwardIter>::value_type>());
    for (std::size_t i = 0; i < ps.size(); ++i)
      for (std::size_t r = 0; r < max_neighbors && idx[i * max_neighbors + r] >= 0; ++r) out[i].push_back(*cand[idx[i * max_neighbors + r]]);
  }

 private:
  template <typename Point, typename ForwardIter, typename PositionMap>
  void search(const std::vector<Point>& ps, ForwardIter first, ForwardIter last, PositionMap position, int k, double radius,
              std::vector<ForwardIter>& cand, std::vector<int32_t>& idx, std::vector<double>& dist) const {
    idx.assign(ps.size() * k, -1);
    dist.assign(ps.size() * k, std::numeric_limits<double>::infinity());
    if (ps.empty()) return;
    const std::size_t dim = ps[0].size();
    std::vector<double> v, q(ps.size() * dim);
    for (; first != last; ++first) {
      cand.push_back(first);
      const Point& pt = get(position, *first);
      if (pt.size() != dim) throw std::range_error("Point dimension mismatch!");
      for (std::size_t c = 0; c < dim; ++c) v.push_back(pt[c]);
    }
    for (std::size_t i = 0; i < ps.size(); ++i) {
      if (ps[i].size() != dim) throw std::range_error("Point dimension mismatch!");
      for (std::size_t c = 0; c < dim; ++c) q[i * dim + c] = ps[i][c];
    }
    const int rc = rkb_nearest(mDevice, cand.size(), v.empty() ? NULL : &v[0], ps.size(), &q[0], static_cast<int>(dim), k, radius, &idx[0],
                               &dist[0], NULL, RKB_MEM_HOST, NULL);
    if (rc != RKB_OK) throw reak_b200::propagator_error(rc, "rkb_nearest");
  }
  int mDevice;
};

}  // namespace pp
}  // namespace ReaK

#endif

"""The chains BASELINE.json's configs name, written with the mirrored ReaK::kte API.

Geometry follows the reference's own model code: the CRS A465 preset of
examples/robot_airship/old/CRS_A465_models.cpp:260-822 (axes z,-y,-y,z,-y,z; link z-offsets
0.3302, 0.3048, 0.1500, 0.1802, 0.0762, 0; unit masses and inertias; gravity as an upward base
acceleration, :299) and the pendulum of ctrl/mbd_kte/test_bm.cpp:46-72.
"""
import math

from . import kte


class kte_system(object):
    """The public data members of ctrl::kte_nl_system (kte_nl_system.hpp:70-78)."""

    def __init__(self, name=""):
        self.name = name
        self.dofs_gen, self.inputs = [], []
        self.dofs_3D = []  # free-joint coordinate frames (kte_nl_system.hpp:72)
        self.chain = kte.kte_map_chain(name + "_chain")
        self.mass_calc = kte.mass_matrix_calc(name + "_mcalc")

    def get_state_dimensions(self):
        return 2 * len(self.dofs_gen) + 13 * len(self.dofs_3D)  # kte_nl_system.hpp:145-147

    def get_input_dimensions(self):
        return sum(a.getInputCount() for a in self.inputs)


CRS_AXES = [(0.0, 0.0, 1.0), (0.0, -1.0, 0.0), (0.0, -1.0, 0.0), (0.0, 0.0, 1.0), (0.0, -1.0, 0.0), (0.0, 0.0, 1.0)]
CRS_LINK_Z = [0.3302, 0.3048, 0.1500, 0.1802, 0.0762, 0.0]
PHYS_MASS = [5.0, 4.0, 3.0, 2.0, 1.0, 0.5]
PHYS_INERTIA = [(0.2, 0.15, 0.1), (0.12, 0.2, 0.08), (0.06, 0.09, 0.05), (0.03, 0.02, 0.04), (0.02, 0.015, 0.01), (0.01, 0.012, 0.011)]


def crs_chain(n_revolute=6, track=False, springs=False, physical=False, actuated=True,
              stiffness=10.0, damping=0.5, saturation=0.0, link_rotation=False, axes=None, link_offsets=None):
    """CRS-A465-style serial arm.  `track` prepends the prismatic x-axis joint (cfg 4);
    `springs` inserts a torsion_spring_3D + torsion_damper_3D across every revolute joint
    (cfg 3); `physical` uses graded masses / anisotropic inertias instead of the preset's
    unit placeholders; `link_rotation` gives the links a fixed twist (exercises R_o != I); `axes` /
    `link_offsets` replace the CRS joint axes / link offsets (other arms of the same build, e.g. the ERA
    and SSRMS geometries of ctrl/kte_models/manip_ERA_arm.cpp:101-227, manip_SSRMS_arm.cpp:105-252)."""
    axes = list(axes) if axes is not None else CRS_AXES
    link_offsets = list(link_offsets) if link_offsets is not None else [(0.0, 0.0, z) for z in CRS_LINK_Z]
    s = kte_system("crs")
    s.joint_end_frames, s.link_end_frames = [], []  # what a proximity model anchors its shapes to
    base = kte.frame_3D()
    base.Acceleration = [0.0, 0.0, 9.81]
    base.Position = [0.0, -3.3, 0.3]
    base.Quat = kte.axis_angle_quat(math.pi * 0.5, (0.0, 0.0, 1.0))
    cur = base
    upstream = []  # (coord, jacobian) of every joint met so far

    def stage(idx, joint_kind, axis, link_pos, mass, tensor, with_spring):
        nonlocal cur
        coord, jac, end, nxt = kte.gen_coord(), kte.jacobian_gen_3D(), kte.frame_3D(), kte.frame_3D()
        if joint_kind == "P":
            joint = kte.prismatic_joint_3D("joint_%d" % idx, coord, axis, cur, end, jac)
        else:
            joint = kte.revolute_joint_3D("joint_%d" % idx, coord, axis, cur, end, jac)
        dep = kte.joint_dependent_gen_coord(coord)
        dep.add_joint(coord, kte.jacobian_gen_gen(1.0, 0.0))
        rotor = kte.inertia_gen("joint_%d_inertia" % idx, dep, 1.0)
        if actuated:
            act = kte.driving_actuator_gen("joint_%d_actuator" % idx, coord, joint)
            s.chain << act
            s.inputs.append(act)
        s.chain << rotor << joint
        if with_spring:
            s.chain << kte.torsion_spring_3D("spring_%d" % idx, cur, end, stiffness, saturation)
            s.chain << kte.torsion_damper_3D("damper_%d" % idx, cur, end, damping)
        q_off = kte.axis_angle_quat(0.3 + 0.1 * idx, (1.0, 2.0, -1.0)) if link_rotation else (1.0, 0.0, 0.0, 0.0)
        link = kte.rigid_link_3D("link_%d" % idx, end, nxt, kte.pose_3D(link_pos, q_off))
        upstream.append((coord, jac))
        depf = kte.joint_dependent_frame_3D(nxt)
        for c, j in upstream:
            depf.add_joint(c, j)
        inertia = kte.inertia_3D("link_%d_inertia" % idx, depf, mass, tensor)
        s.chain << link << inertia
        s.dofs_gen.append(coord)
        s.mass_calc << inertia
        gen_inertias.append(rotor)
        s.joint_end_frames.append(end)
        s.link_end_frames.append(nxt)
        cur = nxt

    gen_inertias = []
    idx = 0
    if track:
        stage(idx, "P", (1.0, 0.0, 0.0), (0.0, 0.0, 0.0), 1.0, (1.0, 0.0, 0.0, 1.0, 0.0, 1.0), False)
        idx += 1
    for k in range(n_revolute):
        if physical:
            ix, iy, iz = PHYS_INERTIA[k % 6]
            mass, tensor = PHYS_MASS[k % 6], (ix, 0.01 * (k + 1), -0.005, iy, 0.002 * (k + 1), iz)
        else:
            mass, tensor = 1.0, (1.0, 0.0, 0.0, 1.0, 0.0, 1.0)
        stage(idx, "R", axes[k % len(axes)], link_offsets[k % len(link_offsets)], mass, tensor, springs)
        idx += 1
    for g in gen_inertias:
        s.mass_calc << g
    for c in s.dofs_gen:
        s.mass_calc << c
    return s


# joint axes and link offsets of the reference's ERA and SSRMS arm geometries (kinematic models there;
# used here with the CRS build of rotors, actuators and link inertias to exercise run-time specialisation)
ERA_AXES = [(0.0, 0.0, 1.0), (0.0, 1.0, 0.0), (1.0, 0.0, 0.0), (1.0, 0.0, 0.0), (1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0)]
ERA_LINKS = [(0.0, 0.0, L) for L in (0.35, 0.30, 4.1, 4.1, 0.30, 0.35, 0.2)]
SSRMS_AXES = [(0.0, 0.0, 1.0), (0.0, -1.0, 0.0), (0.0, 0.0, 1.0), (0.0, 0.0, 1.0), (0.0, 0.0, 1.0), (0.0, -1.0, 0.0), (0.0, 0.0, 1.0)]
SSRMS_LINKS = [(0.38, -0.635, 0.0), (0.504, 0.0, 0.38), (6.85, 0.0, 0.504), (6.85, 0.0, 0.504), (0.38, 0.0, 0.0), (0.504, -0.635, 0.0), (0.0, 0.0, 0.38)]


def planar_chain(lengths=(0.5, 0.4), masses=(1.0, 0.8), moments=(0.1, 0.05), actuated=False,
                 springs=False, stiffness=10.0, damping=0.5):
    """cfg 1: revolute_joint_2D -> rigid_link_2D -> inertia_2D per link; base accel (0, 9.81)."""
    s = kte_system("planar")
    base = kte.frame_2D()
    base.Acceleration = [0.0, 9.81]
    cur, upstream, acts = base, {}, []
    for k, (L, m, J) in enumerate(zip(lengths, masses, moments)):
        coord, jac, end, nxt = kte.gen_coord(), kte.jacobian_gen_2D(), kte.frame_2D(), kte.frame_2D()
        joint = kte.revolute_joint_2D("joint_%d" % k, coord, cur, end, jac)
        if actuated:
            act = kte.driving_actuator_gen("actuator_%d" % k, coord, joint)
            s.chain << act
            s.inputs.append(act)
        s.chain << joint
        if springs:
            s.chain << kte.torsion_spring_2D("spring_%d" % k, cur, end, stiffness)
            s.chain << kte.torsion_damper_2D("damper_%d" % k, cur, end, damping)
        link = kte.rigid_link_2D("link_%d" % k, end, nxt, kte.pose_2D((L, 0.0), 0.0))
        upstream[coord] = jac
        inertia = kte.inertia_2D("mass_%d" % k, kte.joint_dependent_frame_2D(nxt, upstream), m, J)
        s.chain << link << inertia
        s.dofs_gen.append(coord)
        s.mass_calc << inertia
        cur = nxt
    for c in s.dofs_gen:
        s.mass_calc << c
    return s


def free_base_chain(n_revolute=3, actuated=True, link_rotation=False):
    """A platform on a free_joint_3D (free_joints.cpp:119-208) carrying an arm of `n_revolute` revolute joints, the way
    the reference builds its free-floating platforms (ctrl/kte_models/free_floating_platform.cpp, examples/robot_airship):
    every inertia_3D depends on the free frame through the joint's jacobian_3D_3D and on the arm joints below it.  No
    inertia_gen: the reference's mass_matrix_calc crashes on a rotor next to a free joint (mass_matrix_calculator.cpp:226-233).
    n_revolute = 0: a single free rigid body."""
    s = kte_system("free_base")
    base = kte.frame_3D()
    base.Acceleration = [0.0, 0.0, 9.81]
    base.Position = [0.1, -0.2, 0.3]
    base.Quat = kte.axis_angle_quat(0.4, (1.0, 1.0, 0.0))
    coord, jac0, plat = kte.frame_3D(), kte.jacobian_3D_3D(), kte.frame_3D()
    s.chain << kte.free_joint_3D("free_base", coord, base, plat, jac0)
    dep0 = kte.joint_dependent_frame_3D(plat)
    dep0.add_joint(coord, jac0)
    body = kte.inertia_3D("platform", dep0, 5.0, (0.6, 0.02, -0.01, 0.5, 0.03, 0.4))
    s.chain << body
    s.mass_calc << body
    s.dofs_3D.append(coord)
    cur = plat
    upstream = []
    axes = [(0.0, 0.0, 1.0), (0.0, -1.0, 0.0), (1.0, 0.0, 0.0), (0.0, 1.0, 1.0)]
    for idx in range(n_revolute):
        q, jac, end, nxt = kte.gen_coord(), kte.jacobian_gen_3D(), kte.frame_3D(), kte.frame_3D()
        joint = kte.revolute_joint_3D("joint_%d" % idx, q, axes[idx % 4], cur, end, jac)
        if actuated:
            act = kte.driving_actuator_gen("joint_%d_actuator" % idx, q, joint)
            s.chain << act
            s.inputs.append(act)
        s.chain << joint
        q_off = kte.axis_angle_quat(0.2 + 0.1 * idx, (1.0, -1.0, 2.0)) if link_rotation else (1.0, 0.0, 0.0, 0.0)
        s.chain << kte.rigid_link_3D("link_%d" % idx, end, nxt, kte.pose_3D((0.05 * idx, 0.02, 0.25 + 0.05 * idx), q_off))
        upstream.append((q, jac))
        dep = kte.joint_dependent_frame_3D(nxt)
        dep.add_joint(coord, jac0)
        for c, j in upstream:
            dep.add_joint(c, j)
        inertia = kte.inertia_3D("link_%d_inertia" % idx, dep, 1.5 - 0.3 * idx, (0.05, 0.001 * idx, 0.0, 0.04, 0.002, 0.03))
        s.chain << inertia
        s.mass_calc << inertia
        s.dofs_gen.append(q)
        cur = nxt
    for q in s.dofs_gen:
        s.mass_calc << q
    s.mass_calc << coord
    return s


def free_planar_chain(n_revolute=2, actuated=True):
    """A planar platform on a free_joint_2D (free_joints.cpp:33-117) carrying `n_revolute` links: 7 states for the
    platform (position, (cos, sin), velocity, angular velocity) after the joints' (q, qd) pairs.  The system's free
    frames are listed in dofs_3D whatever their dimension (kte_nl_system keeps dofs_2D / dofs_3D apart; a chain has one kind)."""
    s = kte_system("free_planar")
    base = kte.frame_2D()
    base.Acceleration = [0.0, 9.81]
    base.Position = [0.2, -0.1]
    coord, jac0, plat = kte.frame_2D(), kte.jacobian_2D_2D(), kte.frame_2D()
    s.chain << kte.free_joint_2D("free_base", coord, base, plat, jac0)
    dep0 = kte.joint_dependent_frame_2D(plat)
    dep0.add_joint(coord, jac0)
    body = kte.inertia_2D("platform", dep0, 3.0, 0.4)
    s.chain << body
    s.mass_calc << body
    s.dofs_3D.append(coord)
    cur, upstream = plat, []
    for idx in range(n_revolute):
        q, jac, end, nxt = kte.gen_coord(), kte.jacobian_gen_2D(), kte.frame_2D(), kte.frame_2D()
        joint = kte.revolute_joint_2D("joint_%d" % idx, q, cur, end, jac)
        if actuated:
            act = kte.driving_actuator_gen("actuator_%d" % idx, q, joint)
            s.chain << act
            s.inputs.append(act)
        s.chain << joint
        s.chain << kte.rigid_link_2D("link_%d" % idx, end, nxt, kte.pose_2D((0.4 - 0.1 * idx, 0.05), 0.2 * idx))
        upstream.append((q, jac))
        dep = kte.joint_dependent_frame_2D(nxt)
        dep.add_joint(coord, jac0)
        for c, j in upstream:
            dep.add_joint(c, j)
        inertia = kte.inertia_2D("mass_%d" % idx, dep, 1.0 - 0.3 * idx, 0.05)
        s.chain << inertia
        s.mass_calc << inertia
        s.dofs_gen.append(q)
        cur = nxt
    for q in s.dofs_gen:
        s.mass_calc << q
    s.mass_calc << coord
    return s


def pendulum_chain():
    """ctrl/mbd_kte/test_bm.cpp:46-72: 0.5 m massless rod, 1 kg point mass, gravity (0, 9.81)."""
    return planar_chain(lengths=(0.5,), masses=(1.0,), moments=(0.0,))


def torsion_1dof_chain():
    """BASELINE.md known answer: one revolute_joint_3D about -y with torsion_spring_3D k=10 and
    torsion_damper_3D c=0.5, 2 kg at 0.3 m along z, I = diag(0.1,0.2,0.3); base at the origin."""
    s = kte_system("torsion1")
    base, end, tip = kte.frame_3D(), kte.frame_3D(), kte.frame_3D()
    base.Acceleration = [0.0, 0.0, 9.81]
    coord, jac = kte.gen_coord(), kte.jacobian_gen_3D()
    joint = kte.revolute_joint_3D("joint", coord, (0.0, -1.0, 0.0), base, end, jac)
    s.chain << joint
    s.chain << kte.torsion_spring_3D("spring", base, end, 10.0)
    s.chain << kte.torsion_damper_3D("damper", base, end, 0.5)
    s.chain << kte.rigid_link_3D("link", end, tip, kte.pose_3D((0.0, 0.0, 0.3)))
    dep = kte.joint_dependent_frame_3D(tip)
    dep.add_joint(coord, jac)
    inertia = kte.inertia_3D("mass", dep, 2.0, (0.1, 0.0, 0.0, 0.2, 0.0, 0.3))
    s.chain << inertia
    s.dofs_gen.append(coord)
    s.mass_calc << inertia << coord
    return s


def crs_linear_spring_chain():
    """6-DOF arm with two-anchor linear elements (spring.cpp:178-207, damper.cpp:136-149): a spring
    from the robot base to the end of link 3, a saturating spring between links 1 and 5 and a
    damper between the ends of links 2 and 6.  Exercises the interpreter kernels."""
    s = crs_chain(physical=True)
    ends = [k.mEnd for k in s.chain.getKTEs() if isinstance(k, kte.rigid_link_3D)]
    base = s.chain.getKTEs()[2].mBase  # base frame of joint_0 (chain order: actuator, rotor, joint, ...)
    s.chain << kte.spring_3D("spring_a", base, ends[2], 0.3, 50.0)
    s.chain << kte.spring_3D("spring_b", ends[0], ends[4], 0.1, 400.0, 20.0)
    s.chain << kte.damper_3D("damper_a", ends[1], ends[5], 2.0)
    return s


def crs_gen_elements_chain():
    """3-joint arm with the elements that act on generalized coordinates alone (rigid_link.cpp:30-75, spring.cpp:32-96,
    damper.cpp:32-68): a joint-space return spring and damper from joint 0 to a fixed anchor coordinate, a coupling spring
    (saturating) and damper between joints 1 and 2, and a spring acting on an OFFSET copy of joint 2's coordinate
    (rigid_link_gen) whose force flows back through the link."""
    s = crs_chain(n_revolute=3, physical=True)
    q = s.dofs_gen
    ground = kte.gen_coord(0.25, 0.1)          # a gen_coord that is not a state: keeps q = 0.25, q_dot = 0.1
    shifted = kte.gen_coord()                  # written by the rigid_link_gen
    wall = kte.gen_coord(-0.4)
    s.chain << kte.spring_gen("return_spring", q[0], ground, 0.1, 40.0)
    s.chain << kte.damper_gen("return_damper", q[0], ground, 1.5)
    s.chain << kte.spring_gen("coupling_spring", q[1], q[2], 0.2, 300.0, 25.0)
    s.chain << kte.damper_gen("coupling_damper", q[2], q[1], 0.8)
    s.chain << kte.rigid_link_gen("offset", q[2], shifted, 0.35)
    s.chain << kte.spring_gen("offset_spring", wall, shifted, 0.0, 15.0)
    return s


def planar_gen_elements_chain():
    """2-link planar arm with a joint-space spring / damper to a fixed anchor and between its two joints."""
    s = planar_chain(actuated=True)
    q = s.dofs_gen
    ground = kte.gen_coord(0.3)
    s.chain << kte.spring_gen("spring_0", ground, q[0], 0.0, 12.0, 5.0)
    s.chain << kte.damper_gen("damper_01", q[0], q[1], 0.6)
    s.chain << kte.spring_gen("spring_01", q[0], q[1], 0.5, 8.0)
    return s


def planar_linear_spring_chain():
    """3-link planar arm with spring_2D / damper_2D between the base and the tip and a saturating
    spring between link ends (spring.cpp:116-143, damper.cpp:88-102)."""
    s = planar_chain(lengths=(0.5, 0.4, 0.3), masses=(1.0, 0.8, 0.5), moments=(0.1, 0.05, 0.02), actuated=True)
    ends = [k.mEnd for k in s.chain.getKTEs() if isinstance(k, kte.rigid_link_2D)]
    base = [k for k in s.chain.getKTEs() if isinstance(k, kte.revolute_joint_2D)][0].mBase
    s.chain << kte.spring_2D("spring_a", base, ends[2], 0.4, 30.0)
    s.chain << kte.spring_2D("spring_b", ends[0], ends[2], 0.2, 500.0, 15.0)
    s.chain << kte.damper_2D("damper_a", base, ends[1], 1.5)
    return s


def planar_prismatic_revolute_chain():
    """prismatic_joint_2D (axis x) -> link -> inertia_2D -> revolute_joint_2D -> link -> inertia_2D,
    both actuated, base tilted by 0.3 rad (prismatic_joint.cpp:33-123)."""
    s = kte_system("planar_pr")
    base = kte.frame_2D()
    base.Acceleration = [0.0, 9.81]
    base.Rotation = 0.3
    c0, j0, e0, f0 = kte.gen_coord(), kte.jacobian_gen_2D(), kte.frame_2D(), kte.frame_2D()
    slide = kte.prismatic_joint_2D("slide", c0, (1.0, 0.0), base, e0, j0)
    a0 = kte.driving_actuator_gen("slide_act", c0, slide)
    s.chain << a0 << slide << kte.rigid_link_2D("cart", e0, f0, kte.pose_2D((0.0, 0.1), 0.2))
    m0 = kte.inertia_2D("cart_mass", kte.joint_dependent_frame_2D(f0, {c0: j0}), 2.0, 0.3)
    s.chain << m0
    c1, j1, e1, f1 = kte.gen_coord(), kte.jacobian_gen_2D(), kte.frame_2D(), kte.frame_2D()
    pin = kte.revolute_joint_2D("pin", c1, f0, e1, j1)
    a1 = kte.driving_actuator_gen("pin_act", c1, pin)
    s.chain << a1 << pin << kte.rigid_link_2D("rod", e1, f1, kte.pose_2D((0.6, 0.0), 0.0))
    m1 = kte.inertia_2D("bob", kte.joint_dependent_frame_2D(f1, {c0: j0, c1: j1}), 0.7, 0.04)
    s.chain << m1
    s.dofs_gen += [c0, c1]
    s.inputs += [a0, a1]
    s.mass_calc << m0 << m1 << c0 << c1
    return s


def crs_2d_analog_chain():
    """The planar analog of the CRS A465 on its track, examples/robot_airship/old/CRS_A465_2D_analog.cpp:139-330
    and :462-512: prismatic_joint_2D along x, then three revolute_joint_2D with links of 0.3048, 0.3302 and
    0.0762 m along x; every joint carries a driving_actuator_gen and a unit rotor inertia_gen, every link end a
    unit inertia_2D (m = 1, J = 1); no gravity (the base acceleration is commented out there, :139)."""
    s = kte_system("crs2d")
    cur = kte.frame_2D()
    upstream, rotors = {}, []
    for k, L in enumerate((0.0, 0.3048, 0.3302, 0.0762)):
        coord, jac, end, nxt = kte.gen_coord(), kte.jacobian_gen_2D(), kte.frame_2D(), kte.frame_2D()
        if k == 0:
            joint = kte.prismatic_joint_2D("track_joint", coord, (1.0, 0.0), cur, end, jac)
        else:
            joint = kte.revolute_joint_2D("arm_joint_%d" % k, coord, cur, end, jac)
        act = kte.driving_actuator_gen("actuator_%d" % k, coord, joint)
        dep = kte.joint_dependent_gen_coord(coord)
        dep.add_joint(coord, kte.jacobian_gen_gen(1.0, 0.0))
        rotor = kte.inertia_gen("joint_%d_inertia" % k, dep, 1.0)
        link = kte.rigid_link_2D("link_%d" % k, end, nxt, kte.pose_2D((L, 0.0), 0.0))
        upstream[coord] = jac
        inertia = kte.inertia_2D("link_%d_inertia" % k, kte.joint_dependent_frame_2D(nxt, dict(upstream)), 1.0, 1.0)
        s.chain << act << rotor << joint << link << inertia
        s.inputs.append(act)
        s.dofs_gen.append(coord)
        s.mass_calc << inertia
        rotors.append(rotor)
        cur = nxt
    for r in rotors:
        s.mass_calc << r
    for c in s.dofs_gen:
        s.mass_calc << c
    return s


PRESETS = {
    "pendulum": pendulum_chain,
    "planar2": planar_chain,                                          # cfg 1
    "planar2_act": lambda: planar_chain(actuated=True),
    "planar3_sd": lambda: planar_chain(lengths=(0.5, 0.4, 0.3), masses=(1.0, 0.8, 0.5),
                                       moments=(0.1, 0.05, 0.02), actuated=True, springs=True),
    "crs6": crs_chain,                                                # cfg 2 / 5
    "crs6_phys": lambda: crs_chain(physical=True),
    "crs6_sd": lambda: crs_chain(springs=True),                       # cfg 3
    "crs6_sd_sat": lambda: crs_chain(springs=True, saturation=3.0),
    "crs6_twist": lambda: crs_chain(physical=True, link_rotation=True),
    "crs7": lambda: crs_chain(track=True),                            # cfg 4
    "crs7_phys_sd": lambda: crs_chain(track=True, physical=True, springs=True),
    "torsion1": torsion_1dof_chain,
    "crs3": lambda: crs_chain(n_revolute=3),
    "crs6_passive": lambda: crs_chain(actuated=False),
    "crs6_lin_sd": crs_linear_spring_chain,
    "planar2_lin_sd": planar_linear_spring_chain,
    "planar_pr": planar_prismatic_revolute_chain,
    "crs2d": crs_2d_analog_chain,                                     # the reference's own planar dynamic model
    "crs3_gen": crs_gen_elements_chain,                               # rigid_link_gen / spring_gen / damper_gen
    "planar2_gen": planar_gen_elements_chain,
}


# chains with a free_joint_3D: their state is 2 n + 13 doubles (not in PRESETS, whose users assume 2 n)
FREE_PRESETS = {
    "free_body": lambda: free_base_chain(0),                          # one free rigid body (13 states)
    "free_arm3": lambda: free_base_chain(3),                          # cfg 4's "free" variant, as far as the reference evaluates it
    "free_arm2_twist": lambda: free_base_chain(2, actuated=False, link_rotation=True),
    "free_arm6": lambda: free_base_chain(6),                          # six-joint arm on a free-floating base: 25 states, 12 accelerations
    "free_planar_body": lambda: free_planar_chain(0),                 # free_joint_2D: one planar rigid body (7 states)
    "free_planar2": lambda: free_planar_chain(2),
}


def make(name):
    return (PRESETS.get(name) or FREE_PRESETS[name])()


def crs_proxy_models(system, track=False):
    """(robot, lab) proximity models of the CRS A465 in the MD148 lab: the shapes of
    examples/robot_airship/old/CRS_A465_geom_model.cpp:88-148 (joint2 / link2 / link3 / link5 capsules and the
    end-effector sphere, each on its joint's end frame) and of examples/robot_airship/build_MD148_lab.cpp:103-149
    (floor, north and west walls, the two capsules along the track).  `system` is a crs_chain(...) with at least
    six revolute joints; `track` says whether joint 0 is the prismatic track."""
    from . import proximity as px
    je = system.joint_end_frames[1:] if track else system.joint_end_frames
    half_pi = math.pi * 0.5
    robot = px.proxy_query_model_3D("CRS_A465_model_proxy")
    robot.addShape(px.capped_cylinder("joint2_cyl", je[0], px.pose_3D.axis_angle(half_pi, (1.0, 0.0, 0.0), (0.0, 0.0, 0.3302)), 0.34, 0.09))
    robot.addShape(px.capped_cylinder("link2_cyl", je[1], px.pose_3D((0.0, 0.0, 0.15)), 0.3, 0.07))
    robot.addShape(px.capped_cylinder("link3_cyl", je[2], px.pose_3D((0.0, 0.0, 0.165)), 0.33, 0.07))
    robot.addShape(px.capped_cylinder("link5_cyl", je[4], px.pose_3D((0.0, 0.0, 0.0381)), 0.0762, 0.05))
    robot.addShape(px.sphere("EE_sphere", je[5], px.pose_3D((-0.04, 0.0, 0.05)), 0.11))
    lab = px.proxy_query_model_3D("MD148_basic_lab_proxy")
    lab.addShape(px.plane("MD148_floor", None, px.pose_3D((-0.8, -1.0, 0.0)), (4.0, 6.0)))
    lab.addShape(px.plane("MD148_north_wall", None, px.pose_3D.axis_angle(half_pi, (0.0, -1.0, 0.0), (1.2, -1.0, 1.5)), (3.0, 6.0)))
    lab.addShape(px.plane("MD148_west_wall", None, px.pose_3D.axis_angle(half_pi, (1.0, 0.0, 0.0), (-0.8, 2.0, 1.5)), (4.0, 3.0)))
    lab.addShape(px.capped_cylinder("MD148_robot_track_left", None, px.pose_3D.axis_angle(half_pi, (1.0, 0.0, 0.0), (0.1, -1.71, 0.15)), 3.42, 0.18))
    lab.addShape(px.capped_cylinder("MD148_robot_track_right", None, px.pose_3D.axis_angle(half_pi, (1.0, 0.0, 0.0), (-0.1, -1.71, 0.15)), 3.42, 0.18))
    return robot, lab

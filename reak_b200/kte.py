"""Host-side mirror of the ReaK::kte modelling API for the in-scope element set.

Same class names, constructor argument order and chain-building idiom (`chain << element`)
as the reference (ctrl/mbd_kte/*.hpp), so that a model is written here the way it is written
in examples/robot_airship/old/CRS_A465_models.cpp:260-822.  These objects carry *no*
numerics: `compile_chain()` walks a kte_map_chain + mass_matrix_calc + dofs + inputs (the
members kte_nl_system holds, ctrl/ctrl_sys/kte_nl_system.hpp:70-78) and emits the flat
`rkb_chain_desc` the CUDA library (and the test oracles) consume.
"""
import os
import math

from . import _abi


# --------------------------------------------------------------------------- kinetostatics
class gen_coord(object):
    """core/kinetostatics/gen_coord.hpp:44-178 (q, q_dot, q_ddot, f)."""

    def __init__(self, q=0.0, q_dot=0.0, q_ddot=0.0, f=0.0):
        self.q, self.q_dot, self.q_ddot, self.f = q, q_dot, q_ddot, f


class frame_3D(object):
    """core/kinetostatics/frame_3D.hpp:49-418.  Quat is (w, x, y, z)."""

    def __init__(self):
        self.Position = [0.0, 0.0, 0.0]
        self.Quat = [1.0, 0.0, 0.0, 0.0]
        self.Velocity = [0.0, 0.0, 0.0]
        self.AngVelocity = [0.0, 0.0, 0.0]
        self.Acceleration = [0.0, 0.0, 0.0]
        self.AngAcceleration = [0.0, 0.0, 0.0]


class frame_2D(object):
    """core/kinetostatics/frame_2D.hpp.  Rotation is held as an angle."""

    def __init__(self):
        self.Position = [0.0, 0.0]
        self.Rotation = 0.0
        self.Velocity = [0.0, 0.0]
        self.AngVelocity = 0.0
        self.Acceleration = [0.0, 0.0]
        self.AngAcceleration = 0.0


class pose_3D(object):
    """core/kinetostatics/pose_3D.hpp:47-313 (Position, Quat)."""

    def __init__(self, Position=(0.0, 0.0, 0.0), Quat=(1.0, 0.0, 0.0, 0.0)):
        self.Position = list(Position)
        self.Quat = list(Quat)


class pose_2D(object):
    def __init__(self, Position=(0.0, 0.0), Rotation=0.0):
        self.Position = list(Position)
        self.Rotation = float(Rotation)


def axis_angle_quat(angle, axis):
    """axis_angle(angle, axis).getQuaternion() — rotations_3D.hpp:1962-1974, 2107-2115."""
    n = math.sqrt(sum(a * a for a in axis))
    ax = [a / n for a in axis] if n > 1e-7 else [1.0, 0.0, 0.0]
    s = math.sin(0.5 * angle)
    return [math.cos(0.5 * angle), ax[0] * s, ax[1] * s, ax[2] * s]


class jacobian_gen_3D(object):
    """core/kinetostatics/motion_jacobians.hpp:200-330; a holder the joint fills in doMotion."""


class jacobian_gen_2D(object):
    """core/kinetostatics/motion_jacobians.hpp:108-198."""


class jacobian_3D_3D(object):
    """core/kinetostatics/motion_jacobians.hpp:1035-1240; the holder a free_joint_3D fills with identity blocks."""


class jacobian_2D_2D(object):
    """core/kinetostatics/motion_jacobians.hpp:411-520; the holder a free_joint_2D fills with identity blocks."""


class jacobian_gen_gen(object):
    """core/kinetostatics/motion_jacobians.hpp:49-107."""

    def __init__(self, qd_qd=0.0, qd_qdd=0.0):
        self.qd_qd, self.qd_qdd = qd_qd, qd_qdd


# --------------------------------------------------------------------------- KTE elements
class kte_map(object):
    """ctrl/mbd_kte/kte_map.hpp:62-119."""

    def __init__(self, name=""):
        self.name = name

    def getName(self):
        return self.name


class revolute_joint_3D(kte_map):
    """ctrl/mbd_kte/revolute_joint.hpp:247-252 / revolute_joint.cpp:121-213."""

    def __init__(self, name, angle, axis, base, end, jacobian=None):
        kte_map.__init__(self, name)
        self.mAngle, self.mAxis, self.mBase, self.mEnd, self.mJacobian = angle, list(axis), base, end, jacobian


class prismatic_joint_3D(kte_map):
    """ctrl/mbd_kte/prismatic_joint.hpp:262-267 / prismatic_joint.cpp:129-222."""

    def __init__(self, name, coord, axis, base, end, jacobian=None):
        kte_map.__init__(self, name)
        self.mCoord, self.mAxis, self.mBase, self.mEnd, self.mJacobian = coord, list(axis), base, end, jacobian


class revolute_joint_2D(kte_map):
    """ctrl/mbd_kte/revolute_joint.hpp:118-122 / revolute_joint.cpp:32-116."""

    def __init__(self, name, angle, base, end, jacobian=None):
        kte_map.__init__(self, name)
        self.mAngle, self.mBase, self.mEnd, self.mJacobian = angle, base, end, jacobian


class prismatic_joint_2D(kte_map):
    """ctrl/mbd_kte/prismatic_joint.hpp:131-136 / prismatic_joint.cpp:33-123."""

    def __init__(self, name, coord, axis, base, end, jacobian=None):
        kte_map.__init__(self, name)
        self.mCoord, self.mAxis, self.mBase, self.mEnd, self.mJacobian = coord, list(axis), base, end, jacobian


class free_joint_3D(kte_map):
    """ctrl/mbd_kte/free_joints.hpp:205-300 / free_joints.cpp:119-208: End = Base * Coord, the coordinate being a whole
    frame_3D (position, quaternion, velocity, angular velocity: 13 states, kte_nl_system.hpp:205-219)."""

    def __init__(self, name, coord, base, end, jacobian=None):
        kte_map.__init__(self, name)
        self.mCoord, self.mBase, self.mEnd, self.mJacobian = coord, base, end, jacobian


class free_joint_2D(kte_map):
    """ctrl/mbd_kte/free_joints.hpp:46-130 / free_joints.cpp:33-117: End = Base * Coord, the coordinate being a whole
    frame_2D (position, rotation as (cos, sin), velocity, angular velocity: 7 states, kte_nl_system.hpp:194-204)."""

    def __init__(self, name, coord, base, end, jacobian=None):
        kte_map.__init__(self, name)
        self.mCoord, self.mBase, self.mEnd, self.mJacobian = coord, base, end, jacobian


class rigid_link_3D(kte_map):
    """ctrl/mbd_kte/rigid_link.hpp:296-299 / rigid_link.cpp:152-185."""

    def __init__(self, name, base, end, pose_offset):
        kte_map.__init__(self, name)
        self.mBase, self.mEnd, self.mPoseOffset = base, end, pose_offset


class rigid_link_2D(kte_map):
    """ctrl/mbd_kte/rigid_link.hpp:200-203 / rigid_link.cpp:87-139."""

    def __init__(self, name, base, end, pose_offset):
        kte_map.__init__(self, name)
        self.mBase, self.mEnd, self.mPoseOffset = base, end, pose_offset


class _joint_dependent(object):
    def __init__(self, frame=None, upstream=None):
        self.mFrame = frame
        self.mUpStreamJoints = dict(upstream or {})
        self.mUpStream3DJoints = {}  # free-joint coordinate frames (jacobian_joint_map.hpp:252-331)

    def add_joint(self, joint_coord, joint_jacobian):
        if isinstance(joint_coord, (frame_3D, frame_2D)):   # (mUpStream3DJoints / mUpStream2DJoints: one kind per chain)
            self.mUpStream3DJoints[joint_coord] = joint_jacobian
        else:
            self.mUpStreamJoints[joint_coord] = joint_jacobian
        return self


class joint_dependent_gen_coord(_joint_dependent):
    """ctrl/mbd_kte/jacobian_joint_map.hpp:76-160."""


class joint_dependent_frame_2D(_joint_dependent):
    """ctrl/mbd_kte/jacobian_joint_map.hpp:164-248."""


class joint_dependent_frame_3D(_joint_dependent):
    """ctrl/mbd_kte/jacobian_joint_map.hpp:252-331."""


class inertia_gen(kte_map):
    """ctrl/mbd_kte/inertia.hpp:96-98 / inertia.cpp:47-53."""

    def __init__(self, name, center_of_mass, mass):
        kte_map.__init__(self, name)
        self.mCenterOfMass, self.mMass = center_of_mass, float(mass)


class inertia_2D(kte_map):
    """ctrl/mbd_kte/inertia.hpp:190-193 / inertia.cpp:77-86."""

    def __init__(self, name, center_of_mass, mass, moment_of_inertia):
        kte_map.__init__(self, name)
        self.mCenterOfMass, self.mMass, self.mMomentOfInertia = center_of_mass, float(mass), float(moment_of_inertia)


class inertia_3D(kte_map):
    """ctrl/mbd_kte/inertia.hpp:286-289 / inertia.cpp:111-121.

    `inertia_tensor` = (Ixx, Ixy, Ixz, Iyy, Iyz, Izz), the argument order of
    mat<double,mat_structure::symmetric>(a11,a12,a13,a22,a23,a33)."""

    def __init__(self, name, center_of_mass, mass, inertia_tensor):
        kte_map.__init__(self, name)
        self.mCenterOfMass, self.mMass, self.mInertiaTensor = center_of_mass, float(mass), list(inertia_tensor)


class driving_actuator_gen(kte_map):
    """ctrl/mbd_kte/driving_actuator.hpp:50-105 / driving_actuator.cpp:31-38; a system_input with 1 input."""

    def __init__(self, name, frame, joint):
        kte_map.__init__(self, name)
        self.mFrame, self.mJoint = frame, joint

    def getInputCount(self):
        return 1


class torsion_spring_3D(kte_map):
    """ctrl/mbd_kte/torsion_spring.hpp:235-239 / torsion_spring.cpp:106-129."""

    def __init__(self, name, anchor1, anchor2, stiffness, saturation=0.0):
        kte_map.__init__(self, name)
        self.mAnchor1, self.mAnchor2, self.mStiffness, self.mSaturation = anchor1, anchor2, float(stiffness), float(saturation)


class torsion_spring_2D(torsion_spring_3D):
    """ctrl/mbd_kte/torsion_spring.hpp:121-125 / torsion_spring.cpp:50-71."""


class torsion_damper_3D(kte_map):
    """ctrl/mbd_kte/torsion_damper.hpp:194-197 / torsion_damper.cpp:93-104."""

    def __init__(self, name, anchor1, anchor2, damping):
        kte_map.__init__(self, name)
        self.mAnchor1, self.mAnchor2, self.mDamping = anchor1, anchor2, float(damping)


class torsion_damper_2D(torsion_damper_3D):
    """ctrl/mbd_kte/torsion_damper.hpp:102-105 / torsion_damper.cpp:49-58."""


class spring_3D(kte_map):
    """ctrl/mbd_kte/spring.hpp:396-401 / spring.cpp:178-207."""

    def __init__(self, name, anchor1, anchor2, rest_length, stiffness, saturation=0.0):
        kte_map.__init__(self, name)
        self.mAnchor1, self.mAnchor2 = anchor1, anchor2
        self.mRestLength, self.mStiffness, self.mSaturation = float(rest_length), float(stiffness), float(saturation)


class spring_2D(spring_3D):
    """ctrl/mbd_kte/spring.hpp:265-270 / spring.cpp:116-143."""


class damper_3D(kte_map):
    """ctrl/mbd_kte/damper.hpp:290-293 / damper.cpp:136-149."""

    def __init__(self, name, anchor1, anchor2, damping):
        kte_map.__init__(self, name)
        self.mAnchor1, self.mAnchor2, self.mDamping = anchor1, anchor2, float(damping)


class damper_2D(damper_3D):
    """ctrl/mbd_kte/damper.hpp:197-200 / damper.cpp:88-102."""


class rigid_link_gen(kte_map):
    """ctrl/mbd_kte/rigid_link.hpp:40-118 / rigid_link.cpp:30-75: End.q = Base.q + offset on generalized coordinates."""

    def __init__(self, name, base, end, offset):
        kte_map.__init__(self, name)
        self.mBase, self.mEnd, self.mOffset = base, end, float(offset)


class spring_gen(kte_map):
    """ctrl/mbd_kte/spring.hpp:44-175 / spring.cpp:32-96: linear spring between two generalized coordinates."""

    def __init__(self, name, anchor1, anchor2, rest_length, stiffness, saturation=0.0):
        kte_map.__init__(self, name)
        self.mAnchor1, self.mAnchor2 = anchor1, anchor2
        self.mRestLength, self.mStiffness, self.mSaturation = float(rest_length), float(stiffness), float(saturation)


class damper_gen(kte_map):
    """ctrl/mbd_kte/damper.hpp:44-140 / damper.cpp:32-68: linear damper between two generalized coordinates."""

    def __init__(self, name, anchor1, anchor2, damping):
        kte_map.__init__(self, name)
        self.mAnchor1, self.mAnchor2, self.mDamping = anchor1, anchor2, float(damping)


class kte_map_chain(kte_map):
    """ctrl/mbd_kte/kte_map_chain.hpp:50-120: ordered list of KTEs, `chain << kte` appends."""

    def __init__(self, name=""):
        kte_map.__init__(self, name)
        self.mKTEs = []

    def __lshift__(self, kte):
        if kte is not None:
            self.mKTEs.append(kte)
        return self

    def getKTEs(self):
        return self.mKTEs


class mass_matrix_calc(object):
    """ctrl/mbd_kte/mass_matrix_calculator.hpp / .cpp:30-78: `<<` registers inertias and coordinates."""

    def __init__(self, name=""):
        self.name = name
        self.mGenInertias, self.m2DInertias, self.m3DInertias, self.mCoords = [], [], [], []
        self.mFrames3D = []

    def __lshift__(self, obj):
        if isinstance(obj, inertia_gen):
            self.mGenInertias.append(obj)
        elif isinstance(obj, inertia_2D):
            self.m2DInertias.append(obj)
        elif isinstance(obj, inertia_3D):
            self.m3DInertias.append(obj)
        elif isinstance(obj, gen_coord):
            self.mCoords.append(obj)
        elif isinstance(obj, (frame_3D, frame_2D)):
            self.mFrames3D.append(obj)   # mass_matrix_calculator.cpp:64-78: a free joint's coordinate frame (mFrames2D / mFrames3D)
        else:
            raise TypeError("mass_matrix_calc << %r" % (obj,))
        return self


# --------------------------------------------------------------------------- chain compiler
class UnsupportedChain(ValueError):
    pass


class compiled_chain(object):
    """Owns the ctypes descriptor (and keeps the element array alive)."""

    def __init__(self, desc, elements, frames, coords):
        self.desc, self.elements, self.frames, self.coords = desc, elements, frames, coords
        self.n_coords, self.n_inputs, self.dim = desc.n_coords, desc.n_inputs, desc.dim
        # free joints (free_joint_3D): 13 states and 6 accelerations each, after the generalized coordinates'
        self.n_free, self.nx, self.n_acc = 0, 2 * desc.n_coords, desc.n_coords


def compile_chain(chain, mass_calc, dofs_gen, inputs, dofs_3D=()):
    """Flatten (chain, mass_calc, dofs_gen, inputs) — the public members of kte_nl_system —
    into an `rkb_chain_desc`.  Coordinates are numbered in `dofs_gen` order (that is the state
    layout, kte_nl_system.hpp:189-193) and inputs in `inputs` order (kte_nl_system.hpp:221-224)."""
    if len(dofs_gen) > _abi.RKB_MAX_COORDS:
        raise UnsupportedChain("more than %d generalized coordinates" % _abi.RKB_MAX_COORDS)
    if list(mass_calc.mCoords) != list(dofs_gen):
        raise UnsupportedChain("mass_matrix_calc coordinates must be the system dofs in the same order")
    dofs_3D = list(dofs_3D)
    if [id(f) for f in getattr(mass_calc, "mFrames3D", [])] != [id(f) for f in dofs_3D]:
        raise UnsupportedChain("mass_matrix_calc 3D frames must be the system's dofs_3D in the same order")
    if len(dofs_3D) > _abi.RKB_MAX_FREE:
        raise UnsupportedChain("more than %d free joints" % _abi.RKB_MAX_FREE)
    free_id = {id(f): i for i, f in enumerate(dofs_3D)}
    coord_id = {id(c): i for i, c in enumerate(dofs_gen)}
    input_id = {id(a): i for i, a in enumerate(inputs)}
    frames, frame_id = [], {}
    dim = [0]

    def fid(f):
        d = 3 if isinstance(f, frame_3D) else 2 if isinstance(f, frame_2D) else 0
        if d == 0:
            raise UnsupportedChain("frame of unsupported type %r" % (f,))
        if dim[0] and dim[0] != d:
            raise UnsupportedChain("mixed 2D/3D frames in one chain")
        dim[0] = d
        if id(f) not in frame_id:
            frame_id[id(f)] = len(frames)
            frames.append(f)
        return frame_id[id(f)]

    def cid(c):
        if id(c) not in coord_id:
            raise UnsupportedChain("element refers to a coordinate that is not a system dof")
        return coord_id[id(c)]

    aux_id = {}  # gen_coords that are not system states: anchors and rigid_link_gen ends, numbered after the dofs

    def gid(c):
        """any generalized coordinate: a dof, or an auxiliary one (declared by a COORD_GEN record right before the first
        element that uses it — the rule the C++ bridge follows too)"""
        if not isinstance(c, gen_coord):
            raise UnsupportedChain("a _gen element needs gen_coord anchors, got %r" % (c,))
        if id(c) in coord_id:
            return coord_id[id(c)]
        if id(c) not in aux_id:
            if len(dofs_gen) + len(aux_id) >= _abi.RKB_MAX_COORDS:
                raise UnsupportedChain("more than %d generalized coordinates (dofs + auxiliaries)" % _abi.RKB_MAX_COORDS)
            aux_id[id(c)] = len(dofs_gen) + len(aux_id)
            rec(_abi.COORD_GEN, coord=aux_id[id(c)], p=[c.q, c.q_dot, c.q_ddot])
        return aux_id[id(c)]

    def upstream_mask(dep):
        m = 0
        for c in dep.mUpStreamJoints:
            m |= 1 << cid(c)
        for f in getattr(dep, "mUpStream3DJoints", {}):   # free-joint frames: bits 32 and up
            if id(f) not in free_id:
                raise UnsupportedChain("inertia depends on a free-joint frame that is not in dofs_3D")
            m |= 1 << (32 + free_id[id(f)])
        return m

    recs, written, elem_index, actuator_joint = [], set(), {}, {}
    registered = set(id(x) for x in mass_calc.mGenInertias + mass_calc.m2DInertias + mass_calc.m3DInertias)

    def rec(kind, a=-1, b=-1, coord=-1, aux=0, upstream=0, p=()):
        e = _abi.rkb_element()
        e.kind, e.frame_a, e.frame_b, e.coord, e.aux, e.upstream = kind, a, b, coord, aux, upstream
        for i, v in enumerate(p):
            e.p[i] = float(v)
        recs.append(e)

    for k in chain.getKTEs():
        elem_index[id(k)] = len(recs)
        if isinstance(k, revolute_joint_3D):
            a, b = fid(k.mBase), fid(k.mEnd); written.add(b)
            rec(_abi.REVOLUTE_3D, a, b, cid(k.mAngle), p=k.mAxis)
        elif isinstance(k, prismatic_joint_3D):
            a, b = fid(k.mBase), fid(k.mEnd); written.add(b)
            rec(_abi.PRISMATIC_3D, a, b, cid(k.mCoord), p=k.mAxis)
        elif isinstance(k, revolute_joint_2D):
            a, b = fid(k.mBase), fid(k.mEnd); written.add(b)
            rec(_abi.REVOLUTE_2D, a, b, cid(k.mAngle))
        elif isinstance(k, prismatic_joint_2D):
            a, b = fid(k.mBase), fid(k.mEnd); written.add(b)
            rec(_abi.PRISMATIC_2D, a, b, cid(k.mCoord), p=k.mAxis)
        elif isinstance(k, (free_joint_3D, free_joint_2D)):
            if id(k.mCoord) not in free_id:
                raise UnsupportedChain("free joint %s: its coordinate frame is not among the system's free-frame dofs" % k.name)
            a, b = fid(k.mBase), fid(k.mEnd); written.add(b)
            rec(_abi.FREE_3D if isinstance(k, free_joint_3D) else _abi.FREE_2D, a, b, free_id[id(k.mCoord)])
        elif isinstance(k, rigid_link_gen):
            a, b = gid(k.mBase), gid(k.mEnd)
            if b < len(dofs_gen):
                raise UnsupportedChain("rigid_link_gen %s ends on a system dof (the state would be overwritten)" % k.name)
            rec(_abi.RIGID_LINK_GEN, coord=a, aux=b, p=[k.mOffset])
        elif isinstance(k, spring_gen):
            a, b = gid(k.mAnchor1), gid(k.mAnchor2)
            rec(_abi.SPRING_GEN, coord=a, aux=b, p=[k.mRestLength, k.mStiffness, k.mSaturation])
        elif isinstance(k, damper_gen):
            a, b = gid(k.mAnchor1), gid(k.mAnchor2)
            rec(_abi.DAMPER_GEN, coord=a, aux=b, p=[k.mDamping])
        elif isinstance(k, rigid_link_3D):
            a, b = fid(k.mBase), fid(k.mEnd); written.add(b)
            rec(_abi.RIGID_LINK_3D, a, b, p=list(k.mPoseOffset.Position) + list(k.mPoseOffset.Quat))
        elif isinstance(k, rigid_link_2D):
            a, b = fid(k.mBase), fid(k.mEnd); written.add(b)
            rec(_abi.RIGID_LINK_2D, a, b, p=list(k.mPoseOffset.Position) + [k.mPoseOffset.Rotation])
        elif isinstance(k, inertia_3D):
            if id(k) not in registered:
                raise UnsupportedChain("inertia %s is in the chain but not in the mass_matrix_calc" % k.name)
            rec(_abi.INERTIA_3D, fid(k.mCenterOfMass.mFrame), upstream=upstream_mask(k.mCenterOfMass),
                p=[k.mMass] + list(k.mInertiaTensor))
        elif isinstance(k, inertia_2D):
            if id(k) not in registered:
                raise UnsupportedChain("inertia %s is in the chain but not in the mass_matrix_calc" % k.name)
            rec(_abi.INERTIA_2D, fid(k.mCenterOfMass.mFrame), upstream=upstream_mask(k.mCenterOfMass),
                p=[k.mMass, k.mMomentOfInertia])
        elif isinstance(k, inertia_gen):
            if id(k) not in registered:
                raise UnsupportedChain("inertia %s is in the chain but not in the mass_matrix_calc" % k.name)
            c = cid(k.mCenterOfMass.mFrame)
            jac = k.mCenterOfMass.mUpStreamJoints.get(k.mCenterOfMass.mFrame)
            if upstream_mask(k.mCenterOfMass) != (1 << c) or jac is None or jac.qd_qd != 1.0 or jac.qd_qdd != 0.0:
                raise UnsupportedChain("inertia_gen must depend on its own coordinate through jacobian_gen_gen(1,0)")
            rec(_abi.INERTIA_GEN, coord=c, upstream=1 << c, p=[k.mMass])
        elif isinstance(k, driving_actuator_gen):
            if id(k) not in input_id:
                raise UnsupportedChain("actuator %s is not listed in the system inputs" % k.name)
            # the joint may come later in the chain (CRS order: actuator, rotor, joint, ...)
            rec(_abi.ACTUATOR_GEN, b=-1, coord=cid(k.mFrame), aux=input_id[id(k)])
            actuator_joint[len(recs) - 1] = k.mJoint
        elif isinstance(k, (torsion_spring_2D, torsion_spring_3D)):
            kind = _abi.TORSION_SPRING_2D if isinstance(k.mAnchor1, frame_2D) else _abi.TORSION_SPRING_3D
            rec(kind, fid(k.mAnchor1), fid(k.mAnchor2), p=[k.mStiffness, k.mSaturation])
        elif isinstance(k, (torsion_damper_2D, torsion_damper_3D)):
            kind = _abi.TORSION_DAMPER_2D if isinstance(k.mAnchor1, frame_2D) else _abi.TORSION_DAMPER_3D
            rec(kind, fid(k.mAnchor1), fid(k.mAnchor2), p=[k.mDamping])
        elif isinstance(k, (spring_2D, spring_3D)):
            kind = _abi.SPRING_2D if isinstance(k.mAnchor1, frame_2D) else _abi.SPRING_3D
            rec(kind, fid(k.mAnchor1), fid(k.mAnchor2), p=[k.mRestLength, k.mStiffness, k.mSaturation])
        elif isinstance(k, (damper_2D, damper_3D)):
            kind = _abi.DAMPER_2D if isinstance(k.mAnchor1, frame_2D) else _abi.DAMPER_3D
            rec(kind, fid(k.mAnchor1), fid(k.mAnchor2), p=[k.mDamping])
        else:
            raise UnsupportedChain("KTE %r (%s) is outside the compiled element set" % (k, type(k).__name__))

    # the converse: M comes from mass_calc (kte_nl_system.hpp:271) and the forces from the chain, so an inertia
    # registered with mass_calc but absent from the chain would enter the reference's M and be dropped here
    seen = set(id(k) for k in chain.getKTEs())
    for x in mass_calc.mGenInertias + mass_calc.m2DInertias + mass_calc.m3DInertias:
        if id(x) not in seen:
            raise UnsupportedChain("inertia %s is registered with the mass_matrix_calc but is not in the chain" % x.name)

    for i, joint in actuator_joint.items():  # resolve actuator -> joint element index
        j = elem_index.get(id(joint))
        if j is None:
            raise UnsupportedChain("actuator drives a joint that is not in the chain")
        recs[i].frame_b = j

    if len(input_id) != sum(1 for e in recs if e.kind == _abi.ACTUATOR_GEN):
        raise UnsupportedChain("system inputs and chain actuators differ")
    if dofs_3D:
        # the reference's mass_matrix_calc dereferences a null Jacobian when a gen inertia depends on mCoords[i] for a
        # free-frame index i (mass_matrix_calculator.cpp:226-233): there is no behaviour to match for such a model
        n_free_used = sum(1 for e in recs if e.kind in (_abi.FREE_3D, _abi.FREE_2D))
        if n_free_used != len(dofs_3D):
            raise UnsupportedChain("every frame of dofs_3D needs exactly one free_joint_3D in the chain")
        for e in recs:
            if e.kind == _abi.INERTIA_GEN and e.coord < len(dofs_3D):
                raise UnsupportedChain("inertia_gen on coordinate %d next to %d free joint(s): the reference itself crashes on this "
                                       "model (mass_matrix_calculator.cpp:226-233)" % (e.coord, len(dofs_3D)))
    roots = [i for i in range(len(frames)) if i not in written]
    if len(roots) != 1:
        raise UnsupportedChain("chain must have exactly one un-driven base frame (found %d)" % len(roots))

    arr = (_abi.rkb_element * len(recs))()
    for i, e in enumerate(recs):
        arr[i] = e
    d = _abi.rkb_chain_desc()
    d.dim, d.n_elements, d.n_frames = dim[0], len(recs), len(frames)
    d.n_coords, d.n_inputs, d.base_frame = len(dofs_gen), len(inputs), roots[0]
    B = frames[roots[0]]
    if dim[0] == 3:
        for i in range(3):
            d.base.position[i] = B.Position[i]; d.base.velocity[i] = B.Velocity[i]
            d.base.ang_velocity[i] = B.AngVelocity[i]; d.base.acceleration[i] = B.Acceleration[i]
            d.base.ang_acceleration[i] = B.AngAcceleration[i]
        for i in range(4):
            d.base.quat[i] = B.Quat[i]
    else:
        for i in range(2):
            d.base.position[i] = B.Position[i]; d.base.velocity[i] = B.Velocity[i]
            d.base.acceleration[i] = B.Acceleration[i]
        d.base.quat[0] = B.Rotation
        d.base.ang_velocity[0] = B.AngVelocity
        d.base.ang_acceleration[0] = B.AngAcceleration
    import ctypes as C
    d.elements = C.cast(arr, C.POINTER(_abi.rkb_element))
    cc = compiled_chain(d, arr, frames, list(dofs_gen))
    cc.n_free = len(dofs_3D)
    fs, fa = (13, 6) if d.dim == 3 else (7, 3)   # states / accelerations per free joint (kte_nl_system.hpp:145-147)
    cc.nx, cc.n_acc = 2 * len(dofs_gen) + fs * len(dofs_3D), len(dofs_gen) + fa * len(dofs_3D)
    return cc


def read_rkx(path, max_elements=512):
    """The flat descriptor of the kte_nl_system in a ReaK XML archive (rkb_rkx_read: the library's own reader of the
    `.rkx` files ReaK::serialization::xml_oarchive writes).  Returns a compiled_chain without Python-side objects
    (frames / coords are None: address frames by id)."""
    import ctypes as C
    lib = _abi.load_library()
    d = _abi.rkb_chain_desc()
    arr = (_abi.rkb_element * max_elements)()
    err = C.create_string_buffer(512)
    n = lib.rkb_rkx_read(os.fsencode(path), C.byref(d), arr, max_elements, err, 512)
    if n < 0:
        raise UnsupportedChain("%s: %s" % (path, err.value.decode() or "rkb_rkx_read failed (%d)" % n))
    own = (_abi.rkb_element * max(n, 1))()
    C.memmove(own, arr, n * C.sizeof(_abi.rkb_element))
    d.elements = C.cast(own, C.POINTER(_abi.rkb_element))
    cc = compiled_chain(d, (_abi.rkb_element * n).from_buffer(own) if n else own, [None] * d.n_frames, [None] * d.n_coords)
    cc._keep = own
    cc.n_free = sum(1 for i in range(n) if own[i].kind in (_abi.FREE_3D, _abi.FREE_2D))
    fs, fa = (13, 6) if d.dim == 3 else (7, 3)
    cc.nx, cc.n_acc = 2 * d.n_coords + fs * cc.n_free, d.n_coords + fa * cc.n_free
    return cc

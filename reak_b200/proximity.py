"""Proximity models: the host-side mirror of ReaK's geometry/shapes and proxy_query_model vocabulary.

    robot = proxy_query_model_3D("robot").addShape(capped_cylinder("link2", anchor=frame_id, pose=..., length=.3, radius=.05))
    lab   = proxy_query_model_3D("lab").addShape(plane("floor", dims=(4, 6)))
    pair  = proxy_query_pair_3D("robot-lab", robot, lab)
    d, finder, pts = propagator.get_min_distances(pair, x)        # rkb_min_distance

Names and argument order follow geometry/shapes/{plane,sphere,capped_cylinder,cylinder,box}.hpp and
geometry/proximity/proxy_query_model.hpp; `anchor` is a frame id of the chain (the integer the chain builder
hands out) or the frame_3D object itself instead of a shared_ptr< pose_3D >, None for a shape fixed in the world.
"""
import ctypes as C

import numpy as np

from . import _abi


class pose_3D:
    """Position + unit quaternion (w, x, y, z) relative to the anchor (core/kinetostatics/pose_3D.hpp)."""

    def __init__(self, position=(0.0, 0.0, 0.0), quat=(1.0, 0.0, 0.0, 0.0)):
        self.position = tuple(float(v) for v in position)
        self.quat = tuple(float(v) for v in quat)

    @staticmethod
    def axis_angle(angle, axis, position=(0.0, 0.0, 0.0)):
        """axis_angle(angle, axis).getQuaternion() (rotations_3D.hpp:2107-2115)."""
        a = np.asarray(axis, dtype=np.float64)
        a = a / np.linalg.norm(a)
        h = 0.5 * float(angle)
        return pose_3D(position, (np.cos(h), a[0] * np.sin(h), a[1] * np.sin(h), a[2] * np.sin(h)))


class shape_3D:
    kind = 0

    def __init__(self, name="", anchor=None, pose=None, dims=()):
        self.name = name
        self.anchor = anchor  # None (world), a frame id, or a frame_3D object of the chain
        self.pose = pose if pose is not None else pose_3D()
        self.dims = tuple(float(v) for v in dims) + (0.0,) * (3 - len(dims))

    def to_c(self, frames=None):
        s = _abi.rkb_shape()
        s.kind = self.kind
        if self.anchor is None:
            s.anchor = -1
        elif isinstance(self.anchor, (int, np.integer)):
            s.anchor = int(self.anchor)
        else:
            ids = [k for k, f in enumerate(frames or []) if f is self.anchor]
            if not ids:
                raise ValueError("shape %r is anchored to a frame that is not part of the chain" % (self.name,))
            s.anchor = ids[0]
        s.position[:] = self.pose.position
        s.quat[:] = self.pose.quat
        s.dims[:] = self.dims
        return s


class plane(shape_3D):
    kind = _abi.SHAPE_PLANE

    def __init__(self, name="", anchor=None, pose=None, dims=(1.0, 1.0)):
        super().__init__(name, anchor, pose, dims)


class sphere(shape_3D):
    kind = _abi.SHAPE_SPHERE

    def __init__(self, name="", anchor=None, pose=None, radius=1.0):
        super().__init__(name, anchor, pose, (radius,))


class capped_cylinder(shape_3D):
    kind = _abi.SHAPE_CCYLINDER

    def __init__(self, name="", anchor=None, pose=None, length=1.0, radius=1.0):
        super().__init__(name, anchor, pose, (length, radius))


class cylinder(shape_3D):
    kind = _abi.SHAPE_CYLINDER

    def __init__(self, name="", anchor=None, pose=None, length=1.0, radius=1.0):
        super().__init__(name, anchor, pose, (length, radius))


class box(shape_3D):
    kind = _abi.SHAPE_BOX

    def __init__(self, name="", anchor=None, pose=None, dims=(1.0, 1.0, 1.0)):
        super().__init__(name, anchor, pose, dims)


class pose_2D:
    """Position + rotation angle relative to the anchor (core/kinetostatics/pose_2D.hpp: Position, rot_mat_2D(angle))."""

    def __init__(self, position=(0.0, 0.0), angle=0.0):
        self.position = tuple(float(v) for v in position)
        self.angle = float(angle)


class shape_2D(shape_3D):
    """planar shapes ride on planar chains; rkb_shape carries position[0..1] and quat[0..1] = (cos, sin)"""

    def __init__(self, name="", anchor=None, pose=None, dims=()):
        self.name = name
        self.anchor = anchor
        self.pose2 = pose if pose is not None else pose_2D()
        # rot_mat_2D(angle) stores cos(angle), sin(angle) (rotations_2D.hpp:104-109)
        self.pose = pose_3D(self.pose2.position + (0.0,), (np.cos(self.pose2.angle), np.sin(self.pose2.angle), 0.0, 0.0))
        self.dims = tuple(float(v) for v in dims) + (0.0,) * (3 - len(dims))


class circle(shape_2D):
    kind = _abi.SHAPE_CIRCLE

    def __init__(self, name="", anchor=None, pose=None, radius=1.0):
        super().__init__(name, anchor, pose, (radius,))


class capped_rectangle(shape_2D):
    kind = _abi.SHAPE_CRECT

    def __init__(self, name="", anchor=None, pose=None, dims=(1.0, 1.0)):
        super().__init__(name, anchor, pose, dims)


class rectangle(shape_2D):
    kind = _abi.SHAPE_RECTANGLE

    def __init__(self, name="", anchor=None, pose=None, dims=(1.0, 1.0)):
        super().__init__(name, anchor, pose, dims)


class proxy_query_model_3D:
    def __init__(self, name=""):
        self.name = name
        self.mShapeList = []

    def addShape(self, shape):
        self.mShapeList.append(shape)
        return self

    def to_c(self, frames=None):
        arr = (_abi.rkb_shape * max(len(self.mShapeList), 1))()
        for k, s in enumerate(self.mShapeList):
            arr[k] = s.to_c(frames)
        return arr, len(self.mShapeList)


class proxy_query_pair_3D:
    """Two models and the finder list createProxFinderList derives from them (proxy_query_model.cpp:212-384)."""

    def __init__(self, name, model1, model2):
        self.name = name
        self.model1, self.model2 = model1, model2

    def finder_pairs(self):
        """(i1, i2) of every finder, in order — restated here only so that host code can name a finder index."""
        out = []
        for a, sa in enumerate(self.model1.mShapeList):
            for b, sb in enumerate(self.model2.mShapeList):
                lo, hi = min(sa.kind, sb.kind), max(sa.kind, sb.kind)
                if lo in (_abi.SHAPE_PLANE, _abi.SHAPE_SPHERE) or (lo == _abi.SHAPE_CCYLINDER and hi in (_abi.SHAPE_CCYLINDER, _abi.SHAPE_BOX)):
                    out.append((a, b))
        return out


class proxy_query_model_2D(proxy_query_model_3D):
    pass


class proxy_query_pair_2D(proxy_query_pair_3D):
    """Two planar models; every pair of planar shapes has a finder (proxy_query_model.cpp:73-160)."""

    def finder_pairs(self):
        return [(a, b) for a in range(len(self.model1.mShapeList)) for b in range(len(self.model2.mShapeList))]


class ProxyHandle:
    """rkb_proxy owner."""

    def __init__(self, lib, chain_handle, pair, frames=None):
        self._lib = lib
        self._h = C.c_void_p()
        m1, n1 = pair.model1.to_c(frames)
        m2, n2 = pair.model2.to_c(frames)
        _abi.check(lib.rkb_proxy_create(chain_handle, m1, n1, m2, n2, C.byref(self._h)), "rkb_proxy_create")
        self.n_finders = lib.rkb_proxy_finder_count(self._h)

    def finder(self, k):
        a, b = C.c_int(), C.c_int()
        _abi.check(self._lib.rkb_proxy_finder(self._h, int(k), C.byref(a), C.byref(b)), "rkb_proxy_finder")
        return a.value, b.value

    # run-time specialisation of the pair's query (rkb_proxy_specialize)
    OPT_AUTO_SPECIALIZE, OPT_MIN_BLOCKS = 1, 2

    def set_option(self, option, value):
        _abi.check(self._lib.rkb_proxy_set_option(self._h, int(option), int(value)), "rkb_proxy_set_option")

    def specialize(self, device=0):
        _abi.check(self._lib.rkb_proxy_specialize(self._h, int(device)), "rkb_proxy_specialize")

    def is_specialized(self):
        return bool(self._lib.rkb_proxy_is_specialized(self._h))

    def source(self):
        """the CUDA source rkb_proxy_specialize compiles (test hook)"""
        n = self._lib.rkb_proxy_source(self._h, None, 0)
        _abi.check(min(n, 0), "rkb_proxy_source")
        buf = C.create_string_buffer(n)
        _abi.check(min(self._lib.rkb_proxy_source(self._h, buf, n), 0), "rkb_proxy_source")
        return buf.value.decode()

    def close(self):
        if self._h:
            self._lib.rkb_proxy_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

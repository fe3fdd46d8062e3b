"""ctypes access to the two checkers (see oracle/__init__.py).  numpy in, numpy out, AoS layout."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libreak_ref.so")
ORACLE_SO = os.path.join(_HERE, "libkte_oracle.so")


def build(which=("oracle", "ref")):
    """Run oracle/Makefile (gcc only).  `ref` is a no-op where /root/reference is absent."""
    for target in which:
        subprocess.check_call(["make", "-s", "-C", _HERE, target])


def have_ref():
    return os.path.isfile(REF_SO)


def _dp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class _Checker(object):
    """Common driver for a library exporting <prefix>_{create,destroy,eval,gen_forces,mass,rk4,frames}."""

    def __init__(self, so_path, prefix, compiled):
        if not os.path.isfile(so_path):
            raise RuntimeError("%s is not built (make -C oracle)" % so_path)
        self.lib = C.CDLL(so_path)
        self._prefix = prefix
        self.compiled = compiled  # keeps the descriptor arrays alive
        self.n, self.nu = compiled.n_coords, compiled.n_inputs
        # chains with a free_joint_3D carry 13 more states and 6 more accelerations (kte_nl_system.hpp:145-147)
        self.nx = getattr(compiled, "nx", 2 * compiled.n_coords)
        self.na = getattr(compiled, "n_acc", compiled.n_coords)
        g = lambda name: getattr(self.lib, prefix + name)
        self._create, self._destroy = g("create"), g("destroy")
        self._eval, self._forces, self._mass, self._rk4, self._frames = g("eval"), g("gen_forces"), g("mass"), g("rk4"), g("frames")
        self._create.restype = C.c_void_p
        self._create.argtypes = [C.c_void_p]
        self._destroy.argtypes = [C.c_void_p]
        self._eval.argtypes = [C.c_void_p, C.c_size_t] + [C.c_void_p] * 4
        self._forces.argtypes = [C.c_void_p, C.c_size_t] + [C.c_void_p] * 3
        self._mass.argtypes = [C.c_void_p, C.c_size_t] + [C.c_void_p] * 3
        self._rk4.restype = C.c_double
        self._rk4.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_double, C.c_int,
                              C.c_void_p, C.c_void_p, C.c_int]
        self._integrate = g("integrate")
        self._integrate.restype = C.c_double
        self._integrate.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_int]
        self._frames.argtypes = [C.c_void_p] * 4
        self.h = self._create(C.byref(compiled.desc))
        if not self.h:
            raise RuntimeError("checker could not build the model from the descriptor")

    def close(self):
        if self.h:
            self._destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _xu(self, x, u):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, self.nx)
        N = x.shape[0]
        if u is None:
            u = np.zeros((N, self.nu))
        u = np.ascontiguousarray(u, dtype=np.float64).reshape(N, self.nu)
        return x, u, N

    def eval(self, x, u=None):
        x, u, N = self._xu(x, u)
        xd, st = np.empty_like(x), np.zeros(N, dtype=np.int32)
        self._eval(self.h, N, _dp(x), _dp(u), _dp(xd), _dp(st))
        return xd, st

    def gen_forces(self, x, u=None):
        x, u, N = self._xu(x, u)
        f = np.empty((N, self.na))
        self._forces(self.h, N, _dp(x), _dp(u), _dp(f))
        return f

    def mass(self, x, with_dot=True):
        x, _, N = self._xu(x, None)
        M = np.empty((N, self.na, self.na))
        Md = np.empty((N, self.na, self.na)) if with_dot else None
        self._mass(self.h, N, _dp(x), _dp(M), _dp(Md))
        return (M, Md) if with_dot else M

    def rk4(self, x0, u, dt, n_steps, n_workers=1):
        """Returns (x_out, status, seconds)."""
        x, u, N = self._xu(x0, u)
        out, st = np.empty_like(x), np.zeros(N, dtype=np.int32)
        secs = self._rk4(self.h, N, _dp(x), _dp(u), float(dt), int(n_steps), _dp(out), _dp(st), int(n_workers))
        return out, st, secs

    def manip_state_rate(self, x_blocked, u=None):
        """kte::manipulator_dynamics_model::computeStateRate (ctrl/kte_models/manip_dynamics_model.cpp:152-218) on
        blocked states (q..., qd...): returns (xdot_blocked, status).  Reference checker only."""
        if self._prefix != "rkref_":
            raise NotImplementedError("the legacy model is the compiled reference's")
        x, u, N = self._xu(x_blocked, u)
        xd, st = np.empty_like(x), np.zeros(N, dtype=np.int32)
        fn = self.lib.rkref_manip_state_rate
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_size_t] + [C.c_void_p] * 4
        if fn(self.h, N, _dp(x), _dp(u), _dp(xd), _dp(st)) != 0:
            raise RuntimeError("rkref_manip_state_rate: the legacy model could not be assembled")
        return xd, st

    def rk4_inputs(self, x0, u_nodes, dt):
        """RK4 with an input trajectory sampled at every half step, u_nodes [N][2 n_steps + 1][nu]
        (ctrl::detail::runge_kutta4_integrate_impl).  Returns (x_out, status)."""
        x = np.ascontiguousarray(x0, dtype=np.float64).reshape(-1, self.nx)
        N = x.shape[0]
        n_steps = (np.shape(u_nodes)[1] - 1) // 2
        u_nodes = np.ascontiguousarray(u_nodes, dtype=np.float64).reshape(N, 2 * n_steps + 1, self.nu) if self.nu else np.zeros((N, 1, 0))
        out, st = np.empty_like(x), np.zeros(N, dtype=np.int32)
        fn = getattr(self.lib, self._prefix + "rk4_inputs")
        fn.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        fn(self.h, N, _dp(x), _dp(u_nodes), float(dt), int(n_steps), _dp(out), _dp(st))
        return out, st

    def rk4_inputs_concept(self, x0, u_nodes, dt):
        """The same through ctrl::detail::runge_kutta4_integrate_impl ITSELF (runge_kutta4_integrator_sys.hpp:50-97, compiled in
        oracle/ref_steer_law.cpp) over the live kte_nl_system.  Reference checker only."""
        if self._prefix != "rkref_" or not hasattr(self.lib, "rkref_rk4_inputs_concept"):
            raise NotImplementedError("needs oracle/_ref/libreak_ref.so with ref_steer_law.cpp")
        x = np.ascontiguousarray(x0, dtype=np.float64).reshape(-1, self.nx)
        N = x.shape[0]
        n_steps = (np.shape(u_nodes)[1] - 1) // 2
        u_nodes = np.ascontiguousarray(u_nodes, dtype=np.float64).reshape(N, 2 * n_steps + 1, self.nu) if self.nu else np.zeros((N, 1, 0))
        out, st = np.empty_like(x), np.zeros(N, dtype=np.int32)
        self.lib.rkref_kte_nl_system.restype = C.c_void_p
        self.lib.rkref_kte_nl_system.argtypes = [C.c_void_p]
        fn = self.lib.rkref_rk4_inputs_concept
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        fn(self.lib.rkref_kte_nl_system(self.h), self.nx, self.nu, N, _dp(x), _dp(u_nodes), float(dt), int(n_steps), _dp(out), _dp(st))
        return out, st

    def ihaqr_move_toward(self, x0, goal, u_bias, gain, u_prev, T, horizon, proximity, bounds, bandwidth):
        """IHAQR_topology::move_position_toward_impl ITSELF (examples/misc/IHAQR_topology.hpp:337-381, no collision check,
        fraction 1) over the live kte_nl_system, the points' linearisation / gain payloads filled in from the arguments
        (oracle/ref_steer_law.cpp).  Integrates each interval with steps of T * 1e-2, time-driven, as the reference does.
        Returns the state the loop ended on.  Reference checker only."""
        if self._prefix != "rkref_" or not hasattr(self.lib, "rkref_ihaqr_move_toward"):
            raise NotImplementedError("needs oracle/_ref/libreak_ref.so with ref_steer_law.cpp")
        x = np.ascontiguousarray(x0, dtype=np.float64).reshape(-1, self.nx)
        N = x.shape[0]
        goal = np.ascontiguousarray(goal, dtype=np.float64).reshape(N, self.nx)
        u_bias = np.ascontiguousarray(u_bias, dtype=np.float64).reshape(N, self.nu)
        gain = np.ascontiguousarray(gain, dtype=np.float64).reshape(N, self.nu, self.nx)
        up = np.ascontiguousarray(u_prev, dtype=np.float64).reshape(N, self.nu)
        lo, hi = (np.ascontiguousarray(b, dtype=np.float64).reshape(self.nu) for b in bounds)
        bw = np.ascontiguousarray(bandwidth, dtype=np.float64).reshape(self.nu)
        out = np.empty_like(x)
        self.lib.rkref_kte_nl_system.restype = C.c_void_p
        self.lib.rkref_kte_nl_system.argtypes = [C.c_void_p]
        fn = self.lib.rkref_ihaqr_move_toward
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_size_t] + [C.c_void_p] * 6
        if fn(self.lib.rkref_kte_nl_system(self.h), self.nx, self.nu, _dp(lo), _dp(hi), _dp(bw), float(T), float(horizon), float(proximity), N,
              _dp(x), _dp(goal), _dp(up), _dp(u_bias), _dp(gain), _dp(out)) != 0:
            raise RuntimeError("rkref_ihaqr_move_toward failed")
        return out

    def meaqr_steer(self, x0, goal, u_bias, gain, u_prev, T, time_limit, proximity, bounds, bandwidth, is_free=None):
        """MEAQR_topology::steer_with_constant_control ITSELF (examples/misc/MEAQR_topology.hpp:503-561, no collision check) over
        the live kte_nl_system, with H = I and eta = 0 so that u0 = u_bias and K = gain (oracle/ref_steer_law.cpp).  Each
        interval is integrated with steps of T * 1e-1, time-driven.  is_free: a function of one state (numpy [nx]) the loop's
        virtual is_free_impl is answered by (with_collision_check = true); then `collided` is appended.
        Returns (x_end, u_last, time reached[, collided])."""
        if self._prefix != "rkref_" or not hasattr(self.lib, "rkref_meaqr_steer"):
            raise NotImplementedError("needs oracle/_ref/libreak_ref.so with ref_steer_law.cpp")
        x = np.ascontiguousarray(x0, dtype=np.float64).reshape(-1, self.nx)
        N = x.shape[0]
        goal = np.ascontiguousarray(goal, dtype=np.float64).reshape(N, self.nx)
        u_bias = np.ascontiguousarray(u_bias, dtype=np.float64).reshape(N, self.nu)
        gain = np.ascontiguousarray(gain, dtype=np.float64).reshape(N, self.nu, self.nx)
        up = np.ascontiguousarray(u_prev, dtype=np.float64).reshape(N, self.nu)
        lo, hi = (np.ascontiguousarray(b, dtype=np.float64).reshape(self.nu) for b in bounds)
        bw = np.ascontiguousarray(bandwidth, dtype=np.float64).reshape(self.nu)
        out, uo, to = np.empty_like(x), np.empty((N, self.nu)), np.zeros(N)
        self.lib.rkref_kte_nl_system.restype = C.c_void_p
        self.lib.rkref_kte_nl_system.argtypes = [C.c_void_p]
        fn = self.lib.rkref_meaqr_steer
        fn.restype = C.c_int
        CB = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_double), C.c_int, C.c_void_p)
        fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_size_t] + [C.c_void_p] * 8 + [CB, C.c_void_p, C.c_void_p]
        col = np.zeros(N, dtype=np.int32)
        cb = CB(lambda px, n, ctx: int(bool(is_free(np.array([px[k] for k in range(n)]))))) if is_free is not None else C.cast(None, CB)
        if fn(self.lib.rkref_kte_nl_system(self.h), self.nx, self.nu, _dp(lo), _dp(hi), _dp(bw), float(T), float(time_limit), float(proximity), N,
              _dp(x), _dp(goal), _dp(up), _dp(u_bias), _dp(gain), _dp(out), _dp(uo), _dp(to), cb, None, _dp(col)) != 0:
            raise RuntimeError("rkref_meaqr_steer failed")
        return (out, uo, to, col) if is_free is not None else (out, uo, to)

    def integrate(self, x0, u, scheme, dt, n_steps, n_workers=1):
        """n_steps of euler (1) / midpoint (2) / runge_kutta4 (4) / runge_kutta5 (5); returns (x_out, status, seconds)."""
        x, u, N = self._xu(x0, u)
        out, st = np.empty_like(x), np.zeros(N, dtype=np.int32)
        secs = self._integrate(self.h, N, _dp(x), _dp(u), int(scheme), float(dt), int(n_steps), _dp(out), _dp(st), int(n_workers))
        return out, st, secs

    def rollout(self, x0, u_seq, scheme, dt, steps_per_interval, n_workers=1):
        """The planner-side loop: one get_next_state per control interval (num_int_dtnl_system.hpp:166-180).
        u_seq: [N][n_intervals][nu].  Returns (x_out, x_traj [N][n_intervals][nx], status)."""
        x = np.ascontiguousarray(x0, dtype=np.float64).reshape(-1, self.nx)
        N = x.shape[0]
        u_seq = np.asarray(u_seq, dtype=np.float64)
        J = u_seq.shape[1]
        u_seq = u_seq.reshape(N, J, self.nu)
        traj = np.empty((N, J, self.nx))
        st = np.zeros(N, dtype=np.int32)
        for j in range(J):
            x, sj, _ = self.integrate(x, np.ascontiguousarray(u_seq[:, j, :]), scheme, dt, steps_per_interval, n_workers)
            traj[:, j, :] = x
            st |= sj
        return x, traj, st

    def steer_feedback(self, x0, goal, u_bias, gain, u_prev, T, dt, substeps, max_intervals, proximity,
                       saturate_first=False, bounds=None, rate_bounds=None, proxy_pairs=None):
        """The steering loop of MEAQR_topology.hpp:503-561 (oracle/steer_law.h).  bounds / rate_bounds: (lo, hi) or None.
        Returns (x_out, u_last, n_done, traj [N][max_intervals][nx] (NaN where not written), status); with
        proxy_pairs (reference checker only) the loop runs with its collision test on and `collided` is appended."""
        x = np.ascontiguousarray(x0, dtype=np.float64).reshape(-1, self.nx)
        N = x.shape[0]
        goal = np.ascontiguousarray(goal, dtype=np.float64).reshape(N, self.nx)
        u_bias = np.ascontiguousarray(u_bias, dtype=np.float64).reshape(N, self.nu)
        gain = np.ascontiguousarray(gain, dtype=np.float64).reshape(N, self.nu, self.nx)
        up = np.array(u_prev, dtype=np.float64).reshape(N, self.nu).copy()
        arr = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64).reshape(self.nu)
        lo, hi = (arr(bounds[0]), arr(bounds[1])) if bounds is not None else (None, None)
        dlo, dhi = (arr(rate_bounds[0]), arr(rate_bounds[1])) if rate_bounds is not None else (None, None)
        xo, nd = np.empty_like(x), np.zeros(N, dtype=np.int32)
        traj = np.full((N, max(int(max_intervals), 1), self.nx), np.nan)
        st = np.zeros(N, dtype=np.int32)
        if proxy_pairs:
            if self._prefix != "rkref_":
                raise NotImplementedError("the collision test is checked against the compiled reference only")
            P = len(proxy_pairs)
            keep = [(p.model1.to_c(self.compiled.frames), p.model2.to_c(self.compiled.frames)) for p in proxy_pairs]
            m1s = (C.c_void_p * P)(*[C.cast(k[0][0], C.c_void_p) for k in keep])
            m2s = (C.c_void_p * P)(*[C.cast(k[1][0], C.c_void_p) for k in keep])
            n1s = (C.c_int * P)(*[k[0][1] for k in keep])
            n2s = (C.c_int * P)(*[k[1][1] for k in keep])
            col = np.zeros(N, dtype=np.int32)
            fn = self.lib.rkref_steer_feedback_checked
            fn.restype = C.c_int
            fn.argtypes = ([C.c_void_p, C.c_size_t] + [C.c_void_p] * 5 + [C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, C.c_int]
                           + [C.c_void_p] * 4 + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 5)
            rc = fn(self.h, N, _dp(x), _dp(goal), _dp(u_bias), _dp(gain), _dp(up), float(T), float(dt), int(substeps), int(max_intervals),
                    float(proximity), int(bool(saturate_first)), _dp(lo), _dp(hi), _dp(dlo), _dp(dhi), m1s, n1s, m2s, n2s, P,
                    _dp(xo), _dp(nd), _dp(col), _dp(traj), _dp(st))
            if rc != 0:
                raise RuntimeError("steer_feedback_checked failed")
            return xo, up, nd, traj[:, :int(max_intervals), :], st, col
        fn = getattr(self.lib, self._prefix + "steer_feedback")
        fn.restype = C.c_int
        fn.argtypes = ([C.c_void_p, C.c_size_t] + [C.c_void_p] * 5 + [C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, C.c_int]
                       + [C.c_void_p] * 8)
        rc = fn(self.h, N, _dp(x), _dp(goal), _dp(u_bias), _dp(gain), _dp(up), float(T), float(dt), int(substeps), int(max_intervals),
                float(proximity), int(bool(saturate_first)), _dp(lo), _dp(hi), _dp(dlo), _dp(dhi), _dp(xo), _dp(nd), _dp(traj), _dp(st))
        if rc != 0:
            raise RuntimeError("steer_feedback failed")
        return xo, up, nd, traj[:, :int(max_intervals), :], st

    def tmt(self, x):
        """(Tcm, Mcm, Tcm_dot) of mass_matrix_calc::get_TMT_TdMT for one state."""
        x, _, _ = self._xu(x, None)
        fn = getattr(self.lib, self._prefix + "tmt")
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p] * 5
        x0 = x[0].copy()
        rows = fn(self.h, _dp(x0), None, None, None)
        T, Mc, Td = np.zeros((rows, self.na)), np.zeros((rows, rows)), np.zeros((rows, self.na))
        fn(self.h, _dp(x0), _dp(T), _dp(Mc), _dp(Td))
        return T, Mc, Td

    def min_distance(self, pair, x):
        """proxy_query_pair_3D::findMinimumDistance of the live reference after doMotion at every state:
        (distance [N], finder index [N], points [N][6]).  Reference checker only."""
        if self._prefix != "rkref_":
            raise NotImplementedError("proximity is checked against the compiled reference only")
        x, _, N = self._xu(x, None)
        m1, n1 = pair.model1.to_c(self.compiled.frames)
        m2, n2 = pair.model2.to_c(self.compiled.frames)
        d = np.zeros(N)
        f = np.zeros(N, dtype=np.int32)
        pts = np.zeros((N, 6))
        fn = self.lib.rkref_min_distance
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        xc = np.ascontiguousarray(x)
        rc = fn(self.h, N, _dp(xc), C.cast(m1, C.c_void_p), n1, C.cast(m2, C.c_void_p), n2, _dp(d), _dp(f), _dp(pts))
        if rc < 0:
            raise RuntimeError("rkref_min_distance failed")
        return d, f, pts

    def collision_points(self, pair, x, max_records):
        """proxy_query_pair_3D::gatherCollisionPoints of the live reference after doMotion at every state:
        (count [N], records [N][max_records][7]).  Reference checker only."""
        if self._prefix != "rkref_":
            raise NotImplementedError("proximity is checked against the compiled reference only")
        x, _, N = self._xu(x, None)
        m1, n1 = pair.model1.to_c(self.compiled.frames)
        m2, n2 = pair.model2.to_c(self.compiled.frames)
        cnt = np.zeros(N, dtype=np.int32)
        rec = np.zeros((N, max_records, 7))
        fn = self.lib.rkref_collision_points
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        xc = np.ascontiguousarray(x)
        if fn(self.h, N, _dp(xc), C.cast(m1, C.c_void_p), n1, C.cast(m2, C.c_void_p), n2, int(max_records), _dp(cnt), None, _dp(rec)) != 0:
            raise RuntimeError("rkref_collision_points failed")
        return cnt, rec

    def bridge_proxy(self, model):
        """include/reak_b200/reak_bridge.hpp's compile_proxy_model on live geom:: shapes built from `model`:
        returns the rkb_shape array it derives and, per shape, the anchor translated back to this descriptor's
        frame numbering."""
        arr, n = model.to_c(self.compiled.frames)
        out = type(arr)()
        anchors = (C.c_int * max(n, 1))()
        err = C.create_string_buffer(256)
        fn = self.lib.rkref_bridge_proxy
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]
        rc = fn(self.h, C.cast(arr, C.c_void_p), n, C.cast(out, C.c_void_p), C.cast(anchors, C.c_void_p), err, 256)
        if rc < 0:
            raise RuntimeError("bridge rejected the model: " + err.value.decode())
        return arr, out, list(anchors)[:n]

    def frames(self, x, u=None):
        """[n_frames][25]: Position3 Quat4 Velocity3 AngVelocity3 Acceleration3 AngAcceleration3 Force3 Torque3."""
        x, u, _ = self._xu(x, u)
        out = np.zeros((self.compiled.desc.n_frames, 25))
        self._frames(self.h, _dp(x[0].copy()), _dp(u[0].copy()), _dp(out))
        return out


def nearest(which, vertices, queries, k=1, radius=np.inf):
    """k nearest vertices per query by the reference's linear scan: which = "oracle" (kto_nearest, the C restatement) or
    "ref" (rkref_nearest: ReaK::pp::min_dist_linear_search itself).  Returns (index [Q][k], distance [Q][k], count [Q])."""
    lib = C.CDLL(ORACLE_SO if which == "oracle" else REF_SO)
    fn = lib.kto_nearest if which == "oracle" else lib.rkref_nearest
    fn.argtypes = [C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    v = np.ascontiguousarray(vertices, dtype=np.float64)
    q = np.ascontiguousarray(queries, dtype=np.float64)
    dim = q.shape[1]
    v = v.reshape(-1, dim)
    idx = np.empty((q.shape[0], k), dtype=np.int32)
    dist = np.empty((q.shape[0], k))
    cnt = np.empty(q.shape[0], dtype=np.int32)
    fn(v.shape[0], _dp(v), q.shape[0], _dp(q), dim, int(k), float(radius), _dp(idx), _dp(dist), _dp(cnt))
    return idx, dist, cnt


_PRODUCT_SO = os.path.join(os.path.dirname(_HERE), "reak_b200", "lib", "libreak_b200.so")
_preloaded = []


def preload_product():
    """For the tests that drive the C++ drop-in checks (rkref_bridge_gpu_check, rkref_steer_space_check, ...):
    those call INTO the product library from code compiled against the reference, through weak references that
    are bound when libreak_ref.so is mapped.  The product's symbols must be visible before that, so the test
    session (tests/conftest.py) calls this once, before any Reference exists.  Nothing else does: in
    particular `bench.py --impl reference` and the cpu_baseline leg never map the product library through here."""
    path = os.environ.get("RKB_LIB_PATH") or _PRODUCT_SO
    if not _preloaded and os.path.isfile(path):
        _preloaded.append(C.CDLL(path, mode=C.RTLD_GLOBAL))
    return bool(_preloaded)


class Reference(_Checker):
    """The real ReaK code (kte_nl_system + runge_kutta4_integrator)."""

    def __init__(self, compiled):
        _Checker.__init__(self, REF_SO, "rkref_", compiled)


class Oracle(_Checker):
    """The plain-C restatement oracle/kte_oracle.c."""

    def __init__(self, compiled):
        _Checker.__init__(self, ORACLE_SO, "kto_", compiled)
        self.lib.kto_gen_forces_qdd.argtypes = [C.c_void_p] * 5

    def gen_forces_qdd(self, x, u, qdd):
        """gen_coord::f of one state with a caller-chosen q_ddot (test_bm.cpp:103-121)."""
        x, u, _ = self._xu(x, u)
        qdd = np.ascontiguousarray(qdd, dtype=np.float64).reshape(self.n)
        f = np.empty(self.n)
        self.lib.kto_gen_forces_qdd(self.h, _dp(x[0].copy()), _dp(u[0].copy()), _dp(qdd), _dp(f))
        return f

def cholesky_solve(A, b, tol=1e-8, ldl=False):
    """linsolve_Cholesky / the LDL variant of core/lin_alg/mat_cholesky.hpp on a dense system; returns (x, singular)."""
    lib = C.CDLL(ORACLE_SO)
    A = np.ascontiguousarray(A, dtype=np.float64)
    b = np.array(b, dtype=np.float64)
    b2 = np.ascontiguousarray(b.reshape(A.shape[0], -1)).copy()
    fn = lib.kto_ldl_solve if ldl else lib.kto_cholesky_solve
    fn.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double]
    rc = fn(A.shape[0], _dp(A), _dp(b2), b2.shape[1], float(tol))
    return b2.reshape(b.shape), rc

"""kte_batch_propagator — the batched counterpart of ctrl::kte_nl_system + num_int_dtnl_sys.

It is constructed from the same members kte_nl_system holds (chain, mass_calc, dofs_gen, inputs;
ctrl/ctrl_sys/kte_nl_system.hpp:70-78).  The single-sample methods keep the reference's
state-space-system signatures (SSSystemConcept, ctrl/ctrl_sys/state_space_sys_concept.hpp:111-137;
DiscreteSSSConcept, ctrl/ctrl_sys/discrete_sss_concept.hpp:40-141) and raise the reference's
errors; the batched methods are additive.  All numerics run in libreak_b200.so on the GPU —
this module only moves pointers.  numpy arrays are treated as HOST buffers (staged by the
library), torch CUDA tensors as DEVICE buffers (zero-copy).
"""
import ctypes as C

import numpy as np

from . import _abi, kte


class singularity_error(ArithmeticError):
    """core/lin_alg/mat_num_exceptions.hpp:45-56 — thrown by linsolve_Cholesky on a pivot < 1e-8."""


class impossible_integration(ValueError):
    """core/integrators/integration_exceptions.hpp:38."""


def _is_torch(a):
    return type(a).__module__.startswith("torch")


class kte_batch_propagator(object):
    def __init__(self, chain, mass_calc=None, dofs_gen=None, inputs=None, device=0, time_step=1e-3, blocked=False,
                 interpreter=False, general=False, dofs_3D=None):
        if isinstance(chain, kte.compiled_chain):  # a ready descriptor (from_rkx)
            self.chain = self.mass_calc = None
            self.dofs_gen, self.inputs, self.dofs_3D = [], [], [None] * chain.n_free
            self._create(chain, device, time_step, blocked, interpreter, general)
            return
        if mass_calc is None and hasattr(chain, "chain"):  # a kte_system / kte_nl_system-like object
            sys_ = chain
            chain, mass_calc, dofs_gen, inputs = sys_.chain, sys_.mass_calc, sys_.dofs_gen, sys_.inputs
            dofs_3D = getattr(sys_, "dofs_3D", ()) if dofs_3D is None else dofs_3D
        self.chain, self.mass_calc, self.dofs_gen, self.inputs = chain, mass_calc, list(dofs_gen), list(inputs)
        # coordinate frames of free_joint_3D elements (kte_nl_system::dofs_3D): 13 states and 6 accelerations each,
        # after the generalized coordinates' (kte_nl_system.hpp:145-147)
        self.dofs_3D = list(dofs_3D or ())
        self._create(kte.compile_chain(chain, mass_calc, self.dofs_gen, self.inputs, self.dofs_3D), device, time_step, blocked, interpreter, general)

    @classmethod
    def from_rkx(cls, path, **kw):
        """A propagator for the kte_nl_system stored in a ReaK XML archive (`.rkx`, core/serialization/xml_archiver.cpp) —
        read by the library itself (rkb_rkx_read); no ReaK objects are involved."""
        return cls(kte.read_rkx(path), **kw)

    def _create(self, compiled, device, time_step, blocked, interpreter, general):
        self.compiled = compiled
        self.device = int(device)
        self.dt = float(time_step)
        # True: every state-shaped buffer is (q..., qd...) like manipulator_dynamics_model::computeStateRate
        # (ctrl/mbd_kte/manipulator_model.cpp:292-355) instead of kte_nl_system's interleaved (q, qd) pairs
        self.blocked = bool(blocked)
        self._lib = _abi.load_library()
        h = C.c_void_p()
        # interpreter / general: which kernel family runs the chain (rkb_chain_create_ex); same results up to rounding
        cf = (_abi.CREATE_INTERPRETER if interpreter else 0) | (_abi.CREATE_GENERAL if general else 0)
        _abi.check(self._lib.rkb_chain_create_ex(C.byref(self.compiled.desc), cf, C.byref(h)), "rkb_chain_create_ex")
        self._h = h
        self.n = self._lib.rkb_chain_dof(h)
        self.nx = self._lib.rkb_chain_state_dim(h)
        self.na = self.compiled.n_acc  # accelerations: rows / columns of M, entries of f, columns of Tcm
        self.nu = self._lib.rkb_chain_input_dim(h)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rkb_chain_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, option, value):
        """rkb_chain_set_option: 'split_max_samples', 'fused_steer', 'fused_sequence', 'host_pipeline' (include/reak_b200.h)"""
        code = {"split_max_samples": _abi.OPT_SPLIT_MAX_SAMPLES, "fused_steer": _abi.OPT_FUSED_STEER,
                "fused_sequence": _abi.OPT_FUSED_SEQUENCE, "host_pipeline": _abi.OPT_HOST_PIPELINE,
                "auto_specialize": _abi.OPT_AUTO_SPECIALIZE}[option]
        _abi.check(self._lib.rkb_chain_set_option(self._h, code, int(value)), "rkb_chain_set_option")
        return self

    def get_option(self, option):
        code = {"split_max_samples": _abi.OPT_SPLIT_MAX_SAMPLES, "fused_steer": _abi.OPT_FUSED_STEER,
                "fused_sequence": _abi.OPT_FUSED_SEQUENCE, "host_pipeline": _abi.OPT_HOST_PIPELINE,
                "auto_specialize": _abi.OPT_AUTO_SPECIALIZE}[option]
        return int(self._lib.rkb_chain_get_option(self._h, code))

    def wave_samples(self):
        """samples in one full wave of the RK4 rollout kernel on this device (rkb_chain_wave_samples); 0 = unknown"""
        return max(0, int(self._lib.rkb_chain_wave_samples(self._h, self.device)))

    def is_serial(self):
        """True when the chain runs on the register-resident serial-chain kernels."""
        return bool(self._lib.rkb_chain_is_serial(self._h))

    def specialize(self):
        """Compile the serial kernels for exactly this chain's structure (NVRTC) and use them from now on."""
        _abi.check(self._lib.rkb_chain_specialize(self._h, self.device), "rkb_chain_specialize")
        return self

    def is_specialized(self):
        return bool(self._lib.rkb_chain_is_specialized(self._h))

    def kernel_shape(self):
        return int(self._lib.rkb_chain_kernel_shape(self._h))

    # ---- SSSystemConcept / DiscreteSSSConcept (single sample) ---------------------------
    def get_state_dimensions(self):
        return self.nx

    def get_input_dimensions(self):
        return self.nu

    def get_output_dimensions(self):
        return 0

    def get_time_step(self):
        return self.dt

    def set_time_step(self, dt):
        self.dt = float(dt)

    def get_state_derivative(self, space, p, u, t=0.0):
        """kte_nl_system::get_state_derivative (kte_nl_system.hpp:238-346) for one state."""
        p = np.asarray(p, dtype=np.float64).ravel()
        u = np.asarray(u, dtype=np.float64).ravel()
        if p.size != self.nx:
            raise IndexError("State vector dimension mismatch!")  # std::range_error, :181-185
        if u.size != self.nu:
            raise IndexError("Input vector dimension mismatch!")  # std::range_error, :186-187
        xd, st = self.get_state_derivatives(p[None, :], u[None, :])
        if st[0] & _abi.STATUS_SINGULAR:
            raise singularity_error("Cholesky pivot below 1e-8")
        return xd[0]

    def get_next_state(self, space, p, u, t=0.0):
        """num_int_dtnl_sys::get_next_state (num_int_dtnl_system.hpp:166-180): one RK4 step of get_time_step()."""
        p = np.asarray(p, dtype=np.float64).ravel()
        u = np.asarray(u, dtype=np.float64).ravel()
        if p.size != self.nx:
            raise IndexError("State vector dimension mismatch!")
        if u.size != self.nu:
            raise IndexError("Input vector dimension mismatch!")
        x, st = self.get_next_states(p[None, :], u[None, :], self.dt, 1)
        if st[0] & _abi.STATUS_SINGULAR:
            raise singularity_error("Cholesky pivot below 1e-8")
        return x[0]

    # ---- buffer plumbing --------------------------------------------------------------------
    def _prep(self, arrs, soa):
        """Returns (flags, stream, n_samples-agnostic pointers); all buffers must live in one space."""
        kinds = set(_is_torch(a) for a in arrs if a is not None)
        if len(kinds) != 1:
            raise TypeError("mix of numpy (host) and torch (device) buffers")
        flags = (_abi.LAYOUT_SOA if soa else _abi.LAYOUT_AOS) | (_abi.LAYOUT_BLOCKED if self.blocked else 0)
        if kinds.pop():
            import torch
            for a in arrs:
                if a is None:
                    continue
                if not a.is_cuda or a.device.index != self.device:
                    raise TypeError("torch buffers must be CUDA tensors on device %d" % self.device)
                if not a.is_contiguous():
                    raise TypeError("buffers must be contiguous")
            flags |= _abi.MEM_DEVICE
            stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            ptr = lambda a: C.c_void_p(a.data_ptr()) if a is not None else None
        else:
            stream = None
            ptr = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        return flags, stream, ptr

    def _in(self, a, cols, dtype, soa=False, rows=None):
        if _is_torch(a):
            import torch
            want = torch.float64 if dtype == np.float64 else torch.int32
            if a.dtype != want:
                raise TypeError("expected %s tensor" % want)
            a = a.contiguous()
            shape = tuple(a.shape)
        else:
            a = np.ascontiguousarray(a, dtype=dtype)
            shape = a.shape
        if len(shape) != 2:
            raise IndexError("expected a 2-D batch buffer")
        n, c = (shape[1], shape[0]) if soa else shape
        if c != cols or (rows is not None and n != rows):
            raise IndexError("State vector dimension mismatch!" if cols == self.nx else "Input vector dimension mismatch!")
        return a, n

    def _out(self, buf, ref, shape, dtype=np.float64):
        """A caller-supplied result buffer goes to the C-ABI by raw pointer: it must have exactly the shape and
        dtype the call writes, be C-contiguous and live in the same memory space as the inputs; None allocates."""
        if buf is None:
            return self._like(ref, shape, dtype)
        if _is_torch(buf) != _is_torch(ref):
            raise TypeError("mix of numpy (host) and torch (device) buffers")
        if _is_torch(buf):
            import torch
            want = torch.float64 if dtype == np.float64 else torch.int32
            ok_type, contiguous = buf.dtype == want, buf.is_contiguous()
        else:
            if not isinstance(buf, np.ndarray):
                raise TypeError("result buffers must be numpy arrays or torch tensors")
            ok_type, contiguous = buf.dtype == np.dtype(dtype), buf.flags["C_CONTIGUOUS"] and buf.flags["WRITEABLE"]
        if not ok_type:
            raise TypeError("result buffer must be %s" % np.dtype(dtype).name)
        if tuple(buf.shape) != tuple(shape):
            raise IndexError("result buffer has shape %s, the call writes %s" % (tuple(buf.shape), tuple(shape)))
        if not contiguous:
            raise TypeError("result buffers must be C-contiguous and writeable")
        return buf

    def _like(self, ref, shape, dtype=np.float64):
        if _is_torch(ref):
            import torch
            return torch.empty(shape, dtype=torch.float64 if dtype == np.float64 else torch.int32, device=ref.device)
        return np.empty(shape, dtype=dtype)

    def _u_default(self, x, N, soa):
        if self.nu:
            raise IndexError("Input vector dimension mismatch!")
        return self._like(x, (0, N) if soa else (N, 0))

    # ---- batched API ------------------------------------------------------------------------
    def get_state_derivatives(self, x, u=None, soa=False, out=None, status=None):
        x, N = self._in(x, self.nx, np.float64, soa)
        u = self._u_default(x, N, soa) if u is None else self._in(u, self.nu, np.float64, soa, N)[0]
        xd = self._out(out, x, x.shape)
        st = self._out(status, x, (N,), np.int32)
        flags, stream, ptr = self._prep([x, u if self.nu else None, xd, st], soa)
        _abi.check(self._lib.rkb_eval(self._h, self.device, N, ptr(x), ptr(u) if self.nu else None,
                                      ptr(xd), ptr(st), flags, stream), "rkb_eval")
        return xd, st

    def get_next_states(self, x, u=None, dt=None, n_steps=1, soa=False, out=None, status=None):
        x, N = self._in(x, self.nx, np.float64, soa)
        u = self._u_default(x, N, soa) if u is None else self._in(u, self.nu, np.float64, soa, N)[0]
        dt = self.dt if dt is None else float(dt)
        if dt == 0.0 or n_steps < 0:
            raise impossible_integration("dt == 0 or negative step count")
        xo = self._out(out, x, x.shape)
        st = self._out(status, x, (N,), np.int32)
        flags, stream, ptr = self._prep([x, u if self.nu else None, xo, st], soa)
        _abi.check(self._lib.rkb_rollout_rk4(self._h, self.device, N, ptr(x), ptr(u) if self.nu else None,
                                             dt, int(n_steps), ptr(xo), ptr(st), flags, stream), "rkb_rollout_rk4")
        return xo, st

    def get_next_states_input_trajectory(self, x, u_nodes, dt=None, out=None, status=None):
        """RK4 with an input trajectory (ctrl::detail::runge_kutta4_integrate_impl, runge_kutta4_integrator_sys.hpp:50-97):
        u_nodes [N][2 n_steps + 1][nu] is the trajectory sampled at every half step.  Returns (x_out, status)."""
        x, N = self._in(x, self.nx, np.float64)
        u_nodes = u_nodes.contiguous() if _is_torch(u_nodes) else np.ascontiguousarray(u_nodes, dtype=np.float64)
        if len(u_nodes.shape) != 3 or u_nodes.shape[0] != N or u_nodes.shape[2] != self.nu or u_nodes.shape[1] < 1 or u_nodes.shape[1] % 2 != 1:
            raise IndexError("Input vector dimension mismatch!")
        n_steps = (int(u_nodes.shape[1]) - 1) // 2
        dt = self.dt if dt is None else float(dt)
        if dt == 0.0:
            raise impossible_integration("dt == 0")
        xo = self._out(out, x, x.shape)
        st = self._out(status, x, (N,), np.int32)
        flags, stream, ptr = self._prep([x, u_nodes if self.nu else None, xo, st], False)
        _abi.check(self._lib.rkb_rollout_rk4_inputs(self._h, self.device, N, ptr(x), ptr(u_nodes) if self.nu else None, dt, n_steps,
                                                    ptr(xo), ptr(st), flags, stream), "rkb_rollout_rk4_inputs")
        return xo, st

    def rollout(self, x, u_seq, dt=None, steps_per_interval=1, scheme="rk4", want_traj=False, out=None, status=None):
        """n_intervals = u_seq.shape[1] control intervals, the input constant within each (one
        num_int_dtnl_sys::get_next_state per interval), steps_per_interval steps of `scheme`
        (euler / midpoint / rk4 / rk5, core/integrators/fixed_step_integrators.hpp) per interval.
        x: [N][nx]; u_seq: [N][n_intervals][nu].  Returns (x_out, status) or, with want_traj,
        (x_out, x_traj [N][n_intervals][nx], status)."""
        x, N = self._in(x, self.nx, np.float64)
        if _is_torch(u_seq):
            u_seq = u_seq.contiguous()
        else:
            u_seq = np.ascontiguousarray(u_seq, dtype=np.float64)
        if len(u_seq.shape) != 3 or u_seq.shape[0] != N or u_seq.shape[2] != self.nu or u_seq.shape[1] < 1:
            raise IndexError("Input vector dimension mismatch!")
        J = int(u_seq.shape[1])
        dt = self.dt if dt is None else float(dt)
        if dt == 0.0 or steps_per_interval < 0:
            raise impossible_integration("dt == 0 or negative step count")
        code = _abi.SCHEMES[scheme] if isinstance(scheme, str) else int(scheme)
        opts = _abi.rkb_rollout_opts(code, J, int(steps_per_interval), 0, dt)
        xo = self._out(out, x, x.shape)
        tr = self._like(x, (N, J, self.nx)) if want_traj else None
        st = self._out(status, x, (N,), np.int32)
        flags, stream, ptr = self._prep([x, u_seq if self.nu else None, xo, tr, st], False)
        _abi.check(self._lib.rkb_rollout(self._h, self.device, N, ptr(x), ptr(u_seq) if self.nu else None, C.byref(opts),
                                         ptr(xo), ptr(tr), ptr(st), flags, stream), "rkb_rollout")
        return (xo, tr, st) if want_traj else (xo, st)

    def get_next_states_multi(self, x, u=None, dt=None, n_steps=1, devices=(0,), out=None, status=None):
        """Host (numpy, AoS) buffers sharded over several GPUs from this one process (rkb_rollout_rk4_multi)."""
        x, N = self._in(x, self.nx, np.float64)
        u = self._u_default(x, N, False) if u is None else self._in(u, self.nu, np.float64, False, N)[0]
        if _is_torch(x) or _is_torch(u):
            raise TypeError("get_next_states_multi takes host (numpy) buffers")
        dt = self.dt if dt is None else float(dt)
        if dt == 0.0 or n_steps < 0:
            raise impossible_integration("dt == 0 or negative step count")
        xo = self._out(out, x, x.shape)
        st = self._out(status, x, (N,), np.int32)
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        _abi.check(self._lib.rkb_rollout_rk4_multi(self._h, len(devices), devs, N, p(x), p(u) if self.nu else None, dt, int(n_steps),
                                                   p(xo), p(st)), "rkb_rollout_rk4_multi")
        return xo, st

    def get_gen_forces(self, x, u=None, soa=False):
        x, N = self._in(x, self.nx, np.float64, soa)
        u = self._u_default(x, N, soa) if u is None else self._in(u, self.nu, np.float64, soa, N)[0]
        f = self._like(x, (self.na, N) if soa else (N, self.na))
        flags, stream, ptr = self._prep([x, u if self.nu else None, f], soa)
        _abi.check(self._lib.rkb_gen_forces(self._h, self.device, N, ptr(x), ptr(u) if self.nu else None,
                                            ptr(f), flags, stream), "rkb_gen_forces")
        return f

    def get_mass_matrices(self, x, with_derivative=False, soa=False):
        x, N = self._in(x, self.nx, np.float64, soa)
        shape = (self.na * self.na, N) if soa else (N, self.na, self.na)
        M = self._like(x, shape)
        Md = self._like(x, shape) if with_derivative else None
        flags, stream, ptr = self._prep([x, M, Md], soa)
        _abi.check(self._lib.rkb_mass_matrix(self._h, self.device, N, ptr(x), ptr(M), ptr(Md), flags, stream),
                   "rkb_mass_matrix")
        return (M, Md) if with_derivative else M

    def get_frames(self, x, u=None):
        """Every frame after doMotion / clearForce / doForce: [N][n_frames][25] (layout in include/reak_b200.h)."""
        x, N = self._in(x, self.nx, np.float64)
        u = self._u_default(x, N, False) if u is None else self._in(u, self.nu, np.float64, False, N)[0]
        nf = self._lib.rkb_chain_frame_count(self._h)
        fr = self._like(x, (N, nf, 25))
        flags, stream, ptr = self._prep([x, u if self.nu else None, fr], False)
        _abi.check(self._lib.rkb_frames(self._h, self.device, N, ptr(x), ptr(u) if self.nu else None, ptr(fr), flags, stream), "rkb_frames")
        return fr

    def proxy_handle(self, pair):
        """the rkb_proxy of `pair` for this chain (built once, cached on the pair)"""
        from . import proximity
        h = getattr(pair, "_rkb_handle", None)
        if h is None or getattr(pair, "_rkb_owner", None) is not self:
            h = proximity.ProxyHandle(self._lib, self._h, pair, self.compiled.frames)
            pair._rkb_handle, pair._rkb_owner = h, self
        return h

    def get_min_distances(self, pair, x, with_points=True):
        """proxy_query_pair_3D::findMinimumDistance at every state (rkb_min_distance): distance [N], finder index [N]
        and, with_points, (mPoint1, mPoint2) [N][6].  `pair` is a reak_b200.proximity.proxy_query_pair_3D; its device
        program is built once and cached on the pair."""
        from . import proximity
        h = getattr(pair, "_rkb_handle", None)
        if h is None or getattr(pair, "_rkb_owner", None) is not self:
            h = proximity.ProxyHandle(self._lib, self._h, pair, self.compiled.frames)
            pair._rkb_handle, pair._rkb_owner = h, self
        x, N = self._in(x, self.nx, np.float64)
        d = self._like(x, (N,))
        f = self._like(x, (N,), np.int32)
        pts = self._like(x, (N, 6)) if with_points else None
        flags, stream, ptr = self._prep([x, d, f, pts], False)
        _abi.check(self._lib.rkb_min_distance(self._h, h._h, self.device, N, ptr(x), ptr(d), ptr(f), ptr(pts), flags, stream),
                   "rkb_min_distance")
        return (d, f, pts) if with_points else (d, f)

    def gather_collision_points(self, pair, x, max_records=None):
        """proxy_query_pair_3D::gatherCollisionPoints at every state (rkb_collision_points): count [N] of finders reporting
        a negative distance, their finder indices [N][M] (-1 beyond count) and records [N][M][7] (distance, point 1,
        point 2; +inf beyond count).  M = max_records, default: the pair's finder count (nothing is ever dropped)."""
        from . import proximity
        h = getattr(pair, "_rkb_handle", None)
        if h is None or getattr(pair, "_rkb_owner", None) is not self:
            h = proximity.ProxyHandle(self._lib, self._h, pair, self.compiled.frames)
            pair._rkb_handle, pair._rkb_owner = h, self
        M = int(max_records) if max_records is not None else max(1, int(self._lib.rkb_proxy_finder_count(h._h)))
        x, N = self._in(x, self.nx, np.float64)
        cnt = self._like(x, (N,), np.int32)
        fnd = self._like(x, (N, M), np.int32)
        rec = self._like(x, (N, M, 7))
        flags, stream, ptr = self._prep([x, cnt, fnd, rec], False)
        _abi.check(self._lib.rkb_collision_points(self._h, h._h, self.device, N, ptr(x), M, ptr(cnt), ptr(fnd), ptr(rec), flags, stream),
                   "rkb_collision_points")
        return cnt, fnd, rec

    def is_free(self, pairs, x):
        """manip_dk_proxy_env_impl::is_free (ctrl/topologies/manip_free_workspace.hpp:77-99): no proxy pair reports
        a negative minimum distance.  Returns a bool array [N]."""
        from . import proximity
        hs = []
        for pair in pairs:
            h = getattr(pair, "_rkb_handle", None)
            if h is None or getattr(pair, "_rkb_owner", None) is not self:
                h = proximity.ProxyHandle(self._lib, self._h, pair, self.compiled.frames)
                pair._rkb_handle, pair._rkb_owner = h, self
            hs.append(h._h)
        arr = (C.c_void_p * len(hs))(*hs)
        x, N = self._in(x, self.nx, np.float64)
        free = self._like(x, (N,), np.int32)
        flags, stream, ptr = self._prep([x, free], False)
        _abi.check(self._lib.rkb_is_free(self._h, self.device, N, ptr(x), arr, len(hs), ptr(free), flags, stream), "rkb_is_free")
        return free != 0

    def get_twist_shaping(self, x, with_derivative=True):
        """mass_matrix_calc::get_TMT_TdMT (mass_matrix_calculator.cpp:100-287): Tcm [N][rows][n], Mcm [rows][rows]
        (constant) and, with_derivative, Tcm_dot [N][rows][n]."""
        x, N = self._in(x, self.nx, np.float64)
        rows = self._lib.rkb_twist_shaping_rows(self._h)
        T = self._like(x, (N, rows, self.na))
        Td = self._like(x, (N, rows, self.na)) if with_derivative else None
        Mc = np.zeros((rows, rows))
        _abi.check(self._lib.rkb_twist_shaping_mcm(self._h, Mc.ctypes.data_as(C.c_void_p)), "rkb_twist_shaping_mcm")
        flags, stream, ptr = self._prep([x, T, Td], False)
        _abi.check(self._lib.rkb_twist_shaping(self._h, self.device, N, ptr(x), ptr(T), ptr(Td), flags, stream), "rkb_twist_shaping")
        return (T, Mc, Td) if with_derivative else (T, Mc)

    def get_frame_jacobian(self, x, frame, upstream=None, with_derivative=True, free_joints=()):
        """Jacobian of one frame (a kte frame object of the chain, or its frame id) w.r.t. the coordinates and its time
        derivative: J, Jdot [N][6][n] (3D: v then w, in the frame's own coordinates) or [N][3][n] (2D) — rkb_frame_jacobian,
        manip_kin_mdl_jac_calculator::getJacobianMatrixAndDerivative.  upstream: iterable of coordinate indices (or
        gen_coord objects) that move the frame; default: every coordinate."""
        x, N = self._in(x, self.nx, np.float64)
        fid = frame if isinstance(frame, int) else [id(f) for f in self.compiled.frames].index(id(frame))
        if upstream is None:
            mask = ((1 << self.n) - 1) | (((1 << self.compiled.n_free) - 1) << 32)
        else:
            ids = [id(c) for c in self.compiled.coords]
            mask = 0
            for c in upstream:
                mask |= 1 << (c if isinstance(c, int) else ids.index(id(c)))
            for j in free_joints:  # indices into dofs_3D: six more columns each (jacobian_3D_3D), after the coordinates'
                mask |= 1 << (32 + int(j))
        rows = 6 if self.compiled.desc.dim == 3 else 3
        J = self._like(x, (N, rows, self.na))
        Jd = self._like(x, (N, rows, self.na)) if with_derivative else None
        flags, stream, ptr = self._prep([x, J, Jd], False)
        _abi.check(self._lib.rkb_frame_jacobian(self._h, self.device, N, ptr(x), fid, mask, ptr(J), ptr(Jd), flags, stream), "rkb_frame_jacobian")
        return (J, Jd) if with_derivative else J

    def get_linear_blocks(self, x, u=None, eps=1e-6):
        """A = d xdot / d x [N][nx][nx] and B = d xdot / d u [N][nx][nu] about every (x, u) by central differences
        (rkb_linearize; get_linear_blocks of the LQR steering topologies, examples/misc/IHAQR_topology.hpp:240-258).
        Returns (A, B, status)."""
        if self.blocked:
            raise NotImplementedError("rkb_linearize takes interleaved states")
        x, N = self._in(x, self.nx, np.float64)
        u = self._u_default(x, N, False) if u is None else self._in(u, self.nu, np.float64, False, N)[0]
        A = self._like(x, (N, self.nx, self.nx))
        B = self._like(x, (N, self.nx, self.nu)) if self.nu else None
        st = self._like(x, (N,), np.int32)
        flags, stream, ptr = self._prep([x, u if self.nu else None, A, B, st], False)
        _abi.check(self._lib.rkb_linearize(self._h, self.device, N, ptr(x), ptr(u) if self.nu else None, float(eps), ptr(A), ptr(B), ptr(st),
                                           flags, stream), "rkb_linearize")
        return A, B, st

    def steer_batch(self, x0, goal, u, dt=None, n_steps=10, want_status=False):
        """x0, goal: [P][nx]; u: [P][R][nu].  Returns (best_idx[P], best_x[P][nx], best_cost[P])."""
        x0, P = self._in(x0, self.nx, np.float64)
        goal, _ = self._in(goal, self.nx, np.float64, rows=P)
        if _is_torch(u):
            u = u.contiguous()
        else:
            u = np.ascontiguousarray(u, dtype=np.float64)
        if len(u.shape) != 3 or u.shape[0] != P or u.shape[2] != self.nu:
            raise IndexError("Input vector dimension mismatch!")
        R = u.shape[1]
        dt = self.dt if dt is None else float(dt)
        idx = self._like(x0, (P,), np.int32)
        bx = self._like(x0, (P, self.nx))
        bc = self._like(x0, (P,))
        st = self._like(x0, (P, R), np.int32) if want_status else None
        flags, stream, ptr = self._prep([x0, goal, u, idx, bx, bc, st], False)
        _abi.check(self._lib.rkb_steer_batch(self._h, self.device, P, R, ptr(x0), ptr(goal), ptr(u), dt, int(n_steps),
                                             ptr(idx), ptr(bx), ptr(bc), ptr(st), flags, stream), "rkb_steer_batch")
        return (idx, bx, bc, st) if want_status else (idx, bx, bc)

    def steer_feedback(self, x0, goal, u_bias, gain, u_prev, time_step, dt, substeps, max_intervals, goal_proximity,
                       saturate_first=False, bounds=None, rate_bounds=None, want_traj=False, proxy_pairs=None):
        """The loop of steer_with_constant_control (examples/misc/MEAQR_topology.hpp:503-561) for N tuples:
        u = bounded(u_prev, u_bias, -gain (x - goal)), one RK4 control interval, stop within goal_proximity.
        x0, goal: [N][nx]; u_bias, u_prev: [N][nu]; gain: [N][nu][nx]; bounds / rate_bounds: (lo, hi) or None.
        Returns (x_out, u_last, n_done[, x_traj [N][max_intervals][nx]], status); u_prev is not modified.
        proxy_pairs: a list of reak_b200.proximity.proxy_query_pair_3D turns the collision test of the loop on
        (rkb_steer_feedback_checked); the tuple then ends with `collided` [N]."""
        x0, N = self._in(x0, self.nx, np.float64)
        goal, _ = self._in(goal, self.nx, np.float64, rows=N)
        u_bias, _ = self._in(u_bias, self.nu, np.float64, rows=N)
        up, _ = self._in(u_prev, self.nu, np.float64, rows=N)
        up = up.clone() if _is_torch(up) else up.copy()
        gain = gain.contiguous() if _is_torch(gain) else np.ascontiguousarray(gain, dtype=np.float64)
        if tuple(gain.shape) != (N, self.nu, self.nx):
            raise IndexError("Input vector dimension mismatch!")
        keep = []

        def box(b):
            if b is None:
                return None, None
            lo = np.ascontiguousarray(b[0], dtype=np.float64).reshape(self.nu)
            hi = np.ascontiguousarray(b[1], dtype=np.float64).reshape(self.nu)
            keep.extend([lo, hi])
            return lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p)

        lo, hi = box(bounds)
        dlo, dhi = box(rate_bounds)
        opts = _abi.rkb_steer_opts(float(time_step), float(dt), float(goal_proximity), int(substeps), int(max_intervals),
                                   int(bool(saturate_first)), 0, lo, hi, dlo, dhi)
        xo = self._like(x0, (N, self.nx))
        nd = self._like(x0, (N,), np.int32)
        tr = self._like(x0, (N, max(int(max_intervals), 1), self.nx)) if want_traj else None
        st = self._like(x0, (N,), np.int32)
        flags, stream, ptr = self._prep([x0, goal, u_bias if self.nu else None, gain if self.nu else None, up if self.nu else None,
                                         xo, nd, tr, st], False)
        if proxy_pairs:
            from . import proximity
            hs = []
            for pair in proxy_pairs:
                h = getattr(pair, "_rkb_handle", None)
                if h is None or getattr(pair, "_rkb_owner", None) is not self:
                    h = proximity.ProxyHandle(self._lib, self._h, pair, self.compiled.frames)
                    pair._rkb_handle, pair._rkb_owner = h, self
                hs.append(h._h)
            arr = (C.c_void_p * len(hs))(*hs)
            col = self._like(x0, (N,), np.int32)
            flags, stream, ptr = self._prep([x0, goal, u_bias if self.nu else None, gain if self.nu else None, up if self.nu else None,
                                             xo, nd, tr, st, col], False)
            _abi.check(self._lib.rkb_steer_feedback_checked(self._h, self.device, N, ptr(x0), ptr(goal), ptr(u_bias) if self.nu else None,
                                                            ptr(gain) if self.nu else None, ptr(up) if self.nu else None, C.byref(opts),
                                                            arr, len(hs), ptr(xo), ptr(nd), ptr(col), ptr(tr), ptr(st), flags, stream),
                       "rkb_steer_feedback_checked")
            if want_traj:
                return xo, up, nd, tr[:, :int(max_intervals)], st, col
            return xo, up, nd, st, col
        _abi.check(self._lib.rkb_steer_feedback(self._h, self.device, N, ptr(x0), ptr(goal), ptr(u_bias) if self.nu else None,
                                                ptr(gain) if self.nu else None, ptr(up) if self.nu else None, C.byref(opts),
                                                ptr(xo), ptr(nd), ptr(tr), ptr(st), flags, stream), "rkb_steer_feedback")
        if want_traj:
            return xo, up, nd, tr[:, :int(max_intervals)], st
        return xo, up, nd, st

    # ---- checked steering in one launch (rkb_steer_checked_specialize) ----------------------------
    def _pair_array(self, proxy_pairs):
        hs = [self.proxy_handle(pair)._h for pair in proxy_pairs]
        return (C.c_void_p * len(hs))(*hs), len(hs)

    def specialize_checked_steering(self, proxy_pairs):
        """compile the steering kernel with the collision test of these pairs built in (NVRTC, now)"""
        arr, n = self._pair_array(proxy_pairs)
        _abi.check(self._lib.rkb_steer_checked_specialize(self._h, self.device, arr, n), "rkb_steer_checked_specialize")
        return self

    def checked_steering_is_specialized(self, proxy_pairs):
        arr, n = self._pair_array(proxy_pairs)
        return bool(self._lib.rkb_steer_checked_is_specialized(self._h, arr, n))

    def checked_steering_source(self, proxy_pairs):
        """the CUDA source rkb_steer_checked_specialize compiles; its last line names the kernel (test hook)"""
        arr, n = self._pair_array(proxy_pairs)
        size = self._lib.rkb_steer_checked_source(self._h, arr, n, None, 0)
        _abi.check(min(size, 0), "rkb_steer_checked_source")
        buf = C.create_string_buffer(size)
        _abi.check(min(self._lib.rkb_steer_checked_source(self._h, arr, n, buf, size), 0), "rkb_steer_checked_source")
        return buf.value.decode()

    # ---- instrumentation ----------------------------------------------------------------------
    def last_kernel_ms(self):
        return self._lib.rkb_last_kernel_ms(self._h)

    def launch_count(self):
        return int(self._lib.rkb_launch_count(self._h))

#!/usr/bin/env python
"""bench.py — batched 6-DOF KTE-chain RK4 forward dynamics on B200 (BASELINE.json config 2).

One "step" = one pass of the hot path over one batch: rkb_rollout_rk4 on 2^20 independent states
of the 6-DOF CRS-A465-style chain, 100 RK4 steps of 1 ms each with the torques held constant.
metric = state-steps/s (samples x RK4 steps per second), whole job over all ranks.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path
  python bench.py --impl reference [...]                         the reference's CPU path (oracle/_ref)

N > 1 runs under torchrun (one rank per GPU, NCCL): the batch is sharded by sample index, every
rank integrates its own 2^20 states (weak scaling), no data-path collective while integrating; the
all-gather of the end states the north star names IS inside the timed step (reak_b200.sharded): by default the
rollout kernel itself stores every end state into all ranks' copies of the batch over NVLink peer memory (no collective
kernel at all, only a barrier); --gather nccl (and the fallback where peer mapping is unavailable) integrates the rank's
block in pieces of whole waves and all-gathers piece c with NCCL while piece c + 1 integrates.

Every BASELINE config is in the line: config 2 is the headline, configs 1, 3, 4, 5 are `other_configs`
at BASELINE's full sizes (sharded over the ranks: strong scaling for those), and at N = 1 each of them is
checked against the reference's own CPU path on a strided subset of its batch and
carries the CPU rate measured on those samples.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rk4_state_steps_per_s"
UNIT = "state-steps/s"
PRESET = "crs6"
N_SAMPLES = 1 << 20
RK4_STEPS = 100
DT = 1e-3
TOL = 1e-8             # BASELINE.json: relative error per component after the config's horizon
CPU_SUBSET = 4096      # samples of each large batch the reference integrates
GATHER_WAVES = 2      # the rank's block is integrated and gathered in pieces of this many full waves of the rollout kernel
# Roofline accounting:
#   * ALGORITHMIC FP64 flops per RK4 state-step of the formulation the kernel ships — counted by running that
#     formulation on symbolic scalars with only the kernel's structural promises known (tools/flop_count.py, live
#     in this run from the chain the bench builds).  roofline.achieved / frac are on this number.
#   * EXECUTED FP64 instructions / flops per state-step (ncu's smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}
#     of a capture of this very build: tools/ncu_summary.py -> profiles/roofline_constants.json, which records
#     rkb_build_id()).  Reported as executed_over_algorithmic and fp64_issue_frac; dropped when the capture
#     belongs to another build.
#   * SURVEY.md 8(d)'s 2.26e4 flop figure counts the dense, general-axis formulation of the reference; the
#     shipped formulation does the same mathematics in a fifth of that, so it is kept only as survey_frac.
#   * bytes of one sample per call: read 2n + n doubles, write 2n doubles + status word.
SURVEY_FLOP_PER_STATE_STEP = 2.26e4
BYTES_PER_SAMPLE = (12 + 6) * 8 + 12 * 8 + 4


def load_roofline_constants():
    with open(os.path.join(ROOT, "profiles", "roofline_constants.json")) as f:
        c = json.load(f)
    return c


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs."""

    def __init__(self, index, period=0.2):
        threading.Thread.__init__(self, daemon=True)
        self.index, self.period = index, period
        self.stop_flag = threading.Event()
        self.sm, self.reasons, self.sm_max = [], set(), None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def sample(self):
        nv = self.nv
        self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        for bit, name in names.items():
            if r & bit:
                self.reasons.add(name)

    def run(self):
        if not self.ok:
            return
        while not self.stop_flag.is_set():
            try:
                self.sample()
            except Exception:
                break
            self.stop_flag.wait(self.period)

    def result(self):
        self.stop_flag.set()
        if self.is_alive():
            self.join(timeout=2.0)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": []}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons)}


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


def make_inputs(n, nx, nu, seed):
    """BASELINE config 2 inputs: q, qd, tau ~ U(-1, 1)."""
    rng = np.random.default_rng(seed)
    return rng.uniform(-1.0, 1.0, (n, nx)), rng.uniform(-1.0, 1.0, (n, nu))


def checker_for(compiled):
    """The CPU side of every comparison: the reference itself (kte_nl_system + runge_kutta4_integrator compiled from
    the unmodified sources, oracle/_ref) when it travelled with the repo, else the C restatement of it."""
    from oracle import pyref
    if pyref.have_ref():
        return pyref.Reference(compiled), "reference"
    if not os.path.isfile(pyref.ORACLE_SO):
        pyref.build(("oracle",))
    return pyref.Oracle(compiled), "port"


def cpu_reference_run(compiled, x, u, workers, steps=RK4_STEPS):
    """Times the reference's CPU path on `workers` forked processes (threads do not scale: every
    rk_dynamic_ptr_cast bumps shared atomic reference counts)."""
    chk, kind = checker_for(compiled)
    out, st, secs = chk.rk4(x, u, DT, steps, n_workers=workers)
    if secs <= 0:
        raise RuntimeError("CPU baseline run failed")
    return out, secs, kind


def rel_err(a, b):
    """max |a - b| / max(1, |b|): the per-component relative error BASELINE.json's tolerance is stated in"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))) if a.size else 0.0


def strided(n, m):
    """m indices spread evenly over [0, n) (all of them when n <= m)"""
    return np.arange(n) if n <= m else (np.arange(m, dtype=np.int64) * (n // m))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from reak_b200 import kte, presets
    s = presets.make(PRESET)
    compiled = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
    cores = os.cpu_count() or 1
    per_core = 96  # ~1.5-2 s per step at ~5-7 k state-steps/s/core
    n = per_core * cores
    x, u = make_inputs(n, 12, 6, 12346)
    kind = "reference"
    for _ in range(args.warmup):
        _, _, kind = cpu_reference_run(compiled, x[: 8 * cores], u[: 8 * cores], cores)
    total = 0.0
    for _ in range(args.steps):
        _, secs, kind = cpu_reference_run(compiled, x, u, cores)
        total += secs
    value = n * RK4_STEPS * args.steps / total
    sample = "%d samples x %d RK4 steps per step (%d per core) of the same workload" % (n, RK4_STEPS, per_core)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "6-DOF CRS-A465-style kte_map_chain, %d RK4 steps dt=1ms, constant torques; CPU: %s" % (RK4_STEPS, sample),
                   "preset": PRESET, "rk4_steps": RK4_STEPS, "dt": DT},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def other_configs(torch, presets, kte_batch_propagator, local, world, rank, check):
    """BASELINE configs 1, 3, 4, 5 at their full sizes, sharded over the ranks (device-resident buffers, kernel time
    from CUDA events on the launch stream, best of 3; run_ours takes the maximum over ranks), plus the two entries of
    SURVEY 8(f) rank 2.  With `check` (N = 1) every config is also integrated by the reference's CPU path on a strided
    subset of its batch: the entry then carries max_rel_err (asserted <= 1e-8) and cpu_baseline."""
    dev = torch.device("cuda", local)
    gen = torch.Generator(device=dev)
    gen.manual_seed(777 + rank)
    cores = os.cpu_count() or 1
    out = []

    def uniform(shape, lo, hi):
        return torch.rand(shape, dtype=torch.float64, device=dev, generator=gen) * (hi - lo) + lo

    def best(fn, prop, reps=3):
        fn()
        ms = []
        for _ in range(reps):
            fn()
            ms.append(prop.last_kernel_ms())
        return min(ms)

    def cpu_check(entry, prop, x, u, steps, gpu_out, what):
        """reference on the strided subset: rate + agreement with the GPU result on exactly those samples"""
        idx = strided(x.shape[0], CPU_SUBSET)
        ti = torch.from_numpy(idx).to(dev)
        xs, us = x[ti].cpu().numpy(), (u[ti].cpu().numpy() if u is not None else None)
        ref, secs, kind = cpu_reference_run(prop.compiled, xs, us, cores, steps)
        err = rel_err(gpu_out[ti].cpu().numpy(), ref)
        entry["max_rel_err"] = err
        entry["tolerance"] = TOL
        entry["cpu_baseline"] = {"value": idx.size * steps / secs, "unit": UNIT, "cores": cores, "kind": kind,
                                 "sample": "%d samples (every %d-th of the batch) x %d RK4 steps, %d forked workers; %s"
                                 % (idx.size, max(1, x.shape[0] // max(1, idx.size)), steps, cores, what)}
        assert err <= TOL, "config %s: GPU result disagrees with the reference on the CPU subset: %g" % (entry["config"], err)
        return ti, ref

    # cfg 1: 2-link planar arm, 1024 states x 1000 RK4 steps (tiny: every rank runs all of it; the CPU integrates all 1024)
    p1 = kte_batch_propagator(presets.make("planar2"), device=local)
    g1 = torch.Generator(device=dev)
    g1.manual_seed(12345 + 1)
    x1 = torch.empty((1024, p1.nx), dtype=torch.float64, device=dev)
    x1[:, 0::2] = (torch.rand((1024, p1.n), dtype=torch.float64, device=dev, generator=g1) * 2 - 1) * np.pi
    x1[:, 1::2] = (torch.rand((1024, p1.n), dtype=torch.float64, device=dev, generator=g1) * 2 - 1) * 2.0
    o1 = torch.empty_like(x1)
    s1 = torch.empty((1024,), dtype=torch.int32, device=dev)
    t1 = best(lambda: p1.get_next_states(x1, None, DT, 1000, out=o1, status=s1), p1)
    e1 = {"config": 1, "workload": "2-link planar arm (revolute_joint_2D + rigid_link_2D + inertia_2D): 1024 states x 1000 RK4 steps (not sharded)",
          "rollout_ms": t1, "serial_kernels": bool(p1.is_serial()), "units_total": 1024 * 1000, "replicated": True}
    if check:
        cpu_check(e1, p1, x1, None, 1000, o1, "all 1024 states")
    out.append(e1)
    # cfg 3: 6-DOF + torsion springs/dampers, 2^24 states over the ranks x 100 RK4 steps, then M and Mdot at the final state
    p3 = kte_batch_propagator(presets.make("crs6_sd"), device=local)
    n3 = (1 << 24) // world
    x3, u3 = uniform((n3, p3.nx), -1, 1), uniform((n3, p3.nu), -1, 1)
    o3 = torch.empty_like(x3)
    s3 = torch.empty((n3,), dtype=torch.int32, device=dev)
    t_roll = best(lambda: p3.get_next_states(x3, u3, DT, RK4_STEPS, out=o3, status=s3), p3)
    MM = [None, None]

    def mass3():
        MM[0], MM[1] = p3.get_mass_matrices(o3, with_derivative=True)

    t_mass = best(mass3, p3)
    e3 = {"config": 3, "workload": "6-DOF + torsion springs/dampers: 2^24 states over %d GPU(s) x %d RK4 steps, then M and Mdot at the end state"
          % (world, RK4_STEPS), "rollout_ms": t_roll, "mass_and_derivative_ms": t_mass, "serial_kernels": bool(p3.is_serial()),
          "units_total": (1 << 24) * RK4_STEPS, "scaling": "strong"}
    if check:
        ti, ref = cpu_check(e3, p3, x3, u3, RK4_STEPS, o3, "then getMassMatrixAndDerivative at the reference's end states")
        chk, _ = checker_for(p3.compiled)
        Mr, Mdr = chk.mass(ref)
        e3["max_rel_err_M"], e3["max_rel_err_Mdot"] = rel_err(MM[0][ti].cpu().numpy(), Mr), rel_err(MM[1][ti].cpu().numpy(), Mdr)
        assert e3["max_rel_err_M"] <= TOL and e3["max_rel_err_Mdot"] <= TOL, "config 3: M / Mdot disagree with the reference"
    out.append(e3)
    del x3, u3, o3, s3, MM
    # cfg 4: 7-DOF with prismatic track, 2^26 RRT extensions over the ranks, 10 RK4 steps each
    p4 = kte_batch_propagator(presets.make("crs7"), device=local)
    n4 = (1 << 26) // world
    x4, u4 = uniform((n4, p4.nx), -1, 1), uniform((n4, p4.nu), -1, 1)
    o4 = torch.empty_like(x4)
    s4 = torch.empty((n4,), dtype=torch.int32, device=dev)
    t4 = best(lambda: p4.get_next_states(x4, u4, DT, 10, out=o4, status=s4), p4)
    e4 = {"config": 4, "workload": "7-DOF with prismatic track: 2^26 RRT extensions over %d GPU(s) x 10 RK4 steps" % world, "rollout_ms": t4,
          "serial_kernels": bool(p4.is_serial()), "units_total": (1 << 26) * 10, "scaling": "strong"}
    if check:
        cpu_check(e4, p4, x4, u4, 10, o4, "one control interval of 10 steps")
    out.append(e4)
    del x4, u4, o4, s4
    torch.cuda.empty_cache()
    # cfg 4, "free" variant: a six-joint arm on a free-floating base (free_joint_3D: 25 states, 12 accelerations) — the model
    # the reference evaluates (no rotors next to a free joint).  Interpreter kernels: the compatibility path.
    p4f = kte_batch_propagator(presets.make("free_arm6"), device=local)
    n4f = (1 << 18) // world
    x4f, u4f = uniform((n4f, p4f.nx), -1, 1), uniform((n4f, p4f.nu), -1, 1)
    o4f = torch.empty_like(x4f)
    s4f = torch.empty((n4f,), dtype=torch.int32, device=dev)
    t4f = best(lambda: p4f.get_next_states(x4f, u4f, DT, 10, out=o4f, status=s4f), p4f)
    e4f = {"config": "4-free", "workload": "6-joint arm on a free_joint_3D base (25 states): 2^18 extensions over %d GPU(s) x 10 RK4 steps" % world,
           "rollout_ms": t4f, "serial_kernels": bool(p4f.is_serial()), "units_total": (1 << 18) * 10, "scaling": "strong"}
    if check:
        cpu_check(e4f, p4f, x4f, u4f, 10, o4f, "one control interval of 10 steps")
    out.append(e4f)
    del x4f, u4f, o4f, s4f
    # cfg 5: steer batch, 4096 pairs x 256 controls x 100 steps over all GPUs, pairs sharded
    p5 = kte_batch_propagator(presets.make(PRESET), device=local)
    P, R = max(1, 4096 // world), 256
    x0, goal, uu = uniform((P, p5.nx), -1, 1), uniform((P, p5.nx), -1, 1), uniform((P, R, p5.nu), -5, 5)
    res5 = [None]

    def steer5():
        res5[0] = p5.steer_batch(x0, goal, uu, DT, RK4_STEPS)

    t5 = best(steer5, p5)
    e5 = {"config": 5, "workload": "steer batch: 4096 pairs over %d GPU(s) x %d controls x %d RK4 steps, arg-min per pair" % (world, R, RK4_STEPS),
          "steer_ms": t5, "units_total": 4096 * R * RK4_STEPS, "scaling": "strong"}
    if check:
        # the reference integrates every rollout of 16 pairs spread over the batch; arg-min and winner taken in numpy
        pi = strided(P, CPU_SUBSET // R)
        tpi = torch.from_numpy(pi).to(dev)
        xs = np.repeat(x0[tpi].cpu().numpy(), R, axis=0)
        us = uu[tpi].reshape(pi.size * R, p5.nu).cpu().numpy()
        ref, secs, kind = cpu_reference_run(p5.compiled, xs, us, cores, RK4_STEPS)
        ends = ref.reshape(pi.size, R, p5.nx)
        cost = np.linalg.norm(ends - goal[tpi].cpu().numpy()[:, None, :], axis=2)
        idx_g, bx_g, bc_g = (a[tpi].cpu().numpy() for a in res5[0][:3])
        rows = np.arange(pi.size)
        # the GPU's winner must be (one of) the reference's cheapest rollouts, and carry the reference's end state
        e5["max_rel_err"] = max(rel_err(bx_g, ends[rows, idx_g]), rel_err(bc_g, cost[rows, idx_g]))
        e5["argmin_agrees"] = bool(np.all(cost[rows, idx_g] <= cost.min(axis=1) + 1e-9))
        e5["tolerance"] = TOL
        e5["cpu_baseline"] = {"value": pi.size * R * RK4_STEPS / secs, "unit": UNIT, "cores": cores, "kind": kind,
                              "sample": "all %d rollouts of %d pairs spread over the batch x %d RK4 steps, %d forked workers" % (R, pi.size, RK4_STEPS, cores)}
        assert e5["max_rel_err"] <= TOL and e5["argmin_agrees"], "config 5: steer batch disagrees with the reference"
    out.append(e5)
    del x0, goal, uu
    # the collision test of the steering loops on propagated states (CRS arm's proximity model
    # against the MD148 lab, 25 finders), alone and inside the closed-loop steering
    from reak_b200 import proximity as px
    s6 = presets.make(PRESET)
    p6 = kte_batch_propagator(s6, device=local)
    robot, lab = presets.crs_proxy_models(s6)
    pair = px.proxy_query_pair_3D("robot-lab", robot, lab)
    n6 = 1 << 20
    x6 = uniform((n6, p6.nx), -3, 3)
    d6 = [None]

    def prox6():
        d6[0] = p6.get_min_distances(pair, x6, with_points=False)[0]

    # interpreter kernel first (no background compilation while it is timed), then the kernel generated for this chain and pair
    h6 = p6.proxy_handle(pair)
    h6.set_option(h6.OPT_AUTO_SPECIALIZE, 0)
    t6i = best(prox6, p6)
    kernel6 = "generated for this chain and pair (rkb_proxy_specialize, NVRTC)"
    try:
        h6.specialize(local)
    except Exception as e:  # no libnvrtc on this box: the interpreter kernel is what runs
        kernel6 = "interpreter (%s)" % e
    t6 = best(prox6, p6)
    e6 = {"config": "proximity", "workload": "findMinimumDistance, CRS arm vs MD148 lab (25 finders): %d states per GPU" % n6,
          "min_distance_ms": t6, "states_per_gpu": n6, "kernel": kernel6, "interpreter_kernel": {"ms": t6i}}
    if check:
        from oracle import pyref
        if pyref.have_ref():
            ti = torch.from_numpy(strided(n6, 8192)).to(dev)
            xs = x6[ti].cpu().numpy()
            t0 = time.perf_counter()
            dr, _, _ = pyref.Reference(p6.compiled).min_distance(pair, xs)
            secs_p = time.perf_counter() - t0
            errp = float(np.max(np.abs(d6[0][ti].cpu().numpy() - dr)))
            assert errp < 1e-10, "GPU minimum distances disagree with the reference: %g" % errp
            e6["cpu_baseline"] = {"value": xs.shape[0] / secs_p, "unit": "states/s", "cores": 1, "kind": "reference", "sample": "%d states" % xs.shape[0]}
            e6["max_abs_err"] = errp
    out.append(e6)
    m6, J6 = 1 << 18, 10
    g6 = uniform((m6, p6.nu, p6.nx), -2, 2)
    xs6 = x6[:m6].contiguous() * 0.3
    gl = xs6 + uniform((m6, p6.nx), -1, 1)
    ub = uniform((m6, p6.nu), -1, 1)
    up = torch.zeros_like(ub)
    done = [None]

    def steer():
        done[0] = p6.steer_feedback(xs6, gl, ub, g6, up, 1e-2, DT, 10, J6, 0.25, proxy_pairs=[pair])[2]

    p6.set_option("auto_specialize", 0)
    t7i = best(steer, p6)   # interval by interval: law, rollout, proximity, commit launches
    kernel7 = "one launch: steering kernel compiled with the pair's collision test (rkb_steer_checked_specialize, NVRTC)"
    try:
        p6.specialize_checked_steering([pair])
    except Exception as e:
        kernel7 = "interval by interval (%s)" % e
    t7 = best(steer, p6)
    units7 = int(done[0].sum().item()) * 10

    def steer_unchecked():
        done[0] = p6.steer_feedback(xs6, gl, ub, g6, up, 1e-2, DT, 10, J6, 0.25)[2]

    t7u = best(steer_unchecked, p6)
    out.append({"config": "steer_checked", "workload": "closed-loop steering with collision test: %d tuples x <= %d intervals x 10 RK4 steps per GPU"
                % (m6, J6), "steer_checked_ms": t7, "units_per_gpu": units7, "kernel": kernel7, "interval_by_interval": {"ms": t7i},
                "unchecked": {"ms": t7u, "state_steps": int(done[0].sum().item()) * 10,
                              "note": "the same tuples without the test (they run on where the checked loop stops at a collision)"}})
    # the nearest-neighbour search that precedes every steer (linear_neighbor_search / dvp_tree), for
    # a batch of 4096 samples against 2^20 motion-graph vertices of the 6-DOF state space (12 coordinates)
    from reak_b200.nearest import nearest_neighbors
    nv, nq, nd = 1 << 20, 4096, 12
    vv, qq = uniform((nv, nd), -1, 1), uniform((nq, nd), -1, 1)
    res = [None]

    def nn():
        res[0] = nearest_neighbors(vv, qq, 1)

    nn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ms_nn = []
    for _ in range(3):
        ev[0].record()
        nn()
        ev[1].record()
        torch.cuda.synchronize(dev)
        ms_nn.append(ev[0].elapsed_time(ev[1]))
    t8 = min(ms_nn)
    e8 = {"config": "nearest", "workload": "nearest vertex (k = 1) of %d query points among %d vertices, %d coordinates, per GPU" % (nq, nv, nd),
          "nearest_ms": t8, "pairs_per_gpu": float(nv) * nq,
          "roofline": {"bound": "fp64", "algorithmic_flop_per_pair": 3 * nd, "achieved": float(nv) * nq * 3 * nd / (t8 * 1e-3) * 1e-12,
                       "unit": "TFLOP/s", "note": "sub, mul, add per coordinate and pair as the reference computes; the scan issues sub + fma "
                                                  "(ceiling 3/4 of the DFMA peak), exact arithmetic only for the candidates"}}
    if check:
        from oracle import pyref
        sub = strided(nq, 16)
        which = "ref" if pyref.have_ref() else "oracle"
        t0 = time.perf_counter()
        ri, rd, _ = pyref.nearest(which, vv.cpu().numpy(), qq[torch.from_numpy(sub).to(dev)].cpu().numpy(), 1)
        secs_n = time.perf_counter() - t0
        gi, gd = res[0][0].cpu().numpy()[sub], res[0][1].cpu().numpy()[sub]
        assert np.array_equal(gi, ri) and np.array_equal(gd, rd), "GPU nearest neighbours differ from the reference's linear scan"
        e8["exact_match"] = True
        e8["cpu_baseline"] = {"value": float(nv) * sub.size / secs_n, "unit": "pairs/s", "cores": 1,
                              "kind": "reference" if which == "ref" else "port",
                              "sample": "%d of the queries, min_dist_linear_search over all %d vertices" % (sub.size, nv)}
    out.append(e8)
    return out


def bind_to_gpu_numa(torch, local):
    """Run this rank on the CPUs next to its GPU, so that the pinned staging buffers it allocates (first touch) and the
    threads that fill them sit on the GPU's NUMA node: eight ranks' host<->device copies then do not cross sockets."""
    try:
        props = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bdf) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return "%s -> cpus %s" % (bdf, spec)
    except Exception as e:  # containers without sysfs topology: stay where we are
        return "unbound (%s)" % type(e).__name__
    return "unbound"


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(torch, local) if world > 1 else "single rank"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from reak_b200 import kte_batch_propagator, presets, _abi
    from reak_b200.sharded import sharded_propagator
    prop = kte_batch_propagator(presets.make(PRESET), device=local)
    assert prop.is_serial(), "bench chain must run on the serial-chain kernels"
    lib = _abi.load_library()
    nx, nu = prop.nx, prop.nu
    n = N_SAMPLES
    x_h, u_h = make_inputs(n, nx, nu, 12346 + rank)

    # pinned host buffers for the end-to-end leg
    px = torch.empty((n, nx), dtype=torch.float64).pin_memory()
    pu = torch.empty((n, nu), dtype=torch.float64).pin_memory()
    po = torch.empty((n, nx), dtype=torch.float64).pin_memory()
    ps = torch.empty((n,), dtype=torch.int32).pin_memory()
    px.numpy()[:] = x_h
    pu.numpy()[:] = u_h
    # device-resident buffers for the kernel leg (151 MB of input per step: larger than the 126 MB L2)
    dx, du = px.cuda(), pu.cuda()
    dout = torch.empty_like(dx)
    dst = torch.empty((n,), dtype=torch.int32, device=dx.device)
    sp, gathered, gathered_st = None, None, None
    # pieces of GATHER_WAVES full waves of the rollout kernel: a launch ends with a partial wave, and pieces of e.g. 1.7
    # waves would cost 2 each (measured: 8 equal pieces of 2^17 samples = 1.73 waves cost 1.2 ms per step more)
    wave = prop.wave_samples()
    chunk_samples = GATHER_WAVES * wave if wave > 0 else n // 8
    if world > 1:
        sp = sharded_propagator(prop, comm_device=dx.device)
        gathered = torch.empty((world * n, nx), dtype=torch.float64, device=dx.device)
        gathered_st = torch.empty((world * n,), dtype=torch.int32, device=dx.device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    kernel_ms = []
    gathered_now = [None, None]  # the tensors the last step's gathered batch lives in

    def device_step(record=False):
        if world == 1:
            prop.get_next_states(dx, du, DT, RK4_STEPS, out=dout, status=dst)
            if record:
                kernel_ms.append(prop.last_kernel_ms())  # CUDA events recorded around the kernel on its launch stream
        else:
            # the shipped sharded path: rollout in pieces, every piece all-gathered (NCCL) while the next one integrates
            # (or, where the GPUs can map each other's memory, the rollout kernel stores its end states into every rank's
            # copy of the batch itself and no collective runs at all: --gather p2p, the default)
            res = sp.get_next_states(dx, du, DT, RK4_STEPS, local_input=True, n_total=world * n, chunk_samples=chunk_samples,
                                     out=gathered, status=gathered_st, peer_stores=(args.gather == "p2p"))
            gathered_now[0], gathered_now[1] = res

    def e2e_step():
        prop.get_next_states(px.numpy(), pu.numpy(), DT, RK4_STEPS, out=po.numpy(), status=ps.numpy())

    # ---- kernel leg (N > 1: rollout + all-gather) -----------------------------------------------
    t_warm = time.time()
    done = 0
    while done < max(args.warmup, 3) or time.time() - t_warm < 2.0:  # let the SM clock settle
        device_step()
        torch.cuda.synchronize()
        done += 1
    barrier()
    sampler = ClockSampler(physical_gpu_index(local))
    sampler.start()
    launches0 = prop.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        device_step(record=True)
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    clocks = sampler.result()
    launches = prop.launch_count() - launches0
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dx.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    rollout_only_ms = None
    if world > 1:
        assert int(gathered_now[1].max().item()) == 0, "status word set in the timed region"
        gather_mode = "p2p_stores" if gathered_now[0].data_ptr() != gathered.data_ptr() else "nccl_pieces"
        # the same steps without the gather, for the record (what round 1 reported as `value`)
        prop.get_next_states(dx, du, DT, RK4_STEPS, out=dout, status=dst)
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(args.steps):
            prop.get_next_states(dx, du, DT, RK4_STEPS, out=dout, status=dst)
            kernel_ms.append(prop.last_kernel_ms())
        r1.record()
        barrier()
        t = torch.tensor([r0.elapsed_time(r1)], dtype=torch.float64, device=dx.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rollout_only_ms = float(t.item()) / args.steps
        assert torch.equal(gathered_now[0][rank * n:(rank + 1) * n], dout), "gathered block differs from the local result"
        # ... and every rank holds every block: compare a checksum of the whole gathered batch across ranks
        chk = torch.stack([gathered_now[0].sum(), gathered_now[0][::97].abs().sum()])
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "the ranks' gathered batches differ"
    assert int(dst.max().item()) == 0, "status word set in the timed region"

    # ---- end-to-end leg: pinned host buffers through the public API -------------------------
    for _ in range(2):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dx.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    assert np.array_equal(po.numpy(), dout.cpu().numpy()), "host and device legs disagree"

    # ---- the same call from PAGEABLE host memory (a ReaK caller's std::vector), plain and after rkb_host_pin ----------
    pageable = None
    if world == 1:
        hx, hu = np.array(x_h, copy=True), np.array(u_h, copy=True)
        ho, hs = np.empty_like(hx), np.empty((n,), dtype=np.int32)
        reps = max(2, args.steps // 3)

        def timed_host():
            prop.get_next_states(hx, hu, DT, RK4_STEPS, out=ho, status=hs)
            t0 = time.perf_counter()
            for _ in range(reps):
                prop.get_next_states(hx, hu, DT, RK4_STEPS, out=ho, status=hs)
            return (time.perf_counter() - t0) / reps * 1e3

        ms_plain = timed_host()
        assert np.array_equal(ho, po.numpy())
        t0 = time.perf_counter()
        for a in (hx, hu, ho, hs):
            _abi.check(lib.rkb_host_pin(a.ctypes.data_as(C.c_void_p), a.nbytes), "rkb_host_pin")
        pin_ms = (time.perf_counter() - t0) * 1e3
        ms_pinned = timed_host()
        for a in (hx, hu, ho, hs):
            lib.rkb_host_unpin(a.ctypes.data_as(C.c_void_p))
        pageable = {"value": n * RK4_STEPS / (ms_plain * 1e-3), "unit": UNIT, "ms_per_step": ms_plain,
                    "after_rkb_host_pin": {"value": n * RK4_STEPS / (ms_pinned * 1e-3), "ms_per_step": ms_pinned, "pin_once_ms": pin_ms},
                    "note": "numpy arrays from malloc (pageable), wall clock around the blocking call"}

    # ---- the other BASELINE configs: every rank measures its shard, the slowest rank counts ----------
    others = None
    if not args.no_other_configs:
        others = other_configs(torch, presets, kte_batch_propagator, local, world, rank, check=(world == 1 and not args.no_cpu_baseline))
        keys = [(i, k) for i, o in enumerate(others) for k in sorted(o) if k.endswith("_ms")]
        t = torch.tensor([others[i][k] for i, k in keys], dtype=torch.float64, device=dx.device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        for (i, k), v in zip(keys, t.tolist()):
            others[i][k] = v
        for o in others:
            ms = sum(v for k, v in o.items() if k.endswith("_ms"))
            if "states_per_gpu" in o:
                o["states_per_s"] = world * o.pop("states_per_gpu") / (ms * 1e-3)
            elif "units_total" in o:
                o["state_steps_per_s"] = o.pop("units_total") / (ms * 1e-3)
            elif "pairs_per_gpu" in o:  # nearest neighbours: every rank searches its own queries
                o["pairs_per_s"] = world * o.pop("pairs_per_gpu") / (ms * 1e-3)
            else:
                o["state_steps_per_s"] = world * o.pop("units_per_gpu") / (ms * 1e-3)

    if rank != 0:
        shutdown(torch, dist, sp, world)
        return 0

    state_steps_per_step = float(world) * n * RK4_STEPS
    value = state_steps_per_step * args.steps / (elapsed_ms * 1e-3)
    e2e_value = state_steps_per_step * args.steps / (e2e_ms * 1e-3)

    # ---- roofline of the dominant kernel (serial_rollout_kernel<6,0,arm>), this rank --------------
    peaks, peaks_src = load_peaks()
    tf = C.c_double(0.0)
    clk = C.c_double(0.0)
    _abi.check(lib.rkb_measure_fp64_peak(local, 1.0, C.byref(tf), C.byref(clk)), "rkb_measure_fp64_peak")
    k_ms = float(np.mean(kernel_ms))
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import flop_count
    alg = flop_count.count_chain(flop_count.stages_of(prop))
    rate = n * RK4_STEPS / (k_ms * 1e-3)  # state-steps per second of this kernel
    achieved_tf = alg["flop_per_state_step"] * rate / 1e12
    achieved_gbs = BYTES_PER_SAMPLE * n / (k_ms * 1e-3) / 1e9
    build = lib.rkb_build_id().decode()
    roofline = {"bound": "fp64", "achieved": achieved_tf, "peak": tf.value, "unit": "TFLOP/s", "frac": achieved_tf / tf.value,
                "algorithmic_flop_per_state_step": alg["flop_per_state_step"],
                "algorithmic_fp64_instr_per_state_step": alg["instr_per_state_step"],
                "algorithmic_issue_frac": alg["instr_per_state_step"] * rate / (tf.value * 1e12 / 2.0),
                "algorithmic_source": "tools/flop_count.py on the chain of this run (kernel shape %#x)" % prop.kernel_shape(),
                "survey_flop_per_state_step": SURVEY_FLOP_PER_STATE_STEP,
                "survey_frac": SURVEY_FLOP_PER_STATE_STEP * rate / 1e12 / tf.value,
                "kernel": "serial_rollout_kernel<6,0,arm>", "kernel_ms": k_ms, "build_id": build,
                "peak_source": "DFMA loop measured in this run (rkb_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 entry",
                "traffic": None}
    try:
        rc = load_roofline_constants()
    except Exception:
        rc = {}
    if rc.get("build_id") == build:
        # executed counts of THIS build (ncu): pipe utilisation and how far the compiled code is from the formulation
        roofline.update({"traffic": rc["dram_traffic_bytes_per_launch"],
                         "executed_flop_per_state_step": rc["flop_per_state_step"],
                         "executed_fp64_instr_per_state_step": rc["fp64_instr_per_state_step"],
                         "executed_over_algorithmic": rc["flop_per_state_step"] / alg["flop_per_state_step"],
                         "fp64_issue_frac": rc["fp64_instr_per_state_step"] * rate / (tf.value * 1e12 / 2.0),
                         "executed_source": "profiles/roofline_constants.json <- " + str(rc.get("source"))})
    else:
        roofline["executed_source"] = "none: profiles/roofline_constants.json is of build %s, this library is %s" % (rc.get("build_id"), build)
    # ... and of the generated proximity kernel: executed FP64 flops per state of this build's capture (the
    # count is data-dependent: the bounding-sphere test skips finders), against the same measured DFMA peak
    if others:
        try:
            with open(os.path.join(ROOT, "profiles", "proximity_constants.json")) as f:
                pc = json.load(f)
        except Exception:
            pc = {}
        for o in others:
            if o.get("config") == "proximity" and str(o.get("kernel", "")).startswith("generated") and pc.get("build_id") == build:
                sps = o["states_per_s"] / world
                o["roofline"] = {"bound": "fp64", "achieved": pc["flop_per_state_step"] * sps / 1e12, "peak": tf.value, "unit": "TFLOP/s",
                                 "frac": pc["flop_per_state_step"] * sps / 1e12 / tf.value,
                                 "executed_flop_per_state": pc["flop_per_state_step"], "executed_fp64_instr_per_state": pc["fp64_instr_per_state_step"],
                                 "fp64_pipe_active_pct": pc["fp64_pipe_active_pct"], "traffic": pc["dram_traffic_bytes_per_launch"],
                                 "algorithmic_bytes": 108 * pc["units_per_launch"], "kernel": pc["kernel"],
                                 "source": "profiles/proximity_constants.json <- " + str(pc.get("source"))}
    roofline_hbm = {"bound": "hbm", "achieved": achieved_gbs, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                    "frac": achieved_gbs / peaks.get("hbm_gbs"), "peak_source": peaks_src, "bytes_per_sample": BYTES_PER_SAMPLE}

    # ---- CPU baseline: the reference itself on the host cores, a strided subset of the batch (N = 1 only) ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        idx = strided(n, CPU_SUBSET)
        ref_out, secs, kind = cpu_reference_run(prop.compiled, x_h[idx], u_h[idx], cores)
        err = rel_err(po.numpy()[idx], ref_out)
        cpu = {"value": idx.size * RK4_STEPS / secs, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": "%d samples (every %d-th of the %d) x %d RK4 steps, %d forked workers" % (idx.size, n // idx.size, n, RK4_STEPS, cores),
               "max_rel_err_vs_gpu": err, "tolerance": TOL}
        assert err <= TOL, "GPU result disagrees with the reference on the CPU sample: %g" % err

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "6-DOF CRS-A465-style kte_map_chain (BASELINE config 2): 2^20 states x 100 RK4 steps dt=1ms per GPU, constant torques"
                   + ("; end states all-gathered over NCCL inside the step" if world > 1 else ""),
                   "preset": PRESET, "samples_per_gpu": n, "rk4_steps": RK4_STEPS, "dt": DT, "parallelism": "sample-sharded x%d" % world,
                   "l2": "inputs (151 MB per step) exceed the 126 MB L2; no flush", "numa": numa},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": int(world * n * (nx + nu) * 8), "d2h_bytes_per_step": int(world * n * (nx * 8 + 4))},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "roofline_hbm": roofline_hbm,
    }
    if pageable is not None:
        line["e2e_pageable_host"] = pageable
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if rollout_only_ms is not None:
        line["gather"] = {"included_in_value": True, "mode": gather_mode, "piece_samples": int(chunk_samples) if gather_mode == "nccl_pieces" else None,
                          "peer_error": getattr(sp, "_peer_error", None), "rollout_only_ms_per_step": rollout_only_ms,
                          "exposed_ms_per_step": elapsed_ms / args.steps - rollout_only_ms}
    if others is not None:
        line["other_configs"] = others
    emit(line)
    shutdown(torch, dist, sp, world)
    return 0


def shutdown(torch, dist, sp, world):
    """Orderly exit of a multi-rank run: the symmetric (peer-mapped) buffers are released while every rank is still
    there, then the process group goes."""
    if world <= 1:
        return
    torch.cuda.synchronize()
    dist.barrier()
    if sp is not None:
        sp.release_peer_buffers()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


_JSON_OUT = None


def emit(line):
    """The one JSON line, on the process's original stdout."""
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # Native libraries write to file descriptor 1 behind Python's back (NCCL's version banner under NCCL_DEBUG, for
    # one): keep the original stdout for the JSON line and point descriptor 1 at stderr for everything else.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: p2p = the rollout kernel stores into every rank's buffer over NVLink (falls back to nccl); nccl = pipelined NCCL all-gather")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    args = ap.parse_args()
    if args.steps < 1:
        args.steps = 1
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

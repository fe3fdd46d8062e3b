#!/usr/bin/env python
"""bench.py — batched 6-DOF KTE-chain RK4 forward dynamics on B200 (BASELINE.json config 2).

One "step" = one pass of the hot path over one batch: rkb_rollout_rk4 on 2^20 independent states
of the 6-DOF CRS-A465-style chain, 100 RK4 steps of 1 ms each with the torques held constant.
metric = state-steps/s (samples x RK4 steps per second), whole job over all ranks.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path
  python bench.py --impl reference [...]                         the reference's CPU path (oracle/_ref)

N > 1 runs under torchrun (one rank per GPU, NCCL): the batch is sharded by sample index, every
rank integrates its own 2^20 states (weak scaling), no data-path collective; the final
all-gather of the end states the north star names is timed separately (gather_ms).
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
# Algorithmic work of one RK4 state-step of the 6-DOF chain:
#   * FP64 flops and FP64 instructions the structure-specialised kernel executes per state-step (DFMA = 2
#     flops), taken from ncu's smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on of the committed
#     capture: tools/ncu_summary.py writes them, with the DRAM traffic of that launch, to
#     profiles/roofline_constants.json, so the numbers always belong to the kernel that was profiled.
#     SURVEY.md 8(d)'s 2.26e4 flop figure is the count for the dense, general-axis formulation; the
#     specialised kernel does the same mathematics with the structural zeros removed, so that figure
#     would read as more than the DFMA peak and is reported separately (survey_frac).
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


def cpu_reference_run(compiled, x, u, workers):
    """Times the reference's own CPU path (kte_nl_system + runge_kutta4_integrator compiled from
    the unmodified sources, oracle/_ref) on `workers` forked processes; falls back to the C port."""
    from oracle import pyref
    if pyref.have_ref():
        chk, kind = pyref.Reference(compiled), "reference"
    else:
        if not os.path.isfile(pyref.ORACLE_SO):
            pyref.build(("oracle",))
        chk, kind = pyref.Oracle(compiled), "port"
    out, st, secs = chk.rk4(x, u, DT, RK4_STEPS, n_workers=workers)
    if secs <= 0:
        raise RuntimeError("CPU baseline run failed")
    return out, secs, kind


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


def other_configs(torch, presets, kte_batch_propagator, local, world):
    """BASELINE configs 3-5 on this rank's shard (device-resident buffers, kernel time from CUDA events on
    the launch stream, best of 3; run_ours takes the maximum over ranks).  Parity of these paths is the job
    of tests/test_gpu_parity.py."""
    rng = np.random.default_rng(777)
    out = []

    def best(fn, prop, reps=3):
        fn()
        ms = []
        for _ in range(reps):
            fn()
            ms.append(prop.last_kernel_ms())
        return min(ms)

    # cfg 3: 6-DOF + torsion springs/dampers, 100 RK4 steps, then M and Mdot at the final state (2^24 / 8 per GPU)
    p3 = kte_batch_propagator(presets.make("crs6_sd"), device=local)
    n3 = 1 << 21
    x3 = torch.from_numpy(rng.uniform(-1, 1, (n3, p3.nx))).cuda(local)
    u3 = torch.from_numpy(rng.uniform(-1, 1, (n3, p3.nu))).cuda(local)
    o3 = torch.empty_like(x3)
    s3 = torch.empty((n3,), dtype=torch.int32, device=x3.device)
    t_roll = best(lambda: p3.get_next_states(x3, u3, DT, RK4_STEPS, out=o3, status=s3), p3)
    t_mass = best(lambda: p3.get_mass_matrices(o3, with_derivative=True), p3)
    out.append({"config": 3, "workload": "6-DOF + torsion springs/dampers: %d states x %d RK4 steps per GPU, then M and Mdot" % (n3, RK4_STEPS),
                "rollout_ms": t_roll, "mass_and_derivative_ms": t_mass, "serial_kernels": bool(p3.is_serial()),
                "units_per_gpu": n3 * RK4_STEPS})
    del x3, u3, o3, s3
    # cfg 4: 7-DOF with prismatic track, 10 RK4 steps per extension (2^26 / 8 per GPU)
    p4 = kte_batch_propagator(presets.make("crs7"), device=local)
    n4 = 1 << 23
    x4 = torch.from_numpy(rng.uniform(-1, 1, (n4, p4.nx))).cuda(local)
    u4 = torch.from_numpy(rng.uniform(-1, 1, (n4, p4.nu))).cuda(local)
    o4 = torch.empty_like(x4)
    s4 = torch.empty((n4,), dtype=torch.int32, device=x4.device)
    t4 = best(lambda: p4.get_next_states(x4, u4, DT, 10, out=o4, status=s4), p4)
    out.append({"config": 4, "workload": "7-DOF with prismatic track: %d RRT extensions x 10 RK4 steps per GPU" % n4, "rollout_ms": t4,
                "serial_kernels": bool(p4.is_serial()), "units_per_gpu": n4 * 10})
    del x4, u4, o4, s4
    # cfg 5: steer batch, 4096 pairs x 256 controls x 100 steps over all GPUs, pairs sharded
    p5 = kte_batch_propagator(presets.make(PRESET), device=local)
    P, R = max(1, 4096 // world), 256
    x0 = torch.from_numpy(rng.uniform(-1, 1, (P, p5.nx))).cuda(local)
    goal = torch.from_numpy(rng.uniform(-1, 1, (P, p5.nx))).cuda(local)
    uu = torch.from_numpy(rng.uniform(-5, 5, (P, R, p5.nu))).cuda(local)
    t5 = best(lambda: p5.steer_batch(x0, goal, uu, DT, RK4_STEPS), p5)
    out.append({"config": 5, "workload": "steer batch: %d pairs x %d controls x %d RK4 steps per GPU, arg-min per pair" % (P, R, RK4_STEPS),
                "steer_ms": t5, "units_per_gpu": P * R * RK4_STEPS})
    del x0, goal, uu
    # the collision test of the steering loops on propagated states (CRS arm's proximity model
    # against the MD148 lab, 25 finders), alone and inside the closed-loop steering
    from reak_b200 import proximity as px
    s6 = presets.make(PRESET)
    p6 = kte_batch_propagator(s6, device=local)
    robot, lab = presets.crs_proxy_models(s6)
    pair = px.proxy_query_pair_3D("robot-lab", robot, lab)
    n6 = 1 << 20
    x6 = torch.from_numpy(rng.uniform(-3, 3, (n6, p6.nx))).cuda(local)
    t6 = best(lambda: p6.get_min_distances(pair, x6, with_points=False), p6)
    out.append({"config": "proximity", "workload": "findMinimumDistance, CRS arm vs MD148 lab (25 finders): %d states per GPU" % n6,
                "min_distance_ms": t6, "states_per_gpu": n6})
    m6, J6 = 1 << 18, 10
    g6 = torch.from_numpy(rng.uniform(-2, 2, (m6, p6.nu, p6.nx))).cuda(local)
    xs = x6[:m6].contiguous() * 0.3
    gl = xs + torch.from_numpy(rng.uniform(-1, 1, (m6, p6.nx))).cuda(local)
    ub = torch.from_numpy(rng.uniform(-1, 1, (m6, p6.nu))).cuda(local)
    up = torch.zeros_like(ub)
    done = [None]

    def steer():
        done[0] = p6.steer_feedback(xs, gl, ub, g6, up, 1e-2, DT, 10, J6, 0.25, proxy_pairs=[pair])[2]

    t7 = best(steer, p6)
    out.append({"config": "steer_checked", "workload": "closed-loop steering with collision test: %d tuples x <= %d intervals x 10 RK4 steps per GPU"
                % (m6, J6), "steer_checked_ms": t7, "units_per_gpu": int(done[0].sum().item()) * 10})
    return out


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
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from reak_b200 import kte_batch_propagator, presets, _abi
    prop = kte_batch_propagator(presets.make(PRESET), device=local)
    assert prop.is_serial(), "bench chain must run on the serial-chain kernels"
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        prop.get_next_states(dx, du, DT, RK4_STEPS, out=dout, status=dst)

    def e2e_step():
        prop.get_next_states(px.numpy(), pu.numpy(), DT, RK4_STEPS, out=po.numpy(), status=ps.numpy())

    # ---- kernel-only leg -------------------------------------------------------------------
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
    kernel_ms = []
    ev0.record()
    for _ in range(args.steps):
        device_step()
        kernel_ms.append(prop.last_kernel_ms())  # CUDA events recorded around the kernel on its launch stream
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    clocks = sampler.result()
    launches = prop.launch_count() - launches0
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dx.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
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

    # ---- final gather of the end states (north star: NCCL used only for this) ---------------
    gather_ms = None
    if world > 1:
        gathered = torch.empty((world * n, nx), dtype=torch.float64, device=dx.device)
        dist.all_gather_into_tensor(gathered, dout)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        dist.all_gather_into_tensor(gathered, dout)
        g1.record()
        barrier()
        t = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dx.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gather_ms = float(t.item())

    # ---- the other BASELINE configs: every rank measures its shard, the slowest rank counts ----------
    others = None
    if not args.no_other_configs:
        others = other_configs(torch, presets, kte_batch_propagator, local, world)
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
            else:
                o["state_steps_per_s"] = world * o.pop("units_per_gpu") / (ms * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    state_steps_per_step = float(world) * n * RK4_STEPS
    value = state_steps_per_step * args.steps / (elapsed_ms * 1e-3)
    e2e_value = state_steps_per_step * args.steps / (e2e_ms * 1e-3)

    # ---- roofline of the dominant kernel (serial_rollout_kernel<6,0>), this rank --------------
    peaks, peaks_src = load_peaks()
    tf = C.c_double(0.0)
    clk = C.c_double(0.0)
    _abi.check(_abi.load_library().rkb_measure_fp64_peak(local, 1.0, C.byref(tf), C.byref(clk)), "rkb_measure_fp64_peak")
    k_ms = float(np.mean(kernel_ms))
    rc = load_roofline_constants()
    FLOP_PER_STATE_STEP, FP64_INSTR_PER_STATE_STEP = rc["flop_per_state_step"], rc["fp64_instr_per_state_step"]
    DRAM_TRAFFIC_BYTES = rc["dram_traffic_bytes_per_launch"]
    achieved_tf = FLOP_PER_STATE_STEP * n * RK4_STEPS / (k_ms * 1e-3) / 1e12
    achieved_gbs = BYTES_PER_SAMPLE * n / (k_ms * 1e-3) / 1e9
    instr_rate = FP64_INSTR_PER_STATE_STEP * n * RK4_STEPS / (k_ms * 1e-3)
    roofline = {"bound": "fp64", "achieved": achieved_tf, "peak": tf.value, "unit": "TFLOP/s", "frac": achieved_tf / tf.value,
                "fp64_issue_frac": instr_rate / (tf.value * 1e12 / 2.0),
                "fp64_instr_per_state_step": FP64_INSTR_PER_STATE_STEP,
                "survey_flop_per_state_step": SURVEY_FLOP_PER_STATE_STEP,
                "survey_frac": SURVEY_FLOP_PER_STATE_STEP * n * RK4_STEPS / (k_ms * 1e-3) / 1e12 / tf.value,
                "traffic": DRAM_TRAFFIC_BYTES, "kernel": "serial_rollout_kernel<6,0,arm>", "kernel_ms": k_ms,
                "constants_source": "profiles/roofline_constants.json <- " + str(rc.get("source")),
                "peak_source": "DFMA loop measured in this run (rkb_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 entry",
                "flop_per_state_step": FLOP_PER_STATE_STEP}
    roofline_hbm = {"bound": "hbm", "achieved": achieved_gbs, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                    "frac": achieved_gbs / peaks.get("hbm_gbs"), "peak_source": peaks_src, "bytes_per_sample": BYTES_PER_SAMPLE}

    # ---- CPU baseline: the reference itself on the host cores, bounded sample (N = 1 only) ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        per_core = 128
        m = min(n, per_core * cores)
        ref_out, secs, kind = cpu_reference_run(prop.compiled, x_h[:m], u_h[:m], cores)
        err = float(np.max(np.abs(po.numpy()[:m] - ref_out) / np.maximum(1.0, np.abs(ref_out))))
        cpu = {"value": m * RK4_STEPS / secs, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": "first %d of the %d samples x %d RK4 steps, %d forked workers" % (m, n, RK4_STEPS, cores),
               "max_rel_err_vs_gpu": err, "tolerance": 1e-8}
        assert err < 1e-8, "GPU result disagrees with the reference on the CPU sample: %g" % err
        # the same for the proximity entry: the reference's findMinimumDistance on one host core, checked against the GPU
        prox = [o for o in (others or []) if o.get("config") == "proximity"]
        if prox and kind == "reference":
            from oracle import pyref
            from reak_b200 import proximity as px
            s6 = presets.make(PRESET)
            p6 = kte_batch_propagator(s6, device=local)
            robot, lab = presets.crs_proxy_models(s6)
            pair = px.proxy_query_pair_3D("robot-lab", robot, lab)
            xs = 3.0 * np.ascontiguousarray(x_h[:8192])
            t0 = time.perf_counter()
            dr, _, _ = pyref.Reference(p6.compiled).min_distance(pair, xs)
            secs_p = time.perf_counter() - t0
            dg, _ = p6.get_min_distances(pair, xs, with_points=False)
            errp = float(np.max(np.abs(dg - dr)))
            assert errp < 1e-10, "GPU minimum distances disagree with the reference: %g" % errp
            prox[0]["cpu_baseline"] = {"value": xs.shape[0] / secs_p, "unit": "states/s", "cores": 1, "kind": "reference",
                                       "sample": "%d states" % xs.shape[0], "max_abs_err_vs_gpu": errp}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "6-DOF CRS-A465-style kte_map_chain (BASELINE config 2): 2^20 states x 100 RK4 steps dt=1ms per GPU, constant torques",
                   "preset": PRESET, "samples_per_gpu": n, "rk4_steps": RK4_STEPS, "dt": DT, "parallelism": "sample-sharded x%d" % world,
                   "l2": "inputs (151 MB per step) exceed the 126 MB L2; no flush"},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": int(world * n * (nx + nu) * 8), "d2h_bytes_per_step": int(world * n * (nx * 8 + 4))},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "roofline_hbm": roofline_hbm,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if gather_ms is not None:
        line["gather_ms"] = gather_ms
    if others is not None:
        line["other_configs"] = others
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


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

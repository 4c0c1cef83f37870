#!/usr/bin/env python
"""bench.py -- frame-pair RMSDs per second on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--only NAME] [--no-subs]

A "step" is one pass of the hot path over one batch of synthetic input.  The HEADLINE (top-level keys of the JSON line)
is BASELINE.json configs[1]: one-vs-many md.rmsd, 1,000,000 frames x 1,000 atoms, precentered=False, reference = frame
0.  With N > 1 (torchrun, one rank per GPU) every rank owns its own 1M-frame shard (frames are independent: no data-path
collective, weak scaling) and `value` is the aggregate.

The same line carries, under "sub", one record per remaining BASELINE config, each with its own timed region, roofline,
end-to-end number, CPU baseline and parity block:
  superpose      configs[2]  Trajectory.superpose, 200k x 5,000 atoms, every 5th atom aligned        (N = 1 only)
  allpairs_20k   configs[3]  shape at F = 20,000: all-pairs matrix incl. e2e / parity / CPU baseline   (N = 1 only)
  allpairs_100k  configs[3]  at full size, 100k x 100k x 300 atoms: ONE GPU at N = 1; at N > 1 the frames are
                 NCCL-broadcast, row blocks sharded, every unordered block pair computed once and the transposed blocks
                 exchanged over NVLink (distributed.rmsd_matrix_sharded) -- strong scaling, the one step of the path with
                 a collective, inside its timed region
  ovm25k         configs[4]  per-GPU shard: 62,500 frames x 25,000 atoms (weak scaling)

  value      device-resident throughput: inputs already in HBM, CUDA-event timed, barrier + synchronize both sides,
             max over ranks
  e2e        the same metric through the public API (mdb.rmsd(host Trajectory) etc.) on PAGEABLE numpy input -- what a
             user of mdtraj holds -- with H2D of every frame and D2H of the result inside the timed region;
             e2e_pinned: the same with page-locked input
  roofline   dominant kernel: algorithmic bytes (or issued tensor flops) / CUDA-event launch time, against
             MEASURED_PEAKS.json (HBM) or a TF32 GEMM measured in this run (tensor)
  cpu_baseline  the compiled reference (oracle/_ref) on this box's host cores, all of them, bounded sample
  parity     GPU path vs the CPU reference vs the float64 truth on samples of the same shape (outside the timed region)

--impl reference times the reference's own CPU implementation (oracle/_ref = its C++ sources compiled in place + our
OpenMP loop shell; the C port oracle/liboracle.so if _ref was never built) with every host thread this process may use,
on rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "frame-pair RMSDs/sec"
WORKLOADS = {
    # name: (frames per GPU, atoms, description)
    "ovm": (1_000_000, 1000, "one-vs-many md.rmsd: synthetic 1M frames x 1,000 atoms, precentered=False (BASELINE configs[1])"),
    "ovm25k": (62_500, 25_000, "one-vs-many md.rmsd: 62,500 frames x 25,000 atoms per GPU (BASELINE configs[4] per-GPU shard)"),
    "superpose": (200_000, 5000, "Trajectory.superpose: 200k frames x 5,000 atoms, atom_indices=arange(0,5000,5) (BASELINE configs[2])"),
    "allpairs_20k": (20_000, 300, "all-pairs RMSD matrix: 20k x 20k frames x 300 atoms (BASELINE configs[3] shape, reduced F)"),
    "allpairs_100k": (100_000, 300, "all-pairs RMSD matrix: 100k x 100k frames x 300 atoms (BASELINE configs[3])"),
    "ala2": (100, 22, "md.rmsd 100 frames x 22 atoms (BASELINE configs[0] shape, synthetic)"),
}
TOL = "1e-5 nm abs or 1e-4 rel"


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return float(p["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def ncu_fact(key):
    """Per-launch facts of the dominant kernels from the committed ncu captures (profiles/ncu_traffic.json)."""
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(tpath):
        return None
    with open(tpath) as fh:
        return json.load(fh).get(key)


# ------------------------------------------------------------------------------------------------
# clocks sampler (NVML in a thread; nvidia-smi -lms is too coarse for sub-second timed regions)
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        while not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.005)

    def result(self):
        self.stop_flag = True
        self.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": float(statistics.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm
# ------------------------------------------------------------------------------------------------
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_run(workload, steps, warmup):
    """Time the reference CPU path on a bounded sample of `workload` with every host thread this process may use
    (torchrun exports OMP_NUM_THREADS=1 to its ranks: the thread count is set explicitly, like mdtraj's own
    `parallel=True` default uses all cores, _rmsd.pyx:136-141).  Returns the cpu_baseline record."""
    from oracle import oracle as O
    F_full, N, _ = WORKLOADS[workload]
    kind = "reference" if O.ref_available() else "port"
    threads = 1
    if kind == "reference":
        L = O.ref_lib()
        L.refloops_set_threads(host_threads())
        threads = int(L.refloops_max_threads())
    if workload.startswith("allpairs"):
        F, rows = 20_000, 8
        X = O.synth_iid(F, N, seed=4)
        tr = O.center_and_trace(X, kind)

        def step():
            for i in range(rows):
                O.one_vs_many_centered(X, tr, X[i], tr[i], impl=kind)
        units = rows * F
        sample = f"{rows} rows of md.rmsd(t,t,i,precentered=True) at F={F}, N={N} after center_coordinates()"
    elif workload == "superpose":
        F = 2000
        X = O.synth_iid(F, N, seed=4)
        idx = np.arange(0, N, 5)

        def step():
            O.superpose(X, X, 0, idx, impl=kind)
        units = F
        sample = f"Trajectory.superpose restated (numpy glue + reference C kernels) on {F} x {N}, 1000-atom subset"
    else:
        F = int(min(F_full, max(64, 1.2e9 // (N * 12))))
        X = O.synth_iid(F, N, seed=4)

        def step():
            O.rmsd(X, X, 0, impl=kind, inplace=True)  # centres in place, like the reference's view path
        units = F
        sample = f"md.rmsd(t,t,0) on {F} x {N} float32 host frames (precentered=False)"
    for _ in range(max(1, warmup)):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return {"value": units / dt, "unit": "rmsd/s", "cores": threads, "kind": kind, "sample": sample,
            "host_cpus": os.cpu_count(), "ms_per_step": dt * 1e3}


def _verdict(got, ref, truth, what, kind):
    got, ref, truth = (np.asarray(v, dtype=np.float64) for v in (got, ref, truth))
    e_gr, e_gt, e_rt = np.abs(got - ref), np.abs(got - truth), np.abs(ref - truth)
    ok = bool(np.all((e_gr <= np.maximum(1e-5, 1e-4 * np.abs(ref))) | (e_gt <= np.maximum(1e-5, e_rt))))
    return {"what": what, "n": int(got.size), "max_abs_gpu_vs_ref": float(e_gr.max()), "max_abs_gpu_vs_truth": float(e_gt.max()),
            "max_abs_ref_vs_truth": float(e_rt.max()), "tolerance": TOL, "checker": kind, "pass": ok}


def parity_block(workload, mdb):
    """GPU path (through the public host API) against the CPU reference and the float64 truth on samples of the
    workload's shape.  The oracle is the checker here, never the thing measured.  Tolerance:
    1e-5 nm absolute or 1e-4 relative (BASELINE.json north_star), or at least as close to the truth as the reference."""
    from oracle import oracle as O
    F_full, N, _ = WORKLOADS[workload]
    kind = "reference" if O.ref_available() else "port"
    out = {}
    if workload.startswith("allpairs"):
        # iid (large RMSD, no cancellation), MD-like (RMSD 0.25 nm on Rg 1 nm: G_a + G_b - 2 lambda cancels 30x, where the
        # tensor core's truncating accumulation would show) and MD-like in three basins 1.4 nm apart (clustering input:
        # pairs inside a basin far from frame 0 need the multi-reference operands)
        F, rows_n = 4000, 9
        sets = (("iid", O.synth_iid(F, N, seed=14)), ("md_like", O.synth_md(F, N, seed=14, rg=1.0, sigma=0.1)),
                ("md_like_multibasin", O.synth_md_basins(F, N, 3, seed=15, rg=1.0, sigma=0.1, separation=1.4)[0]))
        for name, X in sets:
            D = mdb.rmsd_matrix(mdb.Trajectory(X.copy()))
            rows = np.linspace(0, F - 1, rows_n).astype(int)   # three rows in every third of the trajectory
            Xc = X.copy()
            tr = O.center_and_trace(Xc, kind)
            ref = np.stack([O.one_vs_many_centered(Xc, tr, Xc[i], tr[i], impl=kind) for i in rows])
            truth = np.stack([O.truth_rmsd_batch(X, X[i]) for i in rows])
            m = np.ones((rows_n, F), bool); m[np.arange(rows_n), rows] = False  # a frame against itself: noise floor
            out[name] = _verdict(D[rows][m], ref[m], truth[m], f"{rows_n} rows of the {F}x{F} matrix, {N} atoms, {name} frames", kind)
    elif workload == "superpose":
        F = 512
        X = O.synth_md(F, N, seed=14)  # MD-like frames: rotations are well conditioned
        idx = np.arange(0, N, 5)
        t = mdb.Trajectory(X.copy())
        t.superpose(mdb.Trajectory(X.copy()), 0, atom_indices=idx)
        ref = O.superpose(X, X, 0, idx, impl=kind)
        truth = O.truth_superpose(X, X, 0, idx)[0]
        out["md_like"] = _verdict(t.xyz, ref, truth, f"superposed coordinates of {F} MD-like frames x {N} atoms, every 5th atom aligned", kind)
    else:
        F = int(min(F_full, max(64, min(10_000, 2.4e8 // (N * 12)))))
        # iid frames (strict parity is meaningful at every N) and MD-like frames (realistic cancellation; the reference
        # itself drifts from the float64 truth there, SURVEY.md Appendix C, hence the three-way verdict)
        for name, X in (("iid", O.synth_iid(F, N, seed=14)), ("md_like", O.synth_md(F, N, seed=14))):
            got = mdb.rmsd(mdb.Trajectory(X.copy()), mdb.Trajectory(X.copy()), 0)
            ref = O.rmsd(X, X, 0, impl=kind)
            truth = O.truth_rmsd_batch(X, X[0])
            out[name] = _verdict(got[1:], ref[1:], truth[1:], f"md.rmsd(t,t,0) on {F} {name} frames x {N} atoms (frame 0 itself left out)", kind)
    out["pass"] = bool(all(v["pass"] for v in out.values()))
    return out


def reference_arm(args, config):
    steps = max(1, min(args.steps, 5))
    info = cpu_reference_run("ovm", steps, 1)
    line = {"impl": "reference", "metric": METRIC, "value": info["value"], "unit": "rmsd/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": 1, "ms_per_step": info["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "cpu_baseline": info,
            "e2e": {"value": info["value"], "unit": "rmsd/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if not args.no_subs:
        sub = {}
        for w in ("superpose", "allpairs_20k", "ovm25k"):
            try:
                c = cpu_reference_run(w, 2, 1)
                sub[w] = {"value": c["value"], "unit": "rmsd/s", "config": {"workload": WORKLOADS[w][2]}, "cpu_baseline": c}
            except Exception as e:  # noqa: BLE001
                sub[w] = {"error": repr(e)}
        sub["allpairs_100k"] = sub.get("allpairs_20k")
        line["sub"] = sub
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class Ctx:
    pass


def sync_all(c):
    c.torch.cuda.synchronize(c.dev)
    if c.world > 1:
        c.dist.barrier()
        c.torch.cuda.synchronize(c.dev)


def max_over_ranks(c, x):
    t = c.torch.tensor([x], dtype=c.torch.float64, device=c.dev)
    if c.world > 1:
        c.dist.all_reduce(t, op=c.dist.ReduceOp.MAX)
    return float(t.item())


def timed_region(c, step, steps, warmup):
    """W untimed steps, barrier + synchronize, exactly K steps between two CUDA events on the launching stream,
    synchronize + barrier, max over ranks.  step(record) may append (start, stop) event pairs around its dominant kernel."""
    torch = c.torch
    events = []
    for _ in range(warmup):
        step(None)
    sync_all(c)
    sampler = ClockSampler(c.local_rank)
    sampler.start()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        step(events)
    t1.record()
    sync_all(c)
    clocks = sampler.result()
    ms = max_over_ranks(c, t0.elapsed_time(t1)) / steps
    kern_ms = statistics.mean(a.elapsed_time(b) for a, b in events) if events else None
    return ms, kern_ms, clocks


def bracket(torch, events):
    """Context manager recording a CUDA-event pair around the dominant kernel of a step."""
    class _B:
        def __enter__(self):
            if events is not None:
                self.e0 = torch.cuda.Event(enable_timing=True); self.e1 = torch.cuda.Event(enable_timing=True)
                self.e0.record()

        def __exit__(self, *a):
            if events is not None:
                self.e1.record()
                events.append((self.e0, self.e1))
    return _B()


def host_traj(mdb, arr):
    t = mdb.Trajectory.__new__(mdb.Trajectory)
    t.topology, t._xyz, t._rmsd_traces = None, arr, None   # no copy on the way in
    return t


def e2e_measure(c, fn, units, h2d, d2h, api, steps, memory):
    """fn() = one call of the public API on host buffers, returning a host scalar read from its result."""
    fn()
    sync_all(c)
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    c.torch.cuda.synchronize(c.dev)
    ms = max_over_ranks(c, (time.perf_counter() - t0) * 1e3 / steps)
    return {"value": units * c.world / (ms * 1e-3), "unit": "rmsd/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "ms_per_step": ms, "steps": steps, "api": api, "host_memory": memory, "h2d_GBs_per_gpu": h2d / ms / 1e6}


def host_copies(c, dev_xyz, N, pinned):
    """The device-resident synthetic frames as a host (F,N,3) float32 array: pageable numpy or page-locked."""
    torch = c.torch
    F = dev_xyz.shape[0]
    if pinned:
        host = torch.empty((F, N, 3), dtype=torch.float32, pin_memory=True)
    else:
        host = torch.from_numpy(np.empty((F, N, 3), dtype=np.float32))
    host.copy_(dev_xyz[:, :N, :])
    torch.cuda.synchronize(c.dev)
    return host


def run_ovm(c, name, steps, warmup, full):
    """md.rmsd(traj, traj, 0) on a device-resident shard; e2e through mdb.rmsd(host Trajectory)."""
    torch, mdb, L, capi = c.torch, c.mdb, c.L, c.capi
    F, N, desc = WORKLOADS[name]
    dt = mdb.DeviceTrajectory.synthetic_iid(F, N, seed=1000 * 2 + c.rank, device=c.dev)
    out = torch.empty(F, dtype=torch.float32, device=c.dev)
    scratch = torch.empty(int(L.b200rmsd_scratch_bytes(F, N)), dtype=torch.uint8, device=c.dev)
    ref = torch.empty(dt.n_pad * 3, dtype=torch.float32, device=c.dev)
    stats = torch.empty(capi.REFSTATS_BYTES, dtype=torch.uint8, device=c.dev)
    n_seg = 1 if N <= 4096 else -(-((N + 3) // 4) // 1024)
    launches = 2 + (1 if n_seg > 1 else 0)

    def step(events):
        # md.rmsd(traj, traj, 0): prepare the reference frame, then the streaming kernel
        capi.check(L.b200rmsd_prepare_reference_dev(dt.xyz_dev.data_ptr(), None, N, 1, 0.0, ref.data_ptr(), stats.data_ptr(),
                                                    c.stream), "prepare_reference")
        with bracket(torch, events):
            capi.check(L.b200rmsd_rmsd_dev(dt.xyz_dev.data_ptr(), F, N, dt.frame_stride, None, N, ref.data_ptr(),
                                           stats.data_ptr(), None, 0, out.data_ptr(), None, None, None, scratch.data_ptr(),
                                           scratch.numel(), c.stream), "rmsd_dev")
    ms, kern_ms, clocks = timed_region(c, step, steps, warmup)
    algo = 12.0 * N * F
    peak, peak_src = hbm_peak()
    rec = {"metric": METRIC, "value": F * c.world / (ms * 1e-3), "unit": "rmsd/s", "n_gpus": c.world, "steps": steps,
           "warmup": warmup, "ms_per_step": ms, "scaling": "weak", "dtype": "f32", "data": "synthetic",
           "config": {"workload": desc, "frames_per_gpu": F, "n_atoms": N, "frame": 0,
                      "l2_policy": "inputs (%.1f GB per GPU) are larger than the 126 MB L2" % (F * N * 12 / 1e9)},
           "clocks": clocks, "gpu_launches": launches * steps,
           "roofline": {"bound": "hbm", "kernel": "ovm_tma_kernel" + (" + ovm_finish_kernel" if n_seg > 1 else ""),
                        "achieved": algo / (kern_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": algo / (kern_ms * 1e-3) / 1e9 / peak, "traffic": ncu_fact(name), "peak_source": peak_src,
                        "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": algo}}
    if full:
        e2e_steps = 3
        staging = ("streamed through cache-resident page-locked slots: %d ranks share the host" % c.world) if c.world > 1 \
            else "whole 16 MB chunks through three page-locked lanes"
        for memory, pinned in (("pageable (%s)" % staging, False), ("pinned", True)):
            if pinned and name != "ovm":
                continue
            host = host_copies(c, dt.xyz_dev, N, pinned)
            ht = host_traj(mdb, host.numpy())
            ref_host = mdb.Trajectory(host.numpy()[:1].copy())
            rec["e2e" if not pinned else "e2e_pinned"] = e2e_measure(
                c, lambda: float(mdb.rmsd(ht, ref_host, 0)[-1]), F, F * N * 12 + N * 12, F * 4,
                "mdtraj_b200.rmsd(host Trajectory) -> b200rmsd_rmsd_host_multi", e2e_steps, memory)
            del host, ht
    del dt, out, scratch
    torch.cuda.empty_cache()
    if full and c.rank == 0 and c.world == 1:
        rec["cpu_baseline"] = _try(lambda: cpu_reference_run(name, 3, 1))
        rec["parity"] = _try(lambda: parity_block(name, mdb))
    return rec


def superpose_kernel_name(c, n_atoms, n_sel):
    """Which single-pass kernel b200rmsd_superpose_dev launches for this frame size (host-side geometry hook)."""
    import ctypes
    geo = (ctypes.c_int * 8)()
    kind = ctypes.CDLL(c.capi.LIB_PATH).b200rmsd_debug_fused_geometry(0, n_atoms, n_sel, 1, 1, geo)
    return {1: "frame_resident_kernel<OP_SUPERPOSE>", 2: "superpose_pipe_kernel (stage-pipelined, frames >= 30 KB)"}.get(
        kind, "two-pass: ovm kernel + apply_transform_kernel")


def run_superpose(c, steps, warmup, full):
    torch, mdb, L, capi = c.torch, c.mdb, c.L, c.capi
    F, N, desc = WORKLOADS["superpose"]
    dt = mdb.DeviceTrajectory.synthetic_iid(F, N, seed=3000 + c.rank, device=c.dev)
    idx_np = np.arange(0, N, 5, dtype=np.int32)
    idx = torch.from_numpy(idx_np).to(c.dev)
    out = torch.empty(F, dtype=torch.float32, device=c.dev)
    rot = torch.empty((F, 9), dtype=torch.float32, device=c.dev)
    scratch = torch.empty(int(L.b200rmsd_scratch_bytes(F, N)), dtype=torch.uint8, device=c.dev)
    ref = torch.empty(((len(idx_np) + 3) // 4 * 4) * 3, dtype=torch.float32, device=c.dev)
    stats = torch.empty(capi.REFSTATS_BYTES, dtype=torch.uint8, device=c.dev)
    ref_frame = dt.xyz_dev[0].clone()

    def step(events):
        capi.check(L.b200rmsd_prepare_reference_dev(ref_frame.data_ptr(), idx.data_ptr(), len(idx_np), 1, 0.0, ref.data_ptr(),
                                                    stats.data_ptr(), c.stream), "prepare_reference")
        with bracket(torch, events):
            capi.check(L.b200rmsd_superpose_dev(dt.xyz_dev.data_ptr(), F, N, dt.frame_stride, idx.data_ptr(), len(idx_np),
                                                ref.data_ptr(), stats.data_ptr(), out.data_ptr(), rot.data_ptr(), None,
                                                scratch.data_ptr(), scratch.numel(), c.stream), "superpose_dev")
    ms, kern_ms, clocks = timed_region(c, step, steps, warmup)
    algo = 24.0 * N * F
    peak, peak_src = hbm_peak()
    rec = {"metric": "frames superposed/sec", "value": F * c.world / (ms * 1e-3), "unit": "frames/s", "n_gpus": c.world,
           "steps": steps, "warmup": warmup, "ms_per_step": ms, "scaling": "weak", "dtype": "f32", "data": "synthetic",
           "config": {"workload": desc, "frames_per_gpu": F, "n_atoms": N, "n_aligned": len(idx_np), "frame": 0,
                      "l2_policy": "inputs (%.1f GB per GPU) are larger than the 126 MB L2" % (F * N * 12 / 1e9)},
           "clocks": clocks, "gpu_launches": 2 * steps,
           "roofline": {"bound": "hbm", "kernel": superpose_kernel_name(c, N, len(idx_np)), "achieved": algo / (kern_ms * 1e-3) / 1e9,
                        "peak": peak, "unit": "GB/s", "frac": algo / (kern_ms * 1e-3) / 1e9 / peak,
                        "traffic": ncu_fact("superpose"), "peak_source": peak_src, "kernel_ms": kern_ms,
                        "algorithmic_bytes_per_launch": algo}}
    if full:
        idx_host = np.arange(0, N, 5)
        for memory, pinned in (("pageable", False), ("pinned", True)):
            host = host_copies(c, dt.xyz_dev, N, pinned)
            ht = host_traj(mdb, host.numpy())
            ref_host = mdb.Trajectory(host.numpy()[:1].copy())

            def call():
                ht.superpose(ref_host, 0, atom_indices=idx_host)
                return float(ht.xyz[0, 0, 0])
            r = e2e_measure(c, call, F, F * N * 12 + N * 12, F * N * 12, "Trajectory.superpose -> b200rmsd_superpose_host_multi",
                            2, memory)
            r["unit"] = "frames/s"
            rec["e2e" if not pinned else "e2e_pinned"] = r
            del host, ht
    del dt, out, rot, scratch
    torch.cuda.empty_cache()
    if full and c.rank == 0:
        rec["cpu_baseline"] = _try(lambda: cpu_reference_run("superpose", 2, 1))
        rec["parity"] = _try(lambda: parity_block("superpose", mdb))
    return rec


def measure_tf32_peak(c):
    """Dense TF32 GEMM throughput of this GPU, measured the way MEASURED_PEAKS.json measures bf16: torch.matmul 8192^3
    (cuBLAS, allow_tf32), best of 10 (burst) and back to back for ~1.5 s (sustained).  cuBLAS is the yardstick only."""
    torch = c.torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn((n, n), device=c.dev); b = torch.randn((n, n), device=c.dev)
        for _ in range(3):
            a @ b
        torch.cuda.synchronize(c.dev)
        best = 1e9
        for _ in range(10):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); a @ b; e1.record(); torch.cuda.synchronize(c.dev)
            best = min(best, e0.elapsed_time(e1))
        reps = max(10, int(1500 / best))
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            a @ b
        e1.record(); torch.cuda.synchronize(c.dev)
        sustained = e0.elapsed_time(e1) / reps
        flop = 2.0 * n ** 3
        return {"burst_tflops": flop / best / 1e9, "sustained_tflops": flop / sustained / 1e9,
                "how": "torch.matmul fp32 inputs with allow_tf32 (cuBLAS TF32), 8192^3, best of 10 / back to back %d" % reps}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def allpairs_tiles(c, F):
    """Tiles (40 x 48 frames) a single-GPU full-matrix launch computes: each unordered pair once (host-side walk)."""
    import ctypes
    hook = ctypes.CDLL(c.capi.LIB_PATH).b200rmsd_debug_allpairs_tiles
    hook.restype = ctypes.c_longlong
    hook.argtypes = [ctypes.c_longlong] * 4 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p]
    return int(hook(0, F, 0, F, 0, None, 0, None))


def run_allpairs(c, name, steps, warmup, full, tf32):
    """All-pairs matrix of F frames: one GPU computes it whole (each unordered pair once, mirrored); several GPUs shard
    row blocks after an NCCL broadcast of the frames and exchange transposed blocks (strong scaling)."""
    torch, mdb = c.torch, c.mdb
    from mdtraj_b200 import allpairs as AP
    from mdtraj_b200 import distributed as DD
    F, N, desc = WORKLOADS[name]
    dt = mdb.DeviceTrajectory.synthetic_iid(F, N, seed=2000, device=c.dev)   # the same frames on every rank
    r0, r1 = DD.shard_bounds(F, c.rank, c.world)
    out = torch.empty((r1 - r0, F), dtype=torch.float32, device=c.dev) if c.world == 1 else None
    state = {}

    def step(events):
        with bracket(torch, events):
            if c.world == 1:
                prep = AP.prepare(dt, None)
                AP.rows(prep, 0, F, out=out)
            else:  # frames broadcast from rank 0 over NCCL, symmetric block plan, transposed blocks exchanged
                _, _, blk = DD.rmsd_matrix_sharded(dt, None, broadcast=True, symmetric=True)
                state["blk"] = blk
                state["peer"] = any(e.ok and e.block.data_ptr() == blk.data_ptr() for e in DD._EXCHANGES.values())
    ms, kern_ms, clocks = timed_region(c, step, steps, warmup)
    tiles = allpairs_tiles(c, F)
    kpad = (N + 31) // 32 * 32
    # per GPU: every unordered pair of row blocks is computed on exactly one rank; iid frames walk no augmentation K blocks
    issued = tiles / c.world * 128 * 144 * kpad * 2 * 3 / (ms * 1e-3) / 1e12
    rec = {"metric": METRIC, "value": float(F) * F / (ms * 1e-3), "unit": "rmsd/s", "n_gpus": c.world, "steps": steps,
           "warmup": warmup, "ms_per_step": ms, "scaling": "strong", "dtype": "tf32x3 products, f32 accumulate, f64 solve",
           "data": "synthetic",
           "config": {"workload": desc, "frames_total": F, "n_atoms": N,
                      "parallelism": "1 GPU" if c.world == 1 else
                      "frames NCCL-broadcast, %d row blocks, symmetric block plan, transposed blocks %s" % (
                          c.world, "written into their owner's row block over NVLink by the computing kernel (peer memory)"
                          if state.get("peer") else "sent over NVLink with NCCL send/recv"),
                      "l2_policy": "the %.1f GB of matrix written per GPU and step flushes the 126 MB L2 between steps; the "
                                   "operands (%.2f GB) are meant to stay L2-resident inside a step" % (
                                       F * F * 4 / 1e9 / c.world, F * kpad * 3 * 4 * 4 / 1e9)},
           "clocks": clocks, "gpu_launches": 15 * steps if c.world == 1 else None,
           "roofline": {"bound": "tensor", "kernel": "allpairs_tc144_kernel", "achieved": issued,
                        "peak": tf32["burst_tflops"], "unit": "TFLOP/s", "frac": issued / tf32["burst_tflops"],
                        "frac_of_sustained": issued / tf32["sustained_tflops"], "peak_source": "measured in this run: " + tf32["how"],
                        "tensor_pipe_active_pct_ncu": ncu_fact("allpairs_tensor_pipe_active_pct"),
                        "traffic": ncu_fact("allpairs"), "step_ms": ms, "useful_tflops": float(F) * F * 18 * N / (ms * 1e-3) / 1e12,
                        "tiles_per_matrix": tiles, "mma_shape": [128, 144, 8], "per": "GPU",
                        "note": "achieved = tensor flops issued (3 tf32 MMAs per K-step over the tiles computed, each unordered "
                                "pair once) / whole step time incl. reference selection, operand preparation and, at N > 1, "
                                "broadcast + exchange; useful = 18 A flops per reported pair"}}
    if c.world == 1:
        info = AP.prepare(dt, None).info()
        rec["config"]["references_chosen"] = info["n_refs"]
    if full and c.world == 1:
        host = host_copies(c, dt.xyz_dev, N, False)
        ht = host_traj(mdb, host.numpy())
        host_out = torch.empty((F, F), dtype=torch.float32, pin_memory=True)

        def call():
            mdb.rmsd_matrix(ht, out=host_out.numpy())  # row blocks, copies overlapped with compute
            return float(host_out[0, 1])
        rec["e2e"] = e2e_measure(c, call, float(F) * F, F * N * 12, F * F * 4,
                                 "mdtraj_b200.rmsd_matrix(host Trajectory, out=page-locked ndarray)", 2, "pageable in, page-locked out")
        # MD-like frames in three basins: the input clustering sees (references are chosen, augmentation blocks walked)
        from oracle import oracle as O
        Xb = O.synth_md_basins(F, N, 3, seed=77, rg=1.0, sigma=0.1, separation=1.4)[0]
        db = mdb.DeviceTrajectory.from_host(Xb, device=c.dev)

        def step_b(events):
            prep = AP.prepare(db, None)
            AP.rows(prep, 0, F, out=out)
        ms_b, _, _ = timed_region(c, step_b, max(3, steps // 2), 2)
        rec["md_like_multibasin"] = {"value": float(F) * F / (ms_b * 1e-3), "unit": "rmsd/s", "ms_per_step": ms_b,
                                     "info": AP.prepare(db, None).info()}
        del host, ht, host_out, db
    del dt, out
    state.clear()
    torch.cuda.empty_cache()
    if full and c.rank == 0 and c.world == 1:
        rec["cpu_baseline"] = _try(lambda: cpu_reference_run(name, 2, 1))
        rec["parity"] = _try(lambda: parity_block(name, mdb))
    return rec


def _try(fn):
    try:
        return fn()
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--only", default=None, choices=sorted(WORKLOADS), help="development: run one workload as the headline")
    ap.add_argument("--no-subs", action="store_true", help="headline only")
    ap.add_argument("--no-e2e", action="store_true", help="development: device-resident numbers only")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    F, N, desc = WORKLOADS["ovm"]
    config = {"workload": desc, "frames_per_gpu": F, "n_atoms": N, "frame": 0,
              "l2_policy": "inputs (%.1f GB per GPU) are larger than the 126 MB L2" % (F * N * 12 / 1e9)}

    if args.impl == "reference":   # CPU only, rank 0 only
        if rank == 0:
            reference_arm(args, config)
        return

    import torch
    import torch.distributed as dist

    import mdtraj_b200 as mdb
    from mdtraj_b200 import _capi
    from mdtraj_b200.device import _stream_ptr

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; mdtraj_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    c = Ctx()
    c.torch, c.dist, c.mdb, c.capi, c.L = torch, dist, mdb, _capi, _capi.lib()
    c.rank, c.world, c.local_rank = rank, world, local_rank
    c.dev = torch.device("cuda", local_rank)
    c.stream = _stream_ptr(torch, c.dev)
    mdb.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=c.dev)
        mdb.set_host_pipeline(copy_threads=max(2, host_threads() // world - 1))   # the ranks share the host's cores

    full = not args.no_e2e
    sub_steps = max(3, min(args.steps, 10))
    if args.only:
        if args.only.startswith("allpairs"):
            line = run_allpairs(c, args.only, args.steps, args.warmup, full, measure_tf32_peak(c))
        elif args.only == "superpose":
            line = run_superpose(c, args.steps, args.warmup, full)
        else:
            line = run_ovm(c, args.only, args.steps, args.warmup, full)
        line["higher_is_better"] = True
        line["vs_baseline"] = None
    else:
        line = run_ovm(c, "ovm", args.steps, args.warmup, full)
        line["higher_is_better"] = True
        line["vs_baseline"] = None
        if not args.no_subs:
            tf32 = measure_tf32_peak(c)
            sub = {}

            def rested(fn):
                # a few idle seconds before each record: the board's power limiter keeps the SM clock down for a while
                # after a tensor-core workload (the 2-GPU run of this bench measured ovm25k at 1305 MHz, 0.77 of the
                # HBM peak, straight after the 100k x 100k matrix; 1.0 when it runs first), and the HBM-bound records go
                # before the tensor-bound ones for the same reason
                time.sleep(3.0)
                return _try(fn)
            sub["ovm25k"] = rested(lambda: run_ovm(c, "ovm25k", sub_steps, 3, full and world == 1))
            if world == 1:
                sub["superpose"] = rested(lambda: run_superpose(c, sub_steps, 3, full))
                sub["allpairs_20k"] = rested(lambda: run_allpairs(c, "allpairs_20k", sub_steps, 3, full, tf32))
            sub["allpairs_100k"] = rested(lambda: run_allpairs(c, "allpairs_100k", 3, 3, False, tf32))
            line["sub"] = sub
            line["tf32_peak"] = tf32
            line["gpu_launches"] = int(line["gpu_launches"] + sum((s.get("gpu_launches") or 0) for s in sub.values()
                                                                  if isinstance(s, dict)))
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        from mdtraj_b200 import distributed as DD
        DD.release_peer_exchanges()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- frame-pair RMSDs per second on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one pass of the hot path over one batch of synthetic input.  Default workload =
BASELINE.json configs[1]: one-vs-many md.rmsd, 1,000,000 frames x 1,000 atoms, precentered=False,
reference = frame 0.  With N > 1 (torchrun, one rank per GPU) every rank owns its own 1M-frame shard
(frames are independent: no data-path collective, weak scaling) and `value` is the aggregate.

  value      device-resident throughput: inputs already in HBM, CUDA-event timed, max over ranks
  e2e        the same metric through the public API mdb.rmsd(host_traj, ...) with PINNED HOST buffers:
             H2D of every frame and D2H of the result inside the timed region
  roofline   dominant kernel (ovm_tma_kernel): algorithmic bytes 12*N per frame / CUDA-event launch time
  cpu_baseline  the compiled reference (oracle/_ref) on this box's host cores, bounded sample
  parity     GPU path vs the CPU reference vs the float64 truth on a sample of the same shape (outside the timed region)

--impl reference times the reference's own CPU implementation (oracle/_ref = its C++ sources compiled
in place + our OpenMP loop shell; falls back to the C port oracle/liboracle.so if _ref was never built).
Other workloads (superpose, allpairs, ovm25k, ala2) exist for profiling; the driver uses the default.
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
    "allpairs": (20_000, 300, "all-pairs RMSD matrix: 20k x 20k frames x 300 atoms (BASELINE configs[3] shape, reduced F)"),
    "ala2": (100, 22, "md.rmsd 100 frames x 22 atoms (BASELINE configs[0] shape, synthetic)"),
}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return float(p["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def tf32_peak():
    """Dense TF32 tensor peak in TFLOP/s.  MEASURED_PEAKS.json carries the cuBLAS bf16 number only; kind::tf32
    issues at exactly half the kind::f16 rate (K=8 vs K=16 per instruction at the same cycle count), so the
    measured bf16 burst figure / 2 is used and labelled as such."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return float(p["bf16_tflops"]) / 2.0, "MEASURED_PEAKS.json bf16_tflops / 2 (tf32 = half the bf16 rate; of measured)"
    return 1590.0 / 2.0, "B200_PROFILING.md fallback 1.59 PFLOP/s bf16 / 2 (of fallback)"


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
def cpu_reference_run(workload, steps, warmup, sample_frames=None):
    """Time the reference CPU path on a bounded sample of `workload`.  Returns (value, info)."""
    from oracle import oracle as O
    F_full, N, _ = WORKLOADS[workload]
    kind = "reference" if O.ref_available() else "port"
    threads = 1
    if kind == "reference":
        L = O.ref_lib()
        threads = int(L.refloops_max_threads())
    if workload == "allpairs":
        F = 20_000 if sample_frames is None else sample_frames
        rows = 8
        X = O.synth_iid(F, N, seed=4)
        tr = O.center_and_trace(X, kind)

        def step():
            for i in range(rows):
                O.one_vs_many_centered(X, tr, X[i], tr[i], impl=kind)
        units = rows * F
        sample = f"{rows} rows of md.rmsd(t,t,i,precentered=True) at F={F}, N={N} after center_coordinates()"
    elif workload == "superpose":
        F = 2000 if sample_frames is None else sample_frames
        X = O.synth_iid(F, N, seed=4)
        idx = np.arange(0, N, 5)

        def step():
            O.superpose(X, X, 0, idx, impl=kind)
        units = F
        sample = f"Trajectory.superpose restated (numpy glue + reference C kernels) on {F} x {N}, 1000-atom subset"
    else:
        budget_bytes = 1.2e9
        F = int(min(F_full, max(64, budget_bytes // (N * 12)))) if sample_frames is None else sample_frames
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
    info = {"value": units / dt, "unit": "rmsd/s", "cores": threads, "kind": kind, "sample": sample,
            "host_cpus": os.cpu_count(), "ms_per_step": dt * 1e3}
    return units / dt, info


def parity_block(workload, mdb):
    """GPU path (through the public host API) against the CPU reference and the float64 truth on a sample of the
    workload's shape (SURVEY.md section 8(d) "parity checks in the bench").  The oracle is the checker here, never the
    thing measured.  Tolerance: 1e-5 nm absolute or 1e-4 relative (BASELINE.json north_star)."""
    from oracle import oracle as O
    F_full, N, _ = WORKLOADS[workload]
    kind = "reference" if O.ref_available() else "port"
    tol_abs, tol_rel = 1e-5, 1e-4

    def verdict(got, ref, truth, what, n):
        got, ref, truth = (np.asarray(v, dtype=np.float64) for v in (got, ref, truth))
        e_gr, e_gt, e_rt = np.abs(got - ref), np.abs(got - truth), np.abs(ref - truth)
        ok = bool(np.all((e_gr <= np.maximum(tol_abs, tol_rel * np.abs(ref))) | (e_gt <= np.maximum(tol_abs, e_rt))))
        return {"what": what, "n": int(n), "max_abs_gpu_vs_ref": float(e_gr.max()), "max_abs_gpu_vs_truth": float(e_gt.max()),
                "max_abs_ref_vs_truth": float(e_rt.max()), "tolerance": "1e-5 nm abs or 1e-4 rel", "checker": kind,
                "pass": ok}
    if workload == "allpairs":
        # two kinds of frames: iid (large RMSD, no cancellation) and MD-like (RMSD 0.25 nm on
        # Rg 1 nm: G_a + G_b - 2 lambda cancels 30x, which is where the tensor core's truncating accumulation would show)
        F, rows = 4000, 8
        out = None
        for name, X in (("iid", O.synth_iid(F, N, seed=14)), ("MD-like", O.synth_md(F, N, seed=14, rg=1.0, sigma=0.1))):
            D = mdb.rmsd_matrix(mdb.Trajectory(X.copy()))
            Xc = X.copy()
            tr = O.center_and_trace(Xc, kind)
            ref = np.stack([O.one_vs_many_centered(Xc, tr, Xc[i], tr[i], impl=kind) for i in range(rows)])
            truth = np.stack([O.truth_rmsd_batch(X, X[i]) for i in range(rows)])
            m = np.ones((rows, F), bool); m[np.arange(rows), np.arange(rows)] = False  # a frame against itself: noise floor
            v = verdict(D[:rows][m], ref[m], truth[m], f"{rows} rows of the {F}x{F} matrix, {N} atoms, {name} frames", m.sum())
            if out is None:
                out = v
            else:
                out["md_like"] = v
                out["pass"] = bool(out["pass"] and v["pass"])
        return out
    if workload == "superpose":
        F = 512
        X = O.synth_md(F, N, seed=14)  # MD-like frames: rotations are well conditioned
        idx = np.arange(0, N, 5)
        t = mdb.Trajectory(X.copy())
        t.superpose(mdb.Trajectory(X.copy()), 0, atom_indices=idx)
        ref = O.superpose(X, X, 0, idx, impl=kind)
        truth = O.truth_superpose(X, X, 0, idx)[0]
        return verdict(t.xyz, ref, truth, f"superposed coordinates of {F} MD-like frames x {N} atoms, every 5th atom aligned",
                       F * N * 3)
    F = int(min(F_full, max(64, min(10_000, 2.4e8 // (N * 12)))))
    out = None
    # iid frames (strict parity is meaningful at every N) and MD-like frames (realistic cancellation; the reference
    # itself drifts from the float64 truth there, SURVEY.md Appendix C, hence the three-way verdict)
    for name, X in (("iid", O.synth_iid(F, N, seed=14)), ("MD-like", O.synth_md(F, N, seed=14))):
        got = mdb.rmsd(mdb.Trajectory(X.copy()), mdb.Trajectory(X.copy()), 0)
        ref = O.rmsd(X, X, 0, impl=kind)
        truth = O.truth_rmsd_batch(X, X[0])
        v = verdict(got[1:], ref[1:], truth[1:], f"md.rmsd(t,t,0) on {F} {name} frames x {N} atoms (frame 0 itself left out)",
                    F - 1)
        if out is None:
            out = v
        else:
            out["md_like"] = v
            out["pass"] = bool(out["pass"] and v["pass"])
    return out


# ------------------------------------------------------------------------------------------------
def ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed ncu capture."""
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(tpath):
        return None
    with open(tpath) as fh:
        return json.load(fh).get(workload)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ovm", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=None, help="override frames per GPU (development)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    F, N, desc = WORKLOADS[args.workload]
    if args.frames:
        F = args.frames
        if args.workload == "allpairs":
            desc = ("all-pairs RMSD matrix: %dk x %dk frames x %d atoms (BASELINE configs[3]%s)"
                    % (F // 1000, F // 1000, N, "" if F == 100_000 else " shape, F overridden"))
    config = {"workload": desc, ("frames_total" if args.workload == "allpairs" else "frames_per_gpu"): F, "n_atoms": N,
              "frame": 0,
              "l2_policy": ("the %.1f GB matrix written by every step flushes the 126 MB L2 between steps; the operands "
                            "(%.2f GB) are meant to stay L2-resident inside a step" % (F * F * 4 / 1e9 / max(world, 1),
                                                                                       F * N * 12 / 1e9))
              if args.workload == "allpairs" else
              "inputs (%.1f GB per GPU) are larger than the 126 MB L2" % (F * N * 12 / 1e9)}

    # ---------------- reference arm: CPU only, rank 0 only
    if args.impl == "reference":
        if rank != 0:
            return
        val, info = cpu_reference_run(args.workload, max(1, min(args.steps, 5)), 1)
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "rmsd/s", "n_gpus": args.gpus,
                "steps": max(1, min(args.steps, 5)), "warmup": 1, "ms_per_step": info["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "cpu_baseline": info,
                "e2e": {"value": val, "unit": "rmsd/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist

    import mdtraj_b200 as mdb
    from mdtraj_b200 import _capi
    from mdtraj_b200.device import _Scratch, _stream_ptr

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; mdtraj_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    mdb.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    L = _capi.lib()
    stream = _stream_ptr(torch, dev)
    dt = mdb.DeviceTrajectory.synthetic_iid(F, N, seed=1000 * 2 + rank, device=dev)
    launches_per_step = 0
    kernel_events = []

    if args.workload in ("ovm", "ovm25k", "ala2"):
        out = torch.empty(F, dtype=torch.float32, device=dev)
        scratch = _Scratch.get(torch, dev, L.b200rmsd_scratch_bytes(F, N))
        ref = torch.empty(dt.n_pad * 3, dtype=torch.float32, device=dev)
        stats = torch.empty(_capi.REFSTATS_BYTES, dtype=torch.uint8, device=dev)
        n_seg = 1 if N <= 4096 else -(-((N + 3) // 4) // 1024)
        launches_per_step = 2 + (1 if n_seg > 1 else 0)

        def step(record=False):
            # md.rmsd(traj, traj, 0): prepare the reference frame, then the streaming kernel
            _capi.check(L.b200rmsd_prepare_reference_dev(dt.xyz_dev.data_ptr(), None, N, 1, 0.0, ref.data_ptr(),
                                                         stats.data_ptr(), stream), "prepare_reference")
            if record:
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
            _capi.check(L.b200rmsd_rmsd_dev(dt.xyz_dev.data_ptr(), F, N, dt.frame_stride, None, N, ref.data_ptr(),
                                            stats.data_ptr(), None, 0, out.data_ptr(), None, None, None,
                                            scratch.data_ptr(), scratch.numel(), stream), "rmsd_dev")
            if record:
                e1.record()
                kernel_events.append((e0, e1))
        units_per_step = F
        algo_bytes = 12.0 * N * F
        dominant = "ovm_tma_kernel"
    elif args.workload == "superpose":
        idx_np = np.arange(0, N, 5, dtype=np.int32)
        idx = torch.from_numpy(idx_np).to(dev)
        out = torch.empty(F, dtype=torch.float32, device=dev)
        rot = torch.empty((F, 9), dtype=torch.float32, device=dev)
        scratch = _Scratch.get(torch, dev, L.b200rmsd_scratch_bytes(F, N))
        ref = torch.empty(((len(idx_np) + 3) // 4 * 4) * 3, dtype=torch.float32, device=dev)
        stats = torch.empty(_capi.REFSTATS_BYTES, dtype=torch.uint8, device=dev)
        ref_frame = dt.xyz_dev[0].clone()
        launches_per_step = 3

        def step(record=False):
            _capi.check(L.b200rmsd_prepare_reference_dev(ref_frame.data_ptr(), idx.data_ptr(), len(idx_np), 1, 0.0,
                                                         ref.data_ptr(), stats.data_ptr(), stream), "prepare_reference")
            if record:
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
            _capi.check(L.b200rmsd_superpose_dev(dt.xyz_dev.data_ptr(), F, N, dt.frame_stride, idx.data_ptr(),
                                                 len(idx_np), ref.data_ptr(), stats.data_ptr(), out.data_ptr(),
                                                 rot.data_ptr(), None, scratch.data_ptr(), scratch.numel(), stream),
                        "superpose_dev")
            if record:
                e1.record()
                kernel_events.append((e0, e1))
        units_per_step = F
        algo_bytes = 24.0 * N * F
        dominant = "frame_resident_kernel"
    else:  # allpairs
        from mdtraj_b200 import allpairs as AP
        rows_per_rank = F // world
        r0 = rank * rows_per_rank
        r1 = F if rank == world - 1 else r0 + rows_per_rank
        out = torch.empty((r1 - r0, F), dtype=torch.float32, device=dev)
        launches_per_step = 2

        from mdtraj_b200 import distributed as DD

        def step(record=False):
            if record:
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
            if world == 1:
                prep = AP.prepare(dt, None)
                AP.rows(prep, r0, r1, out=out)
            else:  # frames broadcast from rank 0 over NCCL, symmetric block plan, transposed blocks exchanged
                DD.rmsd_matrix_sharded(dt, None, broadcast=True, symmetric=True)
            if record:
                e1.record()
                kernel_events.append((e0, e1))
        units_per_step = (r1 - r0) * F
        algo_bytes = None
        dominant = "allpairs_tc144_kernel"

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_start = torch.cuda.Event(enable_timing=True); t_stop = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for _ in range(args.steps):
        step(record=True)
    t_stop.record()
    sync_all()
    clocks = sampler.result()
    elapsed_ms = t_start.elapsed_time(t_stop)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    total_units = units_per_step * world
    value = total_units / (ms_per_step * 1e-3)
    kern_ms = statistics.mean(a.elapsed_time(b) for a, b in kernel_events)

    # ---------------- end-to-end through the public API with pinned host buffers
    e2e = None
    if not args.no_e2e:
        host = torch.empty((F, N, 3), dtype=torch.float32, pin_memory=True)
        host.copy_(dt.xyz_dev[:, :N, :])
        torch.cuda.synchronize(dev)
        ht = mdb.Trajectory.__new__(mdb.Trajectory)
        ht.topology = None
        ht._xyz = host.numpy()  # pinned, C-contiguous float32: no copy is made on the way in
        ht._rmsd_traces = None
        ref_host = mdb.Trajectory(host.numpy()[:1].copy())
        e2e_steps = max(2, min(args.steps, 4))
        if args.workload == "superpose":
            idx_host = np.arange(0, N, 5)

            def e2e_step():
                ht.superpose(ref_host, 0, atom_indices=idx_host)
                return float(ht.xyz[0, 0, 0])
            h2d = F * N * 12 + N * 12
            d2h = F * N * 12
        elif args.workload == "allpairs":
            # host frames in, this rank's rows of the matrix out to page-locked host memory
            host_out = torch.empty((r1 - r0, F), dtype=torch.float32, pin_memory=True)

            def e2e_step():
                if world == 1:
                    mdb.rmsd_matrix(ht, out=host_out.numpy())  # row blocks, copies overlapped with compute
                else:
                    dte = mdb.DeviceTrajectory.from_host(ht.xyz, dev) if rank == 0 else \
                        mdb.DeviceTrajectory(torch.zeros_like(dt.xyz_dev), N)
                    _, _, blk = DD.rmsd_matrix_sharded(dte, None, broadcast=True, symmetric=True)
                    host_out.copy_(blk)
                return float(host_out[0, 1])
            h2d = F * N * 12
            d2h = (r1 - r0) * F * 4
        else:
            def e2e_step():
                return float(mdb.rmsd(ht, ref_host, 0)[-1])
            h2d = F * N * 12 + N * 12
            d2h = F * 4
        e2e_step()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize(dev)
        e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        te = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_ms = float(te.item())
        e2e = {"value": units_per_step * world / (e2e_ms * 1e-3), "unit": "rmsd/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "steps": e2e_steps,
               "api": {"superpose": "Trajectory.superpose -> b200rmsd_superpose_host",
                       "allpairs": "mdtraj_b200.rmsd_matrix(host Trajectory, out=page-locked ndarray)"}.get(
                           args.workload, "mdtraj_b200.rmsd(host Trajectory) -> b200rmsd_rmsd_host"),
               "h2d_GBs": h2d / e2e_ms / 1e6}
        del host

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    roofline = None
    if algo_bytes is not None:
        achieved = algo_bytes / (kern_ms * 1e-3) / 1e9
        traffic = ncu_traffic(args.workload)
        roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": algo_bytes}

    if args.workload == "allpairs":
        tpeak, tsrc = tf32_peak()
        kpad = (N + 31) // 32 * 32
        pairs_per_s = units_per_step / (kern_ms * 1e-3)
        # flops actually issued: three tf32 MMAs per K-step over the tiles the kernel computes, K padded to 32; a
        # single-rank full matrix computes each unordered pair once (tiles holding no pair j >= i are skipped)
        import ctypes
        from mdtraj_b200 import _capi
        hook = ctypes.CDLL(_capi.LIB_PATH).b200rmsd_debug_allpairs_tiles   # host-only: walks the kernel's tile order
        hook.restype = ctypes.c_longlong
        hook.argtypes = [ctypes.c_longlong] * 4 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p]
        tiles, mma_n = int(hook(0, F, 0, F, 0, None, 0, None)), 144      # 128x144 tiles = 40x48 frames
        kpad = (((N + 7) // 8 * 8) + 6 + 31) // 32 * 32                   # atoms + 6 augmentation columns, padded to 32
        if world > 1:
            tiles = tiles / world  # symmetric block plan: every unordered pair of row blocks on exactly one rank
        issued = tiles * 128 * mma_n * kpad * 2 * 3 / (kern_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": dominant, "achieved": issued, "peak": tpeak, "unit": "TFLOP/s",
                    "frac": issued / tpeak, "traffic": ncu_traffic("allpairs"), "peak_source": tsrc, "kernel_ms": kern_ms,
                    "useful_tflops": pairs_per_s * 18 * N / 1e12,
                    "tiles_per_launch": tiles, "mma_shape": [128, mma_n, 8],
                    "note": "achieved = tensor flops actually issued (3 tf32 MMAs per K-step over the computed "
                            "tiles; symmetric tiles computed once); useful = 18 * A flops per reported pair"}

    cpu = None
    parity = None
    if not args.no_cpu and args.gpus == 1:
        try:
            _, cpu = cpu_reference_run(args.workload, 3, 1)
        except Exception as e:  # noqa: BLE001
            cpu = {"error": repr(e)}
        try:
            parity = parity_block(args.workload, mdb)
        except Exception as e:  # noqa: BLE001
            parity = {"error": repr(e)}

    line = {"metric": METRIC, "value": value, "unit": "rmsd/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "roofline": roofline, "cpu_baseline": cpu, "parity": parity}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Sweep the geometry of the frame-resident kernel (frames per slot, groups, ring depth) for superpose / centring
(development aid).   python tools/fused_sweep.py N [N ...]   ->  one JSON line per (N, op, config)"""
import json
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch  # noqa: E402

import mdtraj_b200 as mdb  # noqa: E402
from mdtraj_b200 import _capi  # noqa: E402
from mdtraj_b200.device import _Scratch, _stream_ptr, prepare_reference  # noqa: E402

PEAK = 6540.8
dev = torch.device("cuda", 0)
L = _capi.lib()
quick = os.environ.get("SWEEP_QUICK") == "1"
if "SWEEP_PIPE_MIN" in os.environ:  # frames of at least this many bytes take the stage-pipelined superpose kernel
    import ctypes
    ctypes.CDLL(_capi.LIB_PATH).b200rmsd_debug_fused_pipe_min_bytes(int(os.environ["SWEEP_PIPE_MIN"]))


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def configs(N):
    out = [None]
    if quick:
        return out
    if os.environ.get("SWEEP_WIDE") == "1":  # many single-warp groups
        return out + [(f, G, G, 0) for G in (8, 10, 12, 16) for f in (1, 2, 3, 4, 6, 8)]
    for fpb in [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 16, 18, 20, 22, 24, 28, 32, 40, 48, 56, 64]:
        if fpb * N * 12 > 100000:
            continue
        for G in (4, 5, 8):
            for m in (1, 2):
                for lanes in ((0,) if N > 1000 else (0, 2, 4, 8, 16, 32)):
                    if m == 2 and lanes not in (0,):
                        continue
                    out.append((fpb, G, G * m, lanes))
    return out


for N in [int(a) for a in sys.argv[1:]]:
    n_pad = (N + 3) // 4 * 4
    F = max(1480, int(1.5e9 / (n_pad * 12)) // 1480 * 1480)
    dt = mdb.DeviceTrajectory.synthetic_iid(F, N, 1, dev)
    nbytes = F * n_pad * 12
    out = torch.empty(F, dtype=torch.float32, device=dev)
    rot = torch.empty((F, 9), dtype=torch.float32, device=dev)
    traces = torch.empty(F, dtype=torch.float32, device=dev)
    scratch = _Scratch.get(torch, dev, L.b200rmsd_scratch_bytes(F, N))
    stream = _stream_ptr(torch, dev)
    prep = prepare_reference(dt.xyz_dev[0].clone(), None, N, True)
    idx = torch.arange(0, N, 5, dtype=torch.int32, device=dev)
    prep5 = prepare_reference(dt.xyz_dev[0].clone(), idx, int(idx.numel()), True)

    def center():
        _capi.check(L.b200rmsd_center_trace_dev(dt.xyz_dev.data_ptr(), F, N, dt.frame_stride, traces.data_ptr(), stream), "center")

    def sup5():
        _capi.check(L.b200rmsd_superpose_dev(dt.xyz_dev.data_ptr(), F, N, dt.frame_stride, idx.data_ptr(), int(idx.numel()),
                                             prep5.ref.data_ptr(), prep5.stats.data_ptr(), out.data_ptr(), rot.data_ptr(), None,
                                             scratch.data_ptr(), scratch.numel(), stream), "sup5")

    def sup():
        _capi.check(L.b200rmsd_superpose_dev(dt.xyz_dev.data_ptr(), F, N, dt.frame_stride, None, N, prep.ref.data_ptr(),
                                             prep.stats.data_ptr(), out.data_ptr(), rot.data_ptr(), None, scratch.data_ptr(),
                                             scratch.numel(), stream), "sup")

    best = {}
    for cfg in configs(N):
        keys = ("B200RMSD_FUSED_FPB", "B200RMSD_FUSED_GROUPS", "B200RMSD_FUSED_NBUF", "B200RMSD_FUSED_LANES")
        for k in keys:
            os.environ.pop(k, None)
        if cfg:
            for k, v in zip(keys, cfg):
                os.environ[k] = str(v)
        for name, fn in (("center", center), ("sup5", sup5), ("sup", sup)):
            try:
                ms = timed(fn)
            except _capi.B200RMSDError as e:
                if e.code == _capi.EINVAL:  # geometry does not fit shared memory
                    continue
                raise
            gbs = 2 * nbytes / ms / 1e6
            rec = {"N": N, "F": F, "op": name, "cfg": cfg, "ms": round(ms, 4), "GBs": round(gbs), "frac": round(gbs / PEAK, 3)}
            print(json.dumps(rec), flush=True)
            if cfg is None:
                best[name + "_default"] = rec
            if name not in best or rec["GBs"] > best[name]["GBs"]:
                best[name] = rec
    print("BEST", json.dumps(best), flush=True)

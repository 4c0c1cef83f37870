"""Quick device-resident timing of the streaming kernels (development aid, not the bench contract).

    python tools/quick_time.py [F] [N] [reps]
"""
import json
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))

import torch  # noqa: E402

import mdtraj_b200 as mdb  # noqa: E402
from mdtraj_b200 import _capi  # noqa: E402
from mdtraj_b200.device import _Scratch, _stream_ptr, prepare_reference  # noqa: E402


def time_fn(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    print("  [ok warm-up]", getattr(fn, "__name__", "?"), file=sys.stderr, flush=True)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def main():
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    dev = torch.device("cuda", 0)
    dt = mdb.DeviceTrajectory.synthetic_iid(F, N, seed=1, device=dev)
    L = _capi.lib()
    prep = prepare_reference(dt.xyz_dev[0].clone(), None, N, True)
    out = torch.empty(F, dtype=torch.float32, device=dev)
    scratch = _Scratch.get(torch, dev, L.b200rmsd_scratch_bytes(F, N))
    stream = _stream_ptr(torch, dev)
    res = {"F": F, "N": N, "bytes": F * dt.n_pad * 12}

    def ovm():
        _capi.check(L.b200rmsd_rmsd_dev(dt.xyz_dev.data_ptr(), F, N, dt.frame_stride, None, N, prep.ref.data_ptr(),
                                        prep.stats.data_ptr(), None, 0, out.data_ptr(), None, None, None,
                                        scratch.data_ptr(), scratch.numel(), stream), "rmsd_dev")
    med, best = time_fn(ovm, reps)
    res["ovm_ms"] = med; res["ovm_GBs"] = res["bytes"] / med / 1e6; res["ovm_best_GBs"] = res["bytes"] / best / 1e6
    res["ovm_frames_per_s"] = F / med * 1e3

    traces = torch.empty(F, dtype=torch.float32, device=dev)

    def center():
        _capi.check(L.b200rmsd_center_trace_dev(dt.xyz_dev.data_ptr(), F, N, dt.frame_stride, traces.data_ptr(), stream), "center")
    med, best = time_fn(center, reps)
    res["center_ms"] = med; res["center_GBs_rw"] = 2 * res["bytes"] / med / 1e6

    def pre():
        _capi.check(L.b200rmsd_rmsd_dev(dt.xyz_dev.data_ptr(), F, N, dt.frame_stride, None, N, prep.ref.data_ptr(),
                                        prep.stats.data_ptr(), traces.data_ptr(), 1, out.data_ptr(), None, None, None,
                                        scratch.data_ptr(), scratch.numel(), stream), "rmsd_dev pre")
    med, best = time_fn(pre, reps)
    res["pre_ms"] = med; res["pre_GBs"] = res["bytes"] / med / 1e6

    rot = torch.empty((F, 9), dtype=torch.float32, device=dev)
    stride5 = torch.arange(0, N, 5, dtype=torch.int32, device=dev)
    prep5 = prepare_reference(dt.xyz_dev[0].clone(), stride5, int(stride5.numel()), True)

    def sup5():
        _capi.check(L.b200rmsd_superpose_dev(dt.xyz_dev.data_ptr(), F, N, dt.frame_stride, stride5.data_ptr(),
                                             int(stride5.numel()), prep5.ref.data_ptr(), prep5.stats.data_ptr(),
                                             out.data_ptr(), rot.data_ptr(), None, scratch.data_ptr(), scratch.numel(),
                                             stream), "superpose idx")
    med, best = time_fn(sup5, reps)
    res["superpose_idx5_ms"] = med; res["superpose_idx5_GBs_24N"] = 2 * res["bytes"] / med / 1e6

    def sup():
        _capi.check(L.b200rmsd_superpose_dev(dt.xyz_dev.data_ptr(), F, N, dt.frame_stride, None, N, prep.ref.data_ptr(),
                                             prep.stats.data_ptr(), out.data_ptr(), rot.data_ptr(), None,
                                             scratch.data_ptr(), scratch.numel(), stream), "superpose")
    med, best = time_fn(sup, reps)
    res["superpose_ms"] = med; res["superpose_GBs_24N"] = 2 * res["bytes"] / med / 1e6

    # plain copy for comparison (read + write)
    a = torch.empty(res["bytes"] // 4, dtype=torch.float32, device=dev)
    b = dt.xyz_dev.view(-1)
    med, best = time_fn(lambda: a.copy_(b), reps)
    res["copy_GBs_rw"] = 2 * res["bytes"] / med / 1e6
    print(json.dumps(res))


if __name__ == "__main__":
    main()

"""rmsd with atom_indices: throughput of the gather path (development aid)."""
import json, os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import mdtraj_b200 as mdb
from mdtraj_b200 import _capi
from mdtraj_b200.device import _Scratch, _stream_ptr, prepare_reference
dev = torch.device("cuda", 0)
L = _capi.lib()
cases = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]] or [(5000, 5), (5000, 10), (5000, 2), (1000, 4), (25000, 8)]
for N, stride in cases:
    F = int(2.4e9 // (N * 12))
    dt = mdb.DeviceTrajectory.synthetic_iid(F, N, 1, dev)
    idx = torch.arange(0, N, stride, dtype=torch.int32, device=dev)
    S = int(idx.numel())
    prep = prepare_reference(dt.xyz_dev[0].clone(), idx, S, True)
    out = torch.empty(F, dtype=torch.float32, device=dev)
    scratch = _Scratch.get(torch, dev, L.b200rmsd_scratch_bytes(F, N))
    stream = _stream_ptr(torch, dev)
    def run():
        _capi.check(L.b200rmsd_rmsd_dev(dt.xyz_dev.data_ptr(), F, N, dt.frame_stride, idx.data_ptr(), S, prep.ref.data_ptr(),
                                        prep.stats.data_ptr(), None, 0, out.data_ptr(), None, None, None,
                                        scratch.data_ptr(), scratch.numel(), stream), "rmsd_dev idx")
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(json.dumps({"N": N, "stride": stride, "S": S, "F": F, "ms": round(ms, 3), "frames_per_s": F / ms * 1e3,
                      "GBs_selected": F * S * 12 / ms / 1e6, "GBs_full_frame": F * N * 12 / ms / 1e6}))
    del dt

"""Stress individual kernels with many back-to-back launches (development aid)."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import mdtraj_b200 as mdb
from mdtraj_b200 import _capi
from mdtraj_b200.device import _Scratch, _stream_ptr, prepare_reference
F, N = int(sys.argv[1]), int(sys.argv[2])
which = sys.argv[3]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 100
dev = torch.device("cuda", 0)
dt = mdb.DeviceTrajectory.synthetic_iid(F, N, 1, dev)
L = _capi.lib()
prep = prepare_reference(dt.xyz_dev[0].clone(), None, N, True)
out = torch.empty(F, dtype=torch.float32, device=dev)
traces = torch.empty(F, dtype=torch.float32, device=dev)
rot = torch.empty((F, 9), dtype=torch.float32, device=dev)
scratch = _Scratch.get(torch, dev, L.b200rmsd_scratch_bytes(F, N))
stream = _stream_ptr(torch, dev)
idx = torch.arange(0, N, 5, dtype=torch.int32, device=dev)
prep5 = prepare_reference(dt.xyz_dev[0].clone(), idx, int(idx.numel()), True)
def center():
    _capi.check(L.b200rmsd_center_trace_dev(dt.xyz_dev.data_ptr(), F, N, dt.frame_stride, traces.data_ptr(), stream), "center")
def pre():
    _capi.check(L.b200rmsd_rmsd_dev(dt.xyz_dev.data_ptr(), F, N, dt.frame_stride, None, N, prep.ref.data_ptr(), prep.stats.data_ptr(),
                                    traces.data_ptr(), 1, out.data_ptr(), None, None, None, scratch.data_ptr(), scratch.numel(), stream), "pre")
def sup5():
    _capi.check(L.b200rmsd_superpose_dev(dt.xyz_dev.data_ptr(), F, N, dt.frame_stride, idx.data_ptr(), int(idx.numel()), prep5.ref.data_ptr(),
                                         prep5.stats.data_ptr(), out.data_ptr(), rot.data_ptr(), None, scratch.data_ptr(), scratch.numel(), stream), "sup5")
fn = {"center": center, "pre": pre, "sup5": sup5}[which]
center(); torch.cuda.synchronize()
for r in range(reps):
    fn()
    if r % 10 == 9:
        torch.cuda.synchronize()
torch.cuda.synchronize()
print(which, "ok", reps, "launches")

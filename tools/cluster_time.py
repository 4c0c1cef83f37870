"""Throughput of the matrix reductions of csrc/cluster_ops.cu (development aid): python tools/cluster_time.py [F] [N]"""
import json, os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import mdtraj_b200 as mdb
from mdtraj_b200 import _capi
from mdtraj_b200.device import _stream_ptr

F = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 300
dev = torch.device("cuda", 0)
dt = mdb.DeviceTrajectory.synthetic_iid(F, N, seed=2, device=dev)
D = mdb.rmsd_matrix_device(dt)
L = _capi.lib()
st = _stream_ptr(torch, dev)
mom = torch.zeros(2, dtype=torch.float64, device=dev)
rs = torch.empty(F, dtype=torch.float64, device=dev)
arg = torch.empty(F, dtype=torch.int32, device=dev); val = torch.empty(F, dtype=torch.float32, device=dev)
cond = torch.empty(F * (F - 1) // 2, dtype=torch.float32, device=dev)
ops = {
    "matrix_moments": (lambda: L.b200rmsd_matrix_moments_dev(D.data_ptr(), F, F, F, mom.data_ptr(), st), 4 * F * F),
    "exp_rowsum": (lambda: L.b200rmsd_exp_rowsum_dev(D.data_ptr(), F, F, F, -0.5, 0, rs.data_ptr(), st), 4 * F * F),
    "row_argmin": (lambda: L.b200rmsd_row_argmin_dev(D.data_ptr(), F, F, F, arg.data_ptr(), val.data_ptr(), st), 4 * F * F),
    "condense": (lambda: L.b200rmsd_condense_dev(D.data_ptr(), F, F, cond.data_ptr(), st), 4 * F * (F - 1)),  # read half + write half
}
res = {"F": F, "matrix_GB": 4 * F * F / 1e9}
for name, (fn, nbytes) in ops.items():
    for _ in range(3):
        assert fn() == 0
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)[5]
    res[name] = {"ms": round(ms, 4), "GBs": round(nbytes / ms / 1e6), "frac_hbm_peak": round(nbytes / ms / 1e6 / 6540.8, 3)}
import time
t0 = time.perf_counter(); idx = mdb.centroid_index(dt); torch.cuda.synchronize(); res["centroid_index_s"] = round(time.perf_counter() - t0, 4)
print(json.dumps(res))

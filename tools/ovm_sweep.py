"""One-vs-many throughput vs atom count (development aid): python tools/ovm_sweep.py [N ...]"""
import json, os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import mdtraj_b200 as mdb
from mdtraj_b200 import _capi
from mdtraj_b200.device import _Scratch, _stream_ptr, prepare_reference

def main():
    Ns = [int(a) for a in sys.argv[1:]] or [22, 50, 100, 200, 300, 500, 1000]
    dev = torch.device("cuda", 0)
    L = _capi.lib()
    for N in Ns:
        n_pad = (N + 3) // 4 * 4
        F = int(2.4e9 // (n_pad * 12))
        dt = mdb.DeviceTrajectory.synthetic_iid(F, N, seed=1, device=dev)
        prep = prepare_reference(dt.xyz_dev[0].clone(), None, N, True)
        out = torch.empty(F, dtype=torch.float32, device=dev)
        scratch = _Scratch.get(torch, dev, L.b200rmsd_scratch_bytes(F, N))
        stream = _stream_ptr(torch, dev)
        def run():
            _capi.check(L.b200rmsd_rmsd_dev(dt.xyz_dev.data_ptr(), F, N, dt.frame_stride, None, N, prep.ref.data_ptr(),
                                            prep.stats.data_ptr(), None, 0, out.data_ptr(), None, None, None,
                                            scratch.data_ptr(), scratch.numel(), stream), "rmsd_dev")
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(json.dumps({"N": N, "F": F, "ms": round(ms, 4), "frames_per_s": F / ms * 1e3,
                          "GBs_algorithmic": F * N * 12 / ms / 1e6, "GBs_padded": F * n_pad * 12 / ms / 1e6}))
        del dt, out

if __name__ == "__main__":
    main()

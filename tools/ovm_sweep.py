"""One-vs-many throughput vs atom count (development aid): python tools/ovm_sweep.py [--geom] [N ...]
   --geom also tries the development overrides of the short/mid-frame kernel (lanes per frame, warps per CTA)."""
import json, os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import mdtraj_b200 as mdb
from mdtraj_b200 import _capi
from mdtraj_b200.device import _Scratch, _stream_ptr, prepare_reference

def main():
    geom = "--geom" in sys.argv
    tma = "--tma" in sys.argv   # ring geometry of the chunked kernel (dev build): chunk units x stages
    libs = [a[6:] for a in sys.argv[1:] if a.startswith("--lib=")]   # variants/*.so built by tools/build_variants.sh
    if libs:
        _capi.LIB_PATH = os.path.abspath(libs[0])
    extra_env = dict(a[6:].split("=", 1) for a in sys.argv[1:] if a.startswith("--env="))   # dev-build switches
    os.environ.update(extra_env)
    Ns = [int(a) for a in sys.argv[1:] if not a.startswith("--")] or [22, 50, 100, 200, 300, 500, 1000]
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
        envs = [{}]
        if geom:
            envs.append({"B200RMSD_NO_GROUP": "1"})
            envs += [{"B200RMSD_GROUP_LANES": str(l)} for l in (2, 4, 8, 16)] if n_pad * 12 <= 3072 else \
                    [{"B200RMSD_GROUP_WARPS": str(w), "B200RMSD_GROUP_MAX_BYTES": "14000"} for w in (4, 5, 6, 8, 10, 12)]
        if tma:
            envs += [{"B200RMSD_NO_GROUP": "1", "B200RMSD_CHUNK_UNITS": str(c), "B200RMSD_STAGES": str(st)}
                     for c in (32, 48, 64, 96) for st in (2, 3, 4, 6, 8)]
        for env in envs:
            for k in ("B200RMSD_NO_GROUP", "B200RMSD_GROUP_LANES", "B200RMSD_GROUP_WARPS", "B200RMSD_GROUP_MAX_BYTES",
                      "B200RMSD_CHUNK_UNITS", "B200RMSD_STAGES"):
                os.environ.pop(k, None)
            os.environ.update(env)
            for _ in range(3): run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): run()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print(json.dumps({"N": N, "F": F, "env": {**extra_env, **env}, **({"lib": os.path.basename(libs[0])} if libs else {}), "ms": round(ms, 4), "frames_per_s": F / ms * 1e3,
                              "GBs_algorithmic": F * N * 12 / ms / 1e6, "GBs_padded": F * n_pad * 12 / ms / 1e6,
                              "frac": round(F * n_pad * 12 / ms / 1e6 / 6540.8, 3)}), flush=True)
        del dt, out

if __name__ == "__main__":
    main()

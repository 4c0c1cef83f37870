"""All-pairs tensor-core kernel: timing of the epilogue geometries and accuracy against float64 truth (development aid).

    python tools/ap_variants.py [F] [N]

For every B200RMSD_TC_EPILOGUE variant: the full symmetric matrix and an unsymmetric row block
on iid and MD-like frames, checked against the exact-fp32 SIMT kernel on a corner of the matrix.  One JSON line each.
"""
import json
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch  # noqa: E402

import mdtraj_b200 as mdb  # noqa: E402
from mdtraj_b200 import allpairs as AP  # noqa: E402
from ap_time import md_like  # noqa: E402


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    dev = torch.device("cuda", 0)
    data = {"iid": mdb.DeviceTrajectory.synthetic_iid(F, N, 1, dev), "md": md_like(F, N, dev)}
    out = torch.empty((F, F), dtype=torch.float32, device=dev)
    corner = {}
    os.environ["B200RMSD_ALLPAIRS"] = "simt"
    for name, dt in data.items():
        prep = AP.prepare(dt)
        corner[name] = AP.rows(prep, 0, 256).clone()
    # float64 truth for a few rows (host numpy), to tell which of the two kernels a difference belongs to
    import _truth as O
    truth = {}
    for name, dt in data.items():
        X = dt.xyz_dev[:, :N].cpu().numpy()
        truth[name] = {i: O.truth_rmsd_batch(X, X[i]) for i in (0, 100, 255)}
        for i, tr in truth[name].items():
            e = (corner[name][i].cpu().numpy() - tr); e[i] = 0
            print(json.dumps({"kernel": "simt", "data": name, "row": i, "max_err_vs_truth": float(abs(e).max())}), flush=True)
    os.environ["B200RMSD_ALLPAIRS"] = "tc"
    for epi in ("16x2", "16x1"):
        os.environ["B200RMSD_TC_EPILOGUE"] = epi
        res = {"epilogue": epi, "F": F, "N": N}
        try:
            for name, dt in data.items():
                res[name + "_prepare_ms"] = round(timed(lambda: AP.prepare(dt), 3), 3)
                prep = AP.prepare(dt)
                ms = timed(lambda: AP.rows(prep, 0, F, out=out))
                res[name + "_sym_ms"] = round(ms, 3)
                res[name + "_sym_pairs_per_s"] = F * F / ms * 1e3
                res[name + "_max_err_vs_simt"] = (out[:256] - corner[name]).abs().max().item()
                for i, tr in truth[name].items():
                    e = (out[i].cpu().numpy() - tr); e[i] = 0
                    res[f"{name}_row{i}_max_err_vs_truth"] = float(abs(e).max())
                res[name + "_asym"] = (out[:2048, :2048] - out[:2048, :2048].t()).abs().max().item()
                rb = F // 8
                ms = timed(lambda: AP.rows(prep, rb, 2 * rb, out=out[:rb]))
                res[name + "_rowblock_pairs_per_s"] = rb * F / ms * 1e3
                for flag, key in (("0x100", "nosolve"), ("0x300", "delivery_only")):
                    os.environ["B200RMSD_TC_DEBUG"] = flag
                    res[f"{name}_{key}_ms"] = round(timed(lambda: AP.rows(prep, 0, F, out=out), 3), 3)
                    del os.environ["B200RMSD_TC_DEBUG"]
                ms = timed(lambda: AP.rows(prep, 0, F, out=out, precise=False), 3)
                res[name + "_f32solve_pairs_per_s"] = F * F / ms * 1e3
        except Exception as e:  # noqa: BLE001
            res["error"] = repr(e)
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()

"""CTA-pair (cta_group::2) vs single-CTA tensor-core all-pairs kernel: bit-identical matrices, and their timings
(development aid).   python tools/ap_pair_check.py [F] [N]"""
import json, os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import mdtraj_b200 as mdb
from mdtraj_b200 import allpairs as AP
from ap_time import md_like


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 2100
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    dev = torch.device("cuda", 0)
    res = {"F": F, "N": N}
    for name, dt in (("iid", mdb.DeviceTrajectory.synthetic_iid(F, N, 1, dev)), ("md", md_like(F, N, dev))):
        prep = AP.prepare(dt)
        out = {}
        for mode in (False, True):
            AP.configure(cta_pair=mode)
            D = torch.empty((F, F), dtype=torch.float32, device=dev)
            AP.rows(prep, 0, F, out=D)
            torch.cuda.synchronize()
            print(f"  [{name} pair={mode}] full matrix done", file=sys.stderr, flush=True)
            r0, r1 = F // 3, F // 3 + min(F // 4, 1000)
            blk = AP.rows(prep, r0, r1).clone()     # unsymmetric row block (odd tile offsets)
            torch.cuda.synchronize()
            out[mode] = (D, blk)
            res[f"{name}_ms_pair{int(mode)}"] = round(timed(lambda: AP.rows(prep, 0, F, out=D)), 4)
        res[name + "_full_equal"] = bool(torch.equal(out[False][0], out[True][0]))
        res[name + "_block_equal"] = bool(torch.equal(out[False][1], out[True][1]))
        res[name + "_max_diff"] = float((out[False][0] - out[True][0]).abs().max())
        res[name + "_finite"] = bool(torch.isfinite(out[True][0]).all())
    AP.configure(cta_pair=False)
    print(json.dumps(res))


if __name__ == "__main__":
    main()

"""All-pairs timing on iid vs MD-like synthetic data (development aid)."""
import os, sys, json
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import mdtraj_b200 as mdb
from mdtraj_b200 import allpairs as AP

def md_like(F, N, dev, seed=0, rg=1.0, sigma=0.1):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    base = torch.randn((N, 3), generator=g, device=dev) * rg
    X = base[None] + sigma * torch.randn((F, N, 3), generator=g, device=dev)
    q = torch.randn((F, 4), generator=g, device=dev); q = q / q.norm(dim=1, keepdim=True)
    a, b, c, d = q.unbind(1)
    R = torch.stack([a*a+b*b-c*c-d*d, 2*(b*c-a*d), 2*(b*d+a*c), 2*(b*c+a*d), a*a-b*b+c*c-d*d, 2*(c*d-a*b),
                     2*(b*d-a*c), 2*(c*d+a*b), a*a-b*b-c*c+d*d], dim=1).view(F, 3, 3)
    X = torch.bmm(X, R) + (torch.rand((F, 1, 3), generator=g, device=dev) * 10 - 5)
    n_pad = (N + 3) // 4 * 4
    out = torch.zeros((F, n_pad, 3), device=dev); out[:, :N] = X
    return mdb.DeviceTrajectory(out.contiguous(), N)

def main():
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    dev = torch.device("cuda", 0)
    res = {"F": F, "N": N}
    for name, dt in (("iid", mdb.DeviceTrajectory.synthetic_iid(F, N, 1, dev)), ("md", md_like(F, N, dev))):
        prep = AP.prepare(dt)
        torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(); prep = AP.prepare(dt); p1.record(); torch.cuda.synchronize()
        res[name + "_prepare_ms"] = p0.elapsed_time(p1); res[name + "_info"] = prep.info()
        out = torch.empty((F, F), dtype=torch.float32, device=dev)
        for _ in range(2):
            AP.rows(prep, 0, F, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            AP.rows(prep, 0, F, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        res[name + "_ms"] = ms; res[name + "_pairs_per_s"] = F * F / ms * 1e3
        res[name + "_mean_rmsd"] = out[0, 1:].mean().item()
    print(json.dumps(res))

if __name__ == "__main__":
    main()

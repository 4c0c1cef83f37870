"""All-pairs tensor-core operands, checked on the host (development aid).

Reads the prepared workspace back (layout of csrc/allpairs_layout.cuh: ap_geometry), forms M_ij = A_i . B_j^T in float64
from the tf32 hi+lo operands -- augmentation columns included -- and solves for the RMSD with numpy.  Separates
  operand construction (alignment, differences, G_i pieces)   : r_operands vs float64 truth of the original frames
  tensor-core accumulation + epilogue solve                   : r_gpu      vs r_operands
"""
import json
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import mdtraj_b200 as mdb  # noqa: E402
from mdtraj_b200 import _capi, allpairs as AP  # noqa: E402
import _truth as O  # noqa: E402
from ap_time import md_like  # noqa: E402


def a256(x):
    return (x + 255) // 256 * 256


def geometry(F, n_sel):
    k0 = (n_sel + 7) // 8 * 8
    k_pad = (k0 + 6 + 31) // 32 * 32
    rows = (3 * F + 160 + 7) // 8 * 8
    off = 256 + a256(F * 4)
    op = a256(rows * k_pad * 4)
    g = {"k0": k0, "k_pad": k_pad, "rows": rows, "traces": 256}
    for name in ("a_hi", "a_lo", "b_hi", "b_lo"):
        g[name] = off
        off += op
    return g


def lam_max(M):
    Sxx, Sxy, Sxz, Syx, Syy, Syz, Szx, Szy, Szz = [M[..., i, j] for i in range(3) for j in range(3)]
    K = np.zeros(M.shape[:-2] + (4, 4))
    K[..., 0, 0] = Sxx + Syy + Szz; K[..., 0, 1] = Szy - Syz; K[..., 0, 2] = Sxz - Szx; K[..., 0, 3] = Syx - Sxy
    K[..., 1, 1] = Sxx - Syy - Szz; K[..., 1, 2] = Syx + Sxy; K[..., 1, 3] = Sxz + Szx
    K[..., 2, 2] = -Sxx + Syy - Szz; K[..., 2, 3] = Szy + Syz; K[..., 3, 3] = -Sxx - Syy + Szz
    K = K + np.swapaxes(np.triu(K, 1), -1, -2)
    return np.linalg.eigvalsh(K)[..., -1]


def main():
    F, N = 2000, int(sys.argv[1]) if len(sys.argv) > 1 else 300
    dev = torch.device("cuda", 0)
    os.environ["B200RMSD_ALLPAIRS"] = "tc"
    for name, dt in (("iid", mdb.DeviceTrajectory.synthetic_iid(F, N, 1, dev)), ("md", md_like(F, N, dev))):
        X = dt.xyz_dev[:, :N].cpu().numpy()
        prep = AP.prepare(dt)
        D = AP.rows(prep, 0, F).cpu().numpy().astype(np.float64)
        g = geometry(F, N)
        ws = prep.workspace
        def arr(key):
            n = g["rows"] * g["k_pad"]
            return ws[g[key]: g[key] + 4 * n].view(torch.float32).view(g["rows"], g["k_pad"])[:3 * F].cpu().numpy().astype(np.float64)
        A = (arr("a_hi") + arr("a_lo")).reshape(F, 3, g["k_pad"])
        B = (arr("b_hi") + arr("b_lo")).reshape(F, 3, g["k_pad"])
        tr = ws[256: 256 + 4 * F].view(torch.float32).cpu().numpy().astype(np.float64)
        res = {"data": name, "N": N, "k0": g["k0"], "k_pad": g["k_pad"],
               "aug_A_abs_max": float(np.abs(A[:, :, g["k0"]:g["k0"] + 6]).max()),
               "aug_B_sum": float(B[:, :, g["k0"]:].sum() / F), "pad_nonzero": float(np.abs(A[:, :, g["k0"] + 6:]).max())}
        for i in (0, 100, 1999):
            M = np.einsum("ck,jqk->jcq", A[i], B)
            r_op = np.sqrt(np.maximum(tr[i] + tr - 2 * lam_max(M), 0) / N)
            truth = O.truth_rmsd_batch(X, X[i])
            m = np.arange(F) != i
            res[f"row{i}"] = {"operands_vs_truth": float(np.abs(r_op - truth)[m].max()),
                              "gpu_vs_operands": float(np.abs(D[i] - r_op)[m].max()),
                              "gpu_vs_truth": float(np.abs(D[i] - truth)[m].max()),
                              "gpu_minus_operands_mean": float((D[i] - r_op)[m].mean())}
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()

"""float64 Kabsch RMSD (numpy SVD) for the development tools.  The tools stay clear of the oracle tree, which is reserved for
tests/, __graft_entry__.smoke() and bench.py's CPU legs."""
import numpy as np


def truth_rmsd_batch(target_xyz, ref_frame):
    """RMSD after optimal superposition of every frame of target_xyz (F,N,3) onto ref_frame (N,3), float64."""
    X = np.asarray(target_xyz, dtype=np.float64)
    Q = np.asarray(ref_frame, dtype=np.float64)
    Xc = X - X.mean(1, keepdims=True)
    Qc = Q - Q.mean(0)
    H = np.einsum("fki,kj->fij", Xc, Qc)
    U, S, Vt = np.linalg.svd(H)
    d = np.sign(np.linalg.det(U @ Vt))
    e0 = np.einsum("fki,fki->f", Xc, Xc) + (Qc * Qc).sum()
    return np.sqrt(np.maximum(0.0, (e0 - 2.0 * (S[:, 0] + S[:, 1] + d * S[:, 2])) / X.shape[1]))

"""A numpy model of the all-pairs tensor-core arithmetic (test infrastructure, CPU only).

What it models, and what it was calibrated against:
  * 3xTF32 operands: hi = rna_tf32(x), lo = rna_tf32(x - hi)  (csrc/tc_ptx.cuh: rna_tf32, cvt.rna.tf32.f32);
  * tcgen05.mma.kind::tf32, K = 8 per instruction, three per K-step (lo.hi, hi.lo, hi.hi), fp32 accumulator in TMEM with
    ONE truncation (round toward zero) per instruction.  That single assumption reproduces the bias measured on a B200:
    300-term inner products of magnitude ~324 came out 5.3e-4 short (profiles/r01_tc_accumulation_probe.json); the model
    gives 5.8e-4 on the same kind of data;
  * the operand construction of allpairs_tc144_prepare_kernel and the reference traversal of allpairs_refs.cu: frames
    aligned onto their nearest reference c_r, A = x', B = x' - c_r for frames near it, the 3x3 matrix X'c_r^T added
    through eight augmentation columns per reference in K-steps of their own after the atoms.
It is a design tool: it says what an operand layout does to the RMSD before a GPU is involved.  It is NOT the oracle and
nothing in the product depends on it."""
import numpy as np

f32 = np.float32


def rz32(x):
    """float64 -> float32 rounding toward zero."""
    x = np.asarray(x, np.float64)
    y = x.astype(f32)
    over = np.abs(y.astype(np.float64)) > np.abs(x)
    y[over] = np.nextafter(y[over], f32(0))
    return y


def rna_tf32(x):
    """cvt.rna.tf32.f32: nearest, ties away from zero, 10 explicit mantissa bits."""
    u = np.asarray(x, f32).view(np.uint32).astype(np.uint64)
    return ((u + 0x1000) & 0xFFFFE000).astype(np.uint32).view(f32)


def split(x):
    hi = rna_tf32(x)
    lo = rna_tf32((np.asarray(x, f32) - hi).astype(f32))
    return hi, lo


def tc_gemm(a_hi, a_lo, b_hi, b_lo):
    """(m,K) x (n,K) -> (m,n) float64 view of the fp32 accumulator after the K loop."""
    Ah, Al, Bh, Bl = (np.asarray(v, np.float64) for v in (a_hi, a_lo, b_hi, b_lo))
    acc = np.zeros((Ah.shape[0], Bh.shape[0]), f32)
    for k in range(0, Ah.shape[1], 8):
        s = slice(k, k + 8)
        for X, Y in ((Al, Bh), (Ah, Bl), (Ah, Bh)):
            acc = rz32(acc.astype(np.float64) + X[:, s] @ Y[:, s].T)
    return acc.astype(np.float64)


def kabsch_rotation(mobile, target):
    """R (float64) with mobile @ R ~ target, both centred."""
    U, _, Vt = np.linalg.svd(mobile.T @ target)
    d = np.sign(np.linalg.det(U @ Vt))
    return U @ np.diag([1.0, 1.0, d]) @ Vt


def _align_all(T, c):
    """Every centred frame of T (F,N,3) float32 rotated onto c: (aligned frames float32, msd to c)."""
    out = np.empty_like(T)
    msd = np.empty(len(T))
    for f in range(len(T)):
        R = kabsch_rotation(T[f].astype(np.float64), c.astype(np.float64)).astype(f32)
        out[f] = (T[f] @ R).astype(f32)
        msd[f] = float(((out[f].astype(np.float64) - c) ** 2).sum()) / T.shape[1]
    return out, msd


ROUND_BIAS = True
COVER_STOP = 0.25   # nm, kCoverStop of csrc/allpairs_refs.cu


def choose_references(T, max_refs):
    """The traversal of csrc/allpairs_refs.cu (ap_select_references): greedy farthest-point references, every frame owned
    by its nearest one; stops at max_refs, or when far frames remain but three references in a row captured nothing
    (iid-like data), or when everything is near and the covering radius is below COVER_STOP / has stopped shrinking.
    Returns (reference frame indices, per-frame owner, near mask, aligned frames, msd to the owner)."""
    F, N = len(T), T.shape[1]
    refs, owner = [], np.zeros(F, int)
    aligned, best = None, None
    g = []
    cand, unproductive, hist = 0, 0, []
    productive_min = max(2, F // 1000)
    while True:
        r = len(refs)
        refs.append(cand)
        g.append(float((T[cand].astype(np.float64) ** 2).sum()) / N)
        al, msd = _align_all(T, T[cand])
        moved = np.ones(F, bool) if r == 0 else msd < best
        if r == 0:
            aligned, best = al, msd
        else:
            aligned[moved] = al[moved]
            best[moved] = msd[moved]
            owner[moved] = r
        near = best < 0.25 * np.asarray(g)[owner]
        radius = float(np.sqrt(best.max()))
        R = r + 1
        if R >= max_refs or radius <= 0:
            break
        if (~near).any():
            if R >= 2:
                unproductive = 0 if int((moved & near).sum()) >= productive_min else unproductive + 1
            if unproductive >= 3:
                break
        else:   # everything is near a reference: refine until the covering radius is small or has stopped shrinking
            hist.append(radius)
            if radius <= COVER_STOP or (len(hist) >= 3 and radius > 0.9 * hist[-3]):
                break
        cand = int(np.argmax(best))
    return refs, owner, near, aligned, best


def prepare_operands(X, aligned=True, max_refs=32):
    """The prepare step on frames X (F,N,3) float32 -> dict of (3F, K) operand matrices and traces.
    aligned=False: the plain layout (A = B = centred frames, no augmentation).
    aligned=True: allpairs_tc144_prepare_kernel -- A = frames aligned onto their nearest reference, B = their differences
    from it (near frames), eight augmentation columns per reference after the atoms."""
    X = np.asarray(X, f32)
    F, N, _ = X.shape
    mu = X.astype(np.float64).mean(1, keepdims=True).astype(f32)
    T = (X - mu).astype(f32)                                   # center_generic.h: float64 mean, float32 subtraction
    if aligned:
        refs, owner, near, V, msd = choose_references(T, max_refs)
    else:
        refs, owner, near, V = [], np.zeros(F, int), np.zeros(F, bool), T
    R = max(len(refs), 1)
    k0 = (N + 31) // 32 * 32
    K = k0 + (8 * R + 31) // 32 * 32
    A = np.zeros((F, 3, K), f32)
    B = np.zeros((F, 3, K), f32)
    tr = np.empty(F)
    for f in range(F):
        v = V[f]
        tr[f] = float(f32((v.astype(np.float64) ** 2).sum()))
        A[f, :, :N] = v.T
        B[f, :, :N] = ((v - T[refs[owner[f]]]) if near[f] else v).T
    a_hi, a_lo = split(A.reshape(3 * F, K))
    b_hi, b_lo = split(B.reshape(3 * F, K))
    a_hi = a_hi.reshape(F, 3, K); a_lo = a_lo.reshape(F, 3, K)
    b_hi = b_hi.reshape(F, 3, K)
    for f in range(F):
        va = a_hi[f, :, :N].astype(np.float64) + a_lo[f, :, :N]       # the operand the tensor core multiplies
        for r, rf in enumerate(refs):
            G = va @ T[rf].astype(np.float64)                         # G[c][m] = sum_k x'_k[c] c_k[m]
            if ROUND_BIAS:   # half an fp32 ulp of G towards its sign: the accumulator's truncation then rounds to nearest
                G = G + np.sign(G) * np.exp2(np.floor(np.log2(np.maximum(np.abs(G), 1e-300))) - 24)
            g1 = rna_tf32(G.astype(f32))
            g2 = rna_tf32((G - g1).astype(f32))
            g3 = rna_tf32((G - g1 - g2.astype(np.float64)).astype(f32))
            o = k0 + 8 * r
            a_hi[f, :, o:o + 3] = g1
            a_hi[f, :, o + 3:o + 6] = g2
            a_lo[f, :, o:o + 3] = g3
        if near[f]:
            o = k0 + 8 * owner[f]
            b_hi[f, :, o:o + 3] = np.eye(3, dtype=f32)
            b_hi[f, :, o + 3:o + 6] = np.eye(3, dtype=f32)
    return {"a_hi": a_hi.reshape(3 * F, K), "a_lo": a_lo.reshape(3 * F, K), "b_hi": b_hi.reshape(3 * F, K), "b_lo": b_lo,
            "traces": tr, "n_atoms": N, "references": refs, "owner": owner, "near": near}


def lambda_max(M):
    Sxx, Sxy, Sxz, Syx, Syy, Syz, Szx, Szy, Szz = [M[..., i, j] for i in range(3) for j in range(3)]
    K = np.zeros(M.shape[:-2] + (4, 4))
    K[..., 0, 0] = Sxx + Syy + Szz; K[..., 0, 1] = Szy - Syz; K[..., 0, 2] = Sxz - Szx; K[..., 0, 3] = Syx - Sxy
    K[..., 1, 1] = Sxx - Syy - Szz; K[..., 1, 2] = Syx + Sxy; K[..., 1, 3] = Sxz + Szx
    K[..., 2, 2] = -Sxx + Syy - Szz; K[..., 2, 3] = Szy + Syz; K[..., 3, 3] = -Sxx - Syy + Szz
    K = K + np.swapaxes(np.triu(K, 1), -1, -2)
    return np.linalg.eigvalsh(K)[..., -1]


def rmsd_rows(ops, rows, exact=False):
    """RMSD of frames `rows` against all frames from the modelled accumulator (exact=True: float64 product of the same
    operands, i.e. what an ideal accumulator would give)."""
    F = ops["b_hi"].shape[0] // 3
    sel = np.concatenate([np.arange(3 * i, 3 * i + 3) for i in rows])
    if exact:
        a = ops["a_hi"][sel].astype(np.float64) + ops["a_lo"][sel]
        b = ops["b_hi"].astype(np.float64) + ops["b_lo"]
        acc = a @ b.T
    else:
        acc = tc_gemm(ops["a_hi"][sel], ops["a_lo"][sel], ops["b_hi"], ops["b_lo"])
    M = acc.reshape(len(rows), 3, F, 3).transpose(0, 2, 1, 3)
    tr = ops["traces"]
    return np.sqrt(np.maximum(tr[rows][:, None] + tr[None, :] - 2 * lambda_max(M), 0) / ops["n_atoms"])

"""CPU: host-side logic of the drop-in boundary -- validation, exceptions and messages identical to the
reference's (captured verbatim from the real mdtraj by tests/golden/make_golden.py), Trajectory container
semantics, frame sharding.  None of these reach the GPU."""
import os
import warnings

import numpy as np
import pytest

import mdtraj_b200 as mdb
from mdtraj_b200 import distributed as D


def _t(F=5, N=10, seed=1):
    rng = np.random.default_rng(seed)
    return mdb.Trajectory(rng.standard_normal((F, N, 3)).astype(np.float32))


def _msg(fn):
    try:
        fn()
    except Exception as e:  # noqa: BLE001
        return f"{type(e).__name__}: {e}"
    return "NO ERROR"


def test_error_messages_match_reference(golden):
    t = _t()
    assert _msg(lambda: mdb.rmsd(t, t, 0, atom_indices=[0, 10])) == str(golden["msg_bad_index"])
    assert _msg(lambda: mdb.rmsd(t, t, 0, atom_indices=[0, 1], ref_atom_indices=[0])) == str(golden["msg_len_mismatch"])
    assert _msg(lambda: mdb.rmsd(t, t, 5)) == str(golden["msg_bad_frame"])
    assert _msg(lambda: mdb.rmsd(t, t, 0, ref_atom_indices=[0, 1])) == str(golden["msg_ref_only_indices"])
    assert _msg(lambda: t.superpose(t, 0, atom_indices=[])) == str(golden["msg_empty_superpose"])
    with pytest.raises(ValueError, match="valid positive indices"):
        mdb.rmsd(t, t, 0, atom_indices=[-1, 2])
    with pytest.raises(ValueError, match="ref_atom_indices must be valid"):
        mdb.rmsd(t, t, 0, atom_indices=[1, 2], ref_atom_indices=[1, 99])


def test_readonly_buffer_raises_before_any_work():
    ro = np.zeros((3, 4, 3), np.float32); ro.setflags(write=False)

    class Duck:
        xyz = ro
        _rmsd_traces = None
    with pytest.raises(ValueError, match="buffer source array is read-only"):
        mdb.rmsd(Duck(), Duck(), 0)
    with pytest.raises(ValueError, match="read-only"):
        mdb._center_inplace_atom_major(ro)


def test_float_indices_warn_and_truncate():
    t = _t()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        with pytest.raises(ValueError):  # [0.5, 10.2] -> [0, 10]: 10 is out of range, after the cast warning
            mdb.rmsd(t, t, 0, atom_indices=[0.5, 10.2])
    assert any(issubclass(x.category, mdb.TypeCastPerformanceWarning) for x in w)


def test_trajectory_container_semantics():
    x = np.arange(2 * 3 * 3, dtype=np.float64).reshape(2, 3, 3)
    t = mdb.Trajectory(x)
    assert t.xyz.dtype == np.float32 and t.xyz.flags.c_contiguous and t.n_frames == 2 and t.n_atoms == 3
    t._rmsd_traces = np.array([1.0, 2.0], np.float32)
    t.xyz = x + 1
    assert t._rmsd_traces is None  # the setter resets the cache (trajectory.py:1029)
    one = mdb.Trajectory(np.zeros((4, 3)))
    assert one.xyz.shape == (1, 4, 3)  # add_newaxis_on_deficient_ndim
    with pytest.raises(ValueError):
        mdb.Trajectory(np.zeros((2, 3, 4)))
    t._rmsd_traces = np.array([5.0, 6.0], np.float32)
    s = t[1]
    assert s.n_frames == 1 and s._rmsd_traces.tolist() == [6.0]  # traces follow the slice (upstream bug fixed)
    s.xyz[0, 0, 0] = 99
    assert t.xyz[1, 0, 0] != 99  # slices copy


def test_legacy_entry_validation():
    a = np.zeros((2, 5, 3), np.float32); b = np.zeros((3, 6, 3), np.float32); g = np.zeros(3, np.float32)
    with pytest.raises(ValueError, match="same number of atoms"):
        mdb.getMultipleRMSDs_atom_major(a, b, g, g, 0)
    with pytest.raises(ValueError, match="Cannot calculate RMSD of frame 2"):
        mdb.getMultipleRMSDs_atom_major(a, a, g, g, 2)
    with pytest.raises(ValueError, match="same number of frames"):
        mdb.superpose_atom_major(a, a, g, g, np.zeros((3, 5, 3), np.float32), 0)
    with pytest.raises(ValueError, match="4\\*n"):
        mdb.getMultipleAlignDisplaceRMSDs_atom_major(a, a, g, g, a, a, 5, 5, 0)
    t = _t()
    with pytest.raises(ValueError, match="Mode must be one of"):
        mdb.rmsf(t, t, 0, mode="x")


def test_shard_bounds_cover_exactly_once():
    for n in (0, 1, 7, 100, 1_000_003):
        for w in (1, 2, 3, 8):
            b = D.all_shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(y - x for x, y in b) - min(y - x for x, y in b) <= 1


def test_h5min_reads_reference_fixture(ala2):
    path = "/root/reference/examples/ala2.h5"
    if not os.path.exists(path):
        pytest.skip("reference checkout absent on this box")
    from mdtraj_b200 import h5min
    xyz = h5min.load_coordinates(path)
    assert xyz.shape == (100, 22, 3) and np.array_equal(xyz, ala2)
    assert h5min.H5Min(path).keys() == ["coordinates", "time", "topology"]


def test_symmetric_plan_covers_every_entry_once():
    """Symmetric all-pairs sharding: each rank's row block is tiled exactly once by local blocks (including the mirrored
    lower triangle of its diagonal block) and received transposes; work is balanced within one block."""
    for F, W in ((10, 1), (37, 2), (100, 3), (101, 4), (64, 8), (5, 8)):
        plan = D.symmetric_plan(F, W)
        bounds = D.all_shard_bounds(F, W)
        work = []
        for r in range(W):
            r0, r1 = bounds[r]
            cover = np.zeros((r1 - r0, F), dtype=np.int32)
            area = 0
            for (a0, a1, c0, c1, dst) in plan[r]["compute"]:
                assert r0 <= a0 <= a1 <= r1 and 0 <= c0 <= c1 <= F
                cover[a0 - r0:a1 - r0, c0:c1] += 1
                area += (a1 - a0) * (c1 - c0) if dst is not None else (a1 - a0) * (a1 - a0 + 1) // 2
                if dst is not None:  # the receiver must expect exactly this block, transposed
                    assert (r, c0, c1, a0, a1) in plan[dst]["recv"]
            for (src, a0, a1, c0, c1) in plan[r]["recv"]:
                cover[a0 - r0:a1 - r0, c0:c1] += 1
                assert any(t[4] == r and (t[2], t[3], t[0], t[1]) == (a0, a1, c0, c1) for t in plan[src]["compute"])
            assert np.all(cover == 1), (F, W, r)
            work.append(area)
        if W > 1 and F >= 8 * W:
            assert max(work) <= 1.35 * (sum(work) / W), (F, W, work)


def test_frame_resident_geometry_invariants():
    """The slot/ring geometry of the single-pass superpose / centring kernel (csrc/frame_resident.cu: fused_config),
    through the host-only debug hook: a buffer must belong to one group (nbuf % G == 0, the mbarrier parity rule),
    and fit shared memory."""
    import ctypes

    from mdtraj_b200 import _capi
    L = ctypes.CDLL(_capi.LIB_PATH)
    geo = L.b200rmsd_debug_fused_geometry
    geo.argtypes = [ctypes.c_int] * 5 + [ctypes.POINTER(ctypes.c_int)]
    out = (ctypes.c_int * 8)()
    seen_multi = seen_pipe = False
    for op in (0, 1):
        for n in list(range(1, 70)) + [99, 100, 127, 128, 200, 255, 256, 300, 500, 999, 1000, 1500, 2000, 2500, 3000, 4000,
                                        5000, 6000, 9000, 20000]:
            for has_idx in ((0, 1) if op == 0 else (0,)):
                for contiguous in (1, 0):
                    n_sel = max(1, n // 5) if has_idx else n
                    kind = geo(op, n, n_sel, has_idx, contiguous, out)
                    if not kind:
                        assert n * 12 * 3 > 150000, f"single-pass kernel refused a small frame: op={op} n={n}"
                        continue
                    if kind == 2:   # stage-pipelined superpose (superpose_pipe_kernel)
                        S, nbuf, fpb, W, lanes, smem, depth = list(out)[:7]
                        T = 16 - S
                        assert op == 0 and 1 <= S <= 8 and S <= nbuf and 2 <= nbuf <= 12 and smem <= 232448
                        assert 1 <= fpb <= 32 and (contiguous or fpb == 1)    # one solver lane per frame of a slot
                        assert 1 <= depth <= nbuf - 1                          # the transform of slot i - depth frees a buffer in time
                        if W == 1:
                            assert lanes in (2, 4, 8, 16, 32) and T * (32 // lanes) >= fpb   # one pass of the streaming warps
                        else:
                            assert lanes == 32 and W * fpb <= T
                        seen_multi |= fpb > 1
                        seen_pipe = True
                        continue
                    G, nbuf, fpb, tw, lanes, smem = list(out)[:6]
                    assert 1 <= G <= 16 and nbuf % G == 0 and nbuf >= 2 and nbuf >= G
                    assert smem <= 232448
                    assert fpb >= 1 and (contiguous or fpb == 1)
                    wpf = 16 // G
                    assert tw in (1, wpf) and lanes in (2, 4, 8, 16, 32)
                    assert fpb <= min(64, wpf * 32)  # one solver lane per frame of a slot
                    assert (tw == 1) == (fpb >= wpf) and G * wpf <= 16
                    if tw != 1:
                        assert lanes == 32
                    seen_multi |= fpb > 1
    assert seen_multi and seen_pipe


def test_allpairs_tile_walk_covers_every_pair_once():
    """The tile walk of the tensor-core all-pairs kernel (csrc/allpairs_tc144.cu: tc144_tiling + tile_of_slot, through
    the host-only debug hook): 40 x 48-frame tiles, super-block order, tiles under the diagonal skipped in symmetric
    mode.  Every pair of the block (every pair with j >= i in symmetric mode) must lie in exactly one visited tile."""
    import ctypes

    import numpy as np

    from mdtraj_b200 import _capi
    L = ctypes.CDLL(_capi.LIB_PATH)
    fn = L.b200rmsd_debug_allpairs_tiles
    fn.restype = ctypes.c_longlong
    fn.argtypes = [ctypes.c_longlong] * 4 + [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_longlong,
                                             ctypes.POINTER(ctypes.c_int)]
    cap = 200000
    buf = (ctypes.c_int * (2 * cap))()
    sym = ctypes.c_int()
    cases = [(0, 512, 0, 512, 0), (0, 2000, 0, 2000, 0), (0, 2000, 0, 2000, 1), (50, 1990, 50, 1990, 0),
             (47, 1001, 47, 1001, 0), (960, 2880, 960, 2880, 0), (0, 40, 0, 3000, 0), (123, 777, 0, 3000, 0),
             (0, 3000, 500, 548, 0), (1000, 1500, 1500, 2500, 1), (1500, 2500, 1000, 1500, 1), (0, 1, 0, 1, 0),
             (2999, 3000, 0, 3000, 0), (0, 5000, 0, 5000, 0)]
    for row0, row1, col0, col1, has_t in cases:
        n = fn(row0, row1, col0, col1, has_t, buf, cap, ctypes.byref(sym))
        assert 0 < n <= cap, (row0, row1, col0, col1)
        tiles = np.frombuffer(buf, dtype=np.int32, count=2 * n).reshape(n, 2).astype(np.int64)
        assert len({(a, b) for a, b in tiles.tolist()}) == n, "a tile is visited twice"
        symmetric = bool(sym.value)
        assert symmetric == (row0 == col0 and row1 == col1 and not has_t)
        # count, for every pair of the block, the visited tiles that contain it
        cover = np.zeros((row1 - row0, col1 - col0), np.int32)
        for ti, tj in tiles:
            i0, i1 = max(row0, ti * 40), min(row1, ti * 40 + 40)
            j0, j1 = max(col0, tj * 48), min(col1, tj * 48 + 48)
            assert i0 < i1 and j0 < j1, "a visited tile lies outside the block"
            cover[i0 - row0:i1 - row0, j0 - col0:j1 - col0] += 1
        if symmetric:
            need = np.triu(np.ones_like(cover))
            assert (cover[need == 1] == 1).all()
            # no tile that lies entirely under the diagonal is visited
            assert all(tj * 48 + 47 >= ti * 40 for ti, tj in tiles)
        else:
            assert (cover == 1).all()
        # super-block locality: the ~148 tiles in flight touch few distinct operand tiles (they stay L2-resident)
        for k in range(0, max(1, n - 148), 97):
            win = tiles[k:k + 148]
            assert len(set(win[:, 0].tolist())) + len(set(win[:, 1].tolist())) <= 3 * (24 + 20)


def test_allpairs_workspace_geometry():
    """b200rmsd_allpairs_workspace_bytes (csrc/allpairs_layout.cuh: ap_geometry): the tensor-core path (F >= 512) holds four
    tf32 operand matrices of (3F + 160) rows x K floats, K = atoms rounded up to 32, three augmentation matrices of 256
    columns (8 per reference structure, at most 32 references), the references themselves and the traversal's per-frame
    buffers; the SIMT path one (F,3,K32) copy.  Sizes must grow monotonically and stay 256-byte granular (the prepare entry
    point rejects unaligned workspaces)."""
    from mdtraj_b200 import _capi
    L = _capi.lib()
    wb = L.b200rmsd_allpairs_workspace_bytes
    assert wb(0, 10) == 256 and wb(10, 0) == 256
    for n_sel in (1, 2, 22, 299, 300, 301, 314, 315, 1000, 4097):
        k_pad = (n_sel + 31) // 32 * 32
        prev = 0
        for F in (512, 513, 1000, 20000, 100000):
            n = wb(F, n_sel)
            rows = (3 * F + 160 + 7) // 8 * 8
            operands = 4 * rows * k_pad * 4 + 3 * rows * 256 * 4
            refs = 32 * ((n_sel + 3) // 4 * 4 * 12 + 256)
            assert n % 256 == 0 and n > prev
            assert operands + refs <= n <= operands + refs + F * 256 + (1 << 20) + 16 * F * ((n_sel + 4095) // 4096 + 1) * 4 * 2, \
                (F, n_sel, n)
            prev = n
        small = wb(511, n_sel)   # SIMT path below 512 frames
        assert small % 256 == 0 and small >= 511 * 3 * ((n_sel + 31) // 32 * 32) * 4
    # C4: 100k frames x 300 atoms -> four 300160 x 320 float matrices (1.54 GB) + three 300160 x 256 (0.92 GB) + ~20 MB
    assert 2.45e9 < wb(100000, 300) < 2.50e9
    # the kernel switch is a public setting, not an environment variable
    assert L.b200rmsd_allpairs_configure(1 << 30, 0) == 0
    assert 3.8e8 < wb(100000, 300) < 3.9e8             # SIMT layout: one (F,3,320) float copy + traces
    assert L.b200rmsd_allpairs_configure(512, 0) == 0
    assert 2.45e9 < wb(100000, 300) < 2.50e9


def test_lprmsd_validation_mirrors_the_reference():
    """md.lprmsd argument checks (_lprmsd.pyx:230-273) raise before any device work, with the reference's messages."""
    import mdtraj_b200 as mdb
    X = np.zeros((3, 10, 3), dtype=np.float32)
    T = mdb.Trajectory
    with pytest.raises(ValueError, match="Input trajectories must have same number of atoms. found 10 and 9."):
        mdb.lprmsd(T(X), T(X[:, :9].copy()))
    with pytest.raises(ValueError, match="Cannot calculate RMSD of frame 3: reference has only 3 frames."):
        mdb.lprmsd(T(X), T(X), 3)
    with pytest.raises(ValueError, match="atom_indices must be valid positive indices"):
        mdb.lprmsd(T(X), T(X), atom_indices=[0, 10])
    with pytest.raises(ValueError, match="must be a subset of atom_indices"):
        mdb.lprmsd(T(X), T(X), atom_indices=[0, 1, 2], permute_groups=[[2, 3]])
    with pytest.raises(ValueError, match="permute_groups must be mutually disjoint sets"):
        mdb.lprmsd(T(X), T(X), permute_groups=[[1, 2], [2, 3]])


def test_superpose_pipe_reference_placement():
    """Geometry of superpose_pipe_kernel for large selections (host-side hook): the reference stays in shared memory as long
    as three frame buffers fit next to it, and moves to global memory -- keeping three buffers -- when they do not (all
    atoms selected on ~5,000-atom frames); frames over 100 KB leave the single-pass kernels altogether."""
    import ctypes

    from mdtraj_b200 import _capi
    geo = ctypes.CDLL(_capi.LIB_PATH).b200rmsd_debug_fused_geometry
    geo.argtypes = [ctypes.c_int] * 5 + [ctypes.POINTER(ctypes.c_int)]
    out = (ctypes.c_int * 8)()

    def pipe(n, n_sel, has_idx):
        for i in range(8):
            out[i] = 0
        kind = geo(0, n, n_sel, has_idx, 1, out)
        return kind, list(out)
    for n in (3000, 4000):                       # all atoms: reference (36-48 KB) resident, >= 3 buffers
        kind, g = pipe(n, n, 0)
        assert kind == 2 and g[7] == 0 and g[1] >= 3, (n, g)
    for n in (4800, 5000, 5400, 6200):           # all atoms: reference in global memory, exactly the buffers that fit
        kind, g = pipe(n, n, 0)
        assert kind == 2 and g[7] == 1 and g[1] >= 3 and g[5] <= 232448, (n, g)
    kind, g = pipe(5000, 1000, 1)                # C3: every 5th atom, reference (12 KB) resident
    assert kind == 2 and g[7] == 0 and g[1] >= 3, g
    assert pipe(9000, 9000, 0)[0] == 0           # 108 KB frames: two-pass path

"""Per-frame timeline of CTA 0 of the frame-resident superpose kernel (development aid).
   B200RMSD_FUSED_TRACE=1 python tools/fused_trace.py [F] [N]"""
import ctypes, os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
os.environ["B200RMSD_FUSED_TRACE"] = "1"
import numpy as np, torch
import mdtraj_b200 as mdb
from mdtraj_b200 import _capi
libs = [a[6:] for a in sys.argv[1:] if a.startswith("--lib=")]   # a dev build: tools/build_variants.sh name:""
sys.argv = [a for a in sys.argv if not a.startswith("--lib=")]
if libs:
    _capi.LIB_PATH = os.path.abspath(libs[0])
from mdtraj_b200.device import _Scratch, _stream_ptr, prepare_reference
F = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
dev = torch.device("cuda", 0)
dt = mdb.DeviceTrajectory.synthetic_iid(F, N, 1, dev)
L = _capi.lib()
idx = torch.arange(0, N, 5, dtype=torch.int32, device=dev)
prep = prepare_reference(dt.xyz_dev[0].clone(), idx, int(idx.numel()), True)
out = torch.empty(F, dtype=torch.float32, device=dev); rot = torch.empty((F, 9), dtype=torch.float32, device=dev)
scratch = _Scratch.get(torch, dev, L.b200rmsd_scratch_bytes(F, N))
def run():
    _capi.check(L.b200rmsd_superpose_dev(dt.xyz_dev.data_ptr(), F, N, dt.frame_stride, idx.data_ptr(), int(idx.numel()),
                prep.ref.data_ptr(), prep.stats.data_ptr(), out.data_ptr(), rot.data_ptr(), None, scratch.data_ptr(),
                scratch.numel(), _stream_ptr(torch, dev)), "superpose")
for _ in range(3): run()
torch.cuda.synchronize()
geo = (ctypes.c_int * 8)()
ctypes.CDLL(_capi.LIB_PATH).b200rmsd_debug_fused_geometry(0, N, int(idx.numel()), 1, 1, geo)
print("geometry: G=%d nbuf=%d fpb=%d team_warps=%d lanes=%d smem=%d" % tuple(geo)[:6])
per = (F // 148 + 2) * 8
n = F // 148 // max(1, geo[2]) - 1   # slots per CTA
buf = (ctypes.c_longlong * (per + 148 * 2))()
fn = ctypes.CDLL(_capi.LIB_PATH).b200rmsd_debug_fused_trace
fn.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
assert fn(buf, per + 148 * 2) == 0
cta = np.array(buf[per:per + 296], dtype=np.int64).reshape(148, 2)
t = np.array(buf[:n * 8], dtype=np.int64).reshape(n, 8)
t0 = t[t > 0].min()
t = (t - t0).astype(np.float64)
names = ["arrived", "reduced", "solved", "xformed", "handed", "st_issue", "st_drain", "ld_issue"]
print("frame  " + "  ".join(f"{x:>9s}" for x in names) + "   (cycles since first stamp)")
for i in list(range(0, 12)) + list(range(min(100, n - 14), min(112, n - 2))):
    print(f"{i:5d}  " + "  ".join(f"{v:9.0f}" for v in t[i]))
d = t[min(20, n // 3):n-3]
print("mean per-frame period (cycles):", (d[-1, 4] - d[0, 4]) / (len(d) - 1))
print("mean arrived->reduced %.0f  reduced->solved %.0f  solved->xformed %.0f  xformed->handed %.0f  handed->st_issue %.0f  st_issue->drain %.0f" % tuple(
    np.mean(d[:, b] - d[:, a]) for a, b in ((0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 6))))
print("mean ld_issue->arrived (load latency incl. queueing): %.0f" % np.mean(d[:, 0] - d[:, 7]))
print("mean drain(i) -> ld_issue(i+nbuf): see rows; wait for data per frame = arrived(i) - handed(i-G)")

dur = (cta[:, 1] - cta[:, 0]) / 1e3
st = (cta[:, 0] - cta[:, 0].min()) / 1e3
print("per-CTA duration (us): min %.1f median %.1f max %.1f ; start spread %.1f us ; end-to-end %.1f us" % (dur.min(), np.median(dur), dur.max(), st.max(), (cta[:, 1].max() - cta[:, 0].min()) / 1e3))
print("slowest CTAs:", np.argsort(dur)[-8:], np.sort(dur)[-8:].round(1))
print("fastest CTAs:", np.argsort(dur)[:8], np.sort(dur)[:8].round(1))

def timeit(label):
    for _ in range(3): run()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for a, b in evs:
        a.record(); run(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    print(label, "median %.3f ms min %.3f max %.3f -> %.0f GB/s" % (ts[5], ts[0], ts[-1], 2 * F * N * 12 / ts[5] / 1e6))
timeit("with trace env   ")
del os.environ["B200RMSD_FUSED_TRACE"]
timeit("without trace env")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
print("10 back-to-back launches: %.3f ms each" % (e0.elapsed_time(e1) / 10))

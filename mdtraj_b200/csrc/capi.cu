// capi.cu -- extern "C" device entry points declared in include/b200rmsd.h: thin argument checking + kernel
// configuration.  The host-array entry points live in host_pipeline.cu, the all-pairs ones in allpairs.cu.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/b200rmsd.h"
#include "kernels.cuh"

using namespace b200;

static_assert(sizeof(RefStats) == B200RMSD_REFSTATS_BYTES, "RefStats layout is part of the ABI");

static thread_local char g_err_buf[512] = "";
namespace b200 {
int set_error(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err_buf, sizeof(g_err_buf), fmt, ap);
    va_end(ap);
    return code;
}
}  // namespace b200

namespace {

#define fail b200::set_error

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? B200RMSD_ENODEVICE \
                                                                                     : B200RMSD_ECUDA,   \
                        "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));            \
    } while (0)

struct DevInfo {
    int sm_count = 0;
    bool ok = false;
};
DevInfo g_dev[64];
std::mutex g_dev_mu;

int current_sm_count(int* sm)
{
    int dev = 0;
    CU(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(B200RMSD_EINVAL, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lk(g_dev_mu);
    if (!g_dev[dev].ok) {
        int major = 0;
        CU(cudaDeviceGetAttribute(&g_dev[dev].sm_count, cudaDevAttrMultiProcessorCount, dev));
        CU(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
        if (major != 10)
            return fail(B200RMSD_ENODEVICE, "device %d has compute capability %d.x; this library is sm_100a only", dev,
                        major);
        g_dev[dev].ok = true;
    }
    *sm = g_dev[dev].sm_count;
    return 0;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
// development overrides (kernel / geometry selection for the sweeps under tools/): compiled in only with
// -DB200RMSD_DEV_SWITCHES (python -m mdtraj_b200.build --dev); a release build never reads the environment
inline int env_int(const char* name, int dflt)
{
#ifdef B200RMSD_DEV_SWITCHES
    const char* s = getenv(name);
    return s && *s ? atoi(s) : dflt;
#else
    (void)name;
    return dflt;
#endif
}

// choose segmenting / ring geometry of the TMA kernel for n_atoms
void configure_tma(OvmParams& p)
{
    p.total_units = (p.n_atoms + 3) / 4;
    const int max_seg = env_int("B200RMSD_MAX_SEG_UNITS", kMaxSegUnits);
    p.n_seg = (p.total_units + max_seg - 1) / max_seg;
    p.seg_units = (p.total_units + p.n_seg - 1) / p.n_seg;
    const size_t budget = 232448;  // 227 KB opt-in shared memory per CTA
    const size_t ref_bytes = ((size_t)p.seg_units * 48 + 127) / 128 * 128;
    const size_t fixed = (size_t)kWarpsPerCta * kBatch * kSumStride * sizeof(double) + 2048;
    const size_t per_warp = (budget - ref_bytes - fixed) / kWarpsPerCta;
    int chunk = std::min(p.seg_units, env_int("B200RMSD_CHUNK_UNITS", 64));
    int stages = (int)std::min<size_t>(8, per_warp / ((size_t)chunk * 48));
    if (stages < 3 && chunk > 32) {
        chunk = 32;
        stages = (int)std::min<size_t>(8, per_warp / ((size_t)chunk * 48));
    }
    const int forced = env_int("B200RMSD_STAGES", 0);
    if (forced > 0) stages = std::min<int>(forced, (int)(per_warp / ((size_t)chunk * 48)));
    p.chunk_units = chunk;
    p.stages = std::max(stages, 1);
}

size_t align256(size_t x) { return (x + 255) / 256 * 256; }

// development overrides of the frame-resident kernel geometry
bool apply_fused_env(FusedParams& p, int op)
{
    const int g = env_int("B200RMSD_FUSED_GROUPS", 0), n = env_int("B200RMSD_FUSED_NBUF", 0),
              f = env_int("B200RMSD_FUSED_FPB", 0), l = env_int("B200RMSD_FUSED_LANES", 0);
    if ((g > 0 || n > 0 || f > 0 || l > 0) && !fused_override(p, op, g, n, f, l)) return false;
    if (env_int("B200RMSD_FUSED_VERBOSE", 0))
        fprintf(stderr, "b200rmsd: frame_resident op=%d n_pad=%d G=%d nbuf=%d fpb=%d team_warps=%d lanes=%d\n", op, p.n_pad,
                p.batch, p.nbuf, p.fpb, p.team_warps, p.lanes);
    return true;
}

}  // namespace

extern "C" {

int b200rmsd_abi_version(void) { return B200RMSD_ABI_VERSION; }
const char* b200rmsd_last_error(void) { return g_err_buf; }

int b200rmsd_device_info(int device, int* n_devices, int* sm_count, size_t* hbm_bytes, int* cc_major, int* cc_minor)
{
    int n = 0;
    CU(cudaGetDeviceCount(&n));
    if (n_devices) *n_devices = n;
    if (n == 0) return fail(B200RMSD_ENODEVICE, "no CUDA device");
    if (device < 0 || device >= n) return fail(B200RMSD_EINVAL, "device %d out of range (0..%d)", device, n - 1);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (hbm_bytes) *hbm_bytes = prop.totalGlobalMem;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return 0;
}

size_t b200rmsd_scratch_bytes(int64_t n_frames, int n_atoms)
{
    if (n_frames <= 0 || n_atoms <= 0) return 256;
    OvmParams p{};
    p.n_atoms = n_atoms;
    configure_tma(p);
    size_t partial = p.n_seg > 1 ? (size_t)n_frames * p.n_seg * 16 * sizeof(double) : 0;
    size_t rot = (size_t)n_frames * 9 * sizeof(float);
    size_t cen = (size_t)n_frames * 3 * sizeof(double);
    return align256(partial) + align256(rot) + align256(cen) + 256;
}

int b200rmsd_center_trace_dev(float* xyz, int64_t n_frames, int n_atoms, int64_t frame_stride, float* traces,
                              void* stream)
{
    if (!xyz || n_frames < 0 || n_atoms <= 0) return fail(B200RMSD_EINVAL, "center_trace: bad arguments");
    if (!aligned16(xyz) || frame_stride % 4 != 0 || frame_stride < 3 * (int64_t)((n_atoms + 3) / 4 * 4))
        return fail(B200RMSD_EINVAL, "center_trace: xyz must be the padded atom-major layout (16-byte aligned, "
                                     "frame_stride %% 4 == 0, >= 3*n_pad)");
    int sm = 0;
    if (int rc = current_sm_count(&sm)) return rc;
    FusedParams fp{};
    fp.xyz = xyz;
    fp.n_frames = n_frames;
    fp.frame_stride = frame_stride;
    fp.n_atoms = n_atoms;
    fp.n_pad = (n_atoms + 3) / 4 * 4;
    fp.n_sel = n_atoms;
    fp.traces = traces;
    if (!env_int("B200RMSD_NO_FUSED", 0) && fused_config(fp, OP_CENTER)) {
        if (!apply_fused_env(fp, OP_CENTER)) return fail(B200RMSD_EINVAL, "center_trace: B200RMSD_FUSED_* override does not fit");
        CU(launch_frame_resident(fp, OP_CENTER, sm, (cudaStream_t)stream));
        return 0;
    }
    CU(launch_center_trace(xyz, n_frames, n_atoms, frame_stride, traces, sm, (cudaStream_t)stream));
    return 0;
}

int b200rmsd_prepare_reference_dev(const float* ref_frame, const int32_t* idx, int n_sel, int do_center,
                                   float given_trace, float* ref_out, void* ref_stats, void* stream)
{
    if (!ref_frame || !ref_out || !ref_stats || n_sel <= 0) return fail(B200RMSD_EINVAL, "prepare_reference: bad arguments");
    int sm = 0;
    if (int rc = current_sm_count(&sm)) return rc;
    CU(launch_prepare_ref(ref_frame, idx, n_sel, do_center, given_trace, ref_out, (RefStats*)ref_stats,
                          (cudaStream_t)stream));
    return 0;
}

int b200rmsd_rmsd_dev(const float* xyz, int64_t n_frames, int n_atoms, int64_t frame_stride, const int32_t* idx,
                      int n_sel, const float* ref, const void* ref_stats, const float* traces, unsigned flags,
                      float* out_rmsd, float* out_rot, double* out_centroid, unsigned* n_degenerate, void* scratch,
                      size_t scratch_bytes, void* stream)
{
    if (!xyz || !ref || !ref_stats || !out_rmsd || n_frames < 0 || n_atoms <= 0)
        return fail(B200RMSD_EINVAL, "rmsd: bad arguments");
    if (idx && n_sel <= 0) return fail(B200RMSD_EINVAL, "rmsd: empty selection");
    const bool pre = (flags & B200RMSD_PRECENTERED) && !idx;
    if (pre && !traces) return fail(B200RMSD_EINVAL, "rmsd: B200RMSD_PRECENTERED needs traces");
    int sm = 0;
    if (int rc = current_sm_count(&sm)) return rc;

    OvmParams p{};
    p.xyz = xyz;
    p.n_frames = n_frames;
    p.frame_stride = frame_stride;
    p.n_atoms = idx ? n_sel : n_atoms;
    p.inv_n = 1.0 / (double)p.n_atoms;
    p.idx = idx;
    p.ref = ref;
    p.ref_stats = (const RefStats*)ref_stats;
    p.traces = pre ? traces : nullptr;
    p.out_rmsd = out_rmsd;
    p.out_rot = out_rot;
    p.out_centroid = out_centroid;
    p.degenerate = n_degenerate;
    p.n_seg = 1;

    const int n_pad = (n_atoms + 3) / 4 * 4;
    const bool staged = aligned16(xyz) && aligned16(ref) && frame_stride % 4 == 0 && frame_stride >= 3 * (int64_t)n_pad;
    if (idx && staged && !pre && !env_int("B200RMSD_FORCE_GATHER", 0) && !env_int("B200RMSD_NO_GROUP", 0)) {
        // small frames with a selection: stage whole frames and gather in shared memory
        p.frame_atoms = n_atoms;
        cudaError_t ge = cudaSuccess;
        if (launch_ovm_group(p, false, sm, (cudaStream_t)stream, &ge)) {
            CU(ge);
            return 0;
        }
    }
    if (!idx && staged && !env_int("B200RMSD_FORCE_GATHER", 0)) {
        if (!env_int("B200RMSD_NO_GROUP", 0)) {  // short frames: several frames per warp iteration
            cudaError_t ge = cudaSuccess;
            if (launch_ovm_group(p, pre, sm, (cudaStream_t)stream, &ge)) {
                CU(ge);
                return 0;
            }
        }
        configure_tma(p);
        if (p.stages < 2) return fail(B200RMSD_EINVAL, "rmsd: could not fit a shared-memory ring for n_atoms=%d", n_atoms);
        if (p.n_seg > 1) {
            const size_t need = (size_t)n_frames * p.n_seg * 16 * sizeof(double);
            if (!scratch || scratch_bytes < need || !aligned16(scratch))
                return fail(B200RMSD_EINVAL, "rmsd: n_atoms=%d needs %zu bytes of 16-byte aligned scratch", n_atoms, need);
            p.partials = (double*)scratch;
        }
        CU(launch_ovm_tma(p, pre, sm, (cudaStream_t)stream));
    } else {
        if (!idx && frame_stride < 3 * (int64_t)n_atoms) return fail(B200RMSD_EINVAL, "rmsd: frame_stride too small");
        CU(launch_ovm_gather(p, pre, sm, (cudaStream_t)stream));
    }
    return 0;
}

int b200rmsd_rmsd_nosuperpose_dev(const float* xyz, int64_t n_frames, int n_atoms, int64_t frame_stride,
                                  const int32_t* idx, int n_sel, const float* ref_raw, float* out_rmsd, void* stream)
{
    if (!xyz || !ref_raw || !out_rmsd || n_frames < 0 || n_atoms <= 0) return fail(B200RMSD_EINVAL, "nosuperpose: bad arguments");
    int sm = 0;
    if (int rc = current_sm_count(&sm)) return rc;
    CU(launch_nosuperpose(xyz, n_frames, idx ? n_sel : n_atoms, frame_stride, idx, ref_raw, out_rmsd, sm,
                          (cudaStream_t)stream));
    return 0;
}

int b200rmsd_superpose_dev(float* xyz, int64_t n_frames, int n_atoms, int64_t frame_stride, const int32_t* idx,
                           int n_sel, const float* ref, const void* ref_stats, float* out_rmsd, float* out_rot,
                           unsigned* n_degenerate, void* scratch, size_t scratch_bytes, void* stream)
{
    if (!xyz || !ref || !ref_stats || n_frames < 0 || n_atoms <= 0) return fail(B200RMSD_EINVAL, "superpose: bad arguments");
    if (!aligned16(xyz) || frame_stride % 4 != 0 || frame_stride < 3 * (int64_t)((n_atoms + 3) / 4 * 4))
        return fail(B200RMSD_EINVAL, "superpose: xyz must be the padded atom-major layout");
    const size_t need = b200rmsd_scratch_bytes(n_frames, n_atoms);
    if (!scratch || scratch_bytes < need || !aligned16(scratch))
        return fail(B200RMSD_EINVAL, "superpose: needs %zu bytes of 16-byte aligned scratch (b200rmsd_scratch_bytes)", need);
    // carve scratch: [segment partials][rotations][centroids]
    OvmParams cfg{};
    cfg.n_atoms = n_atoms;
    configure_tma(cfg);
    char* base = (char*)scratch;
    const size_t partial = cfg.n_seg > 1 ? align256((size_t)n_frames * cfg.n_seg * 16 * sizeof(double)) : 0;
    float* rot = out_rot ? out_rot : (float*)(base + partial);
    double* cen = (double*)(base + partial + align256((size_t)n_frames * 9 * sizeof(float)));
    // rmsd output is mandatory for the kernel; park it at the head of the centroid block's tail if not wanted
    float* rms = out_rmsd;
    if (!rms) return fail(B200RMSD_EINVAL, "superpose: out_rmsd must not be NULL in the device API");
    {
        // single-pass path: frames resident in shared memory between the rotation solve and the transform
        int sm1 = 0;
        if (int rc1 = current_sm_count(&sm1)) return rc1;
        FusedParams fp{};
        fp.xyz = xyz;
        fp.n_frames = n_frames;
        fp.frame_stride = frame_stride;
        fp.n_atoms = n_atoms;
        fp.n_pad = (n_atoms + 3) / 4 * 4;
        fp.idx = idx;
        fp.n_sel = idx ? n_sel : n_atoms;
        fp.ref = ref;
        fp.ref_stats = (const RefStats*)ref_stats;
        fp.out_rmsd = out_rmsd;
        fp.out_rot = out_rot;
        fp.degenerate = n_degenerate;
        if (!env_int("B200RMSD_NO_FUSED", 0) && aligned16(ref) && fused_config(fp, OP_SUPERPOSE)) {
            if (!apply_fused_env(fp, OP_SUPERPOSE)) return fail(B200RMSD_EINVAL, "superpose: B200RMSD_FUSED_* override does not fit");
            CU(launch_frame_resident(fp, OP_SUPERPOSE, sm1, (cudaStream_t)stream));
            return 0;
        }
    }
    int rc = b200rmsd_rmsd_dev(xyz, n_frames, n_atoms, frame_stride, idx, n_sel, ref, ref_stats, nullptr, 0u, rms, rot,
                               cen, n_degenerate, partial ? base : nullptr, partial, stream);
    if (rc) return rc;
    int sm = 0;
    if ((rc = current_sm_count(&sm))) return rc;
    ApplyParams a{};
    a.xyz = xyz;
    a.n_frames = n_frames;
    a.frame_stride = frame_stride;
    a.n_atoms = n_atoms;
    a.rot = rot;
    a.centroid = cen;
    a.ref_stats = (const RefStats*)ref_stats;
    CU(launch_apply_transform(a, sm, (cudaStream_t)stream));
    return 0;
}

int b200rmsd_rotate_dev(float* xyz, int64_t n_frames, int n_atoms, int64_t frame_stride, const float* rot,
                        void* scratch, size_t scratch_bytes, void* stream)
{
    (void)scratch;
    (void)scratch_bytes;
    if (!xyz || !rot || n_frames < 0 || n_atoms <= 0) return fail(B200RMSD_EINVAL, "rotate: bad arguments");
    if (!aligned16(xyz) || frame_stride % 4 != 0 || frame_stride < 3 * (int64_t)((n_atoms + 3) / 4 * 4))
        return fail(B200RMSD_EINVAL, "rotate: xyz must be the padded atom-major layout");
    int sm = 0;
    if (int rc = current_sm_count(&sm)) return rc;
    ApplyParams a{};
    a.xyz = xyz;
    a.n_frames = n_frames;
    a.frame_stride = frame_stride;
    a.n_atoms = n_atoms;
    a.rot = rot;
    a.centroid = nullptr;
    a.ref_stats = nullptr;
    CU(launch_apply_transform(a, sm, (cudaStream_t)stream));
    return 0;
}

size_t b200rmsd_rmsf_scratch_bytes(int64_t n_frames, int n_sel)
{
    if (n_frames <= 0 || n_sel <= 0) return 256;
    return (size_t)rmsf_chunks(n_frames) * (size_t)n_sel * 4 * sizeof(double) + 256;
}

int b200rmsd_rmsf_dev(const float* xyz, int64_t n_frames, int n_atoms, int64_t frame_stride, const int32_t* idx,
                      int n_sel, const float* rot, const double* centroid, void* scratch, size_t scratch_bytes,
                      float* out_rmsf, void* stream)
{
    if (!xyz || !out_rmsf || n_frames <= 0 || n_atoms <= 0) return fail(B200RMSD_EINVAL, "rmsf: bad arguments");
    const int ns = idx ? n_sel : n_atoms;
    if (ns <= 0) return fail(B200RMSD_EINVAL, "rmsf: empty selection");
    if (!scratch || scratch_bytes < b200rmsd_rmsf_scratch_bytes(n_frames, ns) || (reinterpret_cast<uintptr_t>(scratch) & 7u))
        return fail(B200RMSD_EINVAL, "rmsf: needs %zu bytes of 8-byte aligned scratch", b200rmsd_rmsf_scratch_bytes(n_frames, ns));
    CU(launch_rmsf(xyz, n_frames, frame_stride, idx, ns, rot, centroid, (double*)scratch, out_rmsf, (cudaStream_t)stream));
    return 0;
}

int b200rmsd_rot_msd_dev(const float* a_frame, const float* b_xyz, int64_t n_frames, int n_atoms, int64_t frame_stride,
                         const float* rot, int transpose, float* rot_out, float* out_rmsd, void* stream)
{
    if (!a_frame || !b_xyz || !rot || !out_rmsd || n_frames < 0 || n_atoms <= 0) return fail(B200RMSD_EINVAL, "rot_msd: bad arguments");
    int sm = 0;
    if (int rc = current_sm_count(&sm)) return rc;
    CU(launch_rot_msd(a_frame, b_xyz, n_frames, n_atoms, frame_stride, rot, transpose, rot_out, out_rmsd, sm, (cudaStream_t)stream));
    return 0;
}

}  // extern "C"

// host_pipeline.cu -- the "_host" entry points of include/b200rmsd.h: md.rmsd / Trajectory.superpose /
// _center_inplace_atom_major on HOST arrays, streamed through one or several GPUs from one process.
//
// Every frame crosses PCIe once (12 bytes per atom per frame), so this path is a bus problem:
//
//   * frames are cut into chunks (64 MB of padded coordinates by default) handed out dynamically -- one atomic counter --
//     to the devices taking part (b200rmsd_*_host_multi; the single-device entry points are the n_devices = 1 case);
//     one host thread per device drives it, frames are independent so there is no exchange between devices;
//   * per device three independent LANES (stream + device chunk buffer + page-locked staging buffers + event), so the
//     copy-in of chunk c+1, the kernel of chunk c and the copy-out of chunk c-1 overlap;
//   * pageable host memory -- what every numpy array is -- cannot be DMA'd asynchronously: cudaMemcpyAsync from it is staged
//     by the driver through a small bounce buffer and serialises the lanes.  Here it is staged explicitly: a pool of
//     memcpy threads fills the lane's page-locked buffer (in the padded device layout) while the previous lane's DMA is in
//     flight, and results / superposed coordinates come back the same way through a finalizer thread.  Page-locked user
//     memory (cudaHostAlloc / cudaHostRegister / torch pin_memory) is detected and DMA'd directly;
//   * on every error path all lane streams are synchronised before returning, so no DMA is still writing into the
//     caller's buffers when the Python wrapper raises (after a failed in-place call the contents of xyz are undefined).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/b200rmsd.h"
#include "kernels.cuh"

using namespace b200;
#define fail b200::set_error

namespace {

// ---------------------------------------------------------------------------------------------
// process-wide settings (b200rmsd_host_configure)
// ---------------------------------------------------------------------------------------------
int g_copy_threads = 0;        // 0: not decided yet
int g_chunk_mb = 64;           // page-locked caller memory: DMA'd directly, large chunks amortise the per-chunk launches
int g_staged_chunk_mb = 16;    // pageable caller memory: three lanes of staging buffers should stay inside the host's L3,
                               // so that the DMA engine reads what the memcpy pool has just written from cache
                               // (measured on a 16-core Xeon, 60 MB L3: 16 MB 52 GB/s, 64 MB 46 GB/s, page-locked 55 GB/s)

int g_stage_piece_kb = 0;       // streamed staging (below): 0 = automatic (stage_piece_bytes), > 0 = KB per piece, < 0 = off (the
                               // whole chunk is staged into the lane's buffer and sent with one copy)

// ---------------------------------------------------------------------------------------------
// Streamed staging of pageable memory.  Staging a whole 16 MB chunk and then sending it costs the host's memory system
// three passes per byte (read the caller's array, write the staging buffer, DMA-read the staging buffer): measured on the
// 8-GPU box (one socket, 32 hardware threads, ~200 GB/s of DRAM bandwidth) the eight ranks of a torchrun job together
// staged 63-67 GB/s whatever the chunk size, against 187 GB/s from page-locked arrays.  Here every pool thread owns two
// small page-locked slots: it copies a PIECE of the chunk (a few hundred KB of whole rows) into a slot, sends that slot
// to its place in the device chunk on the lane's stream right away and records an event; when it comes back to the slot two
// pieces later it waits for that event.  All slots of all threads (and of all ranks sharing the host) together are a few
// tens of MB, so the staging writes and the DMA reads stay in the last-level cache and DRAM sees one pass per byte.
// ---------------------------------------------------------------------------------------------
constexpr size_t kSlotCap = 2u << 20;   // bytes per slot (allocation); a piece is whole frames, at most this large

struct StageSlot {
    char* buf = nullptr;
    cudaEvent_t ev[64] = {};
    int last_dev = -1;
};
std::atomic<bool> g_pool_exiting{false};   // set when the pool is torn down at process exit: CUDA may be gone by then

struct StageSlots {            // thread-local; freed when a (per-call device) thread ends, left to the OS at process exit
    StageSlot slot[2];
    int next = 0;
    ~StageSlots()
    {
        if (g_pool_exiting.load()) return;
        for (StageSlot& s : slot) {
            if (s.last_dev >= 0) cudaEventSynchronize(s.ev[s.last_dev]);
            for (cudaEvent_t e : s.ev)
                if (e) cudaEventDestroy(e);
            if (s.buf) cudaFreeHost(s.buf);
        }
        cudaGetLastError();
    }
};
thread_local StageSlots tl_stage;

struct StageGroup {            // the streamed staging of one chunk
    int dev = 0;
    cudaStream_t stream = nullptr;
    std::atomic<int> err{0};   // first cudaError_t seen by any piece
};

int local_ranks()   // processes torchrun started on this host (they share its memory system); 1 when not under torchrun
{
    const char* lw = getenv("LOCAL_WORLD_SIZE");
    const int n = lw ? atoi(lw) : 1;
    return n > 1 ? n : 1;
}

// Piece size of the streamed staging; 0 = stage whole chunks.  Measured, md.rmsd on 1,000-atom frames
// (profiles/r02_host_staging.jsonl; rmsd/s aggregate, whole chunks -> streamed):
//   8 ranks x 4 threads   5.1e6 -> 9.4e6 (512 KB pieces; 1 MB 8.4e6; 2 MB = 128 MB of slots 6.4e6; page-locked input 15.6e6)
//   4 ranks x 8 threads   5.4e6 -> 7.6e6 (1 MB; 512 KB 7.4e6)
//   2 ranks x 16 threads  5.9e6 -> 6.7e6 (1 MB; 512 KB 3.8e6: every piece costs three CUDA calls on a stream all the
//                         threads of a process contend for, ~24 us together, i.e. ~42k pieces/s per process)
//   1 rank x 16 threads   52 GB/s -> 42 GB/s on one box, 38 -> 42 on another: three 16 MB lanes already fit the 60 MB L3
// so: streamed when several ranks share the host, ~32 MB of slots per host, pieces of 512 KB (few threads per rank) to 1 MB.
size_t stage_piece_bytes(int pool_threads)
{
    if (g_stage_piece_kb < 0) return 0;
    if (g_stage_piece_kb > 0) return std::min<size_t>((size_t)g_stage_piece_kb << 10, kSlotCap);
    const int ranks = local_ranks();
    if (ranks < 2) return 0;
    const size_t piece = ((size_t)32 << 20) / ((size_t)ranks * 2 * (size_t)(pool_threads + 1));
    const size_t lo = pool_threads + 1 >= 12 ? (1u << 20) : (512u << 10);
    return std::min<size_t>(1u << 20, std::max<size_t>(lo, piece));
}

// ---------------------------------------------------------------------------------------------
// memcpy pool: row-wise copies between the caller's (F, n_atoms, 3) array and the padded staging layout
// ---------------------------------------------------------------------------------------------
struct CopyTask {
    char* dst;                 // plain copy: destination; streamed staging: DEVICE address of the piece
    const char* src;
    size_t dst_pitch, src_pitch, width, pad;  // per row: copy `width` bytes, then zero `pad` bytes
    int64_t rows;
    std::atomic<int>* pending;
    StageGroup* stage;         // non-null: rows go through one of this thread's slots and on to the device
};

class CopyPool {
public:
    static CopyPool& get()
    {
        static CopyPool pool;
        return pool;
    }
    // Blocking: returns when every row is copied.  The calling thread works too.
    void copy_rows(char* dst, size_t dst_pitch, const char* src, size_t src_pitch, size_t width, size_t pad, int64_t rows)
    {
        if (rows <= 0) return;
        ensure_started();
        const size_t bytes = (size_t)rows * width;
        int parts = (int)std::min<size_t>(n_threads_ + 1, std::max<size_t>(1, bytes >> 19));  // >= 512 KB per part
        parts = (int)std::min<int64_t>(parts, rows);
        std::atomic<int> pending(parts);
        const int64_t per = (rows + parts - 1) / parts;
        CopyTask mine{};
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (int p = 0; p < parts; ++p) {
                const int64_t r0 = (int64_t)p * per, r1 = std::min(rows, r0 + per);
                CopyTask t{dst + (size_t)r0 * dst_pitch, src + (size_t)r0 * src_pitch, dst_pitch, src_pitch, width, pad,
                           std::max<int64_t>(0, r1 - r0), &pending, nullptr};
                if (p == 0) mine = t; else q_.push_back(t);
            }
        }
        if (parts > n_threads_ / 2) cv_.notify_all();
        else for (int p = 1; p < parts; ++p) cv_.notify_one();
        run(mine);
        while (pending.load(std::memory_order_acquire) > 0) {  // help with whatever is queued, then wait
            CopyTask t;
            if (try_pop(t)) run(t); else std::this_thread::yield();
        }
    }
    // Streamed staging of `rows` rows into the device chunk at `dev_dst` (row pitch dst_pitch there and in the slots), in
    // pieces of `piece_rows` rows; blocking until every piece has been ISSUED on grp.stream (stream order then carries
    // the dependency to whatever the caller enqueues next).  Returns the first CUDA error of any piece.
    cudaError_t stage_rows(StageGroup& grp, char* dev_dst, size_t dst_pitch, const char* src, size_t src_pitch, size_t width,
                           size_t pad, int64_t rows, int64_t piece_rows)
    {
        if (rows <= 0) return cudaSuccess;
        ensure_started();
        const int64_t parts = (rows + piece_rows - 1) / piece_rows;
        std::atomic<int> pending((int)parts);
        CopyTask mine{};
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (int64_t p = 0; p < parts; ++p) {
                const int64_t r0 = p * piece_rows, r1 = std::min(rows, r0 + piece_rows);
                CopyTask t{dev_dst + (size_t)r0 * dst_pitch, src + (size_t)r0 * src_pitch, dst_pitch, src_pitch, width, pad,
                           r1 - r0, &pending, &grp};
                if (p == 0) mine = t; else q_.push_back(t);
            }
        }
        cv_.notify_all();
        run(mine);
        while (pending.load(std::memory_order_acquire) > 0) {
            CopyTask t;
            if (try_pop(t)) run(t); else std::this_thread::yield();
        }
        return (cudaError_t)grp.err.load();
    }
    int threads()
    {
        ensure_started();
        return n_threads_;
    }

private:
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<CopyTask> q_;
    bool stop_ = false, started_ = false;
    int n_threads_ = 0;

    static void run_staged(const CopyTask& t)
    {
        StageGroup& g = *t.stage;
        StageSlots& tl = tl_stage;
        auto check = [&](cudaError_t e) {
            int zero = 0;
            if (e != cudaSuccess) g.err.compare_exchange_strong(zero, (int)e);
            return e == cudaSuccess;
        };
        if (g.err.load() != 0) return;
        // The piece may belong to another device than the one this thread last worked for -- also when the thread is a
        // device's own host thread helping out while it waits (it must find its own device current again afterwards)
        int home = -1;
        if (!check(cudaGetDevice(&home))) return;
        struct Restore {
            int home, dev;
            ~Restore() { if (home != dev) cudaSetDevice(home); }
        } restore{home, g.dev};
        if (home != g.dev && !check(cudaSetDevice(g.dev))) return;
        StageSlot& sl = tl.slot[tl.next];
        tl.next ^= 1;
        if (!sl.buf && !check(cudaHostAlloc((void**)&sl.buf, kSlotCap, cudaHostAllocPortable))) return;
        if (sl.last_dev >= 0 && !check(cudaEventSynchronize(sl.ev[sl.last_dev]))) return;  // the copy that last read it
        if (t.pad == 0 && t.src_pitch == t.width) {
            memcpy(sl.buf, t.src, (size_t)t.rows * t.width);
        } else {
            for (int64_t r = 0; r < t.rows; ++r) {
                memcpy(sl.buf + (size_t)r * t.dst_pitch, t.src + (size_t)r * t.src_pitch, t.width);
                if (t.pad) memset(sl.buf + (size_t)r * t.dst_pitch + t.width, 0, t.pad);
            }
        }
        if (!check(cudaMemcpyAsync(t.dst, sl.buf, (size_t)t.rows * t.dst_pitch, cudaMemcpyHostToDevice, g.stream))) return;
        if (!sl.ev[g.dev] && !check(cudaEventCreateWithFlags(&sl.ev[g.dev], cudaEventDisableTiming))) return;
        if (check(cudaEventRecord(sl.ev[g.dev], g.stream))) sl.last_dev = g.dev;
    }
    static void run(const CopyTask& t)
    {
        if (t.stage) {
            run_staged(t);
            t.pending->fetch_sub(1, std::memory_order_release);
            return;
        }
        if (t.pad == 0 && t.dst_pitch == t.width && t.src_pitch == t.width) {
            memcpy(t.dst, t.src, (size_t)t.rows * t.width);
        } else {
            for (int64_t r = 0; r < t.rows; ++r) {
                memcpy(t.dst + (size_t)r * t.dst_pitch, t.src + (size_t)r * t.src_pitch, t.width);
                if (t.pad) memset(t.dst + (size_t)r * t.dst_pitch + t.width, 0, t.pad);
            }
        }
        t.pending->fetch_sub(1, std::memory_order_release);
    }
    bool try_pop(CopyTask& t)
    {
        std::lock_guard<std::mutex> lk(mu_);
        if (q_.empty()) return false;
        t = q_.front();
        q_.pop_front();
        return true;
    }
    void ensure_started()
    {
        std::lock_guard<std::mutex> lk(mu_);
        if (started_) return;
        int n = g_copy_threads;
        if (n <= 0) {
            const unsigned hw = std::thread::hardware_concurrency();
            n = (int)std::min<unsigned>(24, std::max<unsigned>(2, hw > 2 ? hw - 1 : 1));
        }
        n_threads_ = n;
        for (int i = 0; i < n; ++i)
            th_.emplace_back([this] {
                for (;;) {
                    CopyTask t;
                    {
                        std::unique_lock<std::mutex> lk(mu_);
                        cv_.wait(lk, [this] { return stop_ || !q_.empty(); });
                        if (stop_ && q_.empty()) return;
                        t = q_.front();
                        q_.pop_front();
                    }
                    run(t);
                }
            });
        started_ = true;
    }
    CopyPool() = default;
    ~CopyPool()
    {
        g_pool_exiting.store(true);
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
};

// ---------------------------------------------------------------------------------------------
// per-device workspace
// ---------------------------------------------------------------------------------------------
constexpr int kLanes = 3;

struct Lane {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    float* xyz = nullptr;        // device chunk
    float* out = nullptr;        // device per-frame results
    float* rot = nullptr;
    float* trc = nullptr;
    void* scratch = nullptr;
    char* up = nullptr;          // page-locked staging, host -> device
    char* down = nullptr;        // page-locked staging, device -> host (in-place operations)
    float* res = nullptr;        // page-locked per-frame results: [rmsd | traces (fpc)] [rot (9 fpc)]
    bool busy = false;           // owned by the finalizer until the chunk's results are in the caller's buffers
    int64_t chunk = -1;
};

struct Workspace {
    bool init = false;
    Lane lane[kLanes];
    cudaEvent_t ref_ready = nullptr;
    size_t xyz_bytes = 0, up_bytes = 0, down_bytes = 0, per_frame_cap = 0, scratch_bytes = 0, ref_cap = 0, idx_cap = 0;
    float* ref_raw = nullptr;    // full reference frame as uploaded
    float* ref_sel = nullptr;    // prepared (centred / packed)
    int32_t* idx = nullptr;
    int32_t* ref_idx = nullptr;
    RefStats* stats = nullptr;
    unsigned* degen = nullptr;
    std::mutex mu;               // one host call at a time per device
    std::mutex lane_mu;          // busy flags
    std::condition_variable lane_cv;
};
Workspace g_ws[64];

#define CUW(call)                                                                                           \
    do {                                                                                                    \
        cudaError_t e_ = (call);                                                                            \
        if (e_ != cudaSuccess)                                                                              \
            return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? B200RMSD_ENODEVICE   \
                        : e_ == cudaErrorMemoryAllocation                            ? B200RMSD_ENOMEM      \
                                                                                     : B200RMSD_ECUDA,      \
                        "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));               \
    } while (0)

template <class T>
cudaError_t regrow_dev(T*& p, size_t bytes)
{
    if (p) cudaFree(p);
    p = nullptr;
    return cudaMalloc((void**)&p, bytes);
}
template <class T>
cudaError_t regrow_pinned(T*& p, size_t bytes)
{
    if (p) cudaFreeHost(p);
    p = nullptr;
    return cudaHostAlloc((void**)&p, bytes, cudaHostAllocPortable);
}

int ws_prepare(Workspace& w, size_t chunk_bytes, size_t chunk_frames, size_t scratch_bytes, size_t ref_atoms, size_t n_idx,
               bool need_up, bool need_down)
{
    if (!w.init) {
        for (int l = 0; l < kLanes; ++l) {
            CUW(cudaStreamCreateWithFlags(&w.lane[l].stream, cudaStreamNonBlocking));
            CUW(cudaEventCreateWithFlags(&w.lane[l].done, cudaEventDisableTiming));
        }
        CUW(cudaEventCreateWithFlags(&w.ref_ready, cudaEventDisableTiming));
        CUW(cudaMalloc((void**)&w.stats, sizeof(RefStats)));
        CUW(cudaMalloc((void**)&w.degen, sizeof(unsigned)));
        w.init = true;
    }
    if (chunk_bytes > w.xyz_bytes) {
        for (int l = 0; l < kLanes; ++l) CUW(regrow_dev(w.lane[l].xyz, chunk_bytes));
        w.xyz_bytes = chunk_bytes;
    }
    if (need_up && chunk_bytes > w.up_bytes) {
        for (int l = 0; l < kLanes; ++l) CUW(regrow_pinned(w.lane[l].up, chunk_bytes));
        w.up_bytes = chunk_bytes;
    }
    if (need_down && chunk_bytes > w.down_bytes) {
        for (int l = 0; l < kLanes; ++l) CUW(regrow_pinned(w.lane[l].down, chunk_bytes));
        w.down_bytes = chunk_bytes;
    }
    if (chunk_frames > w.per_frame_cap) {
        for (int l = 0; l < kLanes; ++l) {
            CUW(regrow_dev(w.lane[l].out, chunk_frames * sizeof(float)));
            CUW(regrow_dev(w.lane[l].rot, chunk_frames * 9 * sizeof(float)));
            CUW(regrow_dev(w.lane[l].trc, chunk_frames * sizeof(float)));
            CUW(regrow_pinned(w.lane[l].res, chunk_frames * 10 * sizeof(float)));
        }
        w.per_frame_cap = chunk_frames;
    }
    if (scratch_bytes > w.scratch_bytes) {
        for (int l = 0; l < kLanes; ++l) CUW(regrow_dev(w.lane[l].scratch, scratch_bytes));
        w.scratch_bytes = scratch_bytes;
    }
    if (ref_atoms > w.ref_cap) {
        const size_t b = (ref_atoms + 4) * 3 * sizeof(float);
        CUW(regrow_dev(w.ref_raw, b));
        CUW(regrow_dev(w.ref_sel, b));
        w.ref_cap = ref_atoms;
    }
    if (n_idx > w.idx_cap) {
        CUW(regrow_dev(w.idx, n_idx * sizeof(int32_t)));
        CUW(regrow_dev(w.ref_idx, n_idx * sizeof(int32_t)));
        w.idx_cap = n_idx;
    }
    return 0;
}

bool is_page_locked(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

// ---------------------------------------------------------------------------------------------
// one host call, shared by the device threads
// ---------------------------------------------------------------------------------------------
enum HostOp { HOP_RMSD = 0, HOP_SUPERPOSE = 1, HOP_CENTER = 2 };

struct HostJob {
    HostOp op;
    const float* in = nullptr;   // caller's coordinates (read)
    float* inout = nullptr;      // the same array when the operation writes it back
    int64_t n_frames = 0;
    int n_atoms = 0, n_pad = 0;
    const float* ref_frame = nullptr;
    int n_atoms_ref = 0;
    const int32_t* idx = nullptr;
    const int32_t* ref_idx = nullptr;
    int n_sel = 0, superpose = 1, precentered = 0;
    const float* traces = nullptr;
    float ref_trace = 0.f;
    float* out_rmsd = nullptr;
    float* out_rot = nullptr;
    float* out_traces = nullptr;
    int64_t fpc = 0, n_chunks = 0;
    bool src_locked = false;     // caller's coordinates are page-locked: DMA them directly
    int64_t stage_piece_rows = 0;  // > 0: pageable coordinates go up by streamed staging, this many frames per piece
    std::atomic<int64_t> next_chunk{0};
    std::atomic<int> rc{0};
    std::atomic<unsigned> degenerate{0};
    std::mutex err_mu;
    char err[512] = "";
    void set_error(int code)
    {
        std::lock_guard<std::mutex> lk(err_mu);
        if (rc.load() == 0) {
            snprintf(err, sizeof(err), "%s", b200rmsd_last_error());
            rc.store(code);
        }
    }
};

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int sm_count_of(int dev, int* sm)
{
    int major = 0;
    CUW(cudaDeviceGetAttribute(sm, cudaDevAttrMultiProcessorCount, dev));
    CUW(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) return fail(B200RMSD_ENODEVICE, "device %d has compute capability %d.x; this library is sm_100a only", dev, major);
    return 0;
}

inline int job_device(const Workspace& w) { return (int)(&w - g_ws); }

// results of the chunk in lane L -> the caller's buffers (runs after L.done)
void finalize_lane(HostJob& job, Lane& L)
{
    const int64_t f0 = L.chunk * job.fpc, nf = std::min(job.fpc, job.n_frames - f0);
    if (job.out_rmsd) memcpy(job.out_rmsd + f0, L.res, (size_t)nf * 4);
    if (job.out_traces) memcpy(job.out_traces + f0, L.res, (size_t)nf * 4);
    if (job.out_rot) memcpy(job.out_rot + f0 * 9, L.res + job.fpc, (size_t)nf * 36);
    if (job.inout && !(job.src_locked && job.n_pad == job.n_atoms)) {
        CopyPool::get().copy_rows((char*)(job.inout + (size_t)f0 * job.n_atoms * 3), (size_t)job.n_atoms * 12, L.down,
                                  (size_t)job.n_pad * 12, (size_t)job.n_atoms * 12, 0, nf);
    }
}

// issue one chunk into lane L (stream-ordered; returns after the host-side staging copy)
int issue_chunk(HostJob& job, Workspace& w, Lane& L, int64_t c, int sm)
{
    const int64_t f0 = c * job.fpc, nf = std::min(job.fpc, job.n_frames - f0);
    const int n_atoms = job.n_atoms, n_pad = job.n_pad;
    const float* src = job.in + (size_t)f0 * n_atoms * 3;
    cudaStream_t st = L.stream;
    const size_t padded_bytes = (size_t)nf * n_pad * 12;
    const bool small = padded_bytes < (256u << 10);  // the driver's own bounce buffer is as good for small copies
    if (job.src_locked || small) {
        if (n_pad == n_atoms) {
            CUW(cudaMemcpyAsync(L.xyz, src, padded_bytes, cudaMemcpyHostToDevice, st));
        } else {
            CUW(cudaMemsetAsync(L.xyz, 0, padded_bytes, st));
            CUW(cudaMemcpy2DAsync(L.xyz, (size_t)n_pad * 12, src, (size_t)n_atoms * 12, (size_t)n_atoms * 12, (size_t)nf,
                                  cudaMemcpyHostToDevice, st));
        }
    } else if (job.stage_piece_rows > 0) {  // streamed staging through the pool threads' cache-resident slots
        StageGroup grp;
        grp.dev = job_device(w);
        grp.stream = st;
        CUW(CopyPool::get().stage_rows(grp, (char*)L.xyz, (size_t)n_pad * 12, (const char*)src, (size_t)n_atoms * 12,
                                       (size_t)n_atoms * 12, (size_t)(n_pad - n_atoms) * 12, nf, job.stage_piece_rows));
    } else {
        CopyPool::get().copy_rows(L.up, (size_t)n_pad * 12, (const char*)src, (size_t)n_atoms * 12, (size_t)n_atoms * 12,
                                  (size_t)(n_pad - n_atoms) * 12, nf);
        CUW(cudaMemcpyAsync(L.xyz, L.up, padded_bytes, cudaMemcpyHostToDevice, st));
    }
    const int32_t* didx = job.idx ? w.idx : nullptr;
    int rc = 0;
    if (job.op == HOP_RMSD) {
        if (job.superpose) {
            if (job.precentered) CUW(cudaMemcpyAsync(L.trc, job.traces + f0, (size_t)nf * 4, cudaMemcpyHostToDevice, st));
            rc = b200rmsd_rmsd_dev(L.xyz, nf, n_atoms, (int64_t)n_pad * 3, didx, job.n_sel, w.ref_sel, w.stats,
                                   job.precentered ? L.trc : nullptr, job.precentered ? B200RMSD_PRECENTERED : 0u, L.out,
                                   nullptr, nullptr, nullptr, L.scratch, w.scratch_bytes, st);
        } else {
            rc = b200rmsd_rmsd_nosuperpose_dev(L.xyz, nf, n_atoms, (int64_t)n_pad * 3, didx, job.n_sel, w.ref_sel, L.out, st);
        }
        if (rc) return rc;
        CUW(cudaMemcpyAsync(L.res, L.out, (size_t)nf * 4, cudaMemcpyDeviceToHost, st));
    } else {
        if (job.op == HOP_SUPERPOSE) {
            rc = b200rmsd_superpose_dev(L.xyz, nf, n_atoms, (int64_t)n_pad * 3, didx, job.n_sel, w.ref_sel, w.stats, L.out,
                                        L.rot, w.degen, L.scratch, w.scratch_bytes, st);
            if (rc) return rc;
            if (job.out_rmsd) CUW(cudaMemcpyAsync(L.res, L.out, (size_t)nf * 4, cudaMemcpyDeviceToHost, st));
            if (job.out_rot) CUW(cudaMemcpyAsync(L.res + job.fpc, L.rot, (size_t)nf * 36, cudaMemcpyDeviceToHost, st));
        } else {
            rc = b200rmsd_center_trace_dev(L.xyz, nf, n_atoms, (int64_t)n_pad * 3, L.trc, st);
            if (rc) return rc;
            if (job.out_traces) CUW(cudaMemcpyAsync(L.res, L.trc, (size_t)nf * 4, cudaMemcpyDeviceToHost, st));
        }
        if (job.src_locked && n_pad == n_atoms)  // page-locked caller memory in the device layout: DMA straight back
            CUW(cudaMemcpyAsync(job.inout + (size_t)f0 * n_atoms * 3, L.xyz, padded_bytes, cudaMemcpyDeviceToHost, st));
        else
            CUW(cudaMemcpyAsync(L.down, L.xyz, padded_bytes, cudaMemcpyDeviceToHost, st));
    }
    CUW(cudaEventRecord(L.done, st));
    (void)sm;
    return 0;
}

// one device's share of the job; runs on its own host thread (or inline on the caller's for the first device)
void run_device(HostJob& job, int dev)
{
    if (job.next_chunk.load() >= job.n_chunks || job.rc.load() != 0) return;
    DeviceGuard guard(dev);
    if (!guard.ok) {
        fail(B200RMSD_ENODEVICE, "host pipeline: cannot select CUDA device %d", dev);
        job.set_error(B200RMSD_ENODEVICE);
        return;
    }
    Workspace& w = g_ws[dev];
    std::lock_guard<std::mutex> lk(w.mu);
    int sm = 0;
    int rc = sm_count_of(dev, &sm);
    const int n_use = job.idx ? job.n_sel : job.n_atoms;
    const bool writes_back = job.op != HOP_RMSD;
    if (!rc)
        rc = ws_prepare(w, (size_t)job.fpc * job.n_pad * 12, (size_t)job.fpc,
                        job.op == HOP_CENTER ? 256 : b200rmsd_scratch_bytes(job.fpc, job.n_atoms),
                        job.op == HOP_CENTER ? 4 : (size_t)std::max(job.n_atoms_ref, n_use), job.idx ? (size_t)job.n_sel : 0,
                        !job.src_locked && job.stage_piece_rows == 0,
                        writes_back && !(job.src_locked && job.n_pad == job.n_atoms));
    auto cu = [&](cudaError_t e) {
        if (e != cudaSuccess && !rc) rc = fail(B200RMSD_ECUDA, "host pipeline (device %d): %s", dev, cudaGetErrorString(e));
    };
    if (!rc && job.op != HOP_CENTER) {  // reference frame: upload, gather / centre, statistics (once per call per device)
        cudaStream_t s0 = w.lane[0].stream;
        if (job.op == HOP_SUPERPOSE) cu(cudaMemsetAsync(w.degen, 0, sizeof(unsigned), s0));
        cu(cudaMemcpyAsync(w.ref_raw, job.ref_frame, (size_t)job.n_atoms_ref * 12, cudaMemcpyHostToDevice, s0));
        if (job.idx) {
            cu(cudaMemcpyAsync(w.idx, job.idx, (size_t)job.n_sel * 4, cudaMemcpyHostToDevice, s0));
            cu(cudaMemcpyAsync(w.ref_idx, job.ref_idx, (size_t)job.n_sel * 4, cudaMemcpyHostToDevice, s0));
        }
        // superpose: centred packed reference; no-superpose / precentered: packed as it is (do_center = 0)
        const int do_center = job.op == HOP_SUPERPOSE || (job.superpose && !job.precentered);
        cu(launch_prepare_ref(w.ref_raw, job.idx ? w.ref_idx : nullptr, n_use, do_center, job.ref_trace, w.ref_sel, w.stats, s0));
        cu(cudaEventRecord(w.ref_ready, s0));
        for (int l = 1; l < kLanes; ++l) cu(cudaStreamWaitEvent(w.lane[l].stream, w.ref_ready, 0));
    }
    if (rc) {
        job.set_error(rc);
        return;
    }

    // finalizer: waits for each issued chunk in order and moves its results into the caller's buffers
    std::deque<int> issued;  // lane indices, in issue order; -1 = no more
    std::mutex q_mu;
    std::condition_variable q_cv;
    auto finalizer = [&] {
        cudaSetDevice(dev);
        for (;;) {
            int l;
            {
                std::unique_lock<std::mutex> ql(q_mu);
                q_cv.wait(ql, [&] { return !issued.empty(); });
                l = issued.front();
                issued.pop_front();
            }
            if (l < 0) return;
            Lane& L = w.lane[l];
            const cudaError_t e = cudaEventSynchronize(L.done);
            if (e != cudaSuccess) {
                fail(B200RMSD_ECUDA, "host pipeline (device %d): %s", dev, cudaGetErrorString(e));
                job.set_error(B200RMSD_ECUDA);
            } else if (job.rc.load() == 0) {
                finalize_lane(job, L);
            }
            {
                std::lock_guard<std::mutex> ll(w.lane_mu);
                L.busy = false;
            }
            w.lane_cv.notify_all();
        }
    };
    const bool threaded = job.n_chunks > 1;
    std::thread fin;
    if (threaded) fin = std::thread(finalizer);

    int i = 0;
    for (;;) {
        if (job.rc.load() != 0) break;
        const int64_t c = job.next_chunk.fetch_add(1);
        if (c >= job.n_chunks) break;
        const int l = i++ % kLanes;
        Lane& L = w.lane[l];
        if (threaded) {
            std::unique_lock<std::mutex> ll(w.lane_mu);
            w.lane_cv.wait(ll, [&] { return !L.busy; });
            L.busy = true;
        }
        L.chunk = c;
        rc = issue_chunk(job, w, L, c, sm);
        if (rc) {
            job.set_error(rc);
            if (threaded) {
                std::lock_guard<std::mutex> ll(w.lane_mu);
                L.busy = false;
            }
            break;
        }
        if (threaded) {
            {
                std::lock_guard<std::mutex> ql(q_mu);
                issued.push_back(l);
            }
            q_cv.notify_one();
        } else {
            const cudaError_t e = cudaEventSynchronize(L.done);
            if (e != cudaSuccess) {
                fail(B200RMSD_ECUDA, "host pipeline (device %d): %s", dev, cudaGetErrorString(e));
                job.set_error(B200RMSD_ECUDA);
            } else {
                finalize_lane(job, L);
            }
        }
    }
    if (threaded) {
        {
            std::lock_guard<std::mutex> ql(q_mu);
            issued.push_back(-1);
        }
        q_cv.notify_one();
        fin.join();
    }
    // nothing of this call may still be in flight when we return, on success or on error
    for (int l = 0; l < kLanes; ++l) cudaStreamSynchronize(w.lane[l].stream);
    if (job.op == HOP_SUPERPOSE && job.rc.load() == 0) {
        unsigned d = 0;
        if (cudaMemcpy(&d, w.degen, sizeof(unsigned), cudaMemcpyDeviceToHost) == cudaSuccess) job.degenerate.fetch_add(d);
    }
}

int run_job(HostJob& job, const int* devices, int n_devices, const char* what)
{
    if (!devices || n_devices <= 0) return fail(B200RMSD_EINVAL, "%s: no devices given", what);
    for (int i = 0; i < n_devices; ++i) {
        if (devices[i] < 0 || devices[i] >= 64) return fail(B200RMSD_EINVAL, "%s: device %d", what, devices[i]);
        for (int j = 0; j < i; ++j)
            if (devices[j] == devices[i]) return fail(B200RMSD_EINVAL, "%s: device %d listed twice", what, devices[i]);
    }
    if (job.n_frames == 0) return 0;
    job.n_pad = (job.n_atoms + 3) / 4 * 4;
    const size_t frame_bytes = (size_t)job.n_pad * 12;
    job.src_locked = is_page_locked(job.in);
    // in-place operations keep whole-chunk staging both ways: a streamed download (piece by piece through the same slots)
    // measured slower at every rank count (the thread waits for its own piece's DMA before it can copy it out)
    const size_t piece = (job.src_locked || job.op != HOP_RMSD || frame_bytes > kSlotCap)
                             ? 0 : stage_piece_bytes(CopyPool::get().threads());
    const bool streamed = piece > 0;
    // streamed staging keeps the cache footprint in its slots, so its chunks can be as large as the page-locked ones
    // (64 MB chunks measured 5-10 % faster than 16 MB); whole-chunk staging wants three lanes of chunks inside the L3
    const int chunk_mb = job.src_locked ? g_chunk_mb : streamed ? std::max(g_chunk_mb, g_staged_chunk_mb) : g_staged_chunk_mb;
    job.fpc = (int64_t)std::max<size_t>(1, ((size_t)chunk_mb << 20) / frame_bytes);
    job.fpc = std::min<int64_t>(job.fpc, job.n_frames);
    job.n_chunks = (job.n_frames + job.fpc - 1) / job.fpc;
    if (streamed) job.stage_piece_rows = (int64_t)std::max<size_t>(1, piece / frame_bytes);
    const int n_use = (int)std::min<int64_t>(n_devices, job.n_chunks);
    std::vector<std::thread> th;
    for (int i = 1; i < n_use; ++i) th.emplace_back(run_device, std::ref(job), devices[i]);
    run_device(job, devices[0]);
    for (auto& t : th) t.join();
    if (job.rc.load() != 0) return fail(job.rc.load(), "%s", job.err);
    return 0;
}

}  // namespace

extern "C" {

int b200rmsd_host_configure(int copy_threads, int chunk_mb, int staged_chunk_mb)
{
    if (copy_threads > 0) g_copy_threads = std::min(copy_threads, 64);  // takes effect before the pool's first use
    if (chunk_mb > 0) g_chunk_mb = std::min(chunk_mb, 1024);
    if (staged_chunk_mb > 0) g_staged_chunk_mb = std::min(staged_chunk_mb, 1024);
    return 0;
}

int b200rmsd_host_configure_staging(int piece_kb)
{
    g_stage_piece_kb = piece_kb < 0 ? -1 : std::min(piece_kb, (int)(kSlotCap >> 10));
    return 0;
}

int b200rmsd_rmsd_host_multi(const float* target, int64_t n_frames, int n_atoms_target, const float* ref_frame,
                             int n_atoms_ref, const int32_t* idx, const int32_t* ref_idx, int n_sel, int superpose,
                             int precentered, const float* traces, float ref_trace, float* out, const int* devices,
                             int n_devices)
{
    if (!target || !ref_frame || !out || n_frames < 0 || n_atoms_target <= 0 || n_atoms_ref <= 0)
        return fail(B200RMSD_EINVAL, "rmsd_host: bad arguments");
    if ((idx == nullptr) != (ref_idx == nullptr)) return fail(B200RMSD_EINVAL, "rmsd_host: idx and ref_idx must both be given or both NULL");
    if (!idx && n_atoms_target != n_atoms_ref) return fail(B200RMSD_EINVAL, "rmsd_host: atom counts differ and no index lists given");
    if (idx && n_sel <= 0) return fail(B200RMSD_EINVAL, "rmsd_host: empty selection");
    if (precentered && (!traces || idx)) return fail(B200RMSD_EINVAL, "rmsd_host: precentered needs traces and no index lists");
    HostJob job;
    job.op = HOP_RMSD;
    job.in = target;
    job.n_frames = n_frames;
    job.n_atoms = n_atoms_target;
    job.ref_frame = ref_frame;
    job.n_atoms_ref = n_atoms_ref;
    job.idx = idx;
    job.ref_idx = ref_idx;
    job.n_sel = n_sel;
    job.superpose = superpose;
    job.precentered = superpose ? precentered : 0;
    job.traces = traces;
    job.ref_trace = ref_trace;
    job.out_rmsd = out;
    return run_job(job, devices, n_devices, "rmsd_host");
}

int b200rmsd_rmsd_host(const float* target, int64_t n_frames, int n_atoms_target, const float* ref_frame,
                       int n_atoms_ref, const int32_t* idx, const int32_t* ref_idx, int n_sel, int superpose,
                       int precentered, const float* traces, float ref_trace, float* out, int device)
{
    return b200rmsd_rmsd_host_multi(target, n_frames, n_atoms_target, ref_frame, n_atoms_ref, idx, ref_idx, n_sel, superpose,
                                    precentered, traces, ref_trace, out, &device, 1);
}

int b200rmsd_superpose_host_multi(float* xyz, int64_t n_frames, int n_atoms, const float* ref_frame, int n_atoms_ref,
                                  const int32_t* idx, const int32_t* ref_idx, int n_sel, float* out_rot, float* out_rmsd,
                                  unsigned* n_degenerate, const int* devices, int n_devices)
{
    if (!xyz || !ref_frame || n_frames < 0 || n_atoms <= 0 || n_atoms_ref <= 0) return fail(B200RMSD_EINVAL, "superpose_host: bad arguments");
    if ((idx == nullptr) != (ref_idx == nullptr)) return fail(B200RMSD_EINVAL, "superpose_host: idx and ref_idx must both be given or both NULL");
    if (!idx && n_atoms != n_atoms_ref) return fail(B200RMSD_EINVAL, "superpose_host: atom counts differ and no index lists given");
    if (idx && n_sel <= 0) return fail(B200RMSD_EINVAL, "superpose_host: empty selection");
    if (n_degenerate) *n_degenerate = 0;
    HostJob job;
    job.op = HOP_SUPERPOSE;
    job.in = xyz;
    job.inout = xyz;
    job.n_frames = n_frames;
    job.n_atoms = n_atoms;
    job.ref_frame = ref_frame;
    job.n_atoms_ref = n_atoms_ref;
    job.idx = idx;
    job.ref_idx = ref_idx;
    job.n_sel = n_sel;
    job.out_rmsd = out_rmsd;
    job.out_rot = out_rot;
    const int rc = run_job(job, devices, n_devices, "superpose_host");
    if (n_degenerate) *n_degenerate = job.degenerate.load();
    return rc;
}

int b200rmsd_superpose_host(float* xyz, int64_t n_frames, int n_atoms, const float* ref_frame, int n_atoms_ref,
                            const int32_t* idx, const int32_t* ref_idx, int n_sel, float* out_rot, float* out_rmsd,
                            unsigned* n_degenerate, int device)
{
    return b200rmsd_superpose_host_multi(xyz, n_frames, n_atoms, ref_frame, n_atoms_ref, idx, ref_idx, n_sel, out_rot,
                                         out_rmsd, n_degenerate, &device, 1);
}

int b200rmsd_center_host_multi(float* xyz, int64_t n_frames, int n_atoms, float* traces, const int* devices, int n_devices)
{
    if (!xyz || n_frames < 0 || n_atoms <= 0) return fail(B200RMSD_EINVAL, "center_host: bad arguments");
    HostJob job;
    job.op = HOP_CENTER;
    job.in = xyz;
    job.inout = xyz;
    job.n_frames = n_frames;
    job.n_atoms = n_atoms;
    job.out_traces = traces;
    return run_job(job, devices, n_devices, "center_host");
}

int b200rmsd_center_host(float* xyz, int64_t n_frames, int n_atoms, float* traces, int device)
{
    return b200rmsd_center_host_multi(xyz, n_frames, n_atoms, traces, &device, 1);
}

void b200rmsd_release_workspaces(void)
{
    for (int d = 0; d < 64; ++d) {
        Workspace& w = g_ws[d];
        std::lock_guard<std::mutex> lk(w.mu);
        if (!w.init) continue;
        int prev = -1;
        cudaGetDevice(&prev);
        if (cudaSetDevice(d) == cudaSuccess) {
            for (int l = 0; l < kLanes; ++l) {
                Lane& L = w.lane[l];
                cudaStreamSynchronize(L.stream);
                cudaFree(L.xyz); cudaFree(L.out); cudaFree(L.rot); cudaFree(L.trc); cudaFree(L.scratch);
                cudaFreeHost(L.up); cudaFreeHost(L.down); cudaFreeHost(L.res);
                cudaStreamDestroy(L.stream);
                cudaEventDestroy(L.done);
                L = Lane{};
            }
            cudaFree(w.ref_raw); cudaFree(w.ref_sel); cudaFree(w.idx); cudaFree(w.ref_idx); cudaFree(w.stats); cudaFree(w.degen);
            cudaEventDestroy(w.ref_ready);
        }
        if (prev >= 0) cudaSetDevice(prev);
        w.init = false;
        w.ref_ready = nullptr;
        w.xyz_bytes = w.up_bytes = w.down_bytes = w.per_frame_cap = w.scratch_bytes = w.ref_cap = w.idx_cap = 0;
        w.ref_raw = w.ref_sel = nullptr; w.idx = w.ref_idx = nullptr; w.stats = nullptr; w.degen = nullptr;
    }
}

}  // extern "C"

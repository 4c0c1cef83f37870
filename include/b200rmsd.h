/*
 * b200rmsd.h -- C ABI of the B200-native replacement for mdtraj's `_rmsd` hot path.
 *
 * Everything the reference's Cython boundary (mdtraj/rmsd/_rmsd.pyx) obtains from
 * its three C headers
 *     mdtraj/rmsd/include/center.h:7            inplace_center_and_trace_atom_major
 *     mdtraj/rmsd/include/theobald_rmsd.h:13-18 msd_axis_major / msd_atom_major
 *     mdtraj/rmsd/include/rotation.h:7-12       rot_atom_major / rot_msd_atom_major
 * is provided here, batched over frames (a per-frame call makes no sense across PCIe):
 * each entry point replaces one *frame loop* of _rmsd.pyx together with the per-frame
 * C calls inside it.  The entry cites the loop it replaces.
 *
 * Conventions
 *   - plain C: pointers, sizes, ints.  No torch / numpy types.  Every function returns
 *     0 on success or a negative B200RMSD_E* code; b200rmsd_last_error() gives a
 *     thread-local message.  Nothing throws across the boundary.
 *   - "_dev" entry points take DEVICE pointers, run asynchronously on `stream`
 *     (a cudaStream_t passed as void*; NULL = default stream) of the calling thread's
 *     current device, and never allocate.  Scratch is caller-provided.
 *   - "_host" entry points take HOST pointers, stage through an internal per-device
 *     workspace, pipeline H2D copies against kernels, and return when the results are in
 *     the caller's host buffers.  Pageable memory (any numpy array) is staged through
 *     page-locked buffers by a pool of memcpy threads; page-locked memory is DMA'd directly.
 *     After a failed in-place call (superpose_host, center_host) the contents of xyz are
 *     undefined; no transfer is in flight any more when an error is returned.
 *   - staged trajectory layout ("padded atom-major"): float32, frame f starts at
 *     xyz + f*frame_stride, holds n_atoms x 3 floats followed by zeros up to
 *     n_pad = 4*ceil(n_atoms/4) atoms; frame_stride >= 3*n_pad and frame_stride % 4 == 0;
 *     xyz is 16-byte aligned.  For n_atoms % 4 == 0 this is exactly mdtraj's
 *     (F, N, 3) C-contiguous xyz.  This is the reference's own nrealatoms/npaddedatoms
 *     convention (theobald_rmsd_sse.h:192-222).
 *   - there is no CPU fallback: without a CUDA device every call fails with
 *     B200RMSD_ENODEVICE.
 */
#ifndef B200RMSD_H_
#define B200RMSD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200RMSD_ABI_VERSION 2

#define B200RMSD_OK 0
#define B200RMSD_EINVAL (-1)    /* bad argument (NULL, misaligned, non-positive size)   */
#define B200RMSD_ECUDA (-2)     /* a CUDA runtime call failed; see b200rmsd_last_error() */
#define B200RMSD_ENODEVICE (-3) /* no usable sm_100 device                              */
#define B200RMSD_ENOMEM (-4)    /* workspace allocation failed                          */

/* flags for b200rmsd_rmsd_dev / b200rmsd_rmsd_host */
#define B200RMSD_PRECENTERED 1u /* trust `traces`, skip centring (_rmsd.pyx:203-205)    */

/* Opaque per-reference-frame record produced by b200rmsd_prepare_reference_dev
 * (device memory, B200RMSD_REFSTATS_BYTES bytes, 8-byte aligned). */
#define B200RMSD_REFSTATS_BYTES 56

int b200rmsd_abi_version(void);
const char* b200rmsd_last_error(void);

/* Number of CUDA devices; sm_count/hbm_bytes of `device` (any pointer may be NULL). */
int b200rmsd_device_info(int device, int* n_devices, int* sm_count, size_t* hbm_bytes, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------- device API */

/* Bytes of scratch the _dev entry points below may need for a trajectory of this shape
 * (segment partial sums for very long frames; rotations + centroids for superpose). */
size_t b200rmsd_scratch_bytes(int64_t n_frames, int n_atoms);

/* Centre frames in place and write their traces (traces may be NULL).
 * Replaces: inplace_center_and_trace_atom_major, center.h:7 (called at
 * _rmsd.pyx:212 and :490). */
int b200rmsd_center_trace_dev(float* xyz, int64_t n_frames, int n_atoms, int64_t frame_stride, float* traces,
                              void* stream);

/* Prepare the single reference conformation: gather `idx` (int32, may be NULL = all
 * n_sel atoms), centre it (do_center != 0; otherwise keep as is and take its trace from
 * `given_trace`, the precentered path), write it zero-padded to ref_out
 * (3*4*ceil(n_sel/4) floats) and its statistics to ref_stats.
 * Replaces: the n_frames == 1 call at _rmsd.pyx:213 and the ref_align_xyz preparation
 * at core/trajectory.py:1129-1147,1150. */
int b200rmsd_prepare_reference_dev(const float* ref_frame, const int32_t* idx, int n_sel, int do_center,
                                   float given_trace, float* ref_out, void* ref_stats, void* stream);

/* RMSD of every frame to the prepared reference after optimal superposition.
 *   idx == NULL : all n_atoms atoms of each frame take part (TMA streaming kernel)
 *   idx != NULL : the n_sel listed atoms (int32, any order, repeats allowed) take part
 *   flags & B200RMSD_PRECENTERED : frames are already centred and `traces` holds their
 *                                  traces (only honoured when idx == NULL)
 *   out_rot (F,9) / out_centroid (F,3 doubles) may be NULL; when given they receive the
 *   optimal row-major rotation (apply as x' = x . R) and the removed centroid.
 *   n_degenerate (device uint32, may be NULL) counts identity-fallback rotations.
 * Replaces: the prange loop _rmsd.pyx:217-224 (msd_atom_major + sqrtf per frame) together
 * with the centring pass :212 and the fancy-index copy :197; with out_rot also the
 * msd_atom_major(..., computeRot=1) half of _rmsd.pyx:663-674. */
int b200rmsd_rmsd_dev(const float* xyz, int64_t n_frames, int n_atoms, int64_t frame_stride, const int32_t* idx,
                      int n_sel, const float* ref, const void* ref_stats, const float* traces, unsigned flags,
                      float* out_rmsd, float* out_rot, double* out_centroid, unsigned* n_degenerate, void* scratch,
                      size_t scratch_bytes, void* stream);

/* RMSD without superposition against the raw (uncentred) reference selection `ref_raw`
 * (n_sel x 3 floats).  Replaces: the prange loop _rmsd.pyx:234-241 + msd_nosuperpose
 * (:765-793). */
int b200rmsd_rmsd_nosuperpose_dev(const float* xyz, int64_t n_frames, int n_atoms, int64_t frame_stride,
                                  const int32_t* idx, int n_sel, const float* ref_raw, float* out_rmsd, void* stream);

/* Superpose every frame onto the prepared reference, in place:
 *     xyz'[f] = (xyz[f] - centroid_sel(f)) . R_f + centroid_ref
 * with R_f the optimal rotation of the selected atoms.  out_rmsd / out_rot may be NULL.
 * Replaces: superpose_atom_major, _rmsd.pyx:620-674 (msd_atom_major computeRot=1 +
 * rot_atom_major per frame), and the numpy passes of Trajectory.superpose around it
 * (core/trajectory.py:1127, 1135-1150, 1171). */
int b200rmsd_superpose_dev(float* xyz, int64_t n_frames, int n_atoms, int64_t frame_stride, const int32_t* idx,
                           int n_sel, const float* ref, const void* ref_stats, float* out_rmsd, float* out_rot,
                           unsigned* n_degenerate, void* scratch, size_t scratch_bytes, void* stream);

/* Apply caller-supplied rotations:  xyz[f] <- xyz[f] . rot[f]  (no translation).
 * Replaces: rot_atom_major, rotation.h:7 (loop body _rmsd.pyx:668). */
int b200rmsd_rotate_dev(float* xyz, int64_t n_frames, int n_atoms, int64_t frame_stride, const float* rot,
                        void* scratch, size_t scratch_bytes, void* stream);

/* ------------------------------------------------- "next" rows: rmsf, align/displace, lprmsd */

/* Per-atom root-mean-square fluctuation over frames of  y = (x - c_f) . R_f  for the listed atoms
 * (idx int32 or NULL = all n_atoms).  rot (F,9) and centroid (F,3 doubles) come from b200rmsd_rmsd_dev
 * (out_rot / out_centroid); either may be NULL (no rotation: reference=None "prealigned" mode,
 * _rmsd.pyx:328-330,410; no centring: the index-list path, where the reference rotates the uncentred copy,
 * _rmsd.pyx:408).  scratch: b200rmsd_rmsf_scratch_bytes().  out_rmsf: n_sel floats.
 * Replaces: the rotate + two per-atom reduction loops of rmsf(), _rmsd.pyx:410-444. */
size_t b200rmsd_rmsf_scratch_bytes(int64_t n_frames, int n_sel);
int b200rmsd_rmsf_dev(const float* xyz, int64_t n_frames, int n_atoms, int64_t frame_stride, const int32_t* idx,
                      int n_sel, const float* rot, const double* centroid, void* scratch, size_t scratch_bytes,
                      float* out_rmsf, void* stream);

/* out[i] = sqrt(mean_k |b_i[k] - a[k] . R_i|^2) with R_i = rot[i] (transpose != 0: its transpose); when rot_out is
 * not NULL the rotation actually used is also written there.
 * Replaces: rot_msd_atom_major, rotation.h:10 (loop body of getMultipleAlignDisplaceRMSDs_atom_major,
 * _rmsd.pyx:746-749). */
int b200rmsd_rot_msd_dev(const float* a_frame, const float* b_xyz, int64_t n_frames, int n_atoms, int64_t frame_stride,
                         const float* rot, int transpose, float* rot_out, float* out_rmsd, void* stream);

/* LP-RMSD of every frame to one reference conformation: the RMSD minimised over rigid motions AND over the labels of
 * exchangeable atoms.  Replaces the frame loop of md.lprmsd, _lprmsd.pyx:186-224, and the C++ under it
 * (fancy_index2d, inplace_center_and_trace_atom_major, msd_atom_major, rot_atom_major, euclidean_permutation +
 * Munkres::solve, sgemm33).  Per frame: rotation from the distinguishable atoms -> minimum-cost matching of squared
 * distances inside every permute group -> QCP RMSD of the relabelled selection.
 *   xyz          padded atom-major frames (as above); idx (n_sel sorted unique atom indices) or NULL = all n_atoms;
 *   ref_sel      (n_sel,3) floats: the reference conformation's selected atoms, as they are (not centred);
 *   dis, n_dis   positions INSIDE the selection (0 .. n_sel-1) of the atoms in no permute group; may be empty;
 *   group_atoms  positions inside the selection of the permutable atoms, group after group; group_off: n_groups + 1
 *                offsets into it; max_group: size of the largest group (sizes the shared-memory solver state);
 *   out_rmsd     (F); out_rot (F,9) optional: rot1 . rot2, the rotation md.lprmsd(superpose=True) applies to the frame
 *                after centring ALL its atoms (_lprmsd.pyx:217-220; apply with b200rmsd_center_trace_dev +
 *                b200rmsd_rotate_dev); out_map (F,n_sel) optional: the matching, out_map[f*n_sel + i] = position of the
 *                target atom paired with reference atom i.
 * A frame is solved by one warp in shared memory: B200RMSD_EINVAL when the selection does not fit (about 3,500 atoms
 * with one group of all of them). */
int b200rmsd_lprmsd_dev(const float* xyz, int64_t n_frames, int n_atoms, int64_t frame_stride, const int32_t* idx,
                        int n_sel, const float* ref_sel, const int32_t* dis, int n_dis, const int32_t* group_atoms,
                        const int32_t* group_off, int n_groups, int max_group, float* out_rmsd, float* out_rot,
                        int32_t* out_map, void* stream);

/* ------------------------------------------------------------ all-pairs matrix */

/* The reference has no all-pairs function; clustering code loops md.rmsd(traj, traj, i)
 * over i (examples/clustering.ipynb:78-81, examples/centroids.ipynb:80-82).  These two
 * entry points replace that loop:  D[i][j] = md.rmsd(traj, traj, i, atom_indices=idx)[j].
 *
 * workspace: caller-owned device memory, 256-byte aligned, at least
 * b200rmsd_allpairs_workspace_bytes(n_frames, n_sel) bytes; opaque contents (centred
 * K-major operands, traces).  With multiple GPUs every rank prepares (or receives by
 * broadcast) the same workspace and computes its own row block. */
#define B200RMSD_DIAG_ZERO 1u  /* D[i][i] = 0 exactly, like the reference's same-pointer shortcut */
#define B200RMSD_FAST_SOLVE 2u /* all-float32 QCP solve in the tensor-core epilogue: the precision class of the
                                * reference's own msdFromMandG (float32 polynomial + float32 G_a+G_b-2*lambda,
                                * theobald_rmsd.cpp:249-277) instead of the default float64 polynomial */

size_t b200rmsd_allpairs_workspace_bytes(int64_t n_frames, int n_sel);

/* Centre every frame (selection idx, int32, may be NULL = all atoms; then pass n_sel = n_atoms)
 * as inplace_center_and_trace_atom_major does (center.h:7) and lay it out for the contraction.
 * With n_frames >= 512 (tensor-core path) a few reference structures are chosen among the frames (greedy
 * farthest-point traversal, one one-vs-many pass each) and every frame is rotated onto the nearest one and stored
 * as its difference from it -- the RMSD of a pair does not depend on how either frame is placed -- which keeps the
 * tensor-core accumulators small on multi-basin trajectories too ("tensor-core accumulation").
 * xyz itself is only read.  Unlike the other _dev entry points this one SYNCHRONISES `stream` (once per reference
 * chosen: the traversal is data dependent). */
int b200rmsd_allpairs_prepare_dev(const float* xyz, int64_t n_frames, int n_atoms, int64_t frame_stride,
                                  const int32_t* idx, int n_sel, void* workspace, size_t workspace_bytes, void* stream);

/* Process-wide settings of the all-pairs path; a value <= 0 leaves the setting as it is.
 *   min_tc_frames: trajectories with at least this many frames take the tensor-core kernel, shorter ones the exact
 *                  fp32 SIMT kernel (default 512).  Must not change between prepare and the rows/block calls that
 *                  use its workspace.
 *   max_refs:      upper bound on the reference structures prepare may choose (default and maximum 32). */
int b200rmsd_allpairs_configure(int min_tc_frames, int max_refs);

/* Tensor-core kernel geometry: cta_pair != 0 runs the GEMM on 2-CTA clusters -- the two SMs of a TPC share one
 * M = 256 tcgen05.mma (cta_group::2), each loading its own rows of A and half of the rows of B, which cuts the
 * L2 -> shared-memory operand traffic by a quarter (74.7 -> 54.9 GB per 20k x 20k matrix) and is the default; 0
 * selects the single-CTA kernel (M = 128).  Both produce bit-identical matrices.  Measured on one B200, 20k x 20k x 300
 * atoms: 5.63 vs 5.72 ms on iid frames, 5.31 vs 6.27 ms on MD-like frames (profiles/r02_ap_isolate_uni.jsonl; DESIGN.md
 * section 8.4).  Process-wide; returns the previous setting. */
int b200rmsd_allpairs_set_cta_pair(int cta_pair);

/* What prepare found (synchronises `stream`): number of reference structures, number of frames stored as they are
 * because no reference is near them, covering radius (largest RMSD of a frame to its nearest reference, nm) and the
 * frame indices of the references (up to ref_frames_cap).  Any pointer may be NULL. */
int b200rmsd_allpairs_info_dev(const void* workspace, size_t workspace_bytes, int* n_refs, int* n_far,
                               float* cover_radius, int* ref_frames, int ref_frames_cap, void* stream);

/* Rows [row0, row1) of the matrix:  out[(i-row0)*ld + j], j in [0, n_frames), float32.
 * n_sel must equal the value used in prepare. */
int b200rmsd_allpairs_rows_dev(const void* workspace, size_t workspace_bytes, int64_t n_frames, int n_sel,
                               int64_t row0, int64_t row1, float* out, int64_t ld, unsigned flags, void* stream);

/* One rectangular block of the matrix: rows [row0,row1) x columns [col0,col1), written at
 * out[(i-row0)*ld + j] with the absolute column index j: `out` is the address of element (row0, column 0) of a matrix
 * with leading dimension ld.  Only columns [col0,col1) are touched, so a caller that keeps just that window passes
 * window - col0 and ld >= col1 - col0.  When out_t != NULL the transposed block is also
 * written, out_t[(j-col0)*ld_t + (i-row0)]: D is symmetric, so a rank that computes block (r,s) can ship out_t to
 * the owner of row block s instead of that rank recomputing it (mdtraj_b200.distributed.rmsd_matrix_sharded).
 * A square block on the diagonal (row range == column range, out_t == NULL) computes each unordered pair once and
 * mirrors it.  b200rmsd_allpairs_rows_dev(row0,row1) == block(row0,row1,0,n_frames). */
int b200rmsd_allpairs_block_dev(const void* workspace, size_t workspace_bytes, int64_t n_frames, int n_sel,
                                int64_t row0, int64_t row1, int64_t col0, int64_t col1, float* out, int64_t ld,
                                float* out_t, int64_t ld_t, unsigned flags, void* stream);

/* The same block with the superposition of every pair: besides out (as above) the epilogue writes the rotation that
 * carries frame j onto frame i,
 *     out_rot[((i-row0)*(col1-col0) + (j-col0))*9 + 3a + b] = U[a][b],   (x_j - centroid_j) U ~ x_i - centroid_i,
 * row vector times U as in rot_atom_major (rotation_generic.h:40-42) -- the rotation md.rmsd / Trajectory.superpose
 * would find for target frame j against reference frame i (theobald_rmsd.cpp:280-334), i.e. row i of out_rot equals
 * out_rot of b200rmsd_rmsd_dev(frames, reference = frame i).  Entries (i, i) are the identity when
 * B200RMSD_DIAG_ZERO is set.  36 bytes per pair leave the device, so this is meant for blocks (cluster centres x
 * members), not for a 100k x 100k matrix; every pair is computed (no mirroring).  The inner products still never
 * touch HBM: the rotation comes out of the same accumulator tile as the RMSD. */
int b200rmsd_allpairs_block_rot_dev(const void* workspace, size_t workspace_bytes, int64_t n_frames, int n_sel,
                                    int64_t row0, int64_t row1, int64_t col0, int64_t col1, float* out, int64_t ld,
                                    float* out_rot, unsigned flags, void* stream);

/* ------------------------------------------------------------ peer-visible buffers
 *
 * Multi-process all-pairs (one process per GPU): rank r computes block (r,s) of the symmetric matrix once and its
 * epilogue writes the transposed copy straight into rank s's row block over NVLink -- out_t of
 * b200rmsd_allpairs_block_dev is then an address inside a buffer rank s allocated here and rank r opened.  The transfer
 * rides under the tensor-core work of the same kernel; there is no staging buffer, no separate send/receive kernel
 * and no copy on arrival (mdtraj_b200.distributed.rmsd_matrix_sharded(exchange="peer")).
 *   peer_alloc: cudaMalloc + cudaIpcGetMemHandle; `handle` receives B200RMSD_PEER_HANDLE_BYTES bytes to pass to the
 *               other processes by any means (torch.distributed.all_gather_object in the Python layer);
 *   peer_open:  maps another process's buffer (enables peer access on first use); peer_close unmaps it;
 *   peer_free:  releases a buffer of peer_alloc (after every process that opened it has closed it).
 * The caller orders the accesses: stream-synchronise every writer and meet at a barrier before reading. */
#define B200RMSD_PEER_HANDLE_BYTES 64
int b200rmsd_peer_alloc(size_t bytes, void** dev_ptr, void* handle);
int b200rmsd_peer_open(const void* handle, void** dev_ptr);
int b200rmsd_peer_close(void* dev_ptr);
int b200rmsd_peer_free(void* dev_ptr);

/* ------------------------------------------------------------ consumers of the matrix */

/* The reference's example notebooks post-process the (F,F) matrix on the host with numpy / scipy.  These entry points
 * run the same reductions over a device-resident float32 matrix D (rows x cols, leading dimension ld), one read of the
 * matrix each, so that a 40 GB matrix never has to cross PCIe.
 *
 * matrix_moments: out2[0] = sum d_ij, out2[1] = sum d_ij^2 (float64; zeroed by the call) -- what distances.std() needs
 *   (examples/centroids.ipynb:117).  Call once per row block and add, or once on the whole matrix. */
int b200rmsd_matrix_moments_dev(const float* D, int64_t rows, int64_t cols, int64_t ld, double* out2, void* stream);

/* exp_rowsum: rowsum[i] = (accumulate ? rowsum[i] : 0) + sum_j exp(scale * d_ij), float64 --
 *   np.exp(-beta * distances / distances.std()).sum(axis=1) with scale = -beta / std (examples/centroids.ipynb:117). */
int b200rmsd_exp_rowsum_dev(const float* D, int64_t rows, int64_t cols, int64_t ld, float scale, int accumulate,
                            double* rowsum, void* stream);

/* row_argmin: arg[i] = index of the first minimum of row i, val[i] (may be NULL) = that minimum -- the nearest-leader
 *   assignment np.argmin(md.rmsd(leaders, frame, 0)) of examples/two-pass-clustering.ipynb (cell 13) for every frame. */
int b200rmsd_row_argmin_dev(const float* D, int64_t rows, int64_t cols, int64_t ld, int32_t* arg, float* val, void* stream);

/* condense: out[n*i - i*(i+1)/2 + (j-i-1)] = D[i*ld + j] for i < j < n: the condensed form scipy's linkage functions
 *   take, scipy.spatial.distance.squareform(distances, checks=False) in examples/clustering.ipynb:101.
 *   out: n*(n-1)/2 floats. */
int b200rmsd_condense_dev(const float* D, int64_t n, int64_t ld, float* out, void* stream);

/* ------------------------------------------------------------------ host API */

/* md.rmsd on host arrays.  target: (n_frames, n_atoms_target, 3) float32 C-contiguous
 * (mdtraj's xyz, unpadded); ref_frame: (n_atoms_ref, 3).  idx/ref_idx: int32 selections
 * of length n_sel or both NULL (then n_atoms_target == n_atoms_ref == n_sel).
 * superpose == 0 selects the no-alignment branch.  precentered != 0 requires `traces`
 * (n_frames) and ref_trace and idx == NULL.  Frames are streamed through the device in
 * chunks; host buffers are never modified.  out: (n_frames) float32.
 * Replaces: the body of rmsd(), _rmsd.pyx:197-241. */
int b200rmsd_rmsd_host(const float* target, int64_t n_frames, int n_atoms_target, const float* ref_frame,
                       int n_atoms_ref, const int32_t* idx, const int32_t* ref_idx, int n_sel, int superpose,
                       int precentered, const float* traces, float ref_trace, float* out, int device);

/* Trajectory.superpose on host arrays, in place on `xyz` (n_frames, n_atoms, 3).
 * out_rot (n_frames*9) / out_rmsd (n_frames) may be NULL.
 * Replaces: core/trajectory.py:1127-1172. */
int b200rmsd_superpose_host(float* xyz, int64_t n_frames, int n_atoms, const float* ref_frame, int n_atoms_ref,
                            const int32_t* idx, const int32_t* ref_idx, int n_sel, float* out_rot, float* out_rmsd,
                            unsigned* n_degenerate, int device);

/* _center_inplace_atom_major on a host array (n_frames, n_atoms, 3), in place; traces
 * (n_frames) may be NULL.  Replaces: _rmsd.pyx:487-491. */
int b200rmsd_center_host(float* xyz, int64_t n_frames, int n_atoms, float* traces, int device);

/* The same three calls over SEVERAL devices from one process: the frames are cut into chunks that are handed out
 * dynamically to the listed devices (one host thread each; frames are independent, so there is no exchange between
 * devices -- BASELINE north_star: "a trivial frame split with a host gather").  `devices`: n_devices distinct CUDA
 * device indices.  The single-device entry points above are the n_devices = 1 case.  Results are identical to the
 * single-device call bit for bit (each frame is computed by the same kernel whichever device takes its chunk). */
int b200rmsd_rmsd_host_multi(const float* target, int64_t n_frames, int n_atoms_target, const float* ref_frame,
                             int n_atoms_ref, const int32_t* idx, const int32_t* ref_idx, int n_sel, int superpose,
                             int precentered, const float* traces, float ref_trace, float* out, const int* devices,
                             int n_devices);
int b200rmsd_superpose_host_multi(float* xyz, int64_t n_frames, int n_atoms, const float* ref_frame, int n_atoms_ref,
                                  const int32_t* idx, const int32_t* ref_idx, int n_sel, float* out_rot, float* out_rmsd,
                                  unsigned* n_degenerate, const int* devices, int n_devices);
int b200rmsd_center_host_multi(float* xyz, int64_t n_frames, int n_atoms, float* traces, const int* devices,
                               int n_devices);

/* Process-wide settings of the host pipeline; a value <= 0 leaves the setting as it is.
 *   copy_threads:    threads of the memcpy pool that stages pageable host memory through page-locked buffers (default:
 *                    all hardware threads but one, at most 24; takes effect before the pool's first use);
 *   chunk_mb:        MB of padded coordinates per chunk when the caller's memory is page-locked (default 64);
 *   staged_chunk_mb: the same for pageable memory when whole chunks are staged (default 16: three staging lanes stay inside
 *                    a 60 MB host L3); streamed staging (below) uses max(chunk_mb, staged_chunk_mb). */
int b200rmsd_host_configure(int copy_threads, int chunk_mb, int staged_chunk_mb);

/* How pageable coordinates reach the device in b200rmsd_rmsd_host*.  Whole-chunk staging: the memcpy pool fills a lane's
 * page-locked buffer, one copy sends it (three passes over host memory per byte; three 16 MB lanes stay inside the L3 of
 * ONE process).  Streamed staging: every pool thread copies a piece of the chunk (whole frames) into one of its two small
 * page-locked slots and sends it to its place in the device chunk at once, so that the staging writes and DMA reads of ALL
 * ranks on the host stay in its last-level cache (eight ranks on one 32-thread host: 5.1e6 -> 9.4e6 rmsd/s at 1,000 atoms,
 * four ranks 5.4e6 -> 7.6e6, two 5.9e6 -> 6.7e6; one rank is not faster; profiles/r02_host_staging.jsonl).
 *   piece_kb == 0 (default): streamed when $LOCAL_WORLD_SIZE (set by torchrun) says several ranks share the host, with
 *                 pieces of 512 KB - 1 MB; whole chunks otherwise;
 *   piece_kb > 0: streamed, piece_kb KB per piece (at most 2048);   piece_kb < 0: whole chunks always.
 * In-place operations (superpose, centring) and frames larger than 2 MB always stage whole chunks. */
int b200rmsd_host_configure_staging(int piece_kb);

/* Release the internal per-device workspaces of the host API. */
void b200rmsd_release_workspaces(void);

#ifdef __cplusplus
}
#endif
#endif /* B200RMSD_H_ */

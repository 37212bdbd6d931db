/*
 * psb200.h -- C ABI of libpsb200.so: B200 (sm_100a) kernels for PoreSpy's hot path
 *             edt.edt -> porosimetry / local_thickness (+ trim_disconnected_blobs).
 *
 * This is the drop-in boundary.  PoreSpy is pure Python, so the "FFI" a maintainer
 * would bind is ctypes (shown in INTEGRATION.md); every entry point below replaces one
 * piece of the reference path and cites it:
 *
 *   F  = /root/reference/src/porespy/filters/_funcs.py
 *   T  = /root/reference/src/porespy/tools/_funcs.py
 *   B  = /root/reference/src/porespy/generators/_borders.py
 *   edt = third-party `edt.edt` (seung-lab/euclidean-distance-transform-3d, unpinned in
 *         /root/reference/pyproject.toml:29), call sites F:1126, F:1191, T:1153.
 *
 * Conventions
 *   - Volumes are C-contiguous [nz][ny][nx] (x fastest); 2-D images pass nz = 1,
 *     1-D lines pass nz = ny = 1.  Each dimension must be <= PSB200_MAX_DIM.
 *   - All data pointers are DEVICE pointers unless the comment says "host".
 *   - The library never allocates user-visible memory: scratch space comes from a
 *     caller-provided workspace whose size the matching *_workspace_bytes() reports.
 *   - Calls are asynchronous on the caller's stream (a cudaStream_t passed as void*),
 *     except where "synchronises" is stated.  No global mutable state beyond the ctx.
 *   - Every function returns a status (0 = OK, negative = error); the message of the
 *     last error on the calling thread is returned by psb200_last_error().
 *   - No torch types, no C++ exceptions across this boundary.
 */
#ifndef PSB200_H
#define PSB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSB200_VERSION 100            /* 0.1.0 */
#define PSB200_MAX_DIM 32767          /* per-axis limit: 3*(MAX_DIM-1)^2 < 2^32 */
#define PSB200_MAX_THRESHOLDS 253     /* radii (distinct thresholds) per *_idx call */
#define PSB200_INF_U32 0xFFFFFFFFu    /* squared distance when no background exists */
#define PSB200_IDX_KEEP 255           /* idx value: "already written by an earlier call" */

#define PSB200_OK 0
#define PSB200_ERR_INVALID (-1)       /* bad argument */
#define PSB200_ERR_CUDA (-2)          /* CUDA runtime error (see psb200_last_error) */
#define PSB200_ERR_WORKSPACE (-3)     /* workspace too small / missing */
#define PSB200_ERR_UNSUPPORTED (-4)   /* outside the implemented envelope */

/* inlet_mode of the access-limited paths (F:1128-1129, F:1181-1183) */
#define PSB200_INLETS_NONE 0          /* access_limited=False (local_thickness) */
#define PSB200_INLETS_FACES 1         /* default: get_border(shape,'faces') (B:93-100), never materialised */
#define PSB200_INLETS_MASK 2          /* user mask, uint8 [nz][ny][nx], non-zero = inlet */

/* algorithm selector (ctx option "algo") */
#define PSB200_ALGO_FAST 0            /* bounded u8 pipeline (default) */
#define PSB200_ALGO_GENERIC 1         /* three full u32 EDT passes per radius (slow, unbounded) */

/* flags */
#define PSB200_FLAG_IDX_PREINIT 1     /* local_thickness_idx: idx already holds 0 / PSB200_IDX_KEEP */
#define PSB200_FLAG_EXPAND_MERGE 1    /* expand_idx_f64: leave out[] untouched where idx is 0 or KEEP */

typedef struct psb200_ctx psb200_ctx;
typedef void *psb200_stream;          /* cudaStream_t */

int psb200_version(void);
const char *psb200_last_error(void);

/* One context per device (per rank).  Selects the device for subsequent calls. */
int psb200_create(int device, psb200_ctx **ctx);
int psb200_destroy(psb200_ctx *ctx);
/* name: "algo" (PSB200_ALGO_*), "profile" (0/1: record a cudaEvent pair around every kernel
 * launch), "bit_tmax" (thresholds T <= bit_tmax take the bit-parallel dilation kernels, 0 turns
 * them off; default 200).  Returns PSB200_ERR_INVALID for unknown names. */
int psb200_set_option(psb200_ctx *ctx, const char *name, int64_t value);
/* Number of kernel launches issued through this ctx since creation (bench bookkeeping). */
int64_t psb200_launch_count(const psb200_ctx *ctx);
/* Per-kernel-family device time measured with CUDA events on the launching stream while the
 * "profile" option is on.  psb200_profile_read synchronises, fills two arrays of
 * psb200_profile_kernels() entries (milliseconds, launch counts) and clears the records. */
int psb200_profile_kernels(void);
const char *psb200_profile_name(int kernel_id);
int psb200_profile_read(psb200_ctx *ctx, double *ms_total, int64_t *launches);
/* Per-launch records in launch order; returns how many exist (at most `max` are written). */
int psb200_profile_records(psb200_ctx *ctx, int *kernel_ids, float *ms, int max);

/* ---------------------------------------------------------------- exact squared EDT
 * Replaces edt.edt(data) at F:1126 / F:1191 / T:1153 (black_border=False, isotropic,
 * binary).  d2_out[v] = exact integer squared distance from voxel v to the nearest
 * zero voxel of `in`; 0 where in[v]==0; PSB200_INF_U32 everywhere if `in` has no zero.
 * `in` is read-only.  d2_out may not alias in. */
size_t psb200_edt_workspace_bytes(const psb200_ctx *ctx, int64_t nz, int64_t ny, int64_t nx);
int psb200_edt_sq_u8(psb200_ctx *ctx, const uint8_t *in, uint32_t *d2_out,
                     int64_t nz, int64_t ny, int64_t nx,
                     void *ws, size_t ws_bytes, psb200_stream stream);

/* Same transform with the epilogues fused into the last pass: out_kind 0 -> uint32 squared
 * distances (as above), 1 -> float32 distances (exactly what edt.edt returns: IEEE sqrt of the
 * exact integer, +inf when the image has no zero).  max_out (device uint32, may be NULL)
 * receives max(d2) -- np.amax(dt) at F:1132 without another pass over the volume. */
int psb200_edt_u8(psb200_ctx *ctx, const uint8_t *in, void *out, int out_kind,
                  uint32_t *max_out, int64_t nz, int64_t ny, int64_t nx,
                  void *ws, size_t ws_bytes, psb200_stream stream);
/* The same, with *max_out taken over the planes [zmax0, zmax1) only: a z-slab shard runs the transform on its slab
 * plus input halo planes, but the maximum that defines the radii (F:1132) is the one of its own planes. */
int psb200_edt_u8_zmax(psb200_ctx *ctx, const uint8_t *in, void *out, int out_kind, uint32_t *max_out, int64_t nz,
                       int64_t ny, int64_t nx, int64_t zmax0, int64_t zmax1, void *ws, size_t ws_bytes,
                       psb200_stream stream);

/* Single passes of the separable transform, exposed for z-slab sharded volumes
 * (SURVEY 8(e)): x and y passes run on the local slab, the z pass on the pencil
 * layout [nz][ny/P][nx] after the all-to-all.  axis: 2 = x (reads `in`, writes d2),
 * 1 = y, 0 = z (in place on d2; `in` ignored). */
int psb200_edt_pass(psb200_ctx *ctx, int axis, const uint8_t *in, uint32_t *d2,
                    int64_t nz, int64_t ny, int64_t nx,
                    void *ws, size_t ws_bytes, psb200_stream stream);

/* The same transform for a volume sharded into z-slabs over several GPUs.
 * psb200_edt_xy_u8: x and y passes of the local slab [nz][ny][nx] -> 2-D squared distances.
 *   ysplit = 0: h_out is [nz][ny][nx].  ysplit > 0: h_out is written in the send layout of the
 *   slab->pencil all-to-all, [dest d][nz][rows of d][nx] with dest d owning rows
 *   [d*ysplit, min(ny,(d+1)*ysplit)) -- the pack is fused into the store.
 *   Values >= 0xC000FFFE in h_out mean "infinite" (no zero in that plane); only
 *   psb200_edt_z_u32 consumes them.
 * psb200_edt_z_u32: z pass on a pencil [nz][ny_local][nx] (all planes, a range of rows);
 *   out_kind / max_out as in psb200_edt_u8 (max_out is required here).  out may not alias h. */
int psb200_edt_xy_u8(psb200_ctx *ctx, const uint8_t *in, uint32_t *h_out,
                     int64_t nz, int64_t ny, int64_t nx, int64_t ysplit,
                     void *ws, size_t ws_bytes, psb200_stream stream);
int psb200_edt_z_u32(psb200_ctx *ctx, const uint32_t *h, void *out, int out_kind,
                     uint32_t *max_out, int64_t nz, int64_t ny, int64_t nx,
                     psb200_stream stream);

/* out[i] = float32(sqrt(d2[i])) (IEEE, correctly rounded == np.sqrt(float32));
 * PSB200_INF_U32 -> +inf.  This is the float32 array edt.edt returns. */
int psb200_sqrt_f32(psb200_ctx *ctx, const uint32_t *d2, float *out, int64_t n,
                    psb200_stream stream);

/* *dev_out = max(d2[0..n)) (device scalar, overwritten).  np.amax(dt) at F:1132. */
int psb200_max_u32(psb200_ctx *ctx, const uint32_t *d2, int64_t n, uint32_t *dev_out,
                   psb200_stream stream);

/* ------------------------------------------------ sphere-insertion loop (F:1177-1209)
 * For k = 0..nT-1 (thresholds strictly descending, host array T[k] = min{n : sqrt_f32(n)
 * >= r_k}):   seeds = d2 >= T[k]  [F:1180/1196];  if inlet_mode != NONE keep only seeds
 * connected (6-conn in 3-D, 4-conn in 2-D) to the inlets through seeds|inlets
 * [F:1181-1183 -> F:1252-1270];  fill = {v : exists seed s, |v-s|^2 < T[k]}  [F:1191 ==
 * F:1207];  idx[v] = k+1 where idx[v]==0 and fill  [F:1192/1209].
 * idx is uint8 [nz][ny][nx]; zero-initialised by the call unless PSB200_FLAG_IDX_PREINIT.
 * ndim (1,2,3) is the dimensionality of the squeezed image (selects the faces predicate).
 * d2 is read-only. */
size_t psb200_local_thickness_workspace_bytes(const psb200_ctx *ctx, int64_t nz, int64_t ny,
                                              int64_t nx, int inlet_mode);
int psb200_local_thickness_idx(psb200_ctx *ctx, const uint32_t *d2,
                               const uint32_t *T_host, int nT, uint8_t *idx,
                               const uint8_t *inlets, int inlet_mode, int ndim,
                               int64_t nz, int64_t ny, int64_t nx, int flags,
                               void *ws, size_t ws_bytes, psb200_stream stream);

/* Step-level entry points of the same loop, for z-slab sharded volumes: the xy part of
 * one radius is slab-local; the z part needs up to W = ceil(sqrt(T))-1 halo planes of
 * the reach map from each z-neighbour (m_lo = the nlo planes just below local z=0 in
 * ascending z order, m_hi = the nhi planes just above local z=nz-1; NULL/0 at the ends).
 * psb200_lt_xy workspace: nz*ny*nx bytes (x-distances) select the streaming kernels; with another nz*ny*nx/8 + 512
 * bytes behind them the x pass runs from packed seed bits (faster); without a workspace the fused any-shape kernel runs. */
int psb200_lt_classify(psb200_ctx *ctx, const uint32_t *d2, const uint32_t *T_host, int nT,
                       uint8_t *cls, int64_t n, psb200_stream stream);
int psb200_lt_xy(psb200_ctx *ctx, const uint8_t *cls, int k, uint32_t T, uint8_t *reach,
                 int64_t nz, int64_t ny, int64_t nx, void *ws, size_t ws_bytes,
                 psb200_stream stream);      /* ws: >= nz*ny*nx + 256 bytes of scratch */
/* reach is overwritten (forward cone values) by the call */
int psb200_lt_z(psb200_ctx *ctx, uint8_t *reach, const uint8_t *m_lo, int nlo,
                const uint8_t *m_hi, int nhi, uint8_t *idx, int k, uint32_t T,
                int64_t nz, int64_t ny, int64_t nx, psb200_stream stream);

/* Bit-parallel form of the same step for small radii (nx % 32 == 0; buffers 16-byte aligned):
 *   psb200_lt_pack    : bits[v/32] bit v%32 = (cls[v] <= k)                 (seeds, F:1180)
 *   psb200_lt_wmask   : written[v/32] bit   = (idx[v] != 0)
 *   psb200_lt_bitball : fill = dilation of the seed bits by {o : |o|^2 < T} (F:1191 == F:1207);
 *                       idx[v] = k+1 where fill and not yet written (F:1192); written |= fill.
 * seedbits holds nz_src planes; output plane z reads seed plane z + z_off, so a z-slab shard
 * puts the W = ceil(sqrt(T))-1 halo planes of its neighbours in front of / behind its own. */
int psb200_lt_pack(psb200_ctx *ctx, const uint8_t *cls, int k, uint32_t *bits,
                   int64_t nz, int64_t ny, int64_t nx, psb200_stream stream);
/* seed bits of nk <= 16 consecutive radii k0 .. k0 + nk - 1 from ONE read of the class map (bit-sliced
 * comparator): bits + i * vol_words receives radius k0 + i (vol_words >= nz*ny*nx/32, multiple of 4). */
int psb200_lt_packn(psb200_ctx *ctx, const uint8_t *cls, int k0, int nk, uint32_t *bits, int64_t vol_words,
                    int64_t nz, int64_t ny, int64_t nx, psb200_stream stream);
int psb200_lt_wmask(psb200_ctx *ctx, const uint8_t *idx, uint32_t *written,
                    int64_t nz, int64_t ny, int64_t nx, psb200_stream stream);
int psb200_lt_bitball(psb200_ctx *ctx, const uint32_t *seedbits, int64_t nz_src, int64_t z_off,
                      uint32_t *written, uint8_t *idx, int k, uint32_t T,
                      int64_t nz, int64_t ny, int64_t nx, psb200_stream stream);

/* out[i] = lut_host[idx[i]] as float64 (np.zeros(shape) + radii, F:1178, F:1192, F:1212).
 * lut_host has nlut entries (entry 0 must be 0.0).  With PSB200_FLAG_EXPAND_MERGE only
 * voxels with idx in [1, nlut) are written. */
int psb200_expand_idx_f64(psb200_ctx *ctx, const uint8_t *idx, const double *lut_host,
                          int nlut, double *out, int64_t n, int flags, psb200_stream stream);
/* The same map delivered into HOST memory (`return imresults`, F:1212) -- synchronous, returns when
 * out_host[0, n) is complete.  The device holds one index byte per voxel, so the volume is split:
 * the first cpu_permille/1000 of it crosses PCIe as index bytes (into stage_host, page-locked,
 * stage_bytes >= that many voxels) and is widened to float64 by `nthreads` host threads of this
 * library (table lookup only; 0 = all hardware threads); the rest is widened on the device into two
 * chunk buffers carved from the device workspace `ws` and crosses PCIe as float64.  Both halves
 * overlap.  cpu_permille = 0 (or stage_host NULL) selects the all-device path.  out_host should be
 * page-locked for full PCIe speed; `idx` is produced on `stream`. */
int psb200_expand_idx_f64_to_host(psb200_ctx *ctx, const uint8_t *idx, const double *lut_host,
                                  int nlut, double *out_host, int64_t n, uint8_t *stage_host,
                                  size_t stage_bytes, void *ws, size_t ws_bytes, int cpu_permille,
                                  int nthreads, int flags, psb200_stream stream);
/* Host prologue: upload a one-byte-per-voxel HOST volume (numpy bool / uint8, foreground <=> byte != 0,
 * F:1126 `im > 0`) as 0/1 bytes into dst (device, n bytes).  Host threads pack it to one bit per
 * voxel chunk by chunk, the chunks cross PCIe at an eighth of the size, a kernel spreads the bits.
 * stage_host: page-locked, >= ceil(n/8) bytes; ws: device, >= ceil(n/8) bytes.  The host source and
 * the staging buffer may be reused on return (the call waits for its last host-to-device copy). */
int psb200_upload_mask_u8(psb200_ctx *ctx, const uint8_t *src_host, int64_t n, uint8_t *dst,
                          uint8_t *stage_host, size_t stage_bytes, void *ws, size_t ws_bytes,
                          int nthreads, psb200_stream stream);
/* idx[i] = out[i] != 0 ? PSB200_IDX_KEEP : 0   (continuation across >253 thresholds) */
int psb200_mark_written(psb200_ctx *ctx, const double *out, uint8_t *idx, int64_t n,
                        psb200_stream stream);

/* ---------------------------------------------- trim_disconnected_blobs (F:1215-1270)
 * out[v] = mask[v] && (the conn-component of (mask | inlets) containing v holds an inlet).
 * conn: 6 or 26 (3-D), 4 or 8 (2-D, nz = 1).  inlets: uint8 mask (non-zero = inlet). */
size_t psb200_flood_workspace_bytes(const psb200_ctx *ctx, int64_t nz, int64_t ny, int64_t nx);
int psb200_flood(psb200_ctx *ctx, const uint8_t *mask, const uint8_t *inlets, uint8_t *out,
                 int conn, int64_t nz, int64_t ny, int64_t nx,
                 void *ws, size_t ws_bytes, psb200_stream stream);

/* The same flood for NESTED sets (steps 0 .. nsteps-1): cls[v] = first step at which voxel v is a node (254: never,
 * 255: not a node); rcls[v] = first step at which it is a node connected to the inlets (>= cls[v]; 254 never, 255
 * where cls is 255).  inlets_in_set = 0: inlet voxels are nodes from step 0 on (trim_disconnected_blobs, F:1265);
 * 1: an inlet voxel only counts from its own step on (find_trapped_regions looks the outlets up inside the labelled
 * set `seq >= i`, F:131-137).  One time-ordered union pass with join times replaces one flood per step: drainage's
 * pressure loop `simulations/_drainage.py:133-154`, find_trapped_regions' bin loop `filters/_funcs.py:131-137`.
 * Workspace: psb200_flood_workspace_bytes(). */
int psb200_flood_classes(psb200_ctx *ctx, const uint8_t *cls, const uint8_t *inlets, int inlets_in_set, uint8_t *rcls,
                         int nsteps, int conn, int64_t nz, int64_t ny, int64_t nx, void *ws, size_t ws_bytes,
                         psb200_stream stream);

/* Device-side bit packing of a 0 / non-zero byte mask and its inverse (0 / 1 bytes): bits[i] bit j = (src[8 i + j] != 0).
 * The z-slab shards exchange their EDT input halo planes in this form (an eighth of the bytes).  dst: 8-byte aligned. */
int psb200_mask_pack_u8(psb200_ctx *ctx, const uint8_t *src, uint8_t *bits, int64_t n, psb200_stream stream);
int psb200_mask_unpack_u8(psb200_ctx *ctx, const uint8_t *bits, uint8_t *dst, int64_t n, psb200_stream stream);

/* z-slab shards: the cone value a neighbour's sweep receives from this slab's reach bytes (psb200_lt_xy output),
 * one byte per column: out[y][x] = max_j (reach[plane j from the face] - j), j < depth (the reach W of the radius).
 * side 0: the face towards the lower neighbour (plane 0), side 1: towards the upper one (plane nz-1).  The
 * neighbour passes the plane to psb200_lt_z as a one-plane halo (nhi = 1 resp. nlo = 1). */
int psb200_lt_halo_cone(psb200_ctx *ctx, const uint8_t *reach, int64_t nz, int64_t ny, int64_t nx, int depth,
                        int side, uint8_t *out, psb200_stream stream);

/* Step-level access-limited flooding (F:1181-1183 inside the radius loop) for volumes sharded
 * into z-slabs: this rank holds planes [z0, z0+nz) of nz_global.  The union-find
 * (parent: nz*ny*nx + 1 uint32, node 0 = virtual inlet root) is slab-local and kept across radii;
 * connectivity through a slab face is exchanged by the host as one byte per face voxel:
 *   psb200_uf_begin    : rcls = "not reached" class map, parent = every inlet -> root.  With
 *                        PSB200_INLETS_FACES the faces predicate uses GLOBAL z (z0, nz_global);
 *                        with PSB200_INLETS_MASK `inlets` is the local slab of the mask.
 *   psb200_uf_activate : link the voxels with klo < cls <= khi to their active 6-neighbours.
 *   psb200_uf_face     : flags_out[ny*nx] = node of local plane `zplane` is connected to the inlets.
 *   psb200_uf_inject   : nb_flags = the neighbour rank's facing plane from psb200_uf_face; links the
 *                        nodes under a set flag to the root; *changed_dev = 1 if that was news.
 *   psb200_uf_mark     : rcls[v] = k for seeds (cls <= k) now connected (F:1268-1269); *any_dev = 1
 *                        once any voxel was ever marked.
 * The host repeats face -> exchange -> inject until no rank reports a change, then marks. */
int psb200_uf_begin(psb200_ctx *ctx, const uint8_t *cls, uint8_t *rcls, uint32_t *parent,
                    const uint8_t *inlets, int inlet_mode, int ndim, int64_t nz, int64_t ny,
                    int64_t nx, int64_t z0, int64_t nz_global, psb200_stream stream);
size_t psb200_uf_workspace_bytes(const psb200_ctx *ctx, int64_t nz, int64_t ny, int64_t nx);
int psb200_uf_activate(psb200_ctx *ctx, uint32_t *parent, const uint8_t *cls, const uint8_t *inlets,
                       int inlet_mode, int ndim, int klo, int khi, int64_t nz, int64_t ny,
                       int64_t nx, int64_t z0, int64_t nz_global, void *ws, size_t ws_bytes,
                       psb200_stream stream);   /* ws: psb200_uf_workspace_bytes() of scratch */
/* The same flood on the row-rooted forest + link records + join times of the single-GPU loop
 * (csrc/flood_kernels.cuh).  `rec` (psb200_uf_records_bytes() bytes) is owned by the caller and kept untouched
 * across the radius loop:
 *   psb200_uf_begin_records    : forest with the x chains pre-linked, link records of ALL radius indices bucketed
 *   psb200_uf_activate_records : the unions of the indices klo < k <= khi
 *   psb200_uf_face_records / psb200_uf_inject_records : as psb200_uf_face / _inject; a tree that is connected
 *                                through a slab face at index k records k as its join time
 *   psb200_uf_resolve_records  : after the LAST index (all activations and exchanges done), one pass:
 *                                rcls[v] = first index at which v is a seed connected to the inlets
 *                                = max(cls[v], join time of v's tree); CLS 254 = never, 255 = background.
 * There is no per-radius marking: the radius loop runs on the final rcls map afterwards. */
size_t psb200_uf_records_bytes(const psb200_ctx *ctx, int64_t nz, int64_t ny, int64_t nx);
int psb200_uf_begin_records(psb200_ctx *ctx, const uint8_t *cls, uint32_t *parent, const uint8_t *inlets,
                            int inlet_mode, int ndim, int64_t nz, int64_t ny, int64_t nx, int64_t z0,
                            int64_t nz_global, void *rec, size_t rec_bytes, psb200_stream stream);
int psb200_uf_activate_records(psb200_ctx *ctx, uint32_t *parent, int klo, int khi, int64_t nz, int64_t ny,
                               int64_t nx, void *rec, size_t rec_bytes, psb200_stream stream);
int psb200_uf_face_records(psb200_ctx *ctx, uint32_t *parent, const uint8_t *cls, const uint8_t *inlets,
                           int inlet_mode, int ndim, int k, int64_t zplane, uint8_t *flags_out, int64_t nz,
                           int64_t ny, int64_t nx, int64_t z0, int64_t nz_global, void *rec, size_t rec_bytes,
                           psb200_stream stream);
int psb200_uf_inject_records(psb200_ctx *ctx, uint32_t *parent, const uint8_t *cls, const uint8_t *inlets,
                             int inlet_mode, int ndim, int k, int64_t zplane, const uint8_t *nb_flags,
                             int *changed_dev, int64_t nz, int64_t ny, int64_t nx, int64_t z0,
                             int64_t nz_global, void *rec, size_t rec_bytes, psb200_stream stream);
int psb200_uf_resolve_records(psb200_ctx *ctx, uint32_t *parent, const uint8_t *cls, uint8_t *rcls, int64_t nz,
                              int64_t ny, int64_t nx, void *rec, size_t rec_bytes, psb200_stream stream);
int psb200_uf_face(psb200_ctx *ctx, uint32_t *parent, const uint8_t *cls, const uint8_t *inlets,
                   int inlet_mode, int ndim, int k, int64_t zplane, uint8_t *flags_out, int64_t nz,
                   int64_t ny, int64_t nx, int64_t z0, int64_t nz_global, psb200_stream stream);
int psb200_uf_inject(psb200_ctx *ctx, uint32_t *parent, const uint8_t *cls, const uint8_t *inlets,
                     int inlet_mode, int ndim, int k, int64_t zplane, const uint8_t *nb_flags,
                     int *changed_dev, int64_t nz, int64_t ny, int64_t nx, int64_t z0,
                     int64_t nz_global, psb200_stream stream);
int psb200_uf_mark(psb200_ctx *ctx, uint32_t *parent, const uint8_t *cls, uint8_t *rcls, int k,
                   int *any_dev, int64_t n, psb200_stream stream);

/* ---- device-side `ps.generators.blobs` (generators/_imgen.py:1023-1051; norm_to_uniform T:963-969), the
 * input generator of every benchmark configuration.  float64, scipy's arithmetic
 * order; see porespy_b200/csrc/blobs_kernels.cuh.
 *   psb200_noise_philox_f64 : out[i] = uniform [0,1) noise of GLOBAL element first + i (Philox4x32-10 keyed
 *                             by seed: a pure function of the voxel, shard-invariant)
 *   psb200_gauss_axis_f64   : one 1-D correlation of gaussian_filter (mode='reflect') along `axis` (0 = z,
 *                             1 = y, 2 = x); w_host[0..radius] = the kernel from its outermost tap to the centre
 *                             (host doubles).  out is [nz][ny][nx].  axis 1 / 2: in has the same shape.  axis 0:
 *                             out holds global planes [z_out0, z_out0+nz), in holds global planes
 *                             [z_in0, z_in0+nz_in) (the slab plus the filter reach), the global volume has
 *                             nz_glob planes (reflection at ITS ends).  ws: psb200_gauss_workspace_bytes(radius).
 *                             Synchronises the stream once (weight upload).
 *   psb200_stats_f64        : part[p * psb200_stats_chunks() + c] = partial sum (mode 0), sum of (x-mean)^2
 *                             (mode 1), min (2) or max (3) over chunk c of plane p -- fixed order, so the host
 *                             combines the same numbers whatever the sharding.  part: device doubles.
 *   psb200_blobs_finish     : norm_to_uniform + `< porosity` -> out_u8 (0/1), or the uniformised field ->
 *                             out_f64 when out_u8 is NULL (porosity=None in the reference). */
int psb200_noise_philox_f64(psb200_ctx *ctx, double *out, int64_t n, uint64_t seed, uint64_t first,
                            psb200_stream stream);
size_t psb200_gauss_workspace_bytes(const psb200_ctx *ctx, int radius);
int psb200_gauss_axis_f64(psb200_ctx *ctx, const double *in, double *out, int axis, const double *w_host,
                          int radius, int64_t nz, int64_t ny, int64_t nx, int64_t z_out0, int64_t z_in0,
                          int64_t nz_in, int64_t nz_glob, void *ws, size_t ws_bytes, psb200_stream stream);
int psb200_stats_chunks(void);
int psb200_stats_f64(psb200_ctx *ctx, const double *x, int64_t nplanes, int64_t plane, double mean, int mode,
                     double *part, psb200_stream stream);
int psb200_blobs_finish(psb200_ctx *ctx, const double *f, int64_t n, double mean, double sd, double fmin_,
                        double fmax_, double porosity, uint8_t *out_u8, double *out_f64, psb200_stream stream);

/* ---- radius-map post-processing on the INDEX form: size_to_seq / size_to_satn /
 * seq_to_satn (filters/_size_seq_satn.py:16-221), pore_size_distribution (metrics/_funcs.py:558-632), pc_curve's
 * sizes branch (metrics/_funcs.py:1073-1090).  A map with K <= 65536 distinct values is one index per voxel
 * (1 byte for K <= 256, else 2) plus a table; the functions above reduce to a histogram of the index map and a
 * table expansion (porespy_b200/sizemap.py does the arithmetic on the K values with the reference's numpy code).
 *   psb200_hist_idx    : counts[(mask && mask[i] ? K : 0) + idx[i]] += 1; counts: K (mask NULL) or 2K device
 *                        uint64, zeroed by the call.
 *   psb200_expand_lut8 : out[i] = lut[(mask && mask[i] ? K : 0) + idx[i]], 8-byte payloads (float64 / int64);
 *                        lut: device, K or 2K entries.
 *   psb200_distinct64  : distinct 8-byte patterns of x[0,n) into an open-addressing table of `cap` (power of
 *                        two) device uint64 (empty slots = 0x8000000000000000, i.e. -0.0 / INT64_MIN, which the
 *                        input must not hold; the call initialises the table); *overflow = 1 when it is too small.
 *   psb200_index_of64  : idx[i] = position of x[i] in the sorted keys[0,K) (kind 0: float64 order, 1: int64). */
int psb200_hist_idx(psb200_ctx *ctx, const void *idx, int idx_bytes, const uint8_t *mask, int64_t n, int K,
                    uint64_t *counts, psb200_stream stream);
int psb200_expand_lut8(psb200_ctx *ctx, const void *idx, int idx_bytes, const uint8_t *mask, const uint64_t *lut,
                       void *out, int64_t n, int K, psb200_stream stream);
/* the same with one-byte payloads (masks such as `seq >= i`); lut: device bytes, K or 2K entries */
int psb200_expand_lut1(psb200_ctx *ctx, const void *idx, int idx_bytes, const uint8_t *mask, const uint8_t *lut,
                       uint8_t *out, int64_t n, int K, psb200_stream stream);
int psb200_distinct64(psb200_ctx *ctx, const uint64_t *x, int64_t n, uint64_t *table, uint32_t cap, int *overflow,
                      psb200_stream stream);
int psb200_index_of64(psb200_ctx *ctx, const uint64_t *x, int64_t n, const uint64_t *keys, int K, int kind,
                      void *idx, int idx_bytes, psb200_stream stream);

/* ---- image-based drainage: the pressure loop of simulations/_drainage.py:104-154 and the
 * sphere painter tools/_sphere_insertions.py:327-385.  fn = pc + rho g h is evaluated per voxel in the
 * reference's own precisions from dt (float32 EDT of im), or from a caller-supplied float64 pc map:
 * c0 = -(ndim-1) sigma cos(theta), rho_g = delta_rho * g, inner = voxels per step of the first image axis (the
 * direction of gravity), prec_flags bit 0/1/2: dt*voxel_size / h / rgh are float64 products (numpy scalars) instead
 * of float32 ones (python scalars).
 *   psb200_drain_stats     : partials[2b] = max{fn < inf}, partials[2b+1] = min{fn > -inf over im} per block b
 *                            (F:122-123); partials: device doubles, 2 * nblocks.
 *   psb200_drain_threshold : temp = (fn <= p) * im [+ residual]  (F:137-140), uint8 0/1.
 *   psb200_drain_newly     : rad = int(dt) at the voxels of reached [* mask] that are not in seeds yet, 0 elsewhere;
 *                            seeds |= reached [* mask]; *count_dev = how many, *maxr_dev = largest radius (F:142-152).
 *   psb200_drain_paint     : inv[v] = val wherever a sphere {|o|^2 < rad(s)^2} of a voxel s with rad(s) > 0 covers v
 *                            and inv[v] == 0 (power-diagram min-plus passes; ws: psb200_drain_paint_workspace_bytes).
 *   psb200_set_where_u8 / psb200_set_zero_codes_u8 : code-map bookkeeping of the epilogue (F:157-161). */
int psb200_drain_stats(psb200_ctx *ctx, const float *dt, const uint8_t *im, const double *pc_user, int64_t n,
                       int64_t inner, double c0, double voxel_size, double rho_g, int prec_flags, double *partials,
                       int nblocks, psb200_stream stream);
int psb200_drain_threshold(psb200_ctx *ctx, const float *dt, const uint8_t *im, const double *pc_user,
                           const uint8_t *residual, int64_t n, int64_t inner, double c0, double voxel_size,
                           double rho_g, int prec_flags, double p, uint8_t *temp, psb200_stream stream);
int psb200_drain_newly(psb200_ctx *ctx, const uint8_t *reached, const uint8_t *mask, uint8_t *seeds, const float *dt,
                       uint16_t *rad, int64_t n, uint64_t *count_dev, int *maxr_dev, psb200_stream stream);
/* Ascending pressures: every step from ONE flood.  psb200_drain_classify: cls[v] = first step k whose set
 * (fn <= ps[k]) * im [+ residual] holds v (254 never, 255 outside im and residual); psb200_flood_classes gives the
 * first step at which v is invaded; psb200_drain_newly_rcls is psb200_drain_newly for step k from that map
 * (newly = (rcls == k) [* mask]). */
int psb200_drain_classify(psb200_ctx *ctx, const float *dt, const uint8_t *im, const double *pc_user,
                          const uint8_t *residual, int64_t n, int64_t inner, double c0, double voxel_size,
                          double rho_g, int prec_flags, const double *ps_host, int np, uint8_t *cls,
                          psb200_stream stream);
int psb200_drain_newly_rcls(psb200_ctx *ctx, const uint8_t *rcls, int k, const uint8_t *mask, const float *dt,
                            uint16_t *rad, int64_t n, uint64_t *count_dev, int *maxr_dev, psb200_stream stream);
size_t psb200_drain_paint_workspace_bytes(const psb200_ctx *ctx, int64_t nz, int64_t ny, int64_t nx);
int psb200_drain_paint(psb200_ctx *ctx, const uint16_t *rad, int rmax, uint8_t *inv, int val, int64_t nz, int64_t ny,
                       int64_t nx, void *ws, size_t ws_bytes, psb200_stream stream);
int psb200_set_where_u8(psb200_ctx *ctx, uint8_t *dst, const uint8_t *mask, int value, int64_t n, psb200_stream stream);
int psb200_set_zero_codes_u8(psb200_ctx *ctx, uint8_t *codes, const uint8_t *im, const uint8_t *zero_lut_dev, int value,
                             int64_t n, psb200_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* PSB200_H */

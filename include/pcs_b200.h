/* pcs_b200.h -- C ABI of the B200-native point-cloud stitching hot path.
 *
 * Drop-in boundary for conix-center/pointcloud_stitching.  The reference has no
 * plugin/FFI layer: its seam is a handful of free functions called from main()
 * (SURVEY.md s8b).  Each entry point below names the reference function it
 * replaces (file:line relative to the reference checkout) and keeps that
 * function's buffer layout byte for byte:
 *
 *   wire record   10 bytes  [x_mm:i16][y_mm:i16][z_mm:i16][R:u8][G:u8][B:u8][0:u8]
 *                           (src/pcs-camera-optimized.cpp:581-585)
 *   camera buffer [int32 payload_bytes][records...]   payload at byte offset 4
 *                           (src/pcs-camera-optimized.cpp:690,715-719)
 *   stitched buf  [int32 payload_bytes][cam0 records][cam1 records]...
 *                           (src/pcs-multicamera-client.cpp:378,385-395)
 *
 * Plain C types only; no torch / CUDA types in any signature (CUDA streams are
 * passed as void*).  Every call runs on the context's device and restores the calling
 * thread's current CUDA device before it returns.  Every function returns >= 0 on success (a count or a byte
 * size, as the reference function does) and a negative pcs_status on failure;
 * nothing here ever calls exit().  There is NO CPU fallback: without a CUDA
 * device pcs_b200_create() fails with PCS_ERR_CUDA.
 *
 * Pointer conventions: `*_host` arguments are ordinary host memory (pinned or
 * pageable); `*_dev` arguments are device memory on the context's GPU.  Device
 * frame / payload pointers must be 16-byte aligned (cudaMalloc and torch
 * allocations are).  To get a 16-byte aligned payload inside a
 * [int32][records] buffer, place the buffer at (aligned allocation + 12).
 */
#ifndef PCS_B200_H
#define PCS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PCS_API __attribute__((visibility("default")))
#else
#define PCS_API
#endif

#define PCS_B200_ABI_VERSION 2
#define PCS_B200_RECORD_BYTES 10
#define PCS_B200_HEADER_BYTES 4
/* src/pcs-camera-optimized.cpp:27,157,673: short[BUF_SIZE] buffer, memset(BUF_SIZE bytes) */
#define PCS_B200_CAMERA_BUF_SHORTS 5000000

typedef struct pcs_ctx pcs_ctx;
typedef struct pcs_batch pcs_batch;

typedef enum pcs_status {
    PCS_OK = 0,
    PCS_ERR_INVALID = -1,     /* bad argument (null, misaligned, n % 4 != 0, unknown stream ...) */
    PCS_ERR_CUDA = -2,        /* CUDA runtime / driver error, or no device; see pcs_b200_last_error */
    PCS_ERR_NOMEM = -3,
    PCS_ERR_UNSUPPORTED = -4,
    PCS_ERR_CAPACITY = -5     /* output does not fit the buffer given */
} pcs_status;

/* rs2_intrinsics.  model / coeffs follow rs2_distortion: 0 = none (D400 depth: all coefficients zero),
 * 1 = modified Brown-Conrady (applied when a point is PROJECTED: it matters on the colour intrinsics),
 * 2 = inverse Brown-Conrady (applied when a pixel is DEPROJECTED: it matters on the depth intrinsics),
 * 4 = Brown-Conrady (a rectified image: never applied); coeffs = k1, k2, p1, p2, k3.  As in librealsense's rsutil.h a
 * model on the side of the chain that does not apply it is accepted and has no effect (a D455 colour stream reports
 * inverse Brown-Conrady); F-Theta (3) and Kannala-Brandt (5) are refused with PCS_ERR_UNSUPPORTED.
 * Streams whose distortion does apply run the general kernel (the arithmetic is oracle/SPEC.md s1's restatement of
 * rsutil.h, parity unpinned like the rest of the deprojection).  A zero-initialised tail means "none". */
#define PCS_B200_DISTORTION_NONE 0
#define PCS_B200_DISTORTION_MODIFIED_BROWN_CONRADY 1
#define PCS_B200_DISTORTION_INVERSE_BROWN_CONRADY 2
#define PCS_B200_DISTORTION_BROWN_CONRADY 4
typedef struct pcs_intrinsics {
    int32_t width, height;
    float ppx, ppy, fx, fy;
    int32_t model;
    float coeffs[5];
} pcs_intrinsics;

/* Everything the reference keeps in compile-time constants and globals
 * (tf_mat src/pcs-camera-optimized.cpp:64-67; cutoff bounds :398-401; cached
 * geometry :369-377) plus what librealsense holds for rs2::pointcloud. */
typedef struct pcs_stream_desc {
    pcs_intrinsics depth;      /* z16 frame geometry; width % 8 == 0 */
    pcs_intrinsics color;      /* colour frame geometry */
    float d2c_rotation[9];     /* depth -> colour extrinsics, column-major (rs2_extrinsics) */
    float d2c_translation[3];  /* metres */
    float depth_scale;         /* metres per z16 unit (0.001 on D400) */
    int32_t color_bpp;         /* bytes per colour pixel, >= 3; R,G,B are bytes 0,1,2 */
    int32_t color_stride;      /* bytes per colour row */
    float tf[16];              /* camera -> world, row-major 4x4, row 3 ignored */
    int32_t cutoff;            /* -c: keep only z in (z_lo, z_hi], x in (x_lo, x_hi] (pre-transform) */
    float z_lo, z_hi, x_lo, x_hi; /* reference: 0, 1.5, -2, 2 */
    int32_t cutoff_lane_reversed; /* 1 = reproduce the reference's reversed mask lanes */
} pcs_stream_desc;

typedef struct pcs_config {
    int32_t device;            /* CUDA device ordinal */
    int32_t max_streams;       /* camera streams this context serves (>= 1) */
    int32_t kernel_variant;    /* 0 = auto, 1 = direct (LDG/STG), 2 = bulk-async pipelined (TMA) */
    int32_t voxel_variant;     /* 0 = auto (2, else 4, else 1), 1 = (key, index) pair sort, 2 / 3 = one-sweep sort with
                                  8- / 10-bit digits, 4 = slab partition + bitmap ranking (no sort, any n) */
} pcs_config;

/* One frame of work for the batched device-resident path. */
#define PCS_B200_JOB_REMOTE_FRAME 1   /* flags: z16_dev / color_dev point into a peer GPU's memory (pull exchange) */
typedef struct pcs_frame_job {
    int32_t stream;            /* index given to pcs_b200_set_stream */
    int32_t reserved;          /* flags (0, or PCS_B200_JOB_REMOTE_FRAME: a scheduling hint, results do not depend on it) */
    const uint16_t *z16_dev;   /* depth.height * depth.width */
    const uint8_t *color_dev;  /* color.height * color_stride bytes */
    int16_t *payload_dev;      /* depth.width*depth.height records (10 B each); with -c the kept records, compacted in raster
                                * order by the same launch (bytes past *count_dev records are not written) */
    float *xyzrgb_dev;         /* optional: N x {x,y,z metres (transformed), b,g,r,255} 16 B/pt; or NULL */
    int32_t *count_dev;        /* optional: receives the record count (needed with cutoff); or NULL */
} pcs_frame_job;

PCS_API int pcs_b200_abi_version(void);
PCS_API const char *pcs_b200_status_string(int status);

/* Lifetime.  Replaces the process-lifetime malloc/free of the reference mains
 * (src/pcs-camera-optimized.cpp:157,345; src/pcs-multicamera-client.cpp:499,554). */
PCS_API int pcs_b200_create(const pcs_config *cfg, pcs_ctx **out);
PCS_API void pcs_b200_destroy(pcs_ctx *ctx);
/* Message of the last failure on this context (thread-local when ctx is NULL). */
PCS_API const char *pcs_b200_last_error(const pcs_ctx *ctx);

/* Replaces the globals behind `initialized` (src/pcs-camera-optimized.cpp:349-409)
 * and the hard-coded transform (:64-72).  May be called again at any time for the host-buffer
 * calls.  A batch (pcs_b200_batch_create) freezes its streams' geometry, calibration and cutoff
 * settings: under a live batch only `tf` may change; after any other change pcs_b200_batch_run
 * returns PCS_ERR_INVALID until the batch is destroyed and created again. */
PCS_API int pcs_b200_set_stream(pcs_ctx *ctx, int stream, const pcs_stream_desc *desc);

/* ---- camera side, host buffers (what the reference's main() calls) ---------- */

/* Replaces rs2::pointcloud::calculate + sendXYZRGBPointcloud
 * (src/pcs-camera-optimized.cpp:288-292,669-723): depth + colour frame in,
 * the reference's camera buffer image out.  buffer_host is short[5 000 000]:
 * bytes [0,5 000 000) are zeroed, records start at byte 4, and with
 * write_header != 0 (-s) the int32 payload size is stored at byte 0.
 * Returns the payload size in bytes (:697,722). */
PCS_API int pcs_b200_send_xyzrgb(pcs_ctx *ctx, int stream, const uint16_t *z16_host,
                         const uint8_t *color_host, int16_t *buffer_host, int write_header);

/* The same call split in two so one host thread can keep several camera streams in
 * flight: begin() enqueues H2D + kernel + D2H on the stream's own CUDA stream and
 * returns; end() waits and finishes the buffer image.  The copies only overlap when
 * the host buffers are pinned (pcs_b200_host_alloc).  Calls on distinct `stream`
 * ids are thread-safe; one begin/end pair may be outstanding per stream. */
PCS_API int pcs_b200_send_xyzrgb_begin(pcs_ctx *ctx, int stream, const uint16_t *z16_host,
                               const uint8_t *color_host, int16_t *buffer_host, int write_header);
PCS_API int pcs_b200_send_xyzrgb_end(pcs_ctx *ctx, int stream);

/* Page-locked host memory for the calls above (cudaHostAlloc / cudaFreeHost). */
PCS_API void *pcs_b200_host_alloc(pcs_ctx *ctx, size_t bytes);
PCS_API void pcs_b200_host_free(pcs_ctx *ctx, void *p);

/* Replaces copyPointCloudXYZRGBToBufferSIMD (src/pcs-camera-optimized.cpp:363-616):
 * same inputs as the reference function (rs2::vertex[n], rs2::texture_coordinate[n],
 * colour frame), records out.  n % 4 == 0.  Returns the record count (:612-615). */
PCS_API int pcs_b200_pack_from_vertices(pcs_ctx *ctx, int stream, const float *xyz_host,
                                const float *uv_host, int n, const uint8_t *color_host,
                                int16_t *payload_host);

/* ---- camera side, device-resident and batched ------------------------------- */

/* Same work as pcs_b200_send_xyzrgb's kernel for many frames in one launch.
 * The job table is captured once; run() only launches.  cuda_stream is a
 * cudaStream_t (NULL = default stream). */
PCS_API int pcs_b200_batch_create(pcs_ctx *ctx, const pcs_frame_job *jobs, int n_jobs, pcs_batch **out);
PCS_API int pcs_b200_batch_run(pcs_ctx *ctx, pcs_batch *batch, void *cuda_stream);
PCS_API void pcs_b200_batch_destroy(pcs_ctx *ctx, pcs_batch *batch);
/* Fused compute + exchange (multi-GPU stitch): like pcs_b200_batch_create, but every tile of
 * records is stored by the same kernel to the local payload AND to the same offset of n_peers
 * mirror buffers in peer GPU memory (peer-mapped pointers, e.g. CUDA IPC / symmetric memory),
 * i.e. an all-gather of the packed XYZRGB records over NVLink with no second pass.  All
 * payload_dev pointers must lie inside [local_base, local_base + local_bytes).  Replaces the
 * TCP fan-in of readCloud (src/pcs-multicamera-client.cpp:363-371) + the concat (:385-395).
 * The caller synchronises the ranks after batch_run before reading mirrored data. */
PCS_API int pcs_b200_batch_create_fanout(pcs_ctx *ctx, const pcs_frame_job *jobs, int n_jobs,
                                         const void *local_base, size_t local_bytes,
                                         const void *const *peer_bases, int n_peers, pcs_batch **out);
/* Number of kernel launches one batch_run issues (for launch accounting). */
PCS_API int pcs_b200_batch_launches(const pcs_batch *batch);

/* Multi-GPU hosts that drive several GPUs from ONE process (one pcs_ctx per GPU): enables peer access
 * from this context's device to peer_device, after which a job's z16_dev / color_dev may point at
 * frames that live on the peer (the "pull" exchange: this GPU deprojects the peer's cameras itself,
 * 5 B/pt over NVLink instead of receiving 10 B/pt of records) and the mirrors of
 * pcs_b200_batch_create_fanout may be plain cudaMalloc memory of the peer.  Processes that own one
 * GPU each map peer memory with CUDA IPC / symmetric memory instead. */
PCS_API int pcs_b200_enable_peer(pcs_ctx *ctx, int peer_device);

/* Hosts that run ONE PROCESS PER GPU (the reference's deployment: one process per camera, fan-in over TCP,
 * src/pcs-multicamera-client.cpp:363-371): a process exports the device memory that holds its cameras'
 * frames (or its stitched mirror), sends the 80-byte handle to its peers over whatever channel it has (a
 * socket, MPI, a file) and every peer maps it with pcs_b200_ipc_open.  The mapped pointer is an ordinary
 * device pointer on the opener's GPU: it can be a job's z16_dev / color_dev (pull exchange, the peer's frames
 * are read over NVLink by the fused kernel itself) or a peer base of pcs_b200_batch_create_fanout.  dev_ptr may
 * point anywhere inside a cudaMalloc allocation (the handle records the offset).  The exporter must keep the
 * allocation alive until every opener has called pcs_b200_ipc_close; synchronising the processes around a
 * frame (nobody overwrites a frame that a peer is still reading) stays with the host, as it does for the
 * sockets of the reference. */
typedef struct pcs_ipc_handle {
    uint8_t reserved[64];      /* cudaIpcMemHandle_t of the allocation */
    uint64_t offset;           /* dev_ptr - allocation base */
    uint64_t device;           /* exporter's CUDA device ordinal (informative) */
} pcs_ipc_handle;
PCS_API int pcs_b200_ipc_export(pcs_ctx *ctx, const void *dev_ptr, pcs_ipc_handle *out);
PCS_API int pcs_b200_ipc_open(pcs_ctx *ctx, const pcs_ipc_handle *handle, void **dev_ptr_out);
PCS_API int pcs_b200_ipc_close(pcs_ctx *ctx, void *dev_ptr);

/* Device-pointer form of pcs_b200_pack_from_vertices; asynchronous on cuda_stream.  Without cutoff
 * the return value is the record count (n).  With cutoff (-c) the count is only known on the
 * device: it is written to *count_dev (required then, PCS_ERR_INVALID if NULL) and the return value
 * is the UPPER BOUND n -- records beyond *count_dev in payload_dev are stale. */
PCS_API int pcs_b200_pack_from_vertices_dev(pcs_ctx *ctx, int stream, const float *xyz_dev,
                                    const float *uv_dev, int n, const uint8_t *color_dev,
                                    int16_t *payload_dev, int32_t *count_dev, void *cuda_stream);

/* ---- stitch side ------------------------------------------------------------- */

/* Replaces the concat loop of sendStitchToUnity (src/pcs-multicamera-client.cpp:385-395):
 * for each camera in order, every `downsample`-th record is appended at
 * stitched + 4; the int32 total is stored at stitched + 0.  n_shorts[i] is the
 * camera's payload length in shorts (as readCloud leaves it, :368).
 * Returns the stitched payload bytes. */
PCS_API int pcs_b200_stitch_raw_dev(pcs_ctx *ctx, const int16_t *const *payload_dev, const int32_t *n_shorts,
                            int n_cams, int downsample, uint8_t *stitched_dev, size_t stitched_cap,
                            void *cuda_stream);
PCS_API int pcs_b200_stitch_raw(pcs_ctx *ctx, const int16_t *const *payload_host, const int32_t *n_shorts,
                        int n_cams, int downsample, uint8_t *stitched_host, size_t stitched_cap);

/* Replaces updateCloudXYZRGB -> `+=` -> convertPointCloudXYZRGBToBuffer
 * (src/pcs-multicamera-optimized.cpp:226-265,288-289,364-367,308-310): per camera
 * unpack (int16 / 1000.0f), decimate, 4x4 transform, ordered append, repack
 * (truncate(x * 1000.0f)).  transforms = n_cams x 16 floats row-major (host).
 * cloud32_dev, when not NULL, also receives the stitched cloud as 32-byte
 * pcl::PointXYZRGB records.  Returns the stitched payload bytes. */
PCS_API int pcs_b200_stitch_pcl_dev(pcs_ctx *ctx, const int16_t *const *payload_dev, const int32_t *n_shorts,
                            int n_cams, int downsample, const float *transforms,
                            uint8_t *stitched_dev, size_t stitched_cap, void *cloud32_dev,
                            void *cuda_stream);
PCS_API int pcs_b200_stitch_pcl(pcs_ctx *ctx, const int16_t *const *payload_host, const int32_t *n_shorts,
                        int n_cams, int downsample, const float *transforms,
                        uint8_t *stitched_host, size_t stitched_cap);

/* The whole camera -> stitcher path in one call, host buffers on both sides: depth + colour frames
 * of n_cams cameras in, the reference's stitched buffer out.  Replaces, per frame,
 * n_cams x (rs2::pointcloud::calculate + sendXYZRGBPointcloud, src/pcs-camera-optimized.cpp:288-292),
 * the TCP fan-in of readCloud (src/pcs-multicamera-client.cpp:363-371) and the concat loop of
 * sendStitchToUnity (:373-395): stitched_host receives [int32 payload bytes][records of streams[0]]
 * [records of streams[1]]... with every `downsample`-th record of each camera kept (:388).
 * Nothing returns to the host between the cameras' kernels and the concat: every camera's kernel writes its records
 * into the camera's slot of a device-resident stitched buffer and the records come back into their place in
 * stitched_host.  Without decimation every camera runs on its own CUDA stream (frame up, kernel, records down), so
 * that the first cameras' records travel down while the last cameras' frames still travel up -- both directions of
 * the link stay busy; with downsample > 1 all cameras run in one batched launch, the device decimates, and one copy
 * brings [int32][records] back.
 *   slot          0 .. PCS_B200_STITCH_SLOTS-1: independent pipelines; a host thread keeps two frames
 *                 in flight by alternating slots (begin(0) begin(1) end(0) begin(0) end(1) ...)
 *   streams       n_cams stream ids (pcs_b200_set_stream), in stitched order.  With a cutoff (-c) stream in the set the
 *                 record counts only exist on the device: the concat reads them there, and end() fetches the total
 *                 before the records (one more synchronisation)
 *   z16_host / color_host   n_cams frame pointers each; pinned memory (pcs_b200_host_alloc) lets the
 *                 copies overlap the other slot's work
 *   stitched_cap  bytes available at stitched_host
 * begin() enqueues everything and returns; end() waits and returns the stitched payload bytes. */
#define PCS_B200_STITCH_SLOTS 4
PCS_API int pcs_b200_stitch_frames_begin(pcs_ctx *ctx, int slot, int n_cams, const int32_t *streams,
                                 const uint16_t *const *z16_host, const uint8_t *const *color_host,
                                 int downsample, uint8_t *stitched_host, size_t stitched_cap);
PCS_API int pcs_b200_stitch_frames_end(pcs_ctx *ctx, int slot);
PCS_API int pcs_b200_stitch_frames(pcs_ctx *ctx, int n_cams, const int32_t *streams,
                           const uint16_t *const *z16_host, const uint8_t *const *color_host,
                           int downsample, uint8_t *stitched_host, size_t stitched_cap);

/* Voxel-grid merge of n records (own integer specification, oracle/SPEC.md s3; the
 * reference includes pcl/filters/voxel_grid.h but never calls it).  Returns the
 * number of voxels written to out_dev (capacity n records).  The sort-based variants need
 * n * max(256, leaf_mm) < 2^32 (uint32 sums: 16.7 M points at the 10 mm leaf); larger clouds, up to what
 * the reference's int32 size header can describe (n * 10 < 2^31), take the slab-partition variant when
 * leaf <= 32 mm and a z plane of the occupied box holds <= 2^24 voxels (40 m x 40 m at the 10 mm leaf).
 * The call synchronises cuda_stream once, at the end (the voxel count has to reach the host). */
PCS_API int pcs_b200_voxel_merge_dev(pcs_ctx *ctx, const int16_t *records_dev, int n, int leaf_mm,
                             int16_t *out_dev, void *cuda_stream);
PCS_API int pcs_b200_voxel_merge(pcs_ctx *ctx, const int16_t *records_host, int n, int leaf_mm,
                         int16_t *out_host);

/* Enqueue-only forms of the merge (the one-sweep sort, voxel_variant 0 / 2 / 3): the occupied box, the key
 * layout, the number of sort passes and the voxel count all stay on the device, so a frame loop can queue
 * K1 + merge for many frames and synchronise once.  *count_dev (device int32) receives the voxel count, or
 * a negative pcs_status (PCS_ERR_UNSUPPORTED: the cloud's (key, index) word does not fit 64 bits -- use the
 * synchronous call, which falls back).  Merges of one context share scratch memory: queue them on ONE
 * stream.  Returns PCS_OK or a negative pcs_status. */
PCS_API int pcs_b200_voxel_merge_async_dev(pcs_ctx *ctx, const int16_t *records_dev, int n, int leaf_mm,
                                   int16_t *out_dev, int32_t *count_dev, void *cuda_stream);
PCS_API int pcs_b200_voxel_merge_slab_async_dev(pcs_ctx *ctx, const int16_t *records_dev, int n, int leaf_mm,
                                        int kz_lo, int kz_hi, int16_t *out_dev, int32_t *count_dev,
                                        void *cuda_stream);

/* The same, for records whose number is only known on the device: n_max bounds it (grids and scratch are
 * sized for n_max), *n_dev (device int32, <= n_max) is read by the kernels.  Used on an inbox that peer GPUs
 * fill (pcs_b200_shard_scatter_dev). */
PCS_API int pcs_b200_voxel_merge_counted_async_dev(pcs_ctx *ctx, const int16_t *records_dev, int n_max,
                                           const int32_t *n_dev, int leaf_mm, int16_t *out_dev,
                                           int32_t *count_dev, void *cuda_stream);

/* Multi-GPU voxel merge, sharded BEFORE the exchange ("sharded by voxel-key range with one
 * all-to-all"): cameras are independent up to the merge (src/pcs-multicamera-client.cpp:381-392), so every GPU
 * keeps the records of its own cameras and only the merge's input crosses NVLink, once:
 *   1. pcs_b200_shard_hist_dev     points per z plane of my records -> my zhist; my inbox cursor = 0
 *      -- barrier between the ranks (the host's: symmetric-memory barrier, MPI, a socket) --
 *   2. pcs_b200_shard_plan_dev     every rank adds all ranks' histograms (read over NVLink) and cuts the z axis
 *                                  into n_ranks slabs of equal population: identical cuts everywhere, no collective
 *   3. pcs_b200_shard_scatter_dev  the all-to-all: every record goes to the inbox of the rank that owns its z slab
 *                                  (peer stores; one system-scope atomic per tile and destination reserves the run)
 *      -- barrier --
 *   4. pcs_b200_voxel_merge_counted_async_dev(inbox, capacity, cursor, ...)   each rank merges its slab
 * Slab r of the grid ends on rank r; the slabs concatenated in rank order are the single-GPU merge bit for bit.
 * All pointers in pcs_shard_peers are device pointers valid on THIS rank's GPU (its own memory for `rank`,
 * peer-mapped memory for the others: CUDA IPC via pcs_b200_ipc_open, or symmetric memory).  zhist buffers hold
 * pcs_b200_shard_zbins(leaf_mm) uint32 each, zslab_dev as many bytes, kz_splits_dev n_ranks + 1 int32. */
typedef struct pcs_shard_peers {
    int32_t n_ranks, rank;                 /* n_ranks <= 8 */
    void *inbox_dev[8];                    /* capacity_records x 10 bytes each, 16-byte aligned */
    uint32_t *cursor_dev[8];               /* records received so far (the inbox's fill count) */
    const uint32_t *zhist_dev[8];
    int64_t capacity_records;
} pcs_shard_peers;
PCS_API int pcs_b200_shard_zbins(int leaf_mm);
PCS_API int pcs_b200_shard_hist_dev(pcs_ctx *ctx, const int16_t *records_dev, int n, int leaf_mm, uint32_t *zhist_dev,
                            uint32_t *cursor_dev, void *cuda_stream);
PCS_API int pcs_b200_shard_plan_dev(pcs_ctx *ctx, const pcs_shard_peers *peers, int leaf_mm, int32_t *kz_splits_dev,
                            uint8_t *zslab_dev, void *cuda_stream);
PCS_API int pcs_b200_shard_scatter_dev(pcs_ctx *ctx, const int16_t *records_dev, int n, int leaf_mm,
                               const uint8_t *zslab_dev, const pcs_shard_peers *peers, uint32_t *err_dev,
                               void *cuda_stream);

/* Sharded voxel merge (multi-GPU: SURVEY s8(e) "sharded by voxel-key range").  The grid is cut
 * along z into n_slabs slabs of nearly equal population: slab r holds the points with
 * kz_splits[r] <= floor(z / leaf_mm) < kz_splits[r + 1] (kz_splits: n_slabs + 1 host ints;
 * slab_points: n_slabs host ints or NULL).  The plan depends only on the records, so ranks that hold
 * the same stitched cloud compute the same cuts without communicating.
 * pcs_b200_voxel_merge_slab_dev merges one slab; because voxels are emitted in ascending
 * (kz, ky, kx) order, the outputs of slabs 0 .. n_slabs-1 concatenated are exactly what
 * pcs_b200_voxel_merge_dev returns for the whole cloud.  Returns the slab's voxel count. */
PCS_API int pcs_b200_voxel_slab_plan_dev(pcs_ctx *ctx, const int16_t *records_dev, int n, int leaf_mm,
                                 int n_slabs, int32_t *kz_splits, int32_t *slab_points, void *cuda_stream);
PCS_API int pcs_b200_voxel_merge_slab_dev(pcs_ctx *ctx, const int16_t *records_dev, int n, int leaf_mm,
                                  int kz_lo, int kz_hi, int16_t *out_dev, void *cuda_stream);

/* PLY dump of the stitched pcl::PointXYZRGB cloud (the 32-byte records pcs_b200_stitch_pcl_dev writes
 * to cloud32_dev), replacing pcl::io::savePLYFileBinary in visualize()
 * (src/pcs-multicamera-client.cpp:482-489).  _rows_dev: n 15-byte binary PLY vertices
 * (float x, y, z; uchar red, green, blue) into rows_dev.  save_ply: header + vertices + PCL's
 * one-row camera element into a file.  Both return n. */
PCS_API int pcs_b200_cloud_to_ply_rows_dev(pcs_ctx *ctx, const void *cloud32_dev, int n, uint8_t *rows_dev,
                                   void *cuda_stream);
PCS_API int pcs_b200_save_ply(pcs_ctx *ctx, const void *cloud32_dev, int n, const char *path);

/* Blocks until everything issued on cuda_stream by this context has finished. */
PCS_API int pcs_b200_synchronize(pcs_ctx *ctx, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* PCS_B200_H */

/*
 * bseg.h -- C ABI of libbseg.so, the B200 (sm_100a) implementation of buildingSegment's
 * per-point segmentation hot path.
 *
 * The reference has no FFI: its boundary is plain C++ inside one process
 * (/root/reference/tmc3/TMC3.cpp:202-229 calls my_function.h / my_function.cpp directly).
 * This header is the boundary a maintainer binds instead; every entry point names the
 * reference interface it replaces.  The C++ shims in buildingsegment_b200/host/ present the
 * reference's own signatures on top of it.
 *
 * Conventions
 *   - plain pointers and sizes only; the caller owns every host buffer, the library owns all
 *     device memory; nothing throws across the boundary;
 *   - every function returns 0 on success or a negative bseg_status; bseg_last_error() holds text;
 *   - a context is bound to one CUDA device and one stream, and is not thread-safe;
 *   - there is no CPU fallback: without a CUDA device bseg_create fails with BSEG_E_NODEVICE.
 *   - coordinates are int32 (millimetres after ply::read's x1000); after the shift to the cloud
 *     minimum every extent must be < 2^21 kNN cells and < 2^23 units (BSEG_E_ARG otherwise).
 */
#ifndef BSEG_H
#define BSEG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(BSEG_BUILD)
#define BSEG_API __attribute__((visibility("default")))
#else
#define BSEG_API
#endif

typedef struct bseg_ctx bseg_ctx;

typedef enum bseg_status {
  BSEG_OK = 0,
  BSEG_E_ARG = -1,      /* bad argument / parameter out of the supported range */
  BSEG_E_NODEVICE = -2, /* no usable CUDA device -- there is no CPU path */
  BSEG_E_CUDA = -3,     /* a CUDA call failed; see bseg_last_error */
  BSEG_E_STATE = -4,    /* stage called out of order (e.g. grow before knn_normals) */
  BSEG_E_CAPACITY = -5, /* an internal table overflowed */
  BSEG_E_NOMEM = -6
} bseg_status;

/* Every literal of the reference, as a parameter.  bseg_default_params fills the values below. */
typedef struct bseg_params {
  int32_t K;              /* 15    neighbours per row, self included   TMC3.cpp:215-216        */
  int32_t max_nn;         /* 50    hybrid search cap                   my_function.h:63        */
  double radius;          /* 100.0 hybrid search radius (strict <)     my_function.h:63        */
  int32_t th_thickness;   /* 300   |(p-c).n| <= th                     my_function.h:117       */
  int32_t th_point_count; /* 400   plane kept if pointIdx.size() > th  my_function.h:118       */
  double th_dot;          /* 0.88  n.n_i >= th                         my_function.cpp:230     */
  int32_t bin;            /* 100   raster pixel edge                   TMC3.cpp:177            */
  int32_t bin_height;     /* 1000  z-histogram bin of groundTH         TMC3.cpp:177            */
  double count_bias;      /* 20.0  added to log(count+1) when non-zero TMC3.cpp:163            */
  int32_t cell;           /* 0 = auto: kNN grid cell edge, >= ceil(radius)                      */
  int32_t grow_mode;      /* 0 = speculative parallel engine, 1 = single-warp sequential engine */
  double grow_radius;     /* 0 = off (the reference).  > 0: "radius-search growing" (BASELINE config C2): the grower's
                             neighbour list of a point is the prefix of its K-row with d^2 < grow_radius^2 -- entries
                             beyond the radius do not exist (they are never accepted, Broad at depth 0 still wants
                             all K-1, my_function.cpp:238).  Rows exported by bseg_knn_normals stay the full kNN rows. */
  int32_t reserved[3];
} bseg_params;

/* Per-stage device time of the last call, milliseconds (CUDA events on the context's stream). */
typedef struct bseg_timings {
  float h2d, bbox_keys, sort, cells, knn, knn_fallback, normals, grow, finalize, raster, d2h, total;
  float grow_slice_ms, grow_sweep_ms; /* parallel engine: device time of the grower slices / of sweeper + mark log */
  float raster_host;      /* host half of the raster's count channel (libm log, image B), wall ms; overlaps the grower */
  int64_t n_unresolved;   /* queries that needed the ring-expansion fallback */
  int64_t grow_steps;     /* Broad() calls executed (committed work)          */
  int64_t grow_rounds;    /* rounds of the speculative engine                 */
  int64_t kernel_launches;/* kernels launched since bseg_reset_counters       */
  int64_t n_big_cells;    /* cells whose 27-neighbourhood did not fit the shared-memory stage */
  int64_t grow_wasted_steps; /* Broad() calls of speculative growers that were released (parallel engine) */
  int64_t grow_sweep_iters;  /* batches of the sweeper                                                  */
  int64_t grow_tiny_tx;      /* depth-0 failures committed by the sweeper                               */
  int64_t grow_seq_fallbacks;/* planes grown alone by the sequential engine (slot storage exhausted)    */
  int64_t grow_head_steps;   /* Broad() calls of the head slot: the critical path of the parallel engine */
  int64_t grow_head_ns;      /* device time of those calls, ns                                          */
  int64_t grow_sweep_ns;     /* device time inside the sweeper, ns                                      */
  int64_t grow_at_fails;     /* finished growers re-run because an assumed-taken point was free after all */
} bseg_timings;

/* ---- lifetime ---------------------------------------------------------------------------- */
BSEG_API int bseg_create(bseg_ctx** out, int device);
BSEG_API void bseg_destroy(bseg_ctx* ctx);
BSEG_API int bseg_default_params(bseg_params* p);
BSEG_API const char* bseg_last_error(const bseg_ctx* ctx);
BSEG_API const char* bseg_version(void);

/* ---- stage a3: buildingSeg::buildingSeg (TMC3.cpp:55-79) ------------------------------------
 * Uploads the N x 3 int32 AoS cloud (what &cloud[0][0] points at, PCCPointSet.h:271-275), finds
 * the bounding box on the device and shifts the device copy to min = 0.  out_min/out_max receive
 * the UNSHIFTED box; the caller's cloud is shifted too when xyz_shifted_out != NULL (the reference
 * ctor shifts the caller's cloud, TMC3.cpp:71).  xyz_aos may be pinned or pageable. */
BSEG_API int bseg_set_points(bseg_ctx* ctx, const int32_t* xyz_aos, int64_t n, int32_t out_min[3],
                             int32_t out_max[3], int32_t* xyz_shifted_out);

/* Same, for a cloud already resident in device memory (bench "value" leg, multi-GPU slabs). */
BSEG_API int bseg_set_points_device(bseg_ctx* ctx, const int32_t* d_xyz_aos, int64_t n, int32_t out_min[3],
                                    int32_t out_max[3]);

/* ---- stages a4 + a5: get_Normal_and_K_neighbor<K> (my_function.h:48-85) ----------------------
 * Morton binning (radix sort), exact kNN ordered by (d^2, index), hybrid-radius PCA normals
 * oriented to +z.  Outputs are optional (NULL = keep on device only):
 *   neigh_NxK    row i = the K nearest of point i in ORIGINAL indices, -1 padded when N < K
 *   normals_Nx3  unit normals
 *   curvature_N  lambda0/(l0+l1+l2) -- extension, the reference computes none */
BSEG_API int bseg_knn_normals(bseg_ctx* ctx, const bseg_params* p, int32_t* neigh_NxK, double* normals_Nx3,
                              double* curvature_N);

/* Replace the device-resident normals / neighbour rows with caller-supplied ones (original
 * indexing).  Lets tests drive the grower with adversarial inputs, as the reference's
 * seg_plane constructor accepts arbitrary vectors (my_function.h:98). */
BSEG_API int bseg_override_neigh_normals(bseg_ctx* ctx, const bseg_params* p, const int32_t* neigh_NxK,
                                         const double* normals_Nx3);

/* ---- stages a7-a9: seg_plane::seg_plane + get_planes + Broad (my_function.cpp:180-258) --------
 *   plane_idx_N  the reference's Cloud.planeIdx after get_planes: -1 free, else a plane id
 *                (orphan marks of failed seeds included)
 *   label_N      0, or the id of the last plane whose pointIdx holds the point
 *   n_planes     number of committed planes */
BSEG_API int bseg_grow_planes(bseg_ctx* ctx, const bseg_params* p, int32_t* plane_idx_N, int32_t* label_N,
                              int32_t* n_planes);

/* seg_plane on a cloud that buildingSeg has NOT shifted (my_function.h:98 takes any cloud): the device copy is
 * always shifted to min = 0 for the binning, but the reference's int32 centroid sums (and where they wrap, SURVEY
 * A.2-Q6) are in the caller's coordinates.  offset (normally out_min of bseg_set_points) is added back, with int32
 * wrap, to every coordinate the grower's model arithmetic sees; NULL or zeros = the shifted coordinates (default,
 * what TMC3.cpp:210-218 feeds seg_plane).  plane centres are reported in the same coordinates. */
BSEG_API int bseg_set_grow_offset(bseg_ctx* ctx, const int32_t offset[3]);

/* std::vector<plane> of get_planes (my_function.h:25-30): ids are 1..P in order;
 * offsets_Pp1/point_idx are pointIdx in the reference's order, duplicates kept.
 * Call with point_idx == NULL to size the buffers (total entries = offsets[P]). */
BSEG_API int bseg_get_planes(bseg_ctx* ctx, int32_t* seeds_P, double* normals_Px3, int32_t* centers_Px3,
                             int64_t* offsets_Pp1, int32_t* point_idx);

/* Plane post-processing: plane equations and roof / facade / ground classes (north_star stage 4: "... producing roof
 * and facade labels").  The reference declares the record -- struct DetectedPlane { indices; equation ax+by+cz+d=0;
 * normal; d }, my_function.h:41-46 -- and never fills it, so there is no reference behaviour to match; the definitions
 * here are the record's own: for plane p (id p + 1) of the last bseg_grow_planes
 *   (a, b, c) = the plane's cur_normal (my_function.cpp:249), d = -((a * cx + b * cy) + c * cz) with cur_center
 *   (my_function.cpp:250), doubles, evaluated in that order;
 *   class = BSEG_CLASS_FACADE  if |c| <= facade_max_nz                        (the surface is close to vertical)
 *           BSEG_CLASS_ROOF    if |c| >= roof_min_nz and cz >= ground_z       (flat or pitched, above the ground)
 *           BSEG_CLASS_GROUND  if |c| >= roof_min_nz and cz <  ground_z
 *           BSEG_CLASS_OTHER   otherwise (a NaN model, Q9, lands here).
 * ground_z: normally the threshold bseg_raster reports (buildingSeg::groundTH, TMC3.cpp:181-198), same coordinates as
 * the plane centres.  point_class_N (may be NULL): class of the plane each point is labelled with, 0 for an unlabelled
 * point -- one device pass over the labels.  equations_Px4 / plane_class_P may be NULL. */
enum { BSEG_CLASS_NONE = 0, BSEG_CLASS_ROOF = 1, BSEG_CLASS_FACADE = 2, BSEG_CLASS_GROUND = 3, BSEG_CLASS_OTHER = 4 };
BSEG_API int bseg_plane_classes(bseg_ctx* ctx, double facade_max_nz, double roof_min_nz, double ground_z,
                                double* equations_Px4, uint8_t* plane_class_P, uint8_t* point_class_N);

/* ---- stage a10: seg_plane::set_plane_color (my_function.cpp:260-275) --------------------------
 * Paints exactly the pointIdx of the planes it is given, in the order given (later planes overwrite, :268-274),
 * black everywhere else.  plane_ids_Q names the planes (ids 1..P of the last bseg_grow_planes, any subset in any
 * order -- a caller may filter or reorder the vector before colouring); NULL means all planes in id order, and
 * then n_listed must equal the plane count.  plane_rgb_Qx3 is the host's rand() sequence (55 + rand() % 200,
 * three per LISTED plane, drawn in the listed order).  A count that disagrees is BSEG_E_ARG, never a read past
 * the caller's buffer. */
BSEG_API int bseg_paint(bseg_ctx* ctx, const int32_t* plane_ids_Q, int32_t n_listed, const uint16_t* plane_rgb_Qx3,
                        uint16_t* colors_Nx3);

/* ---- stages a13-a15: groundTH + compute_gird_picture + save_image pixels (TMC3.cpp:81-198) -----
 * image_WxHx3: doubles, pixel(x,y,ch) = image[(y*W+x)*3+ch]; png_rgb[0..2]: the three W*H*3
 * uint8 images handed to stbi_write_png (A: mean height in byte 0, B: log count in byte 1,
 * C: always zero).  Any output may be NULL.  W,H are also available from bseg_raster_size. */
BSEG_API int bseg_raster_size(bseg_ctx* ctx, const bseg_params* p, int32_t* W, int32_t* H);
BSEG_API int bseg_raster(bseg_ctx* ctx, const bseg_params* p, double* image_WxHx3, uint8_t* png_a,
                         uint8_t* png_b, uint8_t* png_c, double* ground_th);

/* The host half of the count channel, for callers that hold the weight sums themselves (bseg_raster_device leaves
 * them in channel 1): values[i] = log(values[i] + 1), plus bias when that is non-zero (TMC3.cpp:159-164), with the
 * platform's std::log on worker threads; *max_out = the channel maximum save_image needs (TMC3.cpp:85-90). */
BSEG_API int bseg_count_channel(double* values, int64_t n, double bias, double* max_out);

/* ---- extracted_contour (my_function.cpp:8-145), without OpenCV -----------------------------------------------------
 * bseg_contour_mask   (device) :19-26: mask = 255 where pixels[.., channel] > thresh (extractChannel + THRESH_BINARY; the
 *                     reference takes channel 1 of the count image and 10), then MORPH_CLOSE with the 5x5 elliptic
 *                     element of getStructuringElement, `iterations` times (2) -- `iterations` dilations followed by
 *                     `iterations` erosions, pixels outside the image not taking part.  pixels / mask_out: host.
 * bseg_find_contours  (host) :30-33: findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE when simple != 0, else _NONE): the
 *                     same contours, in the same order, each with the same first point and orientation.  offsets has
 *                     n_contours + 1 entries (in points); call with points_xy == offsets == NULL to size the buffers.
 * bseg_contour_measure (host) :37-38: contourArea and arcLength(closed) of one contour.
 * bseg_draw_contour   (host) :58: one closed contour, 2 pixels thick: every pixel within one pixel of a segment gets
 *                     color3 (the reference's Scalar(255,255,0) in the image's channel order).  OpenCV's fixed-point
 *                     polygon fill is not reproduced bit for bit; everything else above is. */
BSEG_API int bseg_contour_mask(bseg_ctx* ctx, const uint8_t* pixels, int32_t w, int32_t h, int32_t comp, int32_t channel,
                               int32_t thresh, int32_t iterations, uint8_t* mask_out);
BSEG_API int bseg_find_contours(const uint8_t* mask, int32_t w, int32_t h, int32_t simple, int32_t* points_xy, int64_t cap_points,
                                int64_t* offsets, int64_t cap_contours, int64_t* n_contours, int64_t* n_points);
BSEG_API int bseg_contour_measure(const int32_t* points_xy, int64_t n, double* area, double* perimeter);
BSEG_API int bseg_draw_contour(uint8_t* image, int32_t w, int32_t h, int32_t comp, const int32_t* points_xy, int64_t n,
                               const uint8_t* color3);

/* ---- the PNG files of save_image (TMC3.cpp:98,108,119 call stbi_write_png, stb_image_write.h:1215) --------------
 * Host-only.  The file is BYTE-identical to what stbi_write_png of the reference's vendored stb_image_write v1.16
 * writes with its default settings (per-row filter choice, its own deflate at level 8): comp = bytes per pixel (1..4),
 * stride_bytes = 0 means w * comp.  bseg_png_encode returns the length in *len (out may be NULL to size the buffer).
 * bseg_png_write_async copies the pixels and encodes on a worker thread -- the three images of save_image are
 * encoded concurrently and off the caller's critical path; bseg_png_wait joins every pending write. */
BSEG_API int bseg_png_encode(const uint8_t* pixels, int32_t w, int32_t h, int32_t comp, int32_t stride_bytes, uint8_t* out,
                             int64_t cap, int64_t* len);
BSEG_API int bseg_png_write(const char* path, const uint8_t* pixels, int32_t w, int32_t h, int32_t comp, int32_t stride_bytes);
BSEG_API int bseg_png_write_async(const char* path, const uint8_t* pixels, int32_t w, int32_t h, int32_t comp,
                                  int32_t stride_bytes);
BSEG_API int bseg_png_wait(void);

/* ---- the raster of one slab of a tile (multi-GPU): same kernels as bseg_raster, but the ground
 * threshold of TMC3.cpp:181-198 comes from the caller when ground_th is not NULL (the slabs share the tile's, found
 * from the summed z histograms) and the W*H*3 doubles stay on the device (*d_image, valid until the next
 * bseg_set_points / bseg_raster*).  Channel 1 holds the exact weight SUMS: the caller takes them through
 * bseg_count_channel on the host (the log is the platform libm's, as in the reference).  Point order = the cloud's order, as always: buildingsegment_b200/slabs.py
 * uploads [left halo | owned | right halo] so that it is the tile's order. */
BSEG_API int bseg_raster_device(bseg_ctx* ctx, const bseg_params* p, const double* ground_th, const double** d_image,
                                int32_t* W, int32_t* H);

/* ---- label raster (north_star stage 5; extension: the reference rasters height and count only, TMC3.cpp:127-172,
 * and paints labels per point, my_function.cpp:260-275) ---------------------------------------
 * Same W x H grid as bseg_raster.  label_WxH[y*W+x] = plane label (1..P, 0 = none) of the highest point of
 * the pixel (ties: lower index); rgb_WxHx3 = that plane's colour from plane_rgb_Px3 (the set_plane_color
 * sequence handed to bseg_paint), black where the label is 0 -- the byte image for stbi_write_png.
 * Any of the three buffers may be NULL. */
BSEG_API int bseg_label_raster(bseg_ctx* ctx, const bseg_params* p, const uint16_t* plane_rgb_Px3, int32_t* label_WxH,
                               uint8_t* rgb_WxHx3);

/* ---- whole path, device resident (no host transfers): what bench.py times as `value` --------- */
#define BSEG_RUN_KNN 1
#define BSEG_RUN_GROW 2
#define BSEG_RUN_RASTER 4
#define BSEG_RUN_ALL 7
BSEG_API int bseg_run_device(bseg_ctx* ctx, const bseg_params* p, int stages);

/* ---- whole path through host buffers: what bench.py times as `e2e` --------------------------- */
BSEG_API int bseg_segment_host(bseg_ctx* ctx, const bseg_params* p, const int32_t* xyz_aos, int64_t n,
                               int32_t* xyz_shifted_out, int32_t* label_N, int32_t* n_planes,
                               uint8_t* png_a, uint8_t* png_b, int32_t* W, int32_t* H);

/* ---- introspection ---------------------------------------------------------------------------- */
BSEG_API int bseg_get_timings(bseg_ctx* ctx, bseg_timings* out);
BSEG_API int bseg_reset_counters(bseg_ctx* ctx);
BSEG_API void* bseg_stream(bseg_ctx* ctx); /* cudaStream_t the context launches on */
BSEG_API int64_t bseg_point_count(const bseg_ctx* ctx);
BSEG_API int32_t bseg_plane_count(const bseg_ctx* ctx); /* planes of the last grow stage (vector<plane>::size(), my_function.cpp:216), -1 = none */

/* ---- multi-GPU slabs (one context per rank; the exchange itself is the caller's NCCL) --------- */
/* Declares that the first n_owned points of the cloud passed to bseg_set_points are this rank's
 * own slab and the rest are halo copies: halo points serve as neighbours only. */
BSEG_API int bseg_set_owned(bseg_ctx* ctx, int64_t n_owned);
/* Shift the next clouds by `origin` (the tile's minimum, the same on every rank) instead of each cloud's
 * own minimum (TMC3.cpp:70-72 subtracts the minimum of the WHOLE cloud); NULL restores the default.
 * out_min of bseg_set_points then reports the origin. */
BSEG_API int bseg_set_origin(bseg_ctx* ctx, const int32_t origin[3]);
/* Device pointers (original point order) to label / planeIdx / the shifted cloud of the last run, for the
 * cross-slab label merge without a host round trip; valid until the next bseg_set_points. */
BSEG_API int bseg_device_results(bseg_ctx* ctx, const int32_t** d_label, const int32_t** d_plane_idx,
                                 const int32_t** d_xyz_shifted);
/* Halo sufficiency: counts owned points within `halo` of the slab faces x_lo / x_hi (shifted units) whose
 * K-th neighbour is farther than the nearest point this rank may be missing (halo + its distance to the face) --
 * a closer point beyond the halo could exist; 0 = kNN rows and normals of all owned points equal those of the
 * undivided cloud (given halo >= params.radius).  A point is owned when x_lo <= x < x_hi (and, if bseg_set_owned
 * was called, its index is below n_owned).  An outer face of the tile, which has no neighbour rank and nothing
 * beyond it, is passed as INT32_MIN (x_lo) / INT32_MAX (x_hi) and is not checked. */
BSEG_API int bseg_halo_check(bseg_ctx* ctx, int32_t x_lo, int32_t x_hi, int32_t halo, int64_t* n_unresolved);

/* Rows and normals of the last kNN stage in ORIGINAL index space, left in device memory (valid until the next
 * bseg_set_points / kNN stage): the slab's share of get_Normal_and_K_neighbor (my_function.h:48-85) travels to the
 * rank that grows the tile over NVLink, no host round trip. */
BSEG_API int bseg_knn_device_results(bseg_ctx* ctx, const bseg_params* p, const int32_t** d_neigh_NxK,
                                     const double** d_normals_Nx3);
/* The device-pointer form of bseg_override_neigh_normals that does NOT run the kNN stage: the cloud is binned
 * (sorted order only) and the rows / normals -- computed elsewhere, e.g. by the slabs of a tile -- become the
 * input of bseg_grow_planes, as seg_plane's constructor takes any vectors (my_function.h:98). */
BSEG_API int bseg_import_neigh_normals_device(bseg_ctx* ctx, const bseg_params* p, const int32_t* d_neigh_NxK,
                                              const double* d_normals_Nx3);

/* ---- self-test hooks used by tests/ (exercise the hand-written primitives in isolation) ------- */
BSEG_API int bseg_debug_sort_pairs(bseg_ctx* ctx, uint64_t* keys, uint32_t* vals, int64_t n, int key_bits);
BSEG_API int bseg_debug_exclusive_scan(bseg_ctx* ctx, uint32_t* data, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* BSEG_H */

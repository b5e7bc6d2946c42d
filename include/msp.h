/*
 * msp.h — C ABI of libmsp.so, the B200 (sm_100a) geometry hot path of the MultiSPA data engine.
 *
 * The reference (facebookresearch/Multi-SpatialMLLM) is pure Python and has no FFI; the boundary it
 * offers is the Python surface of spatial_engine/utils/scannet_utils/handler/info_handler.py (IH) and the
 * per-scene workers of make_visibility_info.py (MVI) / calculate_frames_relations.py (CFR).  Each entry
 * point below names the reference lines it replaces; the Python mirror of that surface lives in
 * multi_spatialmllm_b200/ and binds these symbols with ctypes.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the parameter is called host_*; the caller owns all memory,
 *    the library never allocates, frees or synchronises; all work is enqueued on `stream` (a cudaStream_t
 *    passed as void*; NULL = legacy default stream);
 *  - return 0 on success, a negative MSP_E* code otherwise; msp_last_error() gives a thread-local message;
 *  - intrinsics (host_K16) are a HOST pointer to the reference's 4x4 row-major float64 K (IH:98);
 *  - poses are 4x4 row-major float64.  "c2w" is camera->world exactly as the reference stores it
 *    (IH:100-111); the kernels never see a host-inverted matrix: msp_pose_prepare inverts on the device;
 *  - points: pts_dtype 0 = packed float32 xyz (12 B), 1 = float32 xyzw (16 B), 2 = packed float64 xyz (24 B).
 *    All index-determining arithmetic is float64 in the reference's operation order (IH:57-69, IH:359-371);
 *  - visibility masks are bit sets: bit (p & 31) of word mask[f * mask_stride_words + (p >> 5)] is point p in
 *    frame f; mask_stride_words must be a multiple of 4 and >= ceil(n/32); pad bits are written as 0.
 */
#ifndef MSP_H_
#define MSP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MSP_API __attribute__((visibility("default")))
#else
#define MSP_API
#endif

#define MSP_VERSION 101

#define MSP_OK 0
#define MSP_EINVAL (-1)   /* bad argument (null pointer, size, alignment, dtype) */
#define MSP_ELAUNCH (-2)  /* CUDA launch / runtime error */

#define MSP_PTS_F32X3 0
#define MSP_PTS_F32X4 1
#define MSP_PTS_F64X3 2

/* Image geometry shared by the per-frame kernels.
 * image_h/w: colour image size the intrinsics refer to (IH:133-139, IH:341); depth_h/w: depth map size
 * (IH:355); depth_scale: raw depth unit in metres (IH:76, 0.001 for ScanNet). */
typedef struct msp_camera {
  int32_t image_h, image_w;
  int32_t depth_h, depth_w;
  double depth_scale;
} msp_camera;

/* One fully annotated frame pair (128 bytes), written by msp_pair_annotate.
 *  inter/uni/overlap  CFR:128-137          |A&B|, |A|B|, inter/uni*100 (NaN when uni == 0)
 *  distance,dyaw,dpitch CFR:181-183         ||t2-t1||, yaw2-yaw1, pitch2-pitch1 (degrees)
 *  rel_disp           CME:185-189           translation of inv(E1)@E2
 *  point_id           VC-coor:303-315       k-th smallest common visible point id, -1 if none / k out of range
 *  flags              IH:297-300            bit0/bit1: the point passes the visibility re-check in frame 1/2
 *  permille           VC-coor:341-344       round_half_even(u/W*1000), round_half_even(v/H*1000) per frame
 *  depth_mm           DE-dot:218            round_half_even(z*1000) per frame
 *  pix                DE-dot:261            (int)u, (int)v (truncation) per frame */
typedef struct msp_pair_record {
  uint32_t inter, uni;
  double overlap;
  double distance, dyaw, dpitch;
  double rel_disp[3];
  int32_t point_id;
  int32_t flags;
  int32_t permille[4]; /* x1,y1,x2,y2 */
  int32_t depth_mm[2];
  int32_t pix[4]; /* u1,v1,u2,v2 */
  int32_t reserved[4];
} msp_pair_record;

MSP_API int msp_version(void);
MSP_API const char* msp_last_error(void);

/* Profiling aid: when given a device pointer to two zeroed uint64 counters, the K1/K2 fast path adds
 * [0] point-frames evaluated in float32 (members of candidate groups) and [1] point-frames re-evaluated by the exact
 * float64 path.  NULL (default) disables it.  Process-global; not meant for concurrent use. */
MSP_API int msp_debug_set_cull_stats(unsigned long long* dev_counters2);

/* Per-frame pose prologue.  Replaces IH:113-124 (aligned = axis_align @ c2w), the np.linalg.inv at IH:57,
 * CFR:86-100 (extract_yaw_pitch on the aligned pose) and IH:409-418 (a pose holding inf/nan is invalid).
 *  c2w        (F,16) in       axis_align (16) in, may be NULL (= identity, the reference's align=False)
 *  E_out      (F,16) out      aligned camera->world          w2c_out (F,16) out   its inverse
 *  yawpitch_out (F,2) out     degrees                        valid_out (F) out    1/0
 * Invalid frames get w2c = NaN so that every later test on them is false. */
MSP_API int msp_pose_prepare(const double* c2w, const double* axis_align, int32_t F, double* E_out, double* w2c_out,
                     double* yawpitch_out, int32_t* valid_out, void* stream);

/* The MVI:93-100 / CFR:152-157 frame loop for F frames of one scene: project every point
 * (IH:313-335 -> IH:46-72), test image bounds (IH:337-344) and the depth map (IH:346-373), AND them
 * (IH:375-386) and store the result as a bit mask.  xy_out (F,n,2) / z_out (F,n) float64 are optional
 * (NULL = do not materialise; the reference always does, at 24 B per point-frame).
 *  host_K16: (16) intrinsics in HOST memory (passed by value to the kernels).  w2c: (F,16) from msp_pose_prepare.  depth: (F,depth_h,depth_w) uint16. */
MSP_API int msp_project_visibility(const void* pts, int64_t n, int32_t pts_dtype, const double* host_K16, const double* w2c,
                           int32_t F, const uint16_t* depth, const msp_camera* host_cam, uint32_t* mask_bits,
                           int64_t mask_stride_words, double* xy_out, double* z_out, void* stream);

/* msp_project_visibility fused with msp_object_counts (COV:110-136): the masks AND counts[f*n_objects + o] in one
 * pass over the points, for any number of objects (each 1024-point tile keeps a table of its own instance ids; falls
 * back to the two separate kernels only when the intrinsics are not the standard form or n_objects == 0).
 * counts (F,n_objects) is overwritten.  Both visibility entry points use a stream-ordered workspace
 * (cudaMallocAsync / cudaFreeAsync on `stream`, ~14 KB per 1000 points) for the prepared point tiles. */
MSP_API int msp_project_visibility_objects(const void* pts, int64_t n, int32_t pts_dtype, const double* host_K16,
                                           const double* w2c, int32_t F, const uint16_t* depth,
                                           const msp_camera* host_cam, uint32_t* mask_bits, int64_t mask_stride_words,
                                           const uint16_t* instance_ids, int32_t n_objects, uint32_t* counts,
                                           void* stream);

/* Z-buffer depth render — NOT in the reference: per frame, per depth pixel, the minimum
 * float32(z) over points with z > 0 inside the colour image (IH:342-343), binned with the reference's own
 * index rule (IH:359-366).  zbuf (F,depth_h,depth_w) float32 is overwritten (+inf where empty). */
MSP_API int msp_zbuffer_render(const void* pts, int64_t n, int32_t pts_dtype, const double* host_K16, const double* w2c,
                       int32_t F, const msp_camera* host_cam, float* zbuf, void* stream);

/* zbuf -> ScanNet-style uint16 depth: floor(z/depth_scale)+1+margin_mm, clamped to 65535, 0 where empty. */
MSP_API int msp_depth_to_u16(const float* zbuf, int64_t count, int32_t margin_mm, double depth_scale, uint16_t* out,
                     void* stream);

/* CFR:102-137 for P pairs of mask rows: inter = |A&B|, uni = |A|B|, overlap = inter/uni*100 (float64). */
MSP_API int msp_pair_overlap(const uint32_t* mask_bits, int64_t n, int64_t mask_stride_words, const int32_t* pairs,
                     int64_t P, uint32_t* inter, uint32_t* uni, double* overlap_pct, void* stream);

/* The whole per-pair stage in one launch: overlap (CFR:102-137), pose deltas (CFR:181-183), relative
 * displacement (CME:185-189), selection of the k-th common visible point (np.intersect1d + index,
 * VC-coor:303-315) and its annotation in both frames (IH:291-305 with check_visible, VC-coor:341-344,
 * DE-dot:216-218,261).
 *  pairs (P,2) frame rows.  sel (P) uint32; sel_mode 0: sel is k itself (point_id = -1 if k >= inter);
 *  sel_mode 1: sel is a uniform 32-bit word and k = floor(sel * inter / 2^32).
 *  E (F,16), w2c (F,16), yawpitch (F,2) from msp_pose_prepare; depth as in msp_project_visibility. */
MSP_API int msp_pair_annotate(const uint32_t* mask_bits, int64_t n, int64_t mask_stride_words, const int32_t* pairs,
                      int64_t P, const uint32_t* sel, int32_t sel_mode, const void* pts, int32_t pts_dtype,
                      const double* host_K16, const double* E, const double* w2c, const double* yawpitch,
                      const uint16_t* depth, const msp_camera* host_cam, msp_pair_record* records, void* stream);

/* IH:46-72 for n points and ONE pose: uvz (n,3) float64 = (u, v, depth).  Uses the operation order NumPy applies to this
 * shape: the N >= 2 matrix-matrix order for n >= 2, the single-column order for n == 1 (they differ in the last ulp). */
MSP_API int msp_project_points(const void* pts, int64_t n, int32_t pts_dtype, const double* host_K16, const double* w2c_one,
                               double* uvz, void* stream);

/* IH:291-305 (+ the integer maths of DE-dot:216-218,261) for M (point, frame) requests.
 * uvz (M,3) float64; permille (M,2), depth_mm (M), pix (M,2) int32; visible (M) uint8.  Any output may be NULL.
 * Every request is evaluated in the single-column matmul order of the reference's per-point calls. */
MSP_API int msp_annotate_points(const void* pts, int64_t n, int32_t pts_dtype, const int32_t* point_ids,
                        const int32_t* frame_ids, int64_t M, const double* host_K16, const double* w2c,
                        const uint16_t* depth, const msp_camera* host_cam, double* uvz, int32_t* permille,
                        int32_t* depth_mm, int32_t* pix, uint8_t* visible, void* stream);

/* The reference's two-step API on already projected points (one frame): mode 1 = IH:337-344 (image bounds),
 * 2 = IH:346-373 (depth test incl. the round/clip index rule, no bounds AND), 3 = IH:375-386 (both).
 *  xy (n,2), z (n) float64 (z may be NULL for mode 1), depth (depth_h,depth_w) uint16, out (n) uint8. */
MSP_API int msp_check_visibility(const double* xy, const double* z, int64_t n, const uint16_t* depth,
                                 const msp_camera* host_cam, int32_t mode, uint8_t* out, void* stream);

/* COV:110-136 for every (frame, object): counts[f*n_objects + o] = |visible_f ∩ {p : instance_ids[p] == o+1}|
 * (instance ids are 1-indexed, 0 = no object, IH:250-261).  instance_ids (n) uint16. */
MSP_API int msp_object_counts(const uint32_t* mask_bits, int64_t n, int64_t mask_stride_words, int32_t F,
                      const uint16_t* instance_ids, int32_t n_objects, uint32_t* counts, void* stream);

/* SOCF:57-65 building block: extents[(f*n_objects + o)*6 + {0,1,2}] = min x,y,z and [+3..5] = max x,y,z (float64) over
 * the points of object o that are visible in frame f (+inf / -inf when none).  The coverage of a set of frames along an
 * axis is max(max_f) - min(min_f), which replaces the union-mask passes of find_minimal_combinations (SOCF:76-220). */
MSP_API int msp_object_extents(const uint32_t* mask_bits, int64_t n, int64_t mask_stride_words, int32_t F, const void* pts,
                               int32_t pts_dtype, const uint16_t* instance_ids, int32_t n_objects, double* extents,
                               void* stream);

/* TAPVid-3D tracks, OM-coord:446-451 + OM-coord:293-315: world = inv(w2c_t) @ [xyz,1]; pinhole projection
 * u = fx*x/(z+1e-8)+cx normalised by W,H; valid iff 0<=u_n<1, 0<=v_n<1, z>0; permille = round_half_even(.*1000).
 *  tracks (T,N,3) float32 camera coordinates, w2c (T,16), fxfycxcy (4) float64 device.
 *  world_out (T,N,3) float64, uvn_out (T,N,2) float64, permille_out (T,N,2) int32, valid_out (T,N) uint8; NULL ok. */
MSP_API int msp_track_project(const float* tracks, const double* w2c, const double* fxfycxcy, int32_t T, int32_t N,
                      int32_t image_h, int32_t image_w, double* world_out, double* uvn_out, int32_t* permille_out,
                      uint8_t* valid_out, void* stream);

/* OM-coord:49-77 (rigid_body_segmentation's accumulation): condensed_out[pair(i<j) in scipy squareform order] =
 * sum over t >= 1 of |d_t - d_{t-1}| where it exceeds `smoothing`, d_t = ||track_i(t) - track_j(t)|| (float64).
 * tracks (T,N,3) float32; condensed_out (N(N-1)/2) float64.  Linkage / fcluster stay on the host. */
MSP_API int msp_track_cumulative_loss(const float* tracks, int32_t T, int32_t N, double smoothing, double* condensed_out,
                                      void* stream);

/* OM-coord:484-498 / 329-356 for M (frame1, frame2, track) triples: ||w2-w1||, the displacement expressed in
 * camera-1 axes (w2c[f1] @ [d,0]) and the camera baseline ||c2-c1||.  out (M,5): dist, dx, dy, dz, cam_dist. */
MSP_API int msp_track_pair_displacement(const double* world, const double* w2c, int32_t T, int32_t N, const int32_t* triples,
                                int64_t M, double* out, void* stream);

/* OM-coord:484-498 for n_tracks selected tracks at once: the distances between ALL i<j pairs of each track's visible
 * frames, in the reference's pair order (i-major) and in np.linalg.norm(axis=1)'s arithmetic ((dx*dx + dy*dy) + dz*dz).
 *  world (T,N,3) float64 from msp_track_project; track_ids (n_tracks); vis_frames: the ascending visible frame indices
 *  of every track back to back, frame_offsets (n_tracks+1) into it; pair_offsets (n_tracks+1) into dist, with
 *  pair_offsets[t+1]-pair_offsets[t] = V_t(V_t-1)/2; max_pairs_per_track sizes the grid. */
MSP_API int msp_track_all_pair_distances(const double* world, int32_t T, int32_t N, const int32_t* track_ids,
                                         int32_t n_tracks, const int32_t* vis_frames, const int64_t* frame_offsets,
                                         const int64_t* pair_offsets, int64_t max_pairs_per_track, double* dist,
                                         void* stream);

/* ---- visibility table text (the reference's canonical parquet, make_visibility_info.py:59-65) --------------------------
 * The "values" column — json.dumps of every frame's sorted visible point ids (image_to_points, MVI:103-104) and of every
 * point's sorted frame ids (point_to_images, MVI:107-118) — written byte-exactly from the bit masks, so that the host
 * never builds a Python list.  Two phases per table: *_size fills bytes-per-row, the caller makes the exclusive scan
 * (offsets[0] = 0 ... offsets[rows] = total) and allocates `text`, *_write places row i at text + offsets[i].
 * Empty rows are "[]"; separators are ", " (json.dumps defaults).
 *  frame_rows (R) int32: mask rows to emit, in output order (NULL = 0..R-1).  For image_to_points pass the valid frames
 *  in image_ids order (dict insertion order, MVI:93-104); for point_to_images pass them sorted by id string (MVI:116-118).
 *  ids (R, id_pitch) uint8 + id_len (R) uint8: the frame id strings in frame_rows order; they must consist of characters
 *  json.dumps leaves unescaped (printable ASCII except '"' and '\\'). */
MSP_API int msp_visibility_json_points_size(const uint32_t* mask_bits, int64_t n, int64_t mask_stride_words,
                                            const int32_t* frame_rows, int32_t R, int64_t* row_bytes, void* stream);
MSP_API int msp_visibility_json_points_write(const uint32_t* mask_bits, int64_t n, int64_t mask_stride_words,
                                             const int32_t* frame_rows, int32_t R, const int64_t* offsets /* (R+1) */,
                                             uint8_t* text, void* stream);
MSP_API int msp_visibility_json_images_size(const uint32_t* mask_bits, int64_t n, int64_t mask_stride_words,
                                            const int32_t* frame_rows, int32_t R, const uint8_t* id_len,
                                            int64_t* point_bytes /* (n) */, void* stream);
MSP_API int msp_visibility_json_images_write(const uint32_t* mask_bits, int64_t n, int64_t mask_stride_words,
                                             const int32_t* frame_rows, int32_t R, const uint8_t* ids, int32_t id_pitch,
                                             const uint8_t* id_len, const int64_t* offsets /* (n+1) */, uint8_t* text,
                                             void* stream);

/* ---- dot annotation (the `*_dot` streams) --------------------------------------------------------------------------
 * DE-dot:171, DC-dot:337-341, VC-dot:363,390-391, OM-dot:337: `cv2.circle(img, (x, y), r, color, -1)` and
 * `cv2.putText(img, letter, org, FONT_HERSHEY_SIMPLEX, 1.0, color, 2)` on BGR uint8 images.  A mark is one circle
 * (radius < 0: none; radius > MSP_MAX_MARK_RADIUS: skipped) plus one optional glyph (glyph < 0: none) painted in
 * `bgr`.  Marks of one image MUST be contiguous in the array and are painted in array order (later marks overwrite
 * earlier ones, as sequential cv2 calls do); marks of different images are independent.
 * The circle is OpenCV's integer midpoint rasterisation, pixel-exact including clipping at the borders.  Glyphs are
 * 1-bit masks, 32 rows of one uint32 each (bit c of row r = pixel (r, c)), with meta (off_x, off_y, width, height)
 * giving the mask's top-left corner relative to the putText origin — multi_spatialmllm_b200/data/ holds the atlas
 * rendered by cv2 for A-Z0-9.  Pixel-exact with cv2.putText whenever the glyph box lies inside the image; a glyph cut
 * by the border is clipped per pixel (cv2 re-runs its line clipper there and moves a few edge pixels). */
#define MSP_MAX_MARK_RADIUS 1024
typedef struct msp_mark {
  int32_t image;        /* index into the (B,H,W,3) batch; out of range: mark ignored */
  int32_t x, y;         /* circle centre in pixels: int(u), int(v) */
  int32_t radius;
  int32_t glyph;        /* row of the atlas */
  int32_t org_x, org_y; /* putText origin (bottom-left of the text) */
  uint8_t bgr[3];
  uint8_t reserved;
} msp_mark;             /* 32 bytes */

MSP_API int msp_draw_marks(uint8_t* images, int32_t B, int32_t H, int32_t W, const msp_mark* marks, int64_t M,
                           const uint32_t* glyph_rows /* (G,32) or NULL */, const int32_t* glyph_meta /* (G,4) or NULL */,
                           int32_t G, void* stream);

/* JPEG glue for the same streams (cv2.imread / cv2.imwrite of DE-dot:221-229, DC-dot:333-346, VC-dot:357-397), over
 * nvJPEG, which is dlopen'ed on first use (the call fails with MSP_EINVAL when libnvjpeg.so.12 is absent; nothing else
 * in the library depends on it).  Streams are HOST buffers, images are device (H,W,3) BGR uint8.
 * decode: enqueued on `stream` (the output is valid after the stream reaches that point).
 * encode: 4:2:0, baseline Huffman; `host_out == NULL` only reports the size in *length; otherwise the call returns
 * after the bitstream is in host_out (it synchronises `stream`).  Encoded bytes are nvJPEG's, not libjpeg's: the
 * files differ from cv2.imwrite's in bytes, not in what they show. */
MSP_API int msp_jpeg_info(const uint8_t* host_jpeg, size_t length, int32_t* height, int32_t* width);
MSP_API int msp_jpeg_decode_bgr(const uint8_t* host_jpeg, size_t length, uint8_t* image_out, int32_t H, int32_t W,
                                void* stream);
MSP_API int msp_jpeg_encode_bgr(const uint8_t* image, int32_t H, int32_t W, int32_t quality, uint8_t* host_out,
                                size_t capacity, size_t* length, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MSP_H_ */

/* mobgs_b200 — C ABI of the B200-native MoBGS render + deblur hot path.
 *
 * Plain C, no torch / CUDA types in any signature: device pointers are `void*`-compatible
 * raw pointers, the stream is an opaque `void*` (a cudaStream_t), every struct is POD.
 *
 * What this replaces in the reference (KAIST-VICLab/MoBGS @ 0a1e0d3).  The reference is pure
 * Python; its only native boundary on this path is the gsplat==1.4.0 pip extension
 * (README.md:26) reached from gaussian_renderer/__init__.py:15.  Each entry point cites the
 * reference interface it stands in for; INTEGRATION.md shows the Python-side binding.
 *
 * Conventions
 *   - every function returns 0 on success or a negative MOBGS_E* code; the message is
 *     available (thread-local) from mobgs_last_error().  Nothing throws or exits.
 *   - the library never allocates or frees device memory and keeps no global state; all
 *     inputs, outputs and workspaces are caller-owned device buffers of the stated extent.
 *   - all kernels are launched on the passed stream and never synchronise the host.
 *   - fp32 everywhere, row-major, contiguous unless a stride is stated.
 *   - "K" = number of cameras / latent sub-frames of one launch (gsplat's C; MoBGS num_warp).
 *   - packed record = 16 floats per (sub-frame, Gaussian):
 *       [0]=mean2d.x [1]=mean2d.y [2]=opacity [3..5]=conic a,b,c [6..15]=up to 10 colour channels
 *     gradients use the same 16-float layout.
 */
#ifndef MOBGS_B200_H_
#define MOBGS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOBGS_OK 0
#define MOBGS_EINVAL (-1)   /* bad argument */
#define MOBGS_ECUDA (-2)    /* CUDA launch / runtime error */
#define MOBGS_ECAPACITY (-3) /* caller-provided workspace too small */

#define MOBGS_REC_FLOATS 16
#define MOBGS_MAX_COLORS 10
#define MOBGS_TILE 16
#define MOBGS_MAX_CTRL 12   /* GaussianModel.control_num, scene/gaussian_model.py:111 */

const char* mobgs_version(void);
const char* mobgs_last_error(void);

/* Cameras of one launch.  viewmats: [K,4,4] world->camera; Ks: [K,3,3] pinhole intrinsics.
 * Constants are gsplat.rendering.rasterization defaults (eps2d=0.3, near=0.01, far=1e10,
 * radius_clip=0) unless the caller overrides them. */
typedef struct {
  int32_t K;
  int32_t width, height;
  const float* viewmats;
  const float* Ks;
  float eps2d, near_plane, far_plane, radius_clip;
} MobgsCameras;

/* ------------------------------------------------------------------------------------------
 * gsplat.rendering.fully_fused_projection(means, covars=None, quats, scales, viewmats, Ks,
 * width, height)  — explicit call sites gaussian_renderer/__init__.py:190,411,422,513,524 and
 * the projection inside every rasterization() call.
 * Outputs (any may be NULL): radii i32[K,N], means2d[K,N,2], depths[K,N], conics[K,N,3].
 * Culled Gaussians get radii=0 and zeros. */
typedef struct {
  MobgsCameras cams;
  int32_t N;
  const float* means;   /* [N,3] */
  const float* quats;   /* [N,4] wxyz, normalised inside */
  const float* scales;  /* [N,3] */
  int32_t* radii;
  float* means2d;
  float* depths;
  float* conics;
} MobgsProjectFwd;
int mobgs_project_fwd(const MobgsProjectFwd* a, void* stream);

/* VJP of the above (gsplat fully_fused_projection_bwd).  Gradient inputs are addressed with an
 * element stride (in floats) per Gaussian so that views into a packed 16-float gradient record
 * can be consumed without a copy; the stride between sub-frames is N * stride.  v_* inputs may be
 * NULL (treated as zero).  Outputs v_means/v_quats/v_scales are *written* (sum over the K
 * cameras); v_viewmats [K,4,4] is *accumulated* with atomics and must be zeroed by the caller
 * (NULL = not needed). */
typedef struct {
  MobgsCameras cams;
  int32_t N;
  const float* means;
  const float* quats;
  const float* scales;
  const int32_t* radii;        /* [K,N] from the forward */
  const float* v_means2d; int32_t v_means2d_stride;   /* 2 floats per Gaussian */
  const float* v_depths;  int32_t v_depths_stride;    /* 1 float  */
  const float* v_conics;  int32_t v_conics_stride;    /* 3 floats */
  float* v_means;   /* [N,3] */
  float* v_quats;   /* [N,4] */
  float* v_scales;  /* [N,3] */
  float* v_viewmats;
} MobgsProjectBwd;
int mobgs_project_bwd(const MobgsProjectBwd* a, void* stream);

/* A launch of the binning / blend kernels renders K *lists*.  List k takes its Gaussians from
 * record set rec_k[k] (a sub-frame of the projection launch) restricted to the index range
 * [g_begin[k], g_end[k]).  Identity (rec_k[k] = k, full range) is the plain K-sub-frame render;
 * {(0, 0..N), (0, Ns..N), (0, 0..Ns)} renders the combined / dynamic-only / static-only images
 * of render() (gaussian_renderer/__init__.py:143,201,236) from ONE projection in ONE launch. */
#define MOBGS_MAX_K 32
typedef struct {
  int32_t rec_k[MOBGS_MAX_K];
  int32_t g_begin[MOBGS_MAX_K];
  int32_t g_end[MOBGS_MAX_K];
  /* Blend kernels only: list k walks the tile lists that binning produced for its list
   * tile_list[k] (identity = k).  Lists with the same geometry and index range but different
   * colour payloads (the mid-time flow renders of get_flow, gaussian_renderer/__init__.py:456)
   * are binned and sorted ONCE and share the result.  The binning entry points ignore it. */
  int32_t tile_list[MOBGS_MAX_K];
} MobgsLists;

/* ------------------------------------------------------------------------------------------
 * Fused MoBGS attribute synthesis + projection for K latent sub-frames (SURVEY.md §8 a1+a2+a4):
 *   static set  : GaussianModel getters scene/gaussian_model.py:209-257 (xyz, exp(scaling),
 *                 normalize(rotation), sigmoid(opacity), [f_dc | 0*f_t])
 *   dynamic set : interpolate_cubic_hermite gaussian_renderer/__init__.py:23-56 (x 1e-2),
 *                 get_rotation_dy (rotation + dt*omega), get_features(dt) = [f_dc | dt*f_t],
 *                 dt = t_poly[k] - trbf_center (detached)
 * Gaussian index g: [0,Ns) static, [Ns,Ns+Nd) dynamic — the torch.cat order of render():181-185.
 * Writes the packed record (colour channels 0..8 = features, channel 9 = camera depth: the
 * "RGB+ED" layout), radii and depths; optionally the world-space means [K,N,3]. */
typedef struct {
  int32_t Ns;
  const float* xyz;         /* [Ns,3] */
  const float* rotation;    /* [Ns,4] */
  const float* scaling;     /* [Ns,3] log-scale */
  const float* opacity;     /* [Ns]   logit */
  const float* features_dc; /* [Ns,6] */
} MobgsStaticParams;

typedef struct {
  int32_t Nd;
  int32_t n_ctrl_max;            /* control_xyz.shape[1] (12) */
  const float* control_xyz;      /* [Nd,n_ctrl_max,3], units 100x world */
  const int64_t* control_num;    /* [Nd] current_control_num (int64, as load_ply produces) */
  const float* rotation;         /* [Nd,4] */
  const float* omega;            /* [Nd,4] */
  const float* scaling;          /* [Nd,3] */
  const float* opacity;          /* [Nd] */
  const float* features_dc;      /* [Nd,6] */
  const float* features_t;       /* [Nd,3] */
  const float* trbf_center;      /* [Nd] */
  const float* offset;           /* [Nd,3] optional `coherent` offset added to the means, or NULL */
} MobgsDynamicParams;

typedef struct {
  MobgsCameras cams;
  MobgsStaticParams st;
  MobgsDynamicParams dy;
  const float* t_spline;   /* [K] device: spline time per sub-frame (already clamped where the reference clamps) */
  const float* t_poly;     /* [K] device: time used for dt = t - trbf_center */
  float* records;          /* [K,N,16] */
  int32_t* radii;          /* [K,N] */
  float* depths;           /* [K,N] */
  float* means3d;          /* [K,N,3] or NULL */
  /* Optional fused counting pass (bin_tile_counts != NULL): what mobgs_tile_count does for the bin_n_lists lists
   * `bin_lists` over this launch's record sets — per-tile counters and, with bin_entries, the recorded
   * (segment, slot, index, depth) entries — done here, while the projected Gaussian is still in registers, instead of
   * by a second kernel that re-reads records / radii / depths.  bin_tile_counts [bin_n_lists * T] and
   * bin_entry_cursor [1] are zeroed inside; afterwards call mobgs_tile_count with counts_ready = 1 (prefix sum only). */
  int32_t bin_n_lists, bin_tight;
  MobgsLists bin_lists;
  int32_t* bin_tile_counts;
  void* bin_entries;
  int64_t bin_entry_capacity;
  int32_t* bin_entry_cursor;
} MobgsSynthFwd;
int mobgs_synth_project_fwd(const MobgsSynthFwd* a, void* stream);

/* VJP: consumes the packed gradient records [K,N,16] written by mobgs_blend_bwd and writes the
 * parameter gradients (each written exactly once, summed over K; control_xyz gradient buffer
 * must be zeroed by the caller).  v_viewmats accumulated atomically (zeroed by caller) or NULL.
 * Pose-only mode: when EVERY parameter-gradient pointer (v_xyz .. v_offset) is NULL only v_viewmats is
 * produced — eval.py's test-time pose optimisation freezes all Gaussians (eval.py:246-255) and differentiates
 * render(w2c=...) with respect to the pose alone (eval.py:120-150). */
typedef struct {
  MobgsCameras cams;
  MobgsStaticParams st;
  MobgsDynamicParams dy;
  const float* t_spline;
  const float* t_poly;
  const int32_t* radii;     /* [K,N] */
  float* v_records;         /* [K,N,16]; read (and, with zero_v_records, zeroed behind the read) */
  float* v_xyz; float* v_rotation_s; float* v_scaling_s; float* v_opacity_s; float* v_features_dc_s;
  float* v_control_xyz; float* v_rotation_d; float* v_omega; float* v_scaling_d; float* v_opacity_d;
  float* v_features_dc_d; float* v_features_t; float* v_offset;
  float* v_viewmats;
  /* Optional Gaussian range [g_lo, g_hi) in the concatenated (static, dynamic) index space (0, 0 = all).  The
   * gradient pointers keep their meaning (row g of the FULL tensor), so a caller that stores a range's rows in
   * a private block passes block - g0 * row_floats.  Used to split the backward into chunks whose gradient
   * all-reduce (NCCL, side stream) overlaps the next chunk's kernel. */
  int32_t g_lo, g_hi;
  /* != 0: the gradient records of the Gaussian range are overwritten with zeros behind the reads (whole warps zero
   * their 2 KB chunks with coalesced stores), so the [K,N,16] buffer returns to the all-zero state mobgs_blend_bwd
   * needs and a caller that keeps it across steps never issues the 448 MB allocation + memset. */
  int32_t zero_v_records;
} MobgsSynthBwd;
int mobgs_synth_project_bwd(const MobgsSynthBwd* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * Pack gsplat-style SoA tensors into records for the blend kernels (the colour concat of
 * rasterization(render_mode="RGB+ED") is folded in: if depths != NULL it becomes channel D). */
typedef struct {
  int32_t K, N, D;            /* D colour channels in `colors` (<= 10, or <= 9 with depths) */
  const float* means2d;       /* [K,N,2] */
  const float* conics;        /* [K,N,3] */
  const float* opacities;     /* [N] (broadcast over K, like opacities.repeat(C,1)) */
  const float* colors;        /* [N,D] (colors_per_cam=0) or [K,N,D] (colors_per_cam=1) */
  int32_t colors_per_cam;
  const float* depths;        /* [K,N] or NULL */
  float* records;             /* [K,N,16] */
} MobgsPack;
int mobgs_pack_records(const MobgsPack* a, void* stream);


/* ------------------------------------------------------------------------------------------
 * Tile binning + per-tile depth sort (gsplat isect_tiles + radix sort + isect_offset_encode,
 * inside every rasterization()).  Tile lists are per (sub-frame, tile), sorted by
 * (depth, Gaussian index) ascending — the order gsplat's stable 64-bit sort produces.
 *
 * K = number of lists; records / radii / depths hold the record sets the lists refer to.
 * Step 1 counts intersections per tile and writes the exclusive prefix sum
 * tile_offsets[K*T+1] (T = tiles_x*tiles_y); the caller reads tile_offsets[K*T] (= I) to size
 * the lists (or provides a capacity it knows is enough).  Step 2 emits and sorts.
 * tight != 0 additionally drops (tile, Gaussian) pairs that provably reach alpha < 1/255 on
 * every pixel centre of the tile (results are unchanged; only the list shrinks). */
typedef struct {
  int32_t K, N, width, height;
  const float* records;      /* [K,N,16] */
  const int32_t* radii;      /* [K,N] */
  int32_t tight;
  MobgsLists lists;          /* K lists over the record sets (see MobgsLists) */
  int32_t* tile_counts;      /* [K*T] workspace, overwritten */
  int32_t* tile_offsets;     /* [K*T+1] out */
  /* Optional (entries != NULL): the counting pass also records every intersection it finds as a 16-byte entry
   * (segment, slot inside the segment, Gaussian index, depth bits), so that step 2 is a streaming scatter of these
   * entries (MobgsTileSort.entries) instead of a second pass over all K*N records.  entries [entry_capacity,4] u32,
   * entry_cursor [1] (zeroed inside; ends as the number of intersections I), depths [K,N].  Intersections beyond
   * entry_capacity are counted but not recorded: when tile_offsets[K*T] > entry_capacity the caller must run step 2
   * WITHOUT entries (the two-pass path). */
  const float* depths;
  void* entries;
  int64_t entry_capacity;
  int32_t* entry_cursor;
  /* != 0: tile_counts (and the entries) were already produced by mobgs_synth_project_fwd's fused counting pass
   * (MobgsSynthFwd.bin_tile_counts): only the prefix sum runs. */
  int32_t counts_ready;
} MobgsTileCount;
int mobgs_tile_count(const MobgsTileCount* a, void* stream);

typedef struct {
  int32_t K, N, width, height;
  const float* records;
  const int32_t* radii;
  const float* depths;       /* [K,N] sort key */
  int32_t tight;
  MobgsLists lists;
  const int32_t* tile_offsets; /* [K*T+1] from step 1 */
  int32_t* tile_cursor;      /* [K*T] workspace (zeroed inside) */
  int64_t capacity;          /* entries available in keys/keys_tmp/sorted_ids */
  uint64_t* keys;            /* [capacity] workspace */
  uint64_t* keys_tmp;        /* [capacity] workspace */
  int32_t* sorted_ids;       /* [capacity] out: Gaussian index per list entry */
  /* Optional (entries != NULL): emit from the entries recorded by mobgs_tile_count (see there) — records, radii,
   * depths and tile_cursor are then not read.  n_entries = MobgsTileCount.entry_cursor (device). */
  const void* entries;
  int64_t entry_capacity;
  const int32_t* n_entries;
} MobgsTileSort;
int mobgs_tile_emit_sort(const MobgsTileSort* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * Front-to-back alpha compositing (gsplat rasterize_to_pixels fwd), D colour channels:
 *   alpha = min(0.999, o * exp(-sigma)); skipped if sigma < 0 or alpha < 1/255; stops *before*
 *   blending once T(1-alpha) <= 1e-4; out = sum c alpha T + T_final * background.
 * out_colors [K,H,W,D], out_alphas [K,H,W], last_idx [K,H,W] (list position of the last blended
 * entry, needed by the backward). */
typedef struct {
  int32_t K, N, D, width, height;
  MobgsLists lists;
  int64_t list_capacity;      /* entries valid in sorted_ids (the `capacity` given to mobgs_tile_emit_sort) */
  const float* records;
  const int32_t* tile_offsets;
  const int32_t* sorted_ids;
  const float* backgrounds;   /* [K,D] or NULL */
  float* out_colors;
  float* out_alphas;
  int32_t* last_idx;
  /* Optional fused epilogue (dec_rays != NULL; D must be 10): the per-pixel work of
   * mobgs_decode_fwd — expected depth + Sandwich decoder — done while the pixel is still in
   * registers.  out_rgb [K,3,H,W], out_depth [K,H,W] (NULL = skip).
   * dec_rays_per_k: 0 = one shared [1,6,H,W]; 1 = one per list [K,6,H,W]; 2 = one per record set,
   * indexed by lists.rec_k[k] (lists that share a projection share its camera rays). */
  const float* dec_rays; int32_t dec_rays_per_k;
  const float* dec_w1; const float* dec_w2;
  float* out_rgb;
  float* out_depth;
  /* Optional fused flow channels (flow_ref >= 0; fused-epilogue launches only): two more channels are
   * composited in the same walk, with per-Gaussian colour records[flow_ref][g].xy - records[rec_k][g].xy
   * (the screen-space displacement between two projections of the same Gaussian — get_flow's
   * exp2mid rasterisation, gaussian_renderer/__init__.py:426-441, which shares geometry and order with
   * the latent image render :473) and no background.  out_flow [K,H,W,2]. */
  int32_t flow_ref;
  float* out_flow;
  /* Optional camera rays generated in registers (dec_pose != NULL; dec_rays may then be NULL): instead of reading
   * Camera.cam_ray [.,6,H,W] (348 MB per blurry view at 1080p, K = 7) the epilogue evaluates the ray of its pixel from
   * 12 pose floats per camera — dec_pose [n_cam,12] = camera-to-world rotation R (row-major 9) | camera centre c (3) —
   * exactly as mobgs_camera_rays_fwd does (scene/cameras.py:132-146): l = normalise(((x+.5-ppx)/sfx, (y+.5-ppy)/sfy, 1)),
   * ray = [c | normalise(R l)].  The camera of list k is chosen by dec_rays_per_k as for dec_rays. */
  const float* dec_pose;
  float dec_ppx, dec_ppy, dec_sfx, dec_sfy;
  /* Optional out [list_capacity] (16-lane-unit builds): the per-entry unit mask the kernel computes while staging — which
   * of the tile's sixteen 4x4-pixel units the Gaussian can reach with alpha >= 1/255 — kept for mobgs_blend_bwd, which
   * would otherwise recompute it for every entry it stages. */
  uint16_t* list_masks;
} MobgsBlendFwd;
int mobgs_blend_fwd(const MobgsBlendFwd* a, void* stream);

/* Back-to-front VJP (gsplat rasterize_to_pixels bwd).  v_records [K,N,16] is accumulated with
 * vector atomics and must be zeroed by the caller. */
typedef struct {
  int32_t K, N, D, width, height;
  MobgsLists lists;
  int64_t list_capacity;
  const float* records;
  const int32_t* tile_offsets;
  const int32_t* sorted_ids;
  const float* backgrounds;
  const float* out_alphas;
  const int32_t* last_idx;
  const float* v_out_colors;  /* [K,H,W,D] */
  const float* v_out_alphas;  /* [K,H,W] or NULL */
  float* v_records;           /* [record sets, N, 16], accumulated */
  int32_t sep_list;           /* list whose d loss / d means2d is ALSO accumulated into v_means2d_sep, or -1 */
  float* v_means2d_sep;       /* [N,2] zeroed by the caller (densification statistics), or NULL */
  /* Optional fused prologue (dec_rays != NULL; D must be 10): the VJP of the fused epilogue is
   * evaluated per pixel instead of reading v_out_colors / v_out_alphas (which may then be NULL).
   * out_colors = the forward's [K,H,W,10]; g_rgb [K,3,H,W], g_depth [K,H,W], g_alpha [K,H,W],
   * g_mean [3,H,W] (gradient of the K-sub-frame mean; scaled by 1/K inside) — each may be NULL.
   * v_rays as in MobgsDecodeBwd (NULL = not needed).  v_w_partial [MOBGS_DEC_SLOTS,90]: partial
   * sums of (v_w1 | v_w2), zeroed by the caller and summed over the first axis afterwards. */
  const float* dec_rays; int32_t dec_rays_per_k;
  const float* dec_w1; const float* dec_w2;
  const float* out_colors;
  const float* g_rgb; const float* g_depth; const float* g_alpha; const float* g_mean;
  int32_t mean_K;             /* the mean is over the first mean_K lists (0 = all K) */
  float* v_rays;
  float* v_w_partial;
  /* VJP of the fused flow channels: g_flow [K,H,W,2]; the colour gradient is added to
   * v_records[flow_ref][g].xy and subtracted from v_records[rec_k][g].xy. */
  int32_t flow_ref;
  const float* g_flow;
  /* Rays generated in registers (see MobgsBlendFwd.dec_pose).  v_pose_partial [n_cam, MOBGS_POSE_SLOTS, 12] (zeroed by
   * the caller, summed over the slot axis afterwards; NULL = pose gradient not needed) receives the gradient of the
   * 12 pose floats — what mobgs_camera_rays_bwd would have reduced from a [.,6,H,W] ray-gradient image. */
  const float* dec_pose;
  float dec_ppx, dec_ppy, dec_sfx, dec_sfy;
  float* v_pose_partial;
  /* Optional in: MobgsBlendFwd.list_masks of the forward over the same lists (NULL = recompute). */
  const uint16_t* list_masks;
  /* != 0 (fused-prologue launches): the caller expects most tiles to receive no gradient at all — train.py's SECOND
   * backward pass over a blurry view (loss.backward() at :680 after photo_loss.backward(retain_graph=True) at :629)
   * reaches only the centre render's depth / alpha / image, i.e. 2 of its K + 2 lists.  Every CTA then first tests
   * its pixels' g_mean / g_rgb / g_depth / g_alpha / g_flow and leaves before the decoder prologue when all are zero
   * (the VJP is linear in them, so the result is unchanged).  With per-list v_rays (dec_rays_per_k == 1) the caller
   * must then zero v_rays itself. */
  int32_t sparse_grads;
} MobgsBlendBwd;

#define MOBGS_DEC_SLOTS 1024
#define MOBGS_POSE_SLOTS 64

/* mean[i] = (1/K) sum_k rgb[k][i] + 1e-10 for i < n  (train.py:540-541). */
int mobgs_subframe_mean(const float* rgb, float* mean, int32_t K, int64_t n, void* stream);
int mobgs_blend_bwd(const MobgsBlendBwd* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * Per-pixel epilogue of one blurry view (SURVEY.md §8 a8 + a9 + the "ED" division of a6):
 *   depth  = img[9] / max(alpha, 1e-10)                 gsplat render_mode="RGB+ED"
 *   rgb    = sigmoid(albedo + W2 relu(W1 [spec, timefeat, rays]))   helper_model.py:19-28 (Sandwich,
 *            bias-free 1x1 convs: W1 [6,12] = mlp1.weight, W2 [3,6] = mlp2.weight)
 *   mean   = (1/K) sum_k rgb_k + 1e-10                  train.py:540-541 (blur model)
 * img [K,H,W,10] and alpha [K,H,W] are the blend outputs; rays [K,6,H,W] (rays_per_k=1) or
 * [1,6,H,W] shared.  Outputs: rgb [K,3,H,W] (required), depth [K,H,W] and mean [3,H,W] (may be NULL). */
typedef struct {
  int32_t K, width, height;
  const float* img;
  const float* alpha;
  const float* rays; int32_t rays_per_k;
  const float* w1;   /* [6,12] */
  const float* w2;   /* [3,6]  */
  float* rgb;
  float* depth;
  float* mean;
} MobgsDecodeFwd;
int mobgs_decode_fwd(const MobgsDecodeFwd* a, void* stream);

/* VJP.  g_rgb / g_depth / g_mean may be NULL.  v_img [K,H,W,10] and v_alpha [K,H,W] are written;
 * v_rays (same shape as rays; NULL = not needed) is written when rays_per_k=1 and accumulated
 * atomically (zeroed by the caller) when shared; v_w1 / v_w2 are accumulated atomically and
 * must be zeroed by the caller. */
typedef struct {
  int32_t K, width, height;
  const float* img;
  const float* alpha;
  const float* rays; int32_t rays_per_k;
  const float* w1;
  const float* w2;
  const float* g_rgb;
  const float* g_depth;
  const float* g_mean;
  float* v_img;
  float* v_alpha;
  float* v_rays;
  float* v_w1;
  float* v_w2;
} MobgsDecodeBwd;
int mobgs_decode_bwd(const MobgsDecodeBwd* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * deform_network.forward(point, scales, rotations, times_sel) — scene/deformation.py:252-253 ->
 * forward_dynamic2 :158-199 over HexPlaneField scene/hexplane.py:19-108 (not called by render(),
 * but part of the model API).  One fused launch: 6*levels bilinear plane
 * samples (align_corners=True, border), product over planes, Linear(32*levels -> 128), three
 * heads ReLU-Linear(128,128)-ReLU-Linear(128,{7,3,4}) on tcgen05 tensor cores (3xTF32), and the
 * dx / quat2mat / ds-clamp / quaternion-product epilogue.
 *   planes[l*6+p]: channels-LAST copy [H][W][32] of grids[l][p] ([1,32,H,W] in the reference);
 *                  plane order = itertools.combinations(range(4), 2); plane_w/h = W/H of each.
 *   aabb[0..2] = HexPlaneField.aabb[0], aabb[3..5] = aabb[1] (normalize_aabb, hexplane.py:19-21).
 *   Weights are pre-tiled by the host into the kernel's shared-memory operand layout, each block
 *   stored twice (tf32 "hi" part, then fp32 remainder "lo"): a [rows x K] row-major matrix becomes
 *   [K/4][rows][4] floats.
 *     w0: feature_out[0].weight [128, 32*levels] as 2 row-halves x {hi,lo} x [K/4][64][4]
 *     wa: the three heads' first Linear [128,128]: [3] x 2 row-halves x {hi,lo} x [32][64][4]
 *     wb: the three heads' last Linear, rows zero-padded to 16: [3] x {hi,lo} x [32][16][4]
 *     b0 [128], ba [3,128], bb [3,16] (zero padded) plain fp32.
 *   head order: pos_deform (7 outputs), scales_deform (3), rotations_deform (4). */
typedef struct {
  int32_t N;
  const float* pts;      /* [N,3] */
  const float* scales;   /* [N,3] */
  const float* rots;     /* [N,4] */
  const float* times;    /* [N]   */
  float aabb[6];
  int32_t levels, net_width, plane_features;
  const float* planes[24];
  int32_t plane_w[24];
  int32_t plane_h[24];
  const float* w0; const float* b0;
  const float* wa; const float* ba;
  const float* wb; const float* bb;
  float* out_pts;     /* [N,3] */
  float* out_scales;  /* [N,3] */
  float* out_rots;    /* [N,4] */
} MobgsHexMlpFwd;
int mobgs_hexplane_mlp_fwd(const MobgsHexMlpFwd* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * a11 backward on the tensor cores (the reference differentiates deform_network through autograd:
 * scene/deformation.py:158-199 + scene/hexplane.py:19-108; 18 grid_sample_backward + ~16 cuBLAS launches).
 *
 * Step 1, mobgs_hexplane_mlp_bwd — one CTA per 128 points, like the forward: recomputes the forward
 * (features, three heads) on tcgen05, evaluates the post-processing VJP per point, then runs the DATA
 * gradient chain  g_o -> (Wb) -> relu' -> (Wa) -> sum over heads -> relu' -> (W0) -> g_feat  as 3xTF32
 * tcgen05 GEMMs against pre-tiled TRANSPOSED weights, with the accumulators in TMEM.  Besides the input
 * gradients it leaves, feature-major ([rows][ld], ld = padded point count, zero for points >= N), the
 * operands the weight gradients contract over the points.
 * Step 2, mobgs_hexplane_wgrad — C[m][n] += sum_k A[m][k] B[n][k] with k = the points: split-K tcgen05
 * GEMMs (K-major on both sides thanks to the feature-major layout), accumulators resident in TMEM over a
 * CTA's whole K range, one atomic flush per CTA; row sums of one operand = the bias gradient.
 * Step 3 is the existing mobgs_hexplane_features_bwd (plane scatter + coordinate gradients) on g_feat. */
typedef struct {
  int32_t N;
  int32_t ld;                /* multiple of 128, >= N */
  const float* pts;          /* [N,3] */
  const float* rots;         /* [N,4] */
  const float* times;        /* [N] */
  float aabb[6];
  int32_t levels, net_width, plane_features;
  const float* planes[24];   /* channels-last [H,W,32] as MobgsHexMlpFwd */
  int32_t plane_w[24];
  int32_t plane_h[24];
  const float* w0; const float* b0; const float* wa; const float* ba; const float* wb; const float* bb;   /* as MobgsHexMlpFwd */
  /* transposed weights, tiled {hi,lo} x [K/4][rows][4] like the forward's:
   *   w0_t: W0^T rows [0,64) then rows [64, 32*levels), K = 128
   *   wa_t: per head, Wa^T rows [0,64) then [64,128), K = 128
   *   wb_t: per head, (Wb zero-padded to 16 outputs)^T rows [0,64) then [64,128), K = 16 */
  const float* w0_t; const float* wa_t; const float* wb_t;
  const float* g_out_pts;    /* [N,3] or NULL (= zero) */
  const float* g_out_scales; /* [N,3] or NULL */
  const float* g_out_rots;   /* [N,4] or NULL */
  float* g_pts;              /* [N,3] d loss / d pts through out_pts = R (pts + dx) only (the grid path is step 3) */
  float* g_scales;           /* [N,3] */
  float* g_rots;             /* [N,4] */
  float* g_feat;             /* [N, 32*levels] */
  float* featT;              /* [32*levels][ld] */
  float* a1T;                /* [128][ld]     relu(h0)                       */
  float* a2T;                /* [3][128][ld]  relu(z1) per head              */
  float* gz1T;               /* [3][128][ld]  d loss / d z1 per head         */
  float* goT;                /* [3][16][ld]   d loss / d head output (padded)*/
  float* gh0T;               /* [128][ld]     d loss / d h0                  */
} MobgsHexMlpBwd;
int mobgs_hexplane_mlp_bwd(const MobgsHexMlpBwd* a, void* stream);

#define MOBGS_WGRAD_MAX 8
typedef struct {
  int32_t n_problems;
  int32_t ld;                          /* contraction length (padded points), multiple of 128 */
  const float* A[MOBGS_WGRAD_MAX];     /* [128][ld] */
  const float* B[MOBGS_WGRAD_MAX];     /* [n_cols][ld] */
  int32_t n_cols[MOBGS_WGRAD_MAX];     /* 16..128, multiple of 16 */
  float* C[MOBGS_WGRAD_MAX];           /* accumulated (zeroed by the caller): [128][ldc], or [n_cols][ldc] when transposed */
  int32_t ldc[MOBGS_WGRAD_MAX];
  int32_t transpose_out[MOBGS_WGRAD_MAX];
  float* bias[MOBGS_WGRAD_MAX];        /* row sums of A (bias_from = 1: [128]) or B (2: [n_cols]), accumulated; NULL / 0 = none */
  int32_t bias_from[MOBGS_WGRAD_MAX];
} MobgsHexWgrad;
int mobgs_hexplane_wgrad(const MobgsHexWgrad* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * a11: device-side packing of the operands MobgsHexMlpFwd / MobgsHexMlpBwd describe above (tiled {hi,lo} weights,
 * zero-padded biases, channels-last planes) — one table-driven launch instead of ~200 torch ops per optimiser step.
 * Every job reads a logical [rows x cols] matrix m[r][c] = src[r*row_stride + c*col_stride] (zero outside
 * valid_rows x valid_cols: padding; strides express the transposes of the backward's W^T operands) and writes
 *   MOBGS_PACK_TILED:     dst[(k4*rows + r)*4 + q] = hi(m[r][4 k4 + q]) followed, rows*cols floats later, by the lo
 *                         block (hi = fp32 bits & 0xffffe000 — the tf32 part —, lo = m - hi); cols % 4 == 0
 *   MOBGS_PACK_PLAIN:     dst[r*cols + c] = m[r][c]
 *   MOBGS_PACK_TRANSPOSE: dst[c*rows + r] = m[r][c]          ([32][H*W] plane -> channels-last [H*W][32])
 * jobs / chunk_begin live in DEVICE memory (they change only when a parameter tensor is re-allocated);
 * chunk_begin[j] = first chunk of job j when every job's rows*cols elements are cut into
 * mobgs_pack_chunk_elems()-element chunks, chunk_begin[n_jobs] = n_chunks. */
#define MOBGS_PACK_MAX_JOBS 96
#define MOBGS_PACK_TILED 0
#define MOBGS_PACK_PLAIN 1
#define MOBGS_PACK_TRANSPOSE 2
typedef struct {
  const float* src;
  int64_t row_stride, col_stride;
  int32_t rows, cols;
  int32_t valid_rows, valid_cols;
  float* dst;
  int32_t kind;
  int32_t reserved_;
} MobgsPackJob;
typedef struct {
  int32_t n_jobs;
  int32_t n_chunks;
  const MobgsPackJob* jobs;       /* [n_jobs] device */
  const int32_t* chunk_begin;     /* [n_jobs + 1] device */
} MobgsPackOperands;
int mobgs_pack_operands(const MobgsPackOperands* a, void* stream);
int mobgs_pack_chunk_elems(void);

/* ------------------------------------------------------------------------------------------
 * Flow records for get_flow() (gaussian_renderer/__init__.py:435-471), K exposure offsets at once.
 * records [K+1,N,16]: set 0 = geometry at the mid time, sets 1..K = geometry at the K exposure
 * times (same camera).  For every k two record sets are written to flow_records [2K,N,16]:
 *   set 2k   : geometry of exposure k, colours (c0,c1) = means2d_mid - means2d_exp_k   (exp2mid flow)
 *   set 2k+1 : geometry of mid,        colours (c0,c1) = means2d_exp_k - means2d_mid   (mid2exp flow)
 * (the reference builds them with two fully_fused_projection calls + tensor arithmetic per k).
 * The VJP accumulates v_flow_records back into v_records [K+1,N,16] (written, not accumulated). */
typedef struct {
  int32_t K, N;
  const float* records;
  float* flow_records;
} MobgsFlowRecFwd;
int mobgs_flow_records_fwd(const MobgsFlowRecFwd* a, void* stream);

typedef struct {
  int32_t K, N;
  const float* v_flow_records;   /* [2K,N,16] */
  float* v_records;              /* [K+1,N,16] */
  int32_t accumulate;            /* mobgs_midflow_records_bwd only: 0 = write v_records, 1 = add to it (the buffer
                                  * already holds the gradient of the other renders that share the projection) */
} MobgsFlowRecBwd;
int mobgs_flow_records_bwd(const MobgsFlowRecBwd* a, void* stream);

/* Mid-time flow records: the K mid2exp rasterisations of get_flow (gaussian_renderer/__init__.py:444-457,
 * one per exposure offset in train.py:563-579) share the mid-time geometry and differ only in their two
 * colour channels.  fwd packs all 2K channels, ten per record set, over the mid geometry:
 * flow_records [M,N,16], M = ceil(2K/10); channel j = 2k + a (a: 0 = x, 1 = y) = means2d_exp_k[a] -
 * means2d_mid[a] sits in set j/10, colour slot j%10 (unused slots 0).  The M sets are rendered as M lists
 * that share ONE tile binning (MobgsLists.tile_list).  bwd (MobgsFlowRecBwd with v_flow_records [M,N,16])
 * writes v_records [K+1,N,16]. */
int mobgs_midflow_records_fwd(const MobgsFlowRecFwd* a, void* stream);
int mobgs_midflow_records_bwd(const MobgsFlowRecBwd* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * HexPlane feature gather and its VJP as stand-alone kernels (the first stage of a11, used by the
 * backward of the deformation module: scene/hexplane.py:19-108 interpolate_ms_features).
 *   fwd: feat[n, l*32 + c] = prod_p bilinear(planes[l*6+p], coords_p(n))[c]
 *   bwd: g_feat -> g_planes (channels-last, accumulated with vector atomics; zeroed by the caller),
 *        g_pts [N,3] and g_times [N] (written; zero where normalize_aabb / the border clamp clipped).
 * planes / plane_w / plane_h / aabb as in MobgsHexMlpFwd. */
typedef struct {
  int32_t N;
  const float* pts;     /* [N,3] */
  const float* times;   /* [N]   */
  float aabb[6];
  int32_t levels;
  const float* planes[24];
  int32_t plane_w[24];
  int32_t plane_h[24];
  float* feat;           /* fwd out [N, 32*levels] */
  const float* g_feat;   /* bwd in  [N, 32*levels] */
  float* g_planes[24];   /* bwd out, same shapes as planes */
  float* g_pts;          /* bwd out [N,3] */
  float* g_times;        /* bwd out [N]   */
} MobgsHexFeat;
int mobgs_hexplane_features_fwd(const MobgsHexFeat* a, void* stream);
int mobgs_hexplane_features_bwd(const MobgsHexFeat* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * f4: multi-tensor Adam step (replaces the two `torch.optim.Adam(l, lr=0.0, eps=1e-15).step()` calls
 * of train.py:796-800 over the parameter groups of scene/gaussian_model.py:598-641; amsgrad=False,
 * weight_decay=0).  One launch updates up to MOBGS_ADAM_MAX_TENSORS fp32 tensors in place:
 *   m = m + (g - m) one_minus_beta1;  v = v beta2 + one_minus_beta2 g g;
 *   p -= step_size[t] * m / (sqrt(v) / bc2_sqrt[t] + eps)
 * with step_size[t] = lr_t / (1 - beta1^step_t), bc2_sqrt[t] = sqrt(1 - beta2^step_t) computed by the
 * caller.  chunk_begin[t] = sum_{u<t} ceil(numel[u] / mobgs_adam_chunk_elems()), chunk_begin[n_tensors]
 * = total (the kernel's work list).  Tensors must be contiguous; 16-byte alignment enables the
 * vector path. */
#define MOBGS_ADAM_MAX_TENSORS 64
typedef struct {
  int32_t n_tensors;
  float beta1, beta2, eps;
  float one_minus_beta1, one_minus_beta2;   /* rounded from double by the caller, as torch passes them */
  int32_t reserved_;
  float* param[MOBGS_ADAM_MAX_TENSORS];
  const float* grad[MOBGS_ADAM_MAX_TENSORS];
  float* exp_avg[MOBGS_ADAM_MAX_TENSORS];
  float* exp_avg_sq[MOBGS_ADAM_MAX_TENSORS];
  int64_t numel[MOBGS_ADAM_MAX_TENSORS];
  float step_size[MOBGS_ADAM_MAX_TENSORS];
  float bc2_sqrt[MOBGS_ADAM_MAX_TENSORS];
  int32_t chunk_begin[MOBGS_ADAM_MAX_TENSORS + 1];
} MobgsAdam;
int mobgs_adam_step(const MobgsAdam* a, void* stream);
int mobgs_adam_chunk_elems(void);

/* ------------------------------------------------------------------------------------------
 * f2: photometric loss of the training step, forward + backward fused (train.py:621-628):
 *   photo_loss = l1_loss(image, gt) + lambda_dssim (1 - ssim(image, gt))
 * with utils/loss_utils.py:233-239 l1_loss (mask=None) and :351-382 ssim (11x11 Gaussian window, sigma
 * 1.5, zero padding, per channel, mean over everything).  img / gt: [planes, H, W] fp32 contiguous
 * (planes = batch x channels).  window: the 11 normalised 1-D taps (the 2-D window is their outer product).
 *   fwd: sums[0] = sum |x - y|, sums[1] = sum ssim_map (double, zeroed inside the call); d_mu1 / d_x2 /
 *        d_xy [planes,H,W] receive the per-pixel partials of the SSIM map (all NULL = forward only).
 *   bwd: v_img = v_loss[0] * ( scale_l1 sign(x - y) + scale_ssim * dSSIMsum/dx ), v_loss a device scalar
 *        (NULL = 1); the caller passes scale_l1 = 1/numel, scale_ssim = -lambda_dssim/numel. */
typedef struct {
  int32_t planes, H, W;
  const float* img;
  const float* gt;
  float window[11];
  double* sums;      /* [2] */
  float* d_mu1;
  float* d_x2;
  float* d_xy;
} MobgsPhotoLossFwd;
int mobgs_photo_loss_fwd(const MobgsPhotoLossFwd* a, void* stream);

typedef struct {
  int32_t planes, H, W;
  const float* img;
  const float* gt;
  float window[11];
  const float* d_mu1;
  const float* d_x2;
  const float* d_xy;
  const float* v_loss;   /* device scalar or NULL */
  float scale_l1, scale_ssim;
  float* v_img;          /* [planes,H,W] written */
} MobgsPhotoLossBwd;
int mobgs_photo_loss_bwd(const MobgsPhotoLossBwd* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * f3: Camera.cam_ray of K cameras (scene/cameras.py:132-146, :244-284): per pixel p = (x, y)
 *   l = normalise(((x + 0.5 - ppx) / sfx, (y + 0.5 - ppy) / sfy, 1)),  d = normalise(rot[k] l)
 *   rays[k] = [centre[k] broadcast (3 planes) | d (3 planes)]            rays [K,6,H,W]
 * rot [K,3,3] = camera-to-world rotation (Camera.R), centre [K,3] = camera centre, (ppx, ppy, sfx, sfy) =
 * principal point and scale factors of the camera metadata (dycheck_geometry/camera.py).
 *   bwd: v_rays [K,6,H,W] -> v_rot [K,3,3], v_centre [K,3] (zeroed inside the call). */
typedef struct {
  int32_t K, H, W;
  float ppx, ppy, sfx, sfy;
  const float* rot;
  const float* centre;
  float* rays;            /* fwd out */
  const float* v_rays;    /* bwd in  */
  float* v_rot;           /* bwd out */
  float* v_centre;        /* bwd out */
} MobgsCameraRays;
int mobgs_camera_rays_fwd(const MobgsCameraRays* a, void* stream);
int mobgs_camera_rays_bwd(const MobgsCameraRays* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * f4: densify / prune gather-compaction (scene/gaussian_model.py:1044-1123 `_prune_optimizer`,
 * `cat_tensors_to_optimizer`; densification at :1417-1434): every per-Gaussian tensor — parameters, both
 * Adam moments, densification statistics — is rebuilt by the same row gather, in ONE launch:
 *   dst[t][i] = src[t][idx[i]]              if idx[i] <  n_old
 *             = ext[t][idx[i] - n_old]      if idx[i] >= n_old      (ext[t] == NULL -> zero row)
 * for i < n_out.  prune: idx = kept rows, ascending; append (clone / split): idx = 0 .. n_old + n_new - 1,
 * ext = the new rows (NULL for the Adam moments: the reference extends them with zeros).  Rows are moved as
 * row_words[t] 32-bit words (fp32 rows and the int64 current_control_num rows alike) — bit-exact.
 * chunk_begin[t] = first chunk of tensor t when every tensor's n_out * row_words[t] output words are cut into
 * mobgs_compact_chunk_words()-word chunks; chunk_begin[n_tensors] = total. */
#define MOBGS_COMPACT_MAX_TENSORS 64
typedef struct {
  int32_t n_tensors;
  int32_t reserved_;
  int64_t n_old;             /* rows of every src[t] */
  int64_t n_out;             /* rows of every dst[t] = entries of idx */
  const int64_t* idx;        /* [n_out] device */
  const void* src[MOBGS_COMPACT_MAX_TENSORS];
  const void* ext[MOBGS_COMPACT_MAX_TENSORS];
  void* dst[MOBGS_COMPACT_MAX_TENSORS];
  int32_t row_words[MOBGS_COMPACT_MAX_TENSORS];
  int32_t chunk_begin[MOBGS_COMPACT_MAX_TENSORS + 1];
} MobgsCompactRows;
int mobgs_compact_rows(const MobgsCompactRows* a, void* stream);
int mobgs_compact_chunk_words(void);

/* dst[i][0..n_words[i]) = src[i][0..n_words[i]) for up to 64 segments in one launch (32-bit words): unpacks the
 * chunk-major gradient buffer of the overlapped data-parallel backward into the per-parameter gradient tensors. */
#define MOBGS_COPY_MAX_SEGMENTS 64
typedef struct {
  int32_t n_segments;
  int32_t reserved_;
  const void* src[MOBGS_COPY_MAX_SEGMENTS];
  void* dst[MOBGS_COPY_MAX_SEGMENTS];
  int64_t n_words[MOBGS_COPY_MAX_SEGMENTS];
  int32_t chunk_begin[MOBGS_COPY_MAX_SEGMENTS + 1];   /* in chunks of mobgs_compact_chunk_words() words */
} MobgsCopySegments;
int mobgs_copy_segments(const MobgsCopySegments* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * f1 (second half): flow-warp loss of train.py:656-676, forward + backward fused.
 *   term1 = l1_loss(grid_sample(ori (expanded over K), norm(exp2mid)), latent, mask = latent_alpha)
 *   term2 = l1_loss(grid_sample(latent, norm(mid2exp)), ori (expanded), mask = d_alpha (expanded))
 * norm(c) = 2 c / (size - 1) - 1; grid_sample bilinear, padding_mode='border', align_corners=False (the
 * call's default); masked l1_loss as utils/loss_utils.py:233-237.  Layouts (fp32, contiguous):
 * ori [B,3,H,W], latent [B,K,3,H,W], exp2mid / mid2exp [B,K,H,W,2] (x, y in pixels), latent_alpha
 * [B,K,1,H,W], d_alpha [B,1,H,W].
 *   fwd: sums[0..3] = (num1, den1, num2, den2) in double (zeroed inside); loss = num1/(den1+1e-8) + num2/(den2+1e-8)
 *   bwd: needs the sums of the forward; v_loss = device scalar (NULL = 1); writes v_exp2mid, v_mid2exp,
 *        v_latent_alpha, v_d_alpha and accumulates v_latent (zeroed inside).  v_ori (optional, NULL =
 *        skip; zeroed inside) receives d loss / d ori: in the reference ori_image_tensor is the live
 *        centre render (train.py:469, :607), so the loss reaches it through the exp2mid grid_sample
 *        (bilinear scatter) and as the L1 target of the mid2exp term. */
typedef struct {
  int32_t B, K, H, W;
  const float* ori;
  const float* latent;
  const float* exp2mid;
  const float* mid2exp;
  const float* latent_alpha;
  const float* d_alpha;
  double* sums;              /* [4] fwd out / bwd in */
  const float* v_loss;
  float* v_latent;
  float* v_exp2mid;
  float* v_mid2exp;
  float* v_latent_alpha;
  float* v_d_alpha;
  float* v_ori;              /* [B,3,H,W] or NULL */
} MobgsFlowWarp;
int mobgs_flow_warp_loss_fwd(const MobgsFlowWarp* a, void* stream);
int mobgs_flow_warp_loss_bwd(const MobgsFlowWarp* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * f2 (remainder): regularisers of train.py:651-655 in one pass.
 *   sums[0] = sum |depth - gt_depth|           (l1_loss numerator, utils/loss_utils.py:233-239; mean = / n_depth)
 *   sums[1] = entropy_loss(alpha)              (:264-276, eps 1e-6)
 *   sums[2] = sparsity_loss(alpha) = sum alpha^2  (:285-295)
 * g_depth [n_depth] = sign(depth - gt_depth), g_alpha [n_alpha] = d(entropy + sparsity)/d alpha — unscaled
 * gradient maps (both optional); sums are double, zeroed inside the call. */
typedef struct {
  int64_t n_depth, n_alpha;
  const float* depth;
  const float* gt_depth;
  const float* alpha;
  double* sums;        /* [3] */
  float* g_depth;
  float* g_alpha;
} MobgsRegLoss;
int mobgs_reg_loss_fwd(const MobgsRegLoss* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * K9 / f4: simple_knn._C.distCUDA2(points [N,3]) -> [N]: mean squared distance to the 3 nearest
 * neighbours, the point itself excluded (scene/gaussian_model.py:420, :514).  Tiled brute force, O(N^2):
 * for the initialisation point clouds. */
int mobgs_knn3_mean_dist2(const float* points, float* out, int32_t n, void* stream);

/* ------------------------------------------------------------------------------------------
 * main_utils.get_normals (main_utils.py:95-141; called once per view and step at train.py:590 on the centre render's
 * depth): camera-space normals from a depth map.  With the local view direction of pixel (u, v)
 *   y = (v + pixel_offset - ppy) / sfy,  x = (u + pixel_offset - ppx - y skew) / sfx,  c(u, v) = (x, y, 1) z(u, v)
 * (dycheck_geometry Camera.get_pixels / principal_point / scale_factor / skew; pixel_offset = 0.5 when use_center),
 *   n = normalise((c(u+1,v) - c(u-1,v)) x (c(u,v-1) - c(u,v+1)))   (F.normalize, eps 1e-12)
 * for interior pixels and 0 on the one-pixel border.  z [B,H,W] -> normals [B,3,H,W].  Forward only: the reference
 * never uses the result in a loss (train.py:615). */
typedef struct {
  int32_t B, width, height;
  const float* z;
  float ppx, ppy, sfx, sfy, skew, pixel_offset;
  float* normals;
} MobgsNormals;
int mobgs_depth_normals(const MobgsNormals* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MOBGS_B200_H_ */

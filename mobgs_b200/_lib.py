"""ctypes binding of libmobgs_b200.so (include/mobgs_b200.h) + the in-tree nvcc build.

The library is the product: there is no Python/torch fallback.  If the shared object is missing
or a symbol cannot be resolved, importing callers get a RuntimeError — nothing silently routes
around the CUDA path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libmobgs_b200.so")
CSRC = os.path.join(_HERE, "csrc")
SOURCES = ["capi.cu", "synth_project.cu", "bin_sort.cu", "blend.cu", "decode.cu", "hexplane_mlp.cu", "hexplane_mlp_bwd.cu", "flow_records.cu", "hexplane_grid.cu", "adam.cu", "photo_loss.cu", "camera_rays.cu", "flow_warp_loss.cu", "reg_loss.cu", "knn.cu", "compact.cu", "normals.cu", "hexplane_pack.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-extended-lambda", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-shared",
    "--threads", "0",      # one compile job per source file: 39 s -> 10 s on the 8-core build box
]

MAX_K = 32
REC = 16
MAX_COLORS = 10
TILE = 16

c_f32p = C.POINTER(C.c_float)
c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_u64p = C.POINTER(C.c_uint64)


def _sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under mobgs_b200/csrc into libmobgs_b200.so for sm_100a (cross-compiles
    without a GPU).  Rebuilds only when a source / header is newer than the library."""
    srcs = _sources()
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(_ROOT, "include", "mobgs_b200.h"))
    if not force and os.path.exists(LIB_PATH):
        t = os.path.getmtime(LIB_PATH)
        if all(os.path.getmtime(d) <= t for d in deps):
            return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("MOBGS_NVCC_EXTRA", "").split()     # e.g. -DMOBGS_FWD_MIN_CTAS=6 (tuning experiments)
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-o", LIB_PATH] + srcs
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB_PATH


class Cameras(C.Structure):
    _fields_ = [("K", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("viewmats", C.c_void_p), ("Ks", C.c_void_p),
                ("eps2d", C.c_float), ("near_plane", C.c_float), ("far_plane", C.c_float),
                ("radius_clip", C.c_float)]


class ProjectFwd(C.Structure):
    _fields_ = [("cams", Cameras), ("N", C.c_int32), ("means", C.c_void_p), ("quats", C.c_void_p),
                ("scales", C.c_void_p), ("radii", C.c_void_p), ("means2d", C.c_void_p),
                ("depths", C.c_void_p), ("conics", C.c_void_p)]


class ProjectBwd(C.Structure):
    _fields_ = [("cams", Cameras), ("N", C.c_int32), ("means", C.c_void_p), ("quats", C.c_void_p),
                ("scales", C.c_void_p), ("radii", C.c_void_p),
                ("v_means2d", C.c_void_p), ("v_means2d_stride", C.c_int32),
                ("v_depths", C.c_void_p), ("v_depths_stride", C.c_int32),
                ("v_conics", C.c_void_p), ("v_conics_stride", C.c_int32),
                ("v_means", C.c_void_p), ("v_quats", C.c_void_p), ("v_scales", C.c_void_p),
                ("v_viewmats", C.c_void_p)]


class StaticParams(C.Structure):
    _fields_ = [("Ns", C.c_int32), ("xyz", C.c_void_p), ("rotation", C.c_void_p),
                ("scaling", C.c_void_p), ("opacity", C.c_void_p), ("features_dc", C.c_void_p)]


class DynamicParams(C.Structure):
    _fields_ = [("Nd", C.c_int32), ("n_ctrl_max", C.c_int32), ("control_xyz", C.c_void_p),
                ("control_num", C.c_void_p), ("rotation", C.c_void_p), ("omega", C.c_void_p),
                ("scaling", C.c_void_p), ("opacity", C.c_void_p), ("features_dc", C.c_void_p),
                ("features_t", C.c_void_p), ("trbf_center", C.c_void_p), ("offset", C.c_void_p)]


class Lists(C.Structure):
    _fields_ = [("rec_k", C.c_int32 * MAX_K), ("g_begin", C.c_int32 * MAX_K), ("g_end", C.c_int32 * MAX_K),
                ("tile_list", C.c_int32 * MAX_K)]


def make_lists(specs, tile_list=None):
    """specs: [(record_set, g_begin, g_end)] per list; tile_list[k] = index of the binned tile lists list k
    walks (default identity)."""
    if len(specs) > MAX_K:
        raise RuntimeError(f"at most {MAX_K} lists per launch, got {len(specs)}")
    l = Lists()
    for i, (rk, g0, g1) in enumerate(specs):
        l.rec_k[i], l.g_begin[i], l.g_end[i] = int(rk), int(g0), int(g1)
        l.tile_list[i] = i if tile_list is None else int(tile_list[i])
    return l


class SynthFwd(C.Structure):
    _fields_ = [("cams", Cameras), ("st", StaticParams), ("dy", DynamicParams),
                ("t_spline", C.c_void_p), ("t_poly", C.c_void_p), ("records", C.c_void_p),
                ("radii", C.c_void_p), ("depths", C.c_void_p), ("means3d", C.c_void_p),
                ("bin_n_lists", C.c_int32), ("bin_tight", C.c_int32), ("bin_lists", Lists),
                ("bin_tile_counts", C.c_void_p), ("bin_entries", C.c_void_p), ("bin_entry_capacity", C.c_int64),
                ("bin_entry_cursor", C.c_void_p)]


class SynthBwd(C.Structure):
    _fields_ = [("cams", Cameras), ("st", StaticParams), ("dy", DynamicParams),
                ("t_spline", C.c_void_p), ("t_poly", C.c_void_p), ("radii", C.c_void_p),
                ("v_records", C.c_void_p),
                ("v_xyz", C.c_void_p), ("v_rotation_s", C.c_void_p), ("v_scaling_s", C.c_void_p),
                ("v_opacity_s", C.c_void_p), ("v_features_dc_s", C.c_void_p),
                ("v_control_xyz", C.c_void_p), ("v_rotation_d", C.c_void_p), ("v_omega", C.c_void_p),
                ("v_scaling_d", C.c_void_p), ("v_opacity_d", C.c_void_p),
                ("v_features_dc_d", C.c_void_p), ("v_features_t", C.c_void_p),
                ("v_offset", C.c_void_p), ("v_viewmats", C.c_void_p), ("g_lo", C.c_int32), ("g_hi", C.c_int32),
                ("zero_v_records", C.c_int32)]


class Pack(C.Structure):
    _fields_ = [("K", C.c_int32), ("N", C.c_int32), ("D", C.c_int32), ("means2d", C.c_void_p),
                ("conics", C.c_void_p), ("opacities", C.c_void_p), ("colors", C.c_void_p),
                ("colors_per_cam", C.c_int32), ("depths", C.c_void_p), ("records", C.c_void_p)]


class TileCount(C.Structure):
    _fields_ = [("K", C.c_int32), ("N", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("records", C.c_void_p), ("radii", C.c_void_p), ("tight", C.c_int32),
                ("lists", Lists),
                ("tile_counts", C.c_void_p), ("tile_offsets", C.c_void_p),
                ("depths", C.c_void_p), ("entries", C.c_void_p), ("entry_capacity", C.c_int64),
                ("entry_cursor", C.c_void_p), ("counts_ready", C.c_int32)]


class TileSort(C.Structure):
    _fields_ = [("K", C.c_int32), ("N", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("records", C.c_void_p), ("radii", C.c_void_p), ("depths", C.c_void_p),
                ("tight", C.c_int32), ("lists", Lists),
                ("tile_offsets", C.c_void_p), ("tile_cursor", C.c_void_p),
                ("capacity", C.c_int64), ("keys", C.c_void_p), ("keys_tmp", C.c_void_p),
                ("sorted_ids", C.c_void_p),
                ("entries", C.c_void_p), ("entry_capacity", C.c_int64), ("n_entries", C.c_void_p)]


class BlendFwd(C.Structure):
    _fields_ = [("K", C.c_int32), ("N", C.c_int32), ("D", C.c_int32), ("width", C.c_int32),
                ("height", C.c_int32), ("lists", Lists), ("list_capacity", C.c_int64), ("records", C.c_void_p),
                ("tile_offsets", C.c_void_p),
                ("sorted_ids", C.c_void_p), ("backgrounds", C.c_void_p), ("out_colors", C.c_void_p),
                ("out_alphas", C.c_void_p), ("last_idx", C.c_void_p),
                ("dec_rays", C.c_void_p), ("dec_rays_per_k", C.c_int32), ("dec_w1", C.c_void_p),
                ("dec_w2", C.c_void_p), ("out_rgb", C.c_void_p), ("out_depth", C.c_void_p),
                ("flow_ref", C.c_int32), ("out_flow", C.c_void_p),
                ("dec_pose", C.c_void_p), ("dec_ppx", C.c_float), ("dec_ppy", C.c_float), ("dec_sfx", C.c_float),
                ("dec_sfy", C.c_float), ("list_masks", C.c_void_p)]


class BlendBwd(C.Structure):
    _fields_ = [("K", C.c_int32), ("N", C.c_int32), ("D", C.c_int32), ("width", C.c_int32),
                ("height", C.c_int32), ("lists", Lists), ("list_capacity", C.c_int64), ("records", C.c_void_p),
                ("tile_offsets", C.c_void_p),
                ("sorted_ids", C.c_void_p), ("backgrounds", C.c_void_p), ("out_alphas", C.c_void_p),
                ("last_idx", C.c_void_p), ("v_out_colors", C.c_void_p), ("v_out_alphas", C.c_void_p),
                ("v_records", C.c_void_p), ("sep_list", C.c_int32), ("v_means2d_sep", C.c_void_p),
                ("dec_rays", C.c_void_p), ("dec_rays_per_k", C.c_int32), ("dec_w1", C.c_void_p),
                ("dec_w2", C.c_void_p), ("out_colors", C.c_void_p), ("g_rgb", C.c_void_p),
                ("g_depth", C.c_void_p), ("g_alpha", C.c_void_p), ("g_mean", C.c_void_p),
                ("mean_K", C.c_int32), ("v_rays", C.c_void_p), ("v_w_partial", C.c_void_p),
                ("flow_ref", C.c_int32), ("g_flow", C.c_void_p),
                ("dec_pose", C.c_void_p), ("dec_ppx", C.c_float), ("dec_ppy", C.c_float), ("dec_sfx", C.c_float),
                ("dec_sfy", C.c_float), ("v_pose_partial", C.c_void_p), ("list_masks", C.c_void_p),
                ("sparse_grads", C.c_int32)]


class DecodeFwd(C.Structure):
    _fields_ = [("K", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("img", C.c_void_p),
                ("alpha", C.c_void_p), ("rays", C.c_void_p), ("rays_per_k", C.c_int32),
                ("w1", C.c_void_p), ("w2", C.c_void_p), ("rgb", C.c_void_p), ("depth", C.c_void_p),
                ("mean", C.c_void_p)]


class DecodeBwd(C.Structure):
    _fields_ = [("K", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("img", C.c_void_p),
                ("alpha", C.c_void_p), ("rays", C.c_void_p), ("rays_per_k", C.c_int32),
                ("w1", C.c_void_p), ("w2", C.c_void_p), ("g_rgb", C.c_void_p), ("g_depth", C.c_void_p),
                ("g_mean", C.c_void_p), ("v_img", C.c_void_p), ("v_alpha", C.c_void_p),
                ("v_rays", C.c_void_p), ("v_w1", C.c_void_p), ("v_w2", C.c_void_p)]


class HexMlpFwd(C.Structure):
    _fields_ = [("N", C.c_int32), ("pts", C.c_void_p), ("scales", C.c_void_p), ("rots", C.c_void_p),
                ("times", C.c_void_p), ("aabb", C.c_float * 6), ("levels", C.c_int32),
                ("net_width", C.c_int32), ("plane_features", C.c_int32), ("planes", C.c_void_p * 24),
                ("plane_w", C.c_int32 * 24), ("plane_h", C.c_int32 * 24),
                ("w0", C.c_void_p), ("b0", C.c_void_p), ("wa", C.c_void_p), ("ba", C.c_void_p),
                ("wb", C.c_void_p), ("bb", C.c_void_p), ("out_pts", C.c_void_p),
                ("out_scales", C.c_void_p), ("out_rots", C.c_void_p)]


class HexMlpBwd(C.Structure):
    _fields_ = [("N", C.c_int32), ("ld", C.c_int32), ("pts", C.c_void_p), ("rots", C.c_void_p), ("times", C.c_void_p),
                ("aabb", C.c_float * 6), ("levels", C.c_int32), ("net_width", C.c_int32), ("plane_features", C.c_int32),
                ("planes", C.c_void_p * 24), ("plane_w", C.c_int32 * 24), ("plane_h", C.c_int32 * 24),
                ("w0", C.c_void_p), ("b0", C.c_void_p), ("wa", C.c_void_p), ("ba", C.c_void_p), ("wb", C.c_void_p),
                ("bb", C.c_void_p), ("w0_t", C.c_void_p), ("wa_t", C.c_void_p), ("wb_t", C.c_void_p),
                ("g_out_pts", C.c_void_p), ("g_out_scales", C.c_void_p), ("g_out_rots", C.c_void_p),
                ("g_pts", C.c_void_p), ("g_scales", C.c_void_p), ("g_rots", C.c_void_p), ("g_feat", C.c_void_p),
                ("featT", C.c_void_p), ("a1T", C.c_void_p), ("a2T", C.c_void_p), ("gz1T", C.c_void_p),
                ("goT", C.c_void_p), ("gh0T", C.c_void_p)]


WGRAD_MAX = 8


class HexWgrad(C.Structure):
    _fields_ = [("n_problems", C.c_int32), ("ld", C.c_int32), ("A", C.c_void_p * WGRAD_MAX), ("B", C.c_void_p * WGRAD_MAX),
                ("n_cols", C.c_int32 * WGRAD_MAX), ("C", C.c_void_p * WGRAD_MAX), ("ldc", C.c_int32 * WGRAD_MAX),
                ("transpose_out", C.c_int32 * WGRAD_MAX), ("bias", C.c_void_p * WGRAD_MAX),
                ("bias_from", C.c_int32 * WGRAD_MAX)]


class FlowRecFwd(C.Structure):
    _fields_ = [("K", C.c_int32), ("N", C.c_int32), ("records", C.c_void_p), ("flow_records", C.c_void_p)]


class FlowRecBwd(C.Structure):
    _fields_ = [("K", C.c_int32), ("N", C.c_int32), ("v_flow_records", C.c_void_p), ("v_records", C.c_void_p),
                ("accumulate", C.c_int32)]


class HexFeat(C.Structure):
    _fields_ = [("N", C.c_int32), ("pts", C.c_void_p), ("times", C.c_void_p), ("aabb", C.c_float * 6),
                ("levels", C.c_int32), ("planes", C.c_void_p * 24), ("plane_w", C.c_int32 * 24),
                ("plane_h", C.c_int32 * 24), ("feat", C.c_void_p), ("g_feat", C.c_void_p),
                ("g_planes", C.c_void_p * 24), ("g_pts", C.c_void_p), ("g_times", C.c_void_p)]


ADAM_MAX_TENSORS = 64


class Adam(C.Structure):
    _fields_ = [("n_tensors", C.c_int32), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("one_minus_beta1", C.c_float), ("one_minus_beta2", C.c_float), ("reserved_", C.c_int32),
                ("param", C.c_void_p * ADAM_MAX_TENSORS), ("grad", C.c_void_p * ADAM_MAX_TENSORS),
                ("exp_avg", C.c_void_p * ADAM_MAX_TENSORS), ("exp_avg_sq", C.c_void_p * ADAM_MAX_TENSORS),
                ("numel", C.c_int64 * ADAM_MAX_TENSORS), ("step_size", C.c_float * ADAM_MAX_TENSORS),
                ("bc2_sqrt", C.c_float * ADAM_MAX_TENSORS), ("chunk_begin", C.c_int32 * (ADAM_MAX_TENSORS + 1))]


class PhotoLossFwd(C.Structure):
    _fields_ = [("planes", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("img", C.c_void_p), ("gt", C.c_void_p),
                ("window", C.c_float * 11), ("sums", C.c_void_p), ("d_mu1", C.c_void_p), ("d_x2", C.c_void_p),
                ("d_xy", C.c_void_p)]


class PhotoLossBwd(C.Structure):
    _fields_ = [("planes", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("img", C.c_void_p), ("gt", C.c_void_p),
                ("window", C.c_float * 11), ("d_mu1", C.c_void_p), ("d_x2", C.c_void_p), ("d_xy", C.c_void_p),
                ("v_loss", C.c_void_p), ("scale_l1", C.c_float), ("scale_ssim", C.c_float), ("v_img", C.c_void_p)]


class CameraRays(C.Structure):
    _fields_ = [("K", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("ppx", C.c_float), ("ppy", C.c_float),
                ("sfx", C.c_float), ("sfy", C.c_float), ("rot", C.c_void_p), ("centre", C.c_void_p),
                ("rays", C.c_void_p), ("v_rays", C.c_void_p), ("v_rot", C.c_void_p), ("v_centre", C.c_void_p)]


class FlowWarp(C.Structure):
    _fields_ = [("B", C.c_int32), ("K", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("ori", C.c_void_p),
                ("latent", C.c_void_p), ("exp2mid", C.c_void_p), ("mid2exp", C.c_void_p), ("latent_alpha", C.c_void_p),
                ("d_alpha", C.c_void_p), ("sums", C.c_void_p), ("v_loss", C.c_void_p), ("v_latent", C.c_void_p),
                ("v_exp2mid", C.c_void_p), ("v_mid2exp", C.c_void_p), ("v_latent_alpha", C.c_void_p),
                ("v_d_alpha", C.c_void_p), ("v_ori", C.c_void_p)]


class RegLoss(C.Structure):
    _fields_ = [("n_depth", C.c_int64), ("n_alpha", C.c_int64), ("depth", C.c_void_p), ("gt_depth", C.c_void_p),
                ("alpha", C.c_void_p), ("sums", C.c_void_p), ("g_depth", C.c_void_p), ("g_alpha", C.c_void_p)]


class Normals(C.Structure):
    _fields_ = [("B", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("z", C.c_void_p),
                ("ppx", C.c_float), ("ppy", C.c_float), ("sfx", C.c_float), ("sfy", C.c_float), ("skew", C.c_float),
                ("pixel_offset", C.c_float), ("normals", C.c_void_p)]


PACK_MAX_JOBS = 96
PACK_TILED, PACK_PLAIN, PACK_TRANSPOSE = 0, 1, 2


class PackJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("row_stride", C.c_int64), ("col_stride", C.c_int64), ("rows", C.c_int32),
                ("cols", C.c_int32), ("valid_rows", C.c_int32), ("valid_cols", C.c_int32), ("dst", C.c_void_p),
                ("kind", C.c_int32), ("reserved_", C.c_int32)]


class PackOperands(C.Structure):
    _fields_ = [("n_jobs", C.c_int32), ("n_chunks", C.c_int32), ("jobs", C.c_void_p), ("chunk_begin", C.c_void_p)]


COMPACT_MAX_TENSORS = 64


class CompactRows(C.Structure):
    _fields_ = [("n_tensors", C.c_int32), ("reserved_", C.c_int32), ("n_old", C.c_int64), ("n_out", C.c_int64),
                ("idx", C.c_void_p), ("src", C.c_void_p * COMPACT_MAX_TENSORS), ("ext", C.c_void_p * COMPACT_MAX_TENSORS),
                ("dst", C.c_void_p * COMPACT_MAX_TENSORS), ("row_words", C.c_int32 * COMPACT_MAX_TENSORS),
                ("chunk_begin", C.c_int32 * (COMPACT_MAX_TENSORS + 1))]


COPY_MAX_SEGMENTS = 64


class CopySegments(C.Structure):
    _fields_ = [("n_segments", C.c_int32), ("reserved_", C.c_int32), ("src", C.c_void_p * COPY_MAX_SEGMENTS),
                ("dst", C.c_void_p * COPY_MAX_SEGMENTS), ("n_words", C.c_int64 * COPY_MAX_SEGMENTS),
                ("chunk_begin", C.c_int32 * (COPY_MAX_SEGMENTS + 1))]


EXTRA_STRUCTS = {"MobgsHexMlpBwd": HexMlpBwd, "MobgsHexWgrad": HexWgrad, "MobgsCopySegments": CopySegments, "MobgsCompactRows": CompactRows, "MobgsRegLoss": RegLoss, "MobgsFlowWarp": FlowWarp, "MobgsCameraRays": CameraRays, "MobgsAdam": Adam, "MobgsPhotoLossFwd": PhotoLossFwd, "MobgsPhotoLossBwd": PhotoLossBwd, "MobgsNormals": Normals, "MobgsPackJob": PackJob, "MobgsPackOperands": PackOperands}

# name -> argument struct (None = no-arg string getter).  tests/test_abi.py checks that every
# function declared in include/mobgs_b200.h appears here and resolves in the .so.
ENTRY_POINTS = {
    "mobgs_version": None,
    "mobgs_last_error": None,
    "mobgs_project_fwd": ProjectFwd,
    "mobgs_project_bwd": ProjectBwd,
    "mobgs_synth_project_fwd": SynthFwd,
    "mobgs_synth_project_bwd": SynthBwd,
    "mobgs_pack_records": Pack,
    "mobgs_tile_count": TileCount,
    "mobgs_tile_emit_sort": TileSort,
    "mobgs_blend_fwd": BlendFwd,
    "mobgs_blend_bwd": BlendBwd,
    "mobgs_decode_fwd": DecodeFwd,
    "mobgs_decode_bwd": DecodeBwd,
    "mobgs_hexplane_mlp_fwd": HexMlpFwd,
    "mobgs_hexplane_mlp_bwd": HexMlpBwd,
    "mobgs_hexplane_wgrad": HexWgrad,
    "mobgs_flow_records_fwd": FlowRecFwd,
    "mobgs_flow_records_bwd": FlowRecBwd,
    "mobgs_midflow_records_fwd": FlowRecFwd,
    "mobgs_midflow_records_bwd": FlowRecBwd,
    "mobgs_hexplane_features_fwd": HexFeat,
    "mobgs_hexplane_features_bwd": HexFeat,
    "mobgs_adam_step": Adam,
    "mobgs_photo_loss_fwd": PhotoLossFwd,
    "mobgs_photo_loss_bwd": PhotoLossBwd,
    "mobgs_reg_loss_fwd": RegLoss,
    "mobgs_flow_warp_loss_fwd": FlowWarp,
    "mobgs_flow_warp_loss_bwd": FlowWarp,
    "mobgs_camera_rays_fwd": CameraRays,
    "mobgs_camera_rays_bwd": CameraRays,
    "mobgs_adam_chunk_elems": "int",
    "mobgs_compact_rows": CompactRows,
    "mobgs_compact_chunk_words": "int",
    "mobgs_copy_segments": CopySegments,
    "mobgs_depth_normals": Normals,
    "mobgs_pack_operands": PackOperands,
    "mobgs_pack_chunk_elems": "int",
}

_lib = None
_lock = threading.Lock()


def register_entry_point(name, struct):
    ENTRY_POINTS[name] = struct


def load() -> C.CDLL:
    """dlopen the in-tree library (building it first if the toolchain is here and it is stale)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            try:
                build()
            except Exception as e:  # noqa: BLE001
                raise RuntimeError(
                    f"mobgs_b200: {LIB_PATH} is missing and could not be built ({e}). "
                    "Run `python -c 'import __graft_entry__ as g; g.build()'`. There is no fallback path."
                ) from e
        lib = C.CDLL(LIB_PATH)
        lib.mobgs_knn3_mean_dist2.restype = C.c_int
        lib.mobgs_knn3_mean_dist2.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        lib.mobgs_subframe_mean.restype = C.c_int
        lib.mobgs_subframe_mean.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p]
        for name, struct in ENTRY_POINTS.items():
            fn = getattr(lib, name)   # AttributeError => symbol missing: fail loudly
            if struct is None:
                fn.restype = C.c_char_p
                fn.argtypes = []
            elif struct == "int":
                fn.restype = C.c_int
                fn.argtypes = []
            else:
                fn.restype = C.c_int
                fn.argtypes = [C.POINTER(struct), C.c_void_p]
        _lib = lib
        return lib


# kernels launched per entry point (memsets excluded) — bench.py's `gpu_launches` claim
KERNELS_PER_CALL = {
    "mobgs_project_fwd": 1, "mobgs_project_bwd": 1, "mobgs_synth_project_fwd": 1,
    "mobgs_synth_project_bwd": 1, "mobgs_pack_records": 1, "mobgs_tile_count": 2,
    "mobgs_tile_emit_sort": 2, "mobgs_blend_fwd": 1, "mobgs_blend_bwd": 1,
    "mobgs_decode_fwd": 1, "mobgs_decode_bwd": 1, "mobgs_hexplane_mlp_fwd": 1,
}
LAUNCH_COUNT = 0
# optional per-entry-point device timing: TIMING = {} enables it; values are lists of
# (start_event, end_event) recorded on the launching stream (bench.py reads them after a sync)
TIMING = None


DEC_SLOTS = 1024
POSE_SLOTS = 64


def knn3_mean_dist2(points_ptr, out_ptr, n, stream) -> None:
    global LAUNCH_COUNT
    lib = load()
    LAUNCH_COUNT += 1
    rc = lib.mobgs_knn3_mean_dist2(C.c_void_p(points_ptr), C.c_void_p(out_ptr), C.c_int32(n), C.c_void_p(stream))
    if rc != 0:
        raise RuntimeError(f"mobgs_knn3_mean_dist2 failed (code {rc}): {lib.mobgs_last_error().decode()}")


def subframe_mean(rgb_ptr, mean_ptr, K, n, stream) -> None:
    global LAUNCH_COUNT
    lib = load()
    LAUNCH_COUNT += 1
    rc = lib.mobgs_subframe_mean(C.c_void_p(rgb_ptr), C.c_void_p(mean_ptr), C.c_int32(K), C.c_int64(n), C.c_void_p(stream))
    if rc != 0:
        raise RuntimeError(f"mobgs_subframe_mean failed (code {rc}): {lib.mobgs_last_error().decode()}")


def current_stream() -> int:
    """Raw cudaStream_t of torch's current stream.  torch.cuda.current_stream() builds a Python Stream object
    (~20 us, seven times per render() call); the private C getter takes ~1 us.  Falls back to the public API
    if the private one is ever missing."""
    import torch
    try:
        return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())
    except AttributeError:
        return torch.cuda.current_stream().cuda_stream


def call(name: str, args: C.Structure, stream: int) -> None:
    global LAUNCH_COUNT
    lib = load()
    LAUNCH_COUNT += KERNELS_PER_CALL.get(name, 1)
    if TIMING is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream())
        rc = getattr(lib, name)(C.byref(args), C.c_void_p(stream))
        e1.record(torch.cuda.current_stream())
        TIMING.setdefault(name, []).append((e0, e1))
    else:
        rc = getattr(lib, name)(C.byref(args), C.c_void_p(stream))
    if rc != 0:
        msg = lib.mobgs_last_error().decode()
        raise RuntimeError(f"{name} failed (code {rc}): {msg}")

import os, subprocess, sys, hashlib
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


# (name, width_mbs, height_mbs, generator arguments) — small pictures so the CPU checkers finish in seconds
STREAMS = [
    ("i_cabac_4x4",      6, 5, "-n 3 -s 1 --gop I --deblock 1 --pcm 0 --t8x8 0"),
    ("i_cabac_8x8_dbk",  9, 7, "-n 3 -s 7 --gop I --deblock 0 --t8x8 50 --pcm 30"),
    ("i_cavlc_slices",   9, 7, "-n 3 -s 9 --gop I --deblock 2 --t8x8 50 --slices 4 --cavlc"),
    ("i_scaling",        9, 7, "-n 3 -s 8 --gop I --deblock 0 --t8x8 60 --scaling 3 --slices 3"),
    ("p_refs_wp",        9, 7, "-n 8 -s 12 --gop IP --deblock 0 --refs 4 --wp 1"),
    ("p_cavlc",          9, 7, "-n 8 -s 21 --gop IP --deblock 0 --refs 3 --cavlc"),
    ("b_default",        9, 7, "-n 10 -s 13 --gop IPB --deblock 0"),
    ("b_explicit",       9, 7, "-n 10 -s 15 --gop IPB --deblock 0 --wp 1"),
    ("b_implicit_temporal", 9, 7, "-n 10 -s 16 --gop IPB --deblock 0 --wp 2 --temporal"),
    ("b_scaling_mv",     9, 7, "-n 10 -s 17 --gop IPB --deblock 0 --scaling 3 --refs 4 --mvrange 60 --slices 2"),
    ("b_cavlc_all",      11, 6, "-n 10 -s 18 --gop IPB --deblock 0 --wp 2 --temporal --slices 3 --cavlc"),
    # 1-macroblock-wide pictures: intra only — the reference's edge emulation cannot clamp both sides of one
    # 16-byte load (edge264_inter.c:1205-1206), so its inter output on 16-pixel-wide pictures is not the standard's
    ("ragged_1x1_intra", 1, 1, "-n 4 -s 19 --gop I --deblock 0"),
    ("ragged_1xN_intra", 1, 9, "-n 4 -s 20 --gop I --deblock 0 --cavlc"),
    ("ragged_2xN",       2, 9, "-n 6 -s 20 --gop IPB --deblock 0 --wp 1"),
    # 0/255 checkerboard PCM blocks as motion-compensation sources: reproduces the reference's int16 wrap in the
    # centre half-sample (edge264_inter.c:4-9), found by bench.py's bit-exactness check
    ("b_pcm_checker",    10, 8, "-n 12 -s 23 --gop IPB --deblock 0 --pcm 250 --pcm-checker --intra-pct 5 --skip-pct 30 --density 10"),
    ("wide_33x2",        33, 2, "-n 6 -s 22 --gop IPB --deblock 0 --wp 2"),
    # 4096 samples wide: the reference pads such strides (edge264_headers.c:2027-2037); ours pads differently,
    # only the samples and the reported strides matter
    ("wide_256x2",       256, 2, "-n 5 -s 24 --gop IPB --deblock 0"),
    # intra-heavy P/B pictures with inter macroblocks in between: the intra-picture wavefront (one warp per row) runs over
    # pictures where a neighbour may have been reconstructed by either kernel
    ("pb_intra_heavy",   20, 12, "-n 8 -s 77 --gop IPB --intra-pct 70 --deblock 0 --slices 3 --t8x8 50"),
    ("pb_intra_half",    13, 9,  "-n 9 -s 78 --gop IP --intra-pct 50 --deblock 0 --refs 2 --cavlc"),
    # implicit bi-prediction weights at both ends of their range (w1 = -64 and w1 = 128: the reference's separate code path,
    # edge264_inter.c:1152-1161) — reference B pictures + list modification put both references on one side of the picture;
    # tests/test_weight_coverage.py proves the blocks are there
    ("wp_implicit_extremes", 4, 4, "-n 40 -s 17 --gop IPB --bref --refs 3 --idr 23 --dpb --deblock 0 --wp 2"),
    # long-term references, list modification, memory-management operations (more of these, CPU side, in DPB_STREAMS)
    ("dpb_mmco_cabac",   4, 3, "-n 40 -s 101 --gop IP --refs 4 --idr 17 --dpb --deblock 0"),
    # B pictures used as references (verified on the B200 like the one above; its siblings run CPU side)
    ("bref_spatial",     4, 4, "-n 40 -s 901 --gop IPB --bref --refs 3 --idr 17 --deblock 0 --wp 1"),
]


# decoded-picture-buffer logic (host side only): list modification to short/long-term pictures, memory-management
# operations, long-term IDR, frame_num wrap-around, picture order count types 1 and 2 — checked CPU-side against the
# reference and the golden digests (reference behaviours mirrored: edge264_headers.c:611-701, 768-893)
DPB_STREAMS = [
    ("dpb_mmco_cavlc_poc1", 3, 2, "-n 50 -s 113 --gop IP --refs 3 --idr 23 --dpb --poc-type 1 --deblock 0 --cavlc"),
    ("dpb_mmco_poc2",    3, 2, "-n 50 -s 131 --gop IP --refs 2 --idr 29 --dpb --poc-type 2 --deblock 0"),
    ("dpb_mmco_refs5",   3, 2, "-n 60 -s 149 --gop IP --refs 5 --idr 40 --dpb --deblock 0 --wp 1"),
    # long-term pictures in B slices: temporal direct (no vector scaling for long-term references) and implicit weights
    ("dpb_ipb_temporal", 3, 2, "-n 60 -s 211 --gop IPB --refs 4 --idr 31 --dpb --deblock 0 --wp 2 --temporal"),
    ("dpb_ipb_spatial",  3, 2, "-n 60 -s 222 --gop IPB --refs 3 --idr 25 --dpb --deblock 0 --wp 1"),
    # slices of one picture with different slice types (I in P pictures, I/P in B pictures), deblocking across them
    ("mixed_slices_cabac", 5, 6, "-n 24 -s 503 --gop IPB --refs 3 --idr 13 --slices 4 --mixed-slices --deblock 0 --wp 1"),
    # direct prediction per 4x4 block (direct_8x8_inference_flag 0), P pictures that are not references, access unit
    # delimiters / SEI / filler NAL units between pictures
    ("direct4x4_temporal", 4, 4, "-n 20 -s 1203 --gop IPB --direct4x4 --temporal --refs 3 --idr 9 --t8x8 50 --deblock 0 --wp 2"),
    ("direct4x4_spatial", 4, 4, "-n 20 -s 1204 --gop IPB --direct4x4 --refs 2 --idr 9 --t8x8 50 --deblock 0 --wp 1 --cavlc"),
    ("nonref_p",         4, 4, "-n 20 -s 1207 --gop IP --nonref-p --refs 3 --idr 9 --deblock 0"),
    ("extra_nals",       4, 4, "-n 20 -s 1209 --gop IPB --extra-nals --refs 2 --idr 9 --deblock 0"),
    # quantiser extremes (both 8x8 dequantisation branches, saturating paths) with vectors far outside the picture
    ("qp_low_far_mv",    4, 3, "-n 9 -s 1101 --gop IPB --refs 2 --qp 2 --t8x8 50 --scaling 1 --density 60 --mvrange 200 --deblock 0 --wp 2"),
    ("qp_high_far_mv",   4, 3, "-n 9 -s 1104 --gop IPB --refs 2 --qp 50 --t8x8 50 --scaling 3 --density 60 --mvrange 240 --deblock 0 --wp 1 --cavlc"),
    # B pictures used as references (their own marking, lists with references on both sides, B co-located pictures)
    ("bref_implicit",    4, 4, "-n 40 -s 905 --gop IPB --bref --refs 4 --idr 21 --deblock 0 --wp 2 --cavlc"),
    # frame cropping rectangle on all four sides; parameter sets re-sent between pictures (new chroma QP offsets and
    # scaling lists in the picture parameter sets, the unchanged sequence parameter set repeated)
    ("crop_rect",        4, 3, "-n 8 -s 838 --gop IPB --refs 2 --crop-left 10 --crop-right 6 --crop-top 8 --crop-bottom 4 --deblock 0"),
    ("ps_update",        4, 4, "-n 30 -s 701 --gop IP --refs 2 --idr 11 --ps-update --scaling 3 --t8x8 50 --deblock 0"),
    ("mixed_slices_cavlc", 5, 6, "-n 24 -s 505 --gop IPB --refs 2 --idr 13 --slices 3 --mixed-slices --deblock 2 --wp 2 --temporal --cavlc"),
]


@pytest.fixture(scope="session")
def workdir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("streams"))


def make_stream(workdir, name, w, h, args):
    path = os.path.join(workdir, name + ".264")
    if not os.path.exists(path):
        gen = os.path.join(ROOT, "tools", "gen264")
        subprocess.run([gen, "-o", path, "-W", str(w), "-H", str(h)] + args.split(), check=True, stderr=subprocess.DEVNULL, timeout=120)
    return path


def md5_frames(frames):
    return [hashlib.md5(f[3]).hexdigest() for f in frames]


def have(backend):
    from checkers import _LIBS
    return os.path.exists(_LIBS[backend])

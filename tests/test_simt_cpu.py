"""CPU check of kernel variants by running the kernels' OWN source on the host: tests/simt/extract.py slices the kernels out of
dsp-map_b200/csrc/dspmap_frame.cuh, tests/simt/simt_host.h supplies threadIdx, __shared__, warp collectives and atomics (one
OS thread per CUDA thread), and tests/simt/check_*.cpp feeds a variant and the kernel it replaces the same random inputs
and requires bit-identical outputs.  No GPU involved; the GPU-side A/B of the same variants is tests/ab_toggles.py."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIMT = os.path.join(ROOT, "tests", "simt")
BUILD = os.path.join(ROOT, "tests", "_build", "simt")
CUH = os.path.join(ROOT, "dsp-map_b200", "csrc", "dspmap_frame.cuh")


def build_and_run(check, inc, kernels, timeout=600, defines=(), tag=""):
    os.makedirs(BUILD, exist_ok=True)
    subprocess.check_call([sys.executable, os.path.join(SIMT, "extract.py"), CUH, os.path.join(BUILD, inc)] + kernels)
    exe = os.path.join(BUILD, check + tag)
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-I/usr/local/cuda/include", "-I" + SIMT, "-I" + BUILD,
                           "-I" + os.path.join(ROOT, "dsp-map_b200", "csrc")] + ["-D" + d for d in defines] +
                          [os.path.join(SIMT, check + ".cpp"), "-o", exe, "-lpthread"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_warp_level_variants_equal_the_kernels_they_replace():
    out = build_and_run("check_warp_kernels", "warp_kernels.inc", ["k_resample", "k_resample_sm", "k_nb_place", "k_nb_place_redux"])
    assert out.count("identical") == 6 and "DIFFERENT" not in out


def test_observation_pass_variants_equal_the_row_major_kernels():
    """k_cz_chain<STG>, k_cz_chain_tma, k_weight2<QF> and the whole column-major family (k_pair_eval_col -> k_cz_chain_col ->
    k_weight_col, with and without the dsp_quot fast path) reproduce C_z, 1/C_z and the particle weights of k_pair_eval ->
    k_cz_chain -> k_weight2 bit for bit; the emulated cp.async.bulk aborts on a copy that breaks the 16-byte rules."""
    out = build_and_run("check_obs_kernels", "obs_kernels.inc",
                        ["use_pair_buffer", "chunk_to_pyramid", "nb_index_of", "cz_build_order", "scan_block", "k_pair_prep", "k_pair_prep_scan", "k_pair_eval", "k_cz_chain",
                         "k_cz_chain_tma", "k_weight2_t", "dsp_pdf2_f", "k_pair_eval_col", "k_cz_chain_col", "k_weight_col"])
    assert out.count("identical") == 30 and "DIFFERENT" not in out and "does not exercise" not in out


def test_warp_per_chunk_weight_kernel_equals_the_other_weight_kernels():
    """k_weight2w takes the frames with many particle chunks (cfg3, cfg5).  Built with W2_SWITCH = 0 it takes the small test
    scenes too and has to reproduce the weights of the column-major family (which the test above ties to k_weight2)."""
    out = build_and_run("check_obs_kernels", "obs_kernels.inc",
                        ["use_pair_buffer", "chunk_to_pyramid", "nb_index_of", "cz_build_order", "scan_block", "k_pair_prep", "k_pair_prep_scan", "k_pair_eval", "k_cz_chain",
                         "k_cz_chain_tma", "k_weight2_t", "k_weight2w_t", "dsp_pdf2_f", "k_pair_eval_col", "k_cz_chain_col", "k_weight_col"],
                        defines=("CHECK_W2W", "W2_SWITCH=0"), tag="_w2w")
    assert out.count("identical") == 6 and "DIFFERENT" not in out and "does not exercise" not in out


def test_normaliser_sparse_future_and_sort_kernels():
    out = build_and_run("check_misc_kernels", "misc_kernels.inc", ["last_block_done", "scan_block", "k_norm", "k_norm_fast", "k_fut_count", "k_fut_compact", "k_pyr_sort", "k_pyr_sort_w",
                         "k_occ_count", "k_occ_count_fs", "bits_below", "nb_point1_body", "k_nb_point1", "k_nb_point1_fs"])
    assert out.count("identical") == 23 and "DIFFERENT" not in out

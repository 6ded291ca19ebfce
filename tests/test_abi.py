"""C-ABI checks that need no GPU: the shared library loads and exports every symbol include/dram_b200.h declares."""
import ctypes
import os
import re

from dram_native import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_parses_and_covers_the_survey_surface():
    protos = lib.parse_header()
    assert len(protos) >= 30
    for name in ["dram_conv3d_umma_fwd", "dram_conv3d_umma_wgrad", "dram_conv3d_simt_fwd", "dram_conv3d_simt_wgrad",
                 "dram_bn_stats", "dram_bn_finalize", "dram_bn_relu_apply", "dram_bn_relu_bwd_reduce",
                 "dram_bn_relu_bwd_apply", "dram_maxpool2_bwd", "dram_upsample2x_concat_fwd", "dram_upsample2x_concat_bwd",
                 "dram_ram_reduce_fwd", "dram_ram_reduce_bwd", "dram_trilinear_resize_fwd", "dram_trilinear_resize_bwd",
                 "dram_masked_pool_fwd", "dram_masked_pool_bwd", "dram_ram_upsample_mask_scatter", "dram_pcm_fwd",
                 "dram_pcm_bwd", "dram_ncdhw_to_ndhwc", "dram_ndhwc_to_ncdhw", "dram_version", "dram_sm_arch"]:
        assert name in protos, name


def test_every_prototype_cites_the_reference():
    src = open(os.path.join(ROOT, "include", "dram_b200.h")).read()
    assert len(re.findall(r"(parts|models|metrics|job_runner|utils|data_transforms)\.py:\d+", src)) >= 15


def test_library_loads_and_exports_every_declared_symbol():
    L = lib.load()                       # binds each prototype; AttributeError on a missing export
    raw = ctypes.CDLL(lib.LIB_PATH)
    for name in lib.parse_header():
        assert hasattr(raw, name), name
    assert L.dram_version() >= 100
    assert L.dram_sm_arch() == 100
    assert isinstance(L.dram_last_error(), bytes)
    assert L.dram_pcm_num_offsets(2, 0) == 18 and L.dram_pcm_num_offsets(1, 1) == 7 and L.dram_pcm_num_offsets(3, 0) == 26


def test_argument_validation_needs_no_device():
    """bad arguments are rejected before any CUDA call, with a message"""
    L = lib.load()
    assert L.dram_conv3d_simt_fwd(None, None, None, None, 1, 4, 4, 4, 3, 3, 3, None) == -1
    assert b"bad arguments" in L.dram_last_error()
    assert L.dram_conv3d_umma_wgrad_workspace_bytes(1, 8, 8, 8, 60, 64, 3) == 0          # pad not a multiple of 64
    assert L.dram_conv3d_umma_wgrad_workspace_bytes(8, 80, 80, 80, 64, 64, 3) > 0


def test_sass_contains_blackwell_instructions():
    """the tensor-core kernels really are tcgen05 / TMA code (B200_PROFILING.md: UTC*MMA, UTMALDG, LDTM)"""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        import pytest
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic
    # the SM-pair kernels: cta_group::2 MMAs, 2-SM TMA loads, multicast commits, cluster barriers
    for mnemonic in ("UTCHMMA.2CTA", "UTMALDG.5D.2CTA", "UTCBAR.2CTA.MULTICAST", "UCGABAR_ARV"):
        assert mnemonic in sass, mnemonic

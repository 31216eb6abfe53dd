"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol declared in
include/b2icp.h, and refuses to run without a CUDA device (no CPU fallback behind the ABI)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b2icp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2icp_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for s in ("b2icp_create", "b2icp_destroy", "b2icp_set_target", "b2icp_set_source", "b2icp_align",
              "b2icp_fitness", "b2icp_nn_search", "b2icp_transform_cloud", "b2icp_align_batch",
              "b2icp_get_correspondences", "b2icp_promote_source_to_target"):
        assert s in syms


def test_library_exports_every_declared_symbol(b2lib):
    L = b2lib.load_library()
    for s in declared_symbols():
        assert hasattr(L, s), f"libb2icp.so does not export {s}"
    assert set(declared_symbols()) == set(b2lib.EXPORTS)


def test_struct_layout_matches_header(b2lib):
    # sizes implied by include/b2icp.h on LP64
    assert C.sizeof(b2lib.Params) == 88
    assert C.sizeof(b2lib.Result) == 16 * 8 + 4 * 4 + 8 + 8
    assert C.sizeof(b2lib.Timing) == 4 + 4 + 8 * 3 + 8 * 2


def test_default_params_are_the_reference_constants(b2lib):
    # reference include/icpslam/icp_odometer.h:62-65, include/icpslam/octree_mapper.h:53-56
    p = b2lib.default_params(b2lib.PRESET_ODOMETER)
    assert p.max_iterations == 10 and p.transformation_epsilon == 1e-6 and p.max_correspondence_distance == 1.0
    p = b2lib.default_params(b2lib.PRESET_MAPPER)
    assert p.max_iterations == 30
    assert p.k_correspondences == 20 and p.gicp_epsilon == 1e-3 and p.rotation_epsilon == 2e-3
    assert p.max_inner_iterations == 20


def test_status_strings_and_version(b2lib):
    L = b2lib.load_library()
    assert L.b2icp_status_string(0) == b"ok"
    assert b"correspondences" in L.b2icp_status_string(-4)
    assert L.b2icp_version() >= 1


def test_invalid_arguments_do_not_crash(b2lib):
    L = b2lib.load_library()
    assert L.b2icp_create(None, None) == -1
    assert L.b2icp_destroy(None) == -1
    assert L.b2icp_default_params(None, 0) == -1
    p = b2lib.default_params()
    p.max_iterations = 0
    h = C.c_void_p()
    assert L.b2icp_create(C.byref(p), C.byref(h)) == -1


def test_no_cpu_fallback_without_cuda(b2lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(b2lib.B2icpError) as e:
        b2lib.Registration()
    assert e.value.code == -7  # B2ICP_ERR_CUDA


def test_product_does_not_reference_the_oracle():
    """The product path must never import, link or dlopen anything under oracle/."""
    pkg = os.path.join(ROOT, "icpslam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "b2icp_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f

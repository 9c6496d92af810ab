"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol include/vqseg.h
declares (no compute calls without a GPU), and the package refuses to run on CPU."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native():
    from vq_seg_b200 import build, _native
    build.build()
    return _native


def test_header_symbols_are_exported(native):
    header = open(os.path.join(ROOT, "include", "vqseg.h")).read()
    declared = set(re.findall(r"\b(vqseg_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 18
    lib = native.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in vqseg.h but not exported by libvqseg.so"
    assert declared == set(native.SIGNATURES), "ctypes SIGNATURES out of sync with vqseg.h"
    assert lib.vqseg_version() == 100


def test_error_strings_and_sizes(native):
    lib = native.lib()
    assert lib.vqseg_error_string(0) == b"ok"
    assert b"workspace" in lib.vqseg_error_string(-2)
    assert b"sm_100" in lib.vqseg_error_string(-4)
    # blob = header + 2*K_pad fp32 norms + fp16 image of K_pad x D_pad + one 4 KiB |e|^2 limb tile per 128 codes
    #        + one 64-bit fingerprint per code row (the codebook guard)
    assert lib.vqseg_codebook_blob_bytes(512, 256) == 1024 + 4096 + 512 * 256 * 2 + 4 * 4096 + 4096
    assert lib.vqseg_codebook_blob_bytes(300, 100) == 1024 + 4096 + 512 * 128 * 2 + 4 * 4096 + 4096
    assert lib.vqseg_assign_workspace_bytes(32768, 256, 512, 0) >= 32768 * 48        # one 48-byte work record per row
    # atomic statistics: 256 bytes + the packed-row scratch of one chunk (strided maps are packed before the row kernels)
    assert lib.vqseg_code_stats_workspace_bytes(1000, 64, 32, 0) == 256 + 1000 * 64 * 4
    # large packed inputs with K <= 1536 take the counting sort in the unordered mode too (+ one code per sorted position)
    n = 1 << 18
    det, unord = lib.vqseg_code_stats_workspace_bytes(n, 64, 1024, 1), lib.vqseg_code_stats_workspace_bytes(n, 64, 1024, 0)
    assert unord == det + n * 4
    assert lib.vqseg_code_stats_workspace_bytes(n, 64, 2048, 0) == 256 + n * 64 * 4
    assert lib.vqseg_code_stats_workspace_bytes(n - 1, 64, 1024, 0) == 256 + (n - 1) * 64 * 4


def test_sass_is_blackwell_native():
    """The shipped .so must contain tcgen05 / TMEM / bulk-copy / TMA tensor-load SASS (UTCHMMA, LDTM, UBLKCP, UTMALDG)."""
    import shutil
    import subprocess
    from vq_seg_b200 import build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", build.LIB], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UBLKCP", "UTMALDG"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16" not in sass                       # no legacy mma.sync path


def test_no_debug_symbols_in_the_product_library(native):
    """Developer hooks (pipeline trace, micro-benchmarks) live in libvqseg_dev.so / scripts/dev, not in the product."""
    import subprocess
    from vq_seg_b200 import build
    syms = subprocess.run(["nm", "-D", "--defined-only", build.LIB], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (vqseg_[a-z0-9_]+)", syms))
    assert not [s for s in exported if "debug" in s or "kernel_timing" in s], exported
    header = open(os.path.join(ROOT, "include", "vqseg.h")).read()
    declared = set(re.findall(r"\b(vqseg_[a-z0-9_]+)\s*\(", header))
    assert exported - declared <= {"vqseg_internal_gather_ticket"}, exported - declared


def test_no_cpu_fallback():
    import vq_seg_b200 as V
    m = V.VectorQuantizer(dim=16, num_embeddings=8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 16, 4, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        V.ops.assign(torch.randn(1, 4, 16), torch.randn(8, 16), None)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "vq_seg_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle package"
                assert "vq_oracle" not in src and "ref_loader" not in src, f"{f} references the oracle package"

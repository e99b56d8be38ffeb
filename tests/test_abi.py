"""CPU-only: the C-ABI library loads and exports every symbol include/krylov_b200.h declares,
and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "krylov_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(kl_[a-z0-9_]+)\s*\(", hdr)))


@pytest.fixture(scope="module")
def lib():
    import gmres_b200 as kl
    path = kl.library_path()
    if not os.path.exists(path):
        import __graft_entry__ as g
        g.build()
    return ctypes.CDLL(path)


def test_every_declared_symbol_is_exported(lib):
    names = _declared()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_reference_entry_points_present(lib):
    # one entry per public module procedure of the reference
    for n in ("kl_gmres_mgsr_omp", "kl_gmres_mgsr_mf", "kl_gmres_hh_omp", "kl_gmres_hh_prec_omp", "kl_cg",
              "kl_pcg", "kl_cg_omp", "kl_pcg_omp", "kl_bicgstab", "kl_pbicgstab", "kl_pbicgstab_omp",
              "kl_apply_operator", "kl_apply_precond"):
        assert hasattr(lib, n)


def test_no_cpu_fallback():
    import torch
    import gmres_b200 as kl
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(kl.KrylovError):
        kl.Handle(0)


def test_product_does_not_touch_oracle():
    """Nothing under gmres_b200/ may import, link or load oracle/."""
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "gmres_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".f90", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"libkrylov_oracle|from oracle|import oracle|\bko_[a-z]+\s*\(|dlopen\([^)]*oracle", txt):
                    bad.append(f)
    assert not bad, bad


def test_cheb_params_policy(lib):
    out = (ctypes.c_double * 2)()
    lib.kl_cheb_params_from_ritz.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.POINTER(ctypes.c_double)]
    assert lib.kl_cheb_params_from_ritz(0.01, 8.0, out) == 0
    assert out[0] == pytest.approx(8.2) and out[1] == pytest.approx(0.2)   # tests/test_poisson_mf.f90:38


def test_cheb_interval_policy_is_host_arithmetic():
    """kl_cheb_interval_from_ritz needs no GPU: [b/ratio(k), b], b = 1.025*theta_max, ratios 41/100/400/400/1000."""
    import ctypes as C
    from gmres_b200.api import load_library
    L = load_library()
    out = (C.c_double * 2)()
    for k, ratio in ((1, 41.0), (2, 100.0), (3, 400.0), (4, 400.0), (6, 1000.0), (12, 1000.0)):
        assert L.kl_cheb_interval_from_ritz(8.0, k, out) == 0
        assert out[0] == 1.025 * 8.0 and out[1] == out[0] / ratio
    assert L.kl_cheb_interval_from_ritz(-1.0, 2, out) < 0 and L.kl_cheb_interval_from_ritz(8.0, 0, out) < 0
    assert L.kl_cheb_params_from_ritz(0.01, 8.0, out) == 0 and out[0] == 8.2 and abs(out[1] - 0.2) < 1e-15


def test_library_is_sm100a_only_and_links_no_math_library(lib):
    """The device code is sm_100a and nothing else; the host side links no cuBLAS / cuSPARSE / cuSOLVER / NCCL
    (NCCL is dlopen'ed by kl_comm_init only), i.e. every kernel on the path is this repo's own."""
    import shutil
    import subprocess
    import gmres_b200 as kl
    path = kl.library_path()
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("no cuobjdump")
    elfs = subprocess.run([cuobjdump, "-lelf", path], capture_output=True, text=True).stdout.split("\n")
    elfs = [e for e in elfs if "ELF file" in e]
    assert elfs and all("sm_100a" in e for e in elfs), elfs
    needed = subprocess.run(["readelf", "-d", path], capture_output=True, text=True).stdout
    libs = re.findall(r"Shared library: \[([^\]]+)\]", needed)
    assert not [n for n in libs if re.search(r"cublas|cusparse|cusolver|nccl|cudnn|cufft|torch", n)], libs
    assert os.path.getsize(path) < 15 * 1024 * 1024      # round-1 build was 43.5 MB (522 kernels)

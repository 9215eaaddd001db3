"""CPU: the C-ABI library builds, loads and exports every symbol include/gofpid_b200.h declares;
struct layouts agree between the header (compiled with gcc) and the ctypes binding; without a
CUDA device compute entry points fail loudly (no CPU fallback)."""
import ctypes as ct
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gofpid_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"GOFPID_API\s+[\w\s\*]+?\b(gofpid_\w+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from pygofpid_b200 import _native as nat
    nat.build()
    L = nat.lib()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "missing export " + n
        assert n in nat.SIGNATURES, "ctypes binding lacks " + n
    assert sorted(nat.SIGNATURES) == names
    out = subprocess.check_output(["nm", "-D", "--defined-only", nat.LIB_PATH], text=True)
    exported = set(re.findall(r"\b(gofpid_\w+)\b", out))
    assert set(names) <= exported


def test_struct_layouts_match_header():
    from pygofpid_b200 import _native as nat
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "gofpid_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu ", sizeof(gofpid_config), sizeof(gofpid_blob), sizeof(gofpid_frame_info), sizeof(gofpid_morph_op));
  printf("%zu %zu %zu %zu %zu\n", offsetof(gofpid_config, detector), offsetof(gofpid_config, knn_rng_seed),
         offsetof(gofpid_config, morph), offsetof(gofpid_config, blob_area_min), offsetof(gofpid_blob, a00));
  return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(prog)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        vals = [int(v) for v in subprocess.check_output([exe], text=True).split()]
    assert vals[:4] == [ct.sizeof(nat.Config), ct.sizeof(nat.Blob), ct.sizeof(nat.FrameInfo), ct.sizeof(nat.MorphOp)]
    assert vals[4:] == [nat.Config.detector.offset, nat.Config.knn_rng_seed.offset, nat.Config.morph.offset,
                        nat.Config.blob_area_min.offset, nat.Blob.a00.offset]


def test_defaults_are_the_reference_defaults():
    from pygofpid_b200 import _native as nat
    cfg = nat.Config()
    nat.lib().gofpid_config_default(ct.byref(cfg))
    assert (cfg.blur_kind, cfg.blur_kw, cfg.blur_kh, cfg.blur_sigma) == (nat.BLUR_GAUSSIAN, 3, 3, 4.0)
    assert cfg.detector == nat.DETECT_MOG2 and cfg.mog2_history == 500 and cfg.knn_history == 500
    assert (cfg.mog2_var_threshold, cfg.mog2_var_threshold_gen, cfg.mog2_var_init) == (16.0, 9.0, 15.0)
    assert abs(cfg.mog2_background_ratio - 0.9) < 1e-7 and abs(cfg.mog2_ct - 0.05) < 1e-8
    assert (cfg.knn_k, cfg.knn_dist2_threshold, cfg.detect_shadows, cfg.shadow_value) == (2, 400.0, 1, 127)
    assert cfg.blob_area_min == 100.0 and cfg.fd_threshold == 5.0


def test_gaussian_taps_host_helper_matches_oracle():
    from oracle import cport
    from pygofpid_b200.hotpath import gaussian_taps_q8
    for n in (1, 3, 5, 7, 9, 15, 31):
        for s in (0.0, 1.0, 4.0, 6.5):
            assert np.array_equal(gaussian_taps_q8(n, s), cport.gauss_kernel_q8(n, s)), (n, s)
    with pytest.raises(ValueError):
        gaussian_taps_q8(4, 1.0)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from pygofpid_b200 import HotPath
    from pygofpid_b200._native import GofpidError
    with pytest.raises(GofpidError, match="no CUDA device"):
        HotPath((16, 16))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pygofpid_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "libgofpid_oracle" not in src, f

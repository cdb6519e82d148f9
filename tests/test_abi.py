"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/spfsplat.h
declares, the ctypes mirrors have the C layout, and argument errors are reported (no compute calls)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "spfsplat.h")


def _lib():
    from spfsplatv2_b200 import _lib as L
    if not os.path.exists(L.LIB_PATH):
        L.build()
    return L


def test_exports_every_declared_symbol():
    L = _lib()
    declared = re.findall(r"SPF_API\s+[\w\s\*]+?\b(spf_\w+)\s*\(", open(HEADER).read())
    assert set(declared) == set(L.EXPORTS) and len(declared) >= 7
    l = L.lib()
    for name in declared:
        assert hasattr(l, name), name
    assert l.spf_version() >= 100


def test_struct_layout_matches_header(tmp_path):
    """Compile a tiny C program against the header and compare sizeof/offsetof with the ctypes mirrors."""
    L = _lib()
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "spfsplat.h"\nint main(){'
                   'printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(SpfRasterDesc), sizeof(SpfRasterIn),'
                   'sizeof(SpfRasterState), sizeof(SpfRasterOut), sizeof(SpfRasterGradOut), sizeof(SpfRasterGradIn),'
                   'offsetof(SpfRasterDesc, dup_capacity), offsetof(SpfRasterIn, viewmatrix));return 0;}')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(L.SpfRasterDesc), C.sizeof(L.SpfRasterIn), C.sizeof(L.SpfRasterState), C.sizeof(L.SpfRasterOut),
            C.sizeof(L.SpfRasterGradOut), C.sizeof(L.SpfRasterGradIn), L.SpfRasterDesc.dup_capacity.offset,
            L.SpfRasterIn.viewmatrix.offset]
    assert got == want


def test_argument_errors_are_reported_without_a_gpu():
    L = _lib()
    l = L.lib()
    bad = L.SpfRasterDesc(1, 1, 10, 64, 64, 7, 0, 1.0, 100)      # sh_degree 7
    assert l.spf_raster_control_ints(C.byref(bad)) == -1
    assert b"sh_degree" in l.spf_last_error()
    ok = L.SpfRasterDesc(2, 3, 1000, 64, 48, 4, 0, 1.0, 4096)
    n = l.spf_raster_control_ints(C.byref(ok))
    B, T, NB = 6, 4 * 3, 8
    assert n >= 4 + 3 * B * T + 1 + 2 * B * NB + 1
    assert l.spf_rope2d(None, None, 1, 1, 1, 4, 4, 4, 0, 100.0, 1.0, None) == -1
    cin = L.SpfRasterIn()
    st = L.SpfRasterState()
    out = L.SpfRasterOut()
    assert l.spf_raster_forward(C.byref(ok), C.byref(cin), C.byref(st), C.byref(out), None) == -1
    assert b"means3D" in l.spf_last_error()


def test_product_path_refuses_cpu_tensors():
    import torch
    from spfsplatv2_b200 import curope
    from spfsplatv2_b200.decoder import render_cuda
    from spfsplatv2_b200.synthetic import make_scene
    sc = make_scene(seed=0, v_cxt=1, h=32, w=32, grid=(8, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        render_cuda(sc.extrinsics[0], sc.intrinsics[0], sc.near[0], sc.far[0], (32, 32), torch.zeros(1, 3), sc.means,
                    sc.covariances, sc.harmonics, sc.opacities, sc.rotations, sc.scales)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        curope.rope_2d(torch.zeros(1, 4, 2, 8), torch.zeros(1, 4, 2, dtype=torch.int64), 100.0, 1.0)


def test_no_product_import_of_oracle():
    """The product package must never import the oracle (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "spfsplatv2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f


def test_install_shims_registers_the_modules_the_reference_imports():
    """`import diff_gauss_pose` (cuda_splatting.py:5) and `import curope` (curope2d.py:6-9) resolve to the drop-ins, with the
    reference's names; the product path refuses CPU tensors instead of falling back."""
    import importlib
    import sys

    import torch

    import spfsplatv2_b200
    saved = {k: sys.modules.get(k) for k in ("diff_gauss_pose", "curope")}
    try:
        spfsplatv2_b200.install_shims()
        dgp = importlib.import_module("diff_gauss_pose")
        cur = importlib.import_module("curope")
        assert dgp.GaussianRasterizationSettings._fields == (
            "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "projmatrix", "sh_degree",
            "prefiltered", "debug", "enable_cov_grad", "enable_sh_grad")
        assert callable(cur.rope_2d) and hasattr(cur, "cuRoPE2D") and hasattr(cur, "cuRoPE2D_func")
        with pytest.raises(RuntimeError, match="CUDA tensors"):
            cur.rope_2d(torch.zeros(1, 2, 1, 8), torch.zeros(1, 2, 2, dtype=torch.int64), 100.0, 1.0)
        s = dgp.GaussianRasterizationSettings(8, 8, 0.5, 0.5, torch.zeros(3), 1.0, torch.eye(4), 0, False, False)
        with pytest.raises(RuntimeError, match="CUDA tensors"):
            dgp.GaussianRasterizer(s)(means3D=torch.zeros(2, 3), means2D=None, opacities=torch.ones(2, 1),
                                      colors_precomp=torch.ones(2, 3), scales=torch.ones(2, 3),
                                      rotations=torch.tensor([[1.0, 0, 0, 0]] * 2), viewmatrix=torch.eye(4))
        with pytest.raises(Exception, match="SHs or precomputed colors"):
            dgp.GaussianRasterizer(s)(means3D=torch.zeros(2, 3), means2D=None, opacities=torch.ones(2, 1),
                                      scales=torch.ones(2, 3), rotations=torch.ones(2, 4), viewmatrix=torch.eye(4))
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v

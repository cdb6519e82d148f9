"""In-switch (NVLS multimem) gradient all-reduce, csrc/allreduce.cu through the C ABI (`spf_multimem_allreduce_f32`).

Needs >= 2 GPUs behind an NVSwitch: skipped on the single-GPU test box; `gpurun --gpus 2 -- python -m pytest
tests/test_nvls_gpu.py -m gpu` runs it (results of the round-1 runs: profiles/nvls_check_n{2,4,8}.json)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_multimem_allreduce_matches_nccl(tmp_path):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    out = tmp_path / "nvls.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "scripts", "nvls_check.py"),
           "--mib", "16", "--out", str(out)]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    d = json.loads(out.read_text())
    if not d["nvls"]:
        pytest.skip(f"no multicast support on this box: {d['error']}")
    assert d["ranks_identical"]
    assert d["max_abs_diff_vs_nccl"] <= 4e-6          # fp32 sums of `world` unit-normal addends, different order


def test_multimem_entry_rejects_bad_arguments():
    """Argument checks come before any CUDA call, so this runs without a GPU (the pointer is never dereferenced)."""
    import ctypes as C
    from spfsplatv2_b200 import _lib
    lib = _lib.lib()
    assert lib.spf_multimem_allreduce_f32(None, 16, 0, 2, 8, None) != 0
    assert b"NULL" in lib.spf_last_error()
    fake = C.c_void_p(0x10000)
    assert lib.spf_multimem_allreduce_f32(fake, 18, 0, 2, 8, None) != 0                  # numel % 4
    assert lib.spf_multimem_allreduce_f32(C.c_void_p(0x10004), 16, 0, 2, 8, None) != 0   # alignment
    assert lib.spf_multimem_allreduce_f32(fake, 16, 2, 2, 8, None) != 0                  # rank >= world
    assert lib.spf_multimem_allreduce_f32(fake, 16, 0, 2, 0, None) != 0                  # n_blocks

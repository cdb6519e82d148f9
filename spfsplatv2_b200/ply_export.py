"""Gaussian scene -> viewer-compatible ``.ply``, rows packed on the GPU.

Mirrors /root/reference/src/model/ply_export.py:76-141 (``export_ply``, ``construct_list_of_attributes``): same
signature, same attribute order, same shift / rescale / viewer rotation, and the file is what ``plyfile`` writes for the
reference's element array (``format binary_little_endian 1.0``, ``property float <name>``), so anything that reads the
reference's exports reads these.  Mechanism: the per-Gaussian work (shift, scale, rotate, log-scales, quaternion
composition with scipy's matrix -> quaternion pivot rule, xyzw -> wxyz) is ONE kernel (``spf_ply_pack``, csrc/ply.cu)
writing the final [n, 17] fp32 rows; the scene statistics (median, 0.95 quantile, the 3x3 inverse) stay on the device,
and the only device -> host transfer is the finished row block.  The reference round-trips every tensor through numpy,
builds n Python tuples and calls scipy twice.  CUDA tensors only (no CPU fallback on the product path).
"""
from __future__ import annotations

import ctypes as C
import math
from pathlib import Path

import torch
from torch import Tensor

from . import _lib as L

ROW = 17


def construct_list_of_attributes(num_rest: int) -> list:
    """ply_export.py:12-24."""
    attributes = ["x", "y", "z", "nx", "ny", "nz"]
    attributes += [f"f_dc_{i}" for i in range(3)]
    attributes += [f"f_rest_{i}" for i in range(num_rest)]
    attributes.append("opacity")
    attributes += [f"scale_{i}" for i in range(3)]
    attributes += [f"rot_{i}" for i in range(4)]
    return attributes


def _viewer_rotation(device) -> Tensor:
    """+Z up (ply_export.py:95-101) composed with the -45 degree turn about z the Polycam viewer wants (:103-111;
    scipy's from_rotvec([0, 0, -45], degrees=True).as_matrix(), written out)."""
    up = torch.tensor([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]], dtype=torch.float32, device=device)
    a = math.radians(-45.0)
    adj = torch.tensor([[math.cos(a), -math.sin(a), 0.0], [math.sin(a), math.cos(a), 0.0], [0.0, 0.0, 1.0]],
                       dtype=torch.float32, device=device)
    return adj @ up


def ply_rows(extrinsics: Tensor, means: Tensor, scales: Tensor, rotations: Tensor, harmonics: Tensor,
             opacities: Tensor) -> Tensor:
    """The [n, 17] fp32 vertex rows of ``export_ply`` as a CUDA tensor (no host synchronisation)."""
    if not means.is_cuda:
        raise RuntimeError("spfsplatv2_b200.ply_export needs CUDA tensors (no CPU fallback on the product path)")
    dev = means.device
    f = lambda t: t.detach().to(torch.float32).contiguous()
    means, scales, rotations, harmonics, opacities = f(means), f(scales), f(rotations), f(harmonics), f(opacities)
    n = means.shape[0]
    shift = means.median(dim=0).values                                          # ply_export.py:87
    scale_factor = (means - shift).abs().quantile(0.95, dim=0).max()             # :90
    rotation = _viewer_rotation(dev) @ extrinsics.detach().to(torch.float32)[:3, :3].inverse()      # :113-115
    params = torch.cat([rotation.reshape(9), shift.reshape(3), scale_factor.reshape(1)]).contiguous()
    rows = torch.empty(n, ROW, dtype=torch.float32, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    with torch.cuda.device(dev):
        L.check(L.lib().spf_ply_pack(p(means), p(scales), p(rotations), p(harmonics), p(opacities), p(params), n,
                                     harmonics.shape[-1], p(rows), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                "spf_ply_pack")
    return rows


def write_ply_rows(path, rows, attributes=None) -> None:
    """Binary little-endian PLY with one float property per column -- byte for byte what
    ``PlyData([PlyElement.describe(elements, "vertex")]).write(path)`` produces for an all-'f4' element array."""
    import numpy as np
    rows = np.ascontiguousarray(rows, dtype="<f4")
    attributes = attributes or construct_list_of_attributes(0)
    if rows.ndim != 2 or rows.shape[1] != len(attributes):
        raise ValueError(f"rows {rows.shape} do not match {len(attributes)} attributes")
    header = "ply\nformat binary_little_endian 1.0\n" + f"element vertex {rows.shape[0]}\n" + \
             "".join(f"property float {a}\n" for a in attributes) + "end_header\n"
    path = Path(path)
    path.parent.mkdir(exist_ok=True, parents=True)
    with open(path, "wb") as fh:
        fh.write(header.encode("ascii"))
        fh.write(rows.tobytes())


def read_ply_rows(path):
    """(attribute names, [n, k] float32 rows) of a PLY written by ``write_ply_rows`` / plyfile's all-float vertex element."""
    import numpy as np
    with open(path, "rb") as fh:
        data = fh.read()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    lines = data[:end].decode("ascii").splitlines()
    if lines[0] != "ply" or lines[1] != "format binary_little_endian 1.0":
        raise ValueError("not a binary little-endian PLY")
    n = int(next(l for l in lines if l.startswith("element vertex")).split()[-1])
    names = [l.split()[-1] for l in lines if l.startswith("property float")]
    return names, np.frombuffer(data, dtype="<f4", count=n * len(names), offset=end).reshape(n, len(names))


def export_ply(extrinsics: Tensor, means: Tensor, scales: Tensor, rotations: Tensor, harmonics: Tensor, opacities: Tensor,
               path: Path) -> None:
    """Same contract as the reference's ``export_ply``: extrinsics [4,4] (c2w), means [g,3], scales [g,3], rotations
    [g,4] (xyzw), harmonics [g,3,d_sh], opacities [g] -> a DC-band-only PLY at ``path``."""
    rows = ply_rows(extrinsics, means, scales, rotations, harmonics, opacities)
    write_ply_rows(path, rows.cpu().numpy())

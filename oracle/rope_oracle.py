"""CPU oracle for the 2-D rotary position embedding (TEST INFRASTRUCTURE ONLY).

The checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
legs may import this file.

PARITY PINNED: unlike the rasterizer, the reference ships this path's CPU implementation
(``rope_2d_cpu``, /root/reference/src/model/encoder/backbone/croco/curope/curope.cpp:11-47) and a
pure-PyTorch equivalent (``RoPE2D``, /root/reference/src/model/encoder/backbone/croco/pos_embed.py:112-159).
``oracle/build_ref.sh`` compiles the former from the reference sources into ``oracle/_ref/`` and
``tests/golden/make_golden.py`` stores its outputs (and RoPE2D's) as fixtures; ``tests/test_rope_cpu.py``
checks the two restatements below against both.

Two restatements:
  * ``rope_2d_loop``   -- literal quintuple loop of curope.cpp:22-46 (numpy float32, small cases only)
  * ``rope_2d``        -- vectorised numpy float32, same formula: for the y half (x=0) and the x half
                          (x=1) of each head vector [u_Y(Q) v_Y(Q) u_X(Q) v_X(Q)], Q = D/4:
                          ang = fwd * pos / base**(d/Q);  u' = u cos - v sin;  v' = v cos + u sin
"""
from __future__ import annotations

import numpy as np


def rope_2d_loop(tokens: np.ndarray, positions: np.ndarray, base: float, fwd: float) -> np.ndarray:
    """tokens [B,N,H,D] float32, positions [B,N,2] int64 -> rotated copy.  Follows curope.cpp:11-47."""
    tok = np.array(tokens, dtype=np.float32, copy=True)
    B, N, H, D4 = tok.shape
    Q = D4 // 4
    f32 = np.float32
    for b in range(B):
        for x in range(2):
            for n in range(N):
                p = int(positions[b, n, x])
                for h in range(H):
                    for d in range(Q):
                        u = tok[b, n, h, d + 0 + x * 2 * Q]
                        v = tok[b, n, h, d + Q + x * 2 * Q]
                        # curope.cpp:36  inv_freq = fwd * p / powf(base, d/float(D))
                        ang = f32(f32(f32(fwd) * f32(p)) / np.power(f32(base), f32(d) / f32(Q), dtype=f32))
                        c, s = np.cos(ang, dtype=f32), np.sin(ang, dtype=f32)
                        tok[b, n, h, d + 0 + x * 2 * Q] = u * c - v * s
                        tok[b, n, h, d + Q + x * 2 * Q] = v * c + u * s
    return tok


def rope_2d(tokens: np.ndarray, positions: np.ndarray, base: float, fwd: float) -> np.ndarray:
    """Vectorised float32 restatement of curope.cpp:11-47; returns a rotated copy (the reference rotates
    in place).  Any float dtype is computed in float32 and rounded back to the input dtype once, which is
    what the CUDA kernel does for fp16 / bf16 inputs (kernels.cu:17-82 loads to float registers)."""
    tok = np.asarray(tokens, dtype=np.float32)
    B, N, H, D4 = tok.shape
    Q = D4 // 4
    f32 = np.float32
    d = np.arange(Q, dtype=f32)
    denom = np.power(f32(base), d / f32(Q), dtype=f32)                            # [Q]
    out = tok.copy()
    for x in range(2):
        p = positions[:, :, x].astype(f32)                                        # [B,N]
        ang = (f32(fwd) * p)[:, :, None] / denom[None, None, :]                   # [B,N,Q]
        c = np.cos(ang, dtype=f32)[:, :, None, :]
        s = np.sin(ang, dtype=f32)[:, :, None, :]
        u = tok[..., x * 2 * Q: x * 2 * Q + Q]
        v = tok[..., x * 2 * Q + Q: x * 2 * Q + 2 * Q]
        out[..., x * 2 * Q: x * 2 * Q + Q] = u * c - v * s
        out[..., x * 2 * Q + Q: x * 2 * Q + 2 * Q] = v * c + u * s
    return out

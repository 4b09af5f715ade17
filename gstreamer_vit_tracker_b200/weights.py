"""Model configuration and the flat weight-file format ("VTW1").

The reference's network lives in a crate that is not in the reference tree (`vit_tracker`,
/root/reference/Cargo.toml:24) and ships as an `.rknn` file that is not on disk
(/root/reference/src/main.rs:25), so no weights exist offline.  As SURVEY.md §8(c) prescribes,
the architecture is a declared stand-in (one-stream ViT over 64 template + 256 search
tokens, patch 16, three maps out) and the weights are *constructed* random-init values
written to one flat file that the CPU oracle and the CUDA path load identically.

File layout (little endian):
    char[4]  "VTW1"
    int32[7] D, depth, heads, hidden, head_ch, n_tensors, reserved
    float32  tensors, concatenated in `tensor_specs(cfg)` order, each in torch layout
             (Linear: [out, in]; conv: [out, in, kh, kw]).
"""
from __future__ import annotations

import hashlib
import struct
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np

from .synth import hash_u64

MAGIC = b"VTW1"
N_TEMPLATE_TOKENS = 64
N_SEARCH_TOKENS = 256
PATCH_K = 3 * 16 * 16
N_OUT = 5  # conf, size_w, size_h, off_x, off_y


@dataclass(frozen=True)
class ModelConfig:
    name: str
    D: int
    depth: int
    heads: int
    hidden: int
    head_ch: int

    @property
    def head_dim(self) -> int:
        return self.D // self.heads


MODELS = {
    # ~0.15 GFLOP / frame; the size class the upstream 0.7 MB ONNX suggests
    "nano": ModelConfig("nano", D=64, depth=2, heads=4, hidden=256, head_ch=64),
    # ViT-Tiny class, ~4.5 GFLOP / frame; the bench default
    "tiny": ModelConfig("tiny", D=192, depth=12, heads=3, hidden=768, head_ch=128),
}


def tensor_specs(cfg: ModelConfig) -> List[Tuple[str, Tuple[int, ...]]]:
    D, H, C = cfg.D, cfg.hidden, cfg.head_ch
    specs: List[Tuple[str, Tuple[int, ...]]] = [
        ("patch_w", (D, PATCH_K)),
        ("patch_b", (D,)),
        ("pos_z", (N_TEMPLATE_TOKENS, D)),
        ("pos_x", (N_SEARCH_TOKENS, D)),
    ]
    for i in range(cfg.depth):
        p = f"blk{i}."
        specs += [
            (p + "ln1_g", (D,)), (p + "ln1_b", (D,)),
            (p + "qkv_w", (3 * D, D)), (p + "qkv_b", (3 * D,)),
            (p + "proj_w", (D, D)), (p + "proj_b", (D,)),
            (p + "ln2_g", (D,)), (p + "ln2_b", (D,)),
            (p + "fc1_w", (H, D)), (p + "fc1_b", (H,)),
            (p + "fc2_w", (D, H)), (p + "fc2_b", (D,)),
        ]
    specs += [
        ("lnf_g", (D,)), ("lnf_b", (D,)),
        ("head1_w", (C, D, 3, 3)), ("head1_b", (C,)),
        ("head2_w", (N_OUT, C)), ("head2_b", (N_OUT,)),
    ]
    return specs


def n_params(cfg: ModelConfig) -> int:
    return sum(int(np.prod(s)) for _, s in tensor_specs(cfg))


def flops_per_frame(cfg: ModelConfig, n_tok: int = 320) -> int:
    """Dense-contraction FLOPs of one update (SURVEY.md §8(d) formula, head included)."""
    D, H, L, C = cfg.D, cfg.hidden, cfg.depth, cfg.head_ch
    per_blk = 2 * n_tok * D * (3 * D) + 2 * n_tok * D * D + 2 * 2 * n_tok * D * H + 4 * n_tok * n_tok * D
    patch = 2 * N_SEARCH_TOKENS * PATCH_K * D
    head = 2 * N_SEARCH_TOKENS * 9 * D * C + 2 * N_SEARCH_TOKENS * C * N_OUT
    return L * per_blk + patch + head


def _uniform(seed: int, shape: Tuple[int, ...], bound: float, offset: int = 0) -> np.ndarray:
    n = int(np.prod(shape))
    u = (hash_u64(seed, n, offset) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
    return ((u * 2.0 - 1.0) * bound).astype(np.float32).reshape(shape)


def make_weights(cfg: ModelConfig, seed: int = 20231001, variant: str = "stable") -> "OrderedDict[str, np.ndarray]":
    """Constructed random-init weights (SURVEY.md §7 'hard parts'):

    * trunk: uniform fan-in scaling, LayerNorm gains near 1;
    * conf head: weights scaled up so the hann-weighted top-1/top-2 margin is far above
      numeric noise (the margin is reported with every parity run);
    * size head: bias logit(0.2502) -> the box side stays near sqrt(w*h) instead of doubling
      every frame.  A random (untrained) size head has no restoring force, so any input
      dependence makes the box shrink or explode over hundreds of frames: the "stable"
      variant (long sequences, bench) zeroes the size weights; the "wild" variant
      (single-step / short-sequence parity) keeps them input dependent;
    * offset head: bias 0.5 (cell centre), small weights.
    """
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for k, (name, shape) in enumerate(tensor_specs(cfg)):
        s = seed * 1009 + k
        base = name.split(".")[-1]
        if base in ("ln1_g", "ln2_g", "lnf_g"):
            t = 1.0 + _uniform(s, shape, 0.1)
        elif base in ("ln1_b", "ln2_b", "lnf_b"):
            t = _uniform(s, shape, 0.05)
        elif base in ("pos_z", "pos_x"):
            t = _uniform(s, shape, 0.5)
        elif base.endswith("_b"):
            t = _uniform(s, shape, 0.05)
        else:
            fan_in = int(np.prod(shape[1:]))
            gain = 1.7 if base in ("patch_w", "fc1_w", "qkv_w") else 1.0
            t = _uniform(s, shape, gain * (3.0 / fan_in) ** 0.5)
        out[name] = t.astype(np.float32)
    # heads: rows 0 conf, 1-2 size, 3-4 offset
    w2, b2 = out["head2_w"], out["head2_b"]
    w2[0] *= 3.0
    b2[0] = 0.8
    w2[1:3] *= 0.0 if variant == "stable" else 0.05
    b2[1:3] = np.float32(np.log(0.2502 / (1 - 0.2502)))
    w2[3:5] *= 0.3
    b2[3:5] = 0.5
    return out


def save_weights(path: str, cfg: ModelConfig, tensors: Dict[str, np.ndarray]) -> str:
    specs = tensor_specs(cfg)
    h = hashlib.sha256()
    with open(path, "wb") as f:
        hdr = MAGIC + struct.pack("<7i", cfg.D, cfg.depth, cfg.heads, cfg.hidden, cfg.head_ch, len(specs), 0)
        f.write(hdr)
        h.update(hdr)
        for name, shape in specs:
            t = np.ascontiguousarray(tensors[name], dtype="<f4")
            assert t.shape == tuple(shape), (name, t.shape, shape)
            b = t.tobytes()
            f.write(b)
            h.update(b)
    return h.hexdigest()


def load_weights(path: str) -> Tuple[ModelConfig, "OrderedDict[str, np.ndarray]"]:
    with open(path, "rb") as f:
        raw = f.read()
    assert raw[:4] == MAGIC, "not a VTW1 file"
    D, depth, heads, hidden, head_ch, n_t, _ = struct.unpack("<7i", raw[4:32])
    cfg = ModelConfig("file", D, depth, heads, hidden, head_ch)
    specs = tensor_specs(cfg)
    assert n_t == len(specs)
    off = 32
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for name, shape in specs:
        n = int(np.prod(shape))
        out[name] = np.frombuffer(raw, dtype="<f4", count=n, offset=off).reshape(shape).copy()
        off += 4 * n
    assert off == len(raw)
    return cfg, out


def ensure_weight_file(model: str, directory: str, seed: int = 20231001, variant: str = "stable") -> str:
    """Write (once) and return the weight file for a named model."""
    import os

    cfg = MODELS[model]
    os.makedirs(directory, exist_ok=True)
    path = os.path.join(directory, f"vittrack_{model}_{variant}_s{seed}.vtw")
    expect = 32 + 4 * n_params(cfg)
    if not (os.path.exists(path) and os.path.getsize(path) == expect):
        tmp = path + f".tmp{os.getpid()}"
        save_weights(tmp, cfg, make_weights(cfg, seed, variant))
        os.replace(tmp, path)
    return path

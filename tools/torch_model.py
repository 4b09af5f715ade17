"""PyTorch statement of the stand-in VitTrack network, used ONLY to (a) export the weight file to
ONNX so the third-party cv2.TrackerVit can cross-check the oracle and
(b) sanity-check the C oracle's forward pass.  Fixture tooling; never on the product path.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gstreamer_vit_tracker_b200.weights import ModelConfig, load_weights  # noqa: E402


class Block(nn.Module):
    def __init__(self, D, heads, hidden):
        super().__init__()
        self.heads = heads
        self.ln1 = nn.LayerNorm(D, eps=1e-6)
        self.qkv = nn.Linear(D, 3 * D)
        self.proj = nn.Linear(D, D)
        self.ln2 = nn.LayerNorm(D, eps=1e-6)
        self.fc1 = nn.Linear(D, hidden)
        self.fc2 = nn.Linear(hidden, D)

    def forward(self, x):
        B, N, D = x.shape
        dh = D // self.heads
        qkv = self.qkv(self.ln1(x)).reshape(B, N, 3, self.heads, dh).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        a = torch.softmax((q @ k.transpose(-2, -1)) * (dh ** -0.5), dim=-1)
        x = x + self.proj((a @ v).transpose(1, 2).reshape(B, N, D))
        return x + self.fc2(F.gelu(self.fc1(self.ln2(x))))


class VitTrackNet(nn.Module):
    def __init__(self, cfg: ModelConfig):
        super().__init__()
        D = cfg.D
        self.cfg = cfg
        self.patch = nn.Conv2d(3, D, 16, 16)
        self.pos_z = nn.Parameter(torch.zeros(1, 64, D))
        self.pos_x = nn.Parameter(torch.zeros(1, 256, D))
        self.blocks = nn.ModuleList([Block(D, cfg.heads, cfg.hidden) for _ in range(cfg.depth)])
        self.lnf = nn.LayerNorm(D, eps=1e-6)
        self.head1 = nn.Conv2d(D, cfg.head_ch, 3, padding=1)
        self.head2 = nn.Conv2d(cfg.head_ch, 5, 1)

    def forward(self, template, search):
        z = self.patch(template).flatten(2).transpose(1, 2) + self.pos_z
        x = self.patch(search).flatten(2).transpose(1, 2) + self.pos_x
        t = torch.cat([z, x], dim=1)
        for b in self.blocks:
            t = b(t)
        t = self.lnf(t)
        f = t[:, 64:, :].transpose(1, 2).reshape(-1, self.cfg.D, 16, 16)
        o = self.head2(F.relu(self.head1(f)))
        return torch.sigmoid(o[:, 0:1]), torch.sigmoid(o[:, 1:3]), o[:, 3:5]


def from_weight_file(path: str) -> VitTrackNet:
    cfg, w = load_weights(path)
    net = VitTrackNet(cfg)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))  # noqa: E731
    with torch.no_grad():
        net.patch.weight.copy_(T(w["patch_w"]).reshape(cfg.D, 3, 16, 16))
        net.patch.bias.copy_(T(w["patch_b"]))
        net.pos_z.copy_(T(w["pos_z"])[None])
        net.pos_x.copy_(T(w["pos_x"])[None])
        for i, b in enumerate(net.blocks):
            p = f"blk{i}."
            b.ln1.weight.copy_(T(w[p + "ln1_g"])); b.ln1.bias.copy_(T(w[p + "ln1_b"]))
            b.qkv.weight.copy_(T(w[p + "qkv_w"])); b.qkv.bias.copy_(T(w[p + "qkv_b"]))
            b.proj.weight.copy_(T(w[p + "proj_w"])); b.proj.bias.copy_(T(w[p + "proj_b"]))
            b.ln2.weight.copy_(T(w[p + "ln2_g"])); b.ln2.bias.copy_(T(w[p + "ln2_b"]))
            b.fc1.weight.copy_(T(w[p + "fc1_w"])); b.fc1.bias.copy_(T(w[p + "fc1_b"]))
            b.fc2.weight.copy_(T(w[p + "fc2_w"])); b.fc2.bias.copy_(T(w[p + "fc2_b"]))
        net.lnf.weight.copy_(T(w["lnf_g"])); net.lnf.bias.copy_(T(w["lnf_b"]))
        net.head1.weight.copy_(T(w["head1_w"])); net.head1.bias.copy_(T(w["head1_b"]))
        net.head2.weight.copy_(T(w["head2_w"]).reshape(5, cfg.head_ch, 1, 1)); net.head2.bias.copy_(T(w["head2_b"]))
    return net.eval()


def export_onnx(weight_path: str, onnx_path: str) -> None:
    """torch.onnx legacy exporter without the `onnx` package."""
    from torch.onnx._internal.torchscript_exporter import onnx_proto_utils

    onnx_proto_utils._add_onnxscript_fn = lambda b, c: b
    net = from_weight_file(weight_path)
    z, x = torch.zeros(1, 3, 128, 128), torch.zeros(1, 3, 256, 256)
    torch.onnx.export(net, (z, x), onnx_path, input_names=["template", "search"],
                      output_names=["output1", "output2", "output3"], opset_version=17, dynamo=False)

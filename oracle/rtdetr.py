"""RT-DETR-L (the network of PP-DocLayout-L, BASELINE.json configs[4]) on torch-CPU fp32 -- TEST INFRASTRUCTURE ONLY.

The oracle-first step of SURVEY.md 8f item 1: a restatement of the detector that the CUDA kernels of the next round
have to reproduce.  oar-ocr-core runs this model as an opaque .onnx through ONNX Runtime
(domain/adapters/layout_detection_adapter.rs, models/detection/scale_aware_detector.rs); the architecture is taken
from the reference's in-tree description of the same family (oar-ocr-vl/src/models/pp_doclayout/: hgnetv2.rs,
encoder.rs, decoder.rs, model.rs:154-183, 296-345, 459-476, config.rs:279-327) with the plain RT-DETR heads
(per-layer class / box heads, no mask or reading-order head).  Weights are synthetic (seeded); the .onnx file and the
PaddleDetection export are not available offline, so nothing here is pinned by reference outputs: PARITY UNPINNED.

    backbone  HGNetV2-L (models.build_hgnetv2_l, run by oracle/net.py) -> strides 8 / 16 / 32, 512 / 1024 / 2048 ch
    encoder   1x1 projections to 256 -> AIFI (one post-norm transformer layer, 8 heads, GELU, 2-D sin/cos positions on
              q and k) on the stride-32 map -> CCFM: top-down (lateral 1x1, nearest x2, concat, CSPRep block) then
              bottom-up (3x3 s2, concat, CSPRep block); SiLU; RepVGG 3x3 + 1x1 branches summed before the activation
    decoder   1x1 projections, flatten + concat the three levels (8400 tokens at 640 x 640), anchors (0.05 * 2^level),
              top-300 query selection on the max class logit, 6 layers of [self-attention, multi-scale deformable
              attention (8 heads x 3 levels x 4 points, bilinear, zero outside), FFN 1024 ReLU], iterative box refinement
    rows()    the exported model's post-process: sigmoid, top-300 over (query, class), cxcywh -> xyxy in source pixels:
              [class_id, score, x1, y1, x2, y2] -- the rows LayoutDetectionAdapter::postprocess_pp_doclayout reads.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from .net import OracleNet

D, HEADS, FFN, LEVELS, POINTS, QUERIES, DEC_LAYERS = 256, 8, 1024, 3, 4, 300, 6
LN_EPS = 1e-5


def inverse_sigmoid(x, eps=1e-5):
    """decoder.rs:768-779"""
    x = x.clamp(0.0, 1.0)
    return torch.log(x.clamp(min=eps) / (1.0 - x).clamp(min=eps))


def sine_position_embedding(h, w, dim=D, temperature=10000.0):
    """encoder.rs:179-216: [sin(y w_k), cos(y w_k), sin(x w_k), cos(x w_k)], computed in f64"""
    pd = dim // 4
    omega = 1.0 / temperature ** (np.arange(pd, dtype=np.float64) / pd)
    ys, xs = np.meshgrid(np.arange(h, dtype=np.float64), np.arange(w, dtype=np.float64), indexing="ij")
    ph, pw = ys.reshape(-1, 1) * omega, xs.reshape(-1, 1) * omega
    return torch.from_numpy(np.concatenate([np.sin(ph), np.cos(ph), np.sin(pw), np.cos(pw)], 1).astype(np.float32))[None]


def generate_anchors(shapes):
    """model.rs:154-183: (logit-space anchors [1,N,4], valid mask [1,N,1])"""
    anchors, valid = [], []
    for level, (h, w) in enumerate(shapes):
        wh = np.float32(0.05) * np.float32(2.0 ** level)
        for y in range(h):
            for x in range(w):
                c = np.array([(x + 0.5) / w, (y + 0.5) / h, wh, wh], np.float32)
                ok = bool(np.all((c > 0.01) & (c < 0.99)))
                valid.append(1.0 if ok else 0.0)
                anchors.append(np.log(c / (1.0 - c)) if ok else np.full(4, np.finfo(np.float32).max, np.float32))
    return (torch.from_numpy(np.stack(anchors).astype(np.float32))[None],
            torch.tensor(valid, dtype=torch.float32).reshape(1, -1, 1))


def deformable_attention(value, shapes, locations, weights):
    """decoder.rs:212-470.  value [B,N,heads,hd]; locations [B,Q,heads,L,P,2] in [0,1]; weights [B,Q,heads,L,P]
    (softmax over L*P).  Bilinear, align_corners=False, zeros outside = grid_sample's defaults at 2*loc - 1."""
    B, _, heads, hd = value.shape
    Q = locations.shape[1]
    out, off = 0, 0
    for lvl, (h, w) in enumerate(shapes):
        v = value[:, off:off + h * w].permute(0, 2, 3, 1).reshape(B * heads, hd, h, w)
        off += h * w
        grid = (2.0 * locations[:, :, :, lvl] - 1.0).permute(0, 2, 1, 3, 4).reshape(B * heads, Q, POINTS, 2)
        s = F.grid_sample(v, grid, mode="bilinear", padding_mode="zeros", align_corners=False)  # [B*heads,hd,Q,P]
        wgt = weights[:, :, :, lvl].permute(0, 2, 1, 3).reshape(B * heads, 1, Q, POINTS)
        out = out + (s * wgt).sum(-1)
    return out.reshape(B, heads, hd, Q).permute(0, 3, 1, 2).reshape(B, Q, heads * hd)


class RTDetrL:
    def __init__(self, seed: int = 42, num_labels: int = 23):
        from oar_ocr_b200 import models
        self.num_labels = num_labels
        self.taps = []
        self.backbone = OracleNet(models.build_hgnetv2_l(seed, taps=self.taps))
        # the synthetic weight table lives with the other synthetic weights (models.layout_weights); names below
        self.w = {k: torch.from_numpy(v) for k, v in models.layout_weights(seed, num_labels).items()}

    # -- small helpers over the weight table
    def _lin(self, name, x):
        return F.linear(x, self.w[name + ".w"], self.w[name + ".b"])

    def _conv(self, name, x, stride=1, act=True):
        w = self.w[name + ".w"]
        y = F.conv2d(x, w, self.w[name + ".b"], stride, w.shape[-1] // 2)
        return F.silu(y) if act else y

    def _ln(self, name, x):
        return F.layer_norm(x, (D,), self.w[name + ".g"], self.w[name + ".b"], LN_EPS)

    def _mha(self, p, x, pos):
        """q = k = x + pos, v = x (encoder.rs:34-79, decoder.rs:105-160)"""
        B, T, _ = x.shape
        hd = D // HEADS
        qk = x + pos
        q = self._lin(p + ".q", qk).reshape(B, T, HEADS, hd).transpose(1, 2)
        k = self._lin(p + ".k", qk).reshape(B, T, HEADS, hd).transpose(1, 2)
        v = self._lin(p + ".v", x).reshape(B, T, HEADS, hd).transpose(1, 2)
        a = torch.softmax((q @ k.transpose(2, 3)) * hd ** -0.5, -1)
        return self._lin(p + ".o", (a @ v).transpose(1, 2).reshape(B, T, D))

    def _csp(self, p, x):
        """CspRepLayer (encoder.rs:259-316): conv1 -> 3 RepVGG blocks, + conv2 (hidden_expansion 1: no conv3)"""
        y = self._conv(p + ".conv1", x)
        for i in range(3):
            y = F.silu(self._conv(f"{p}.rep{i}.c3", y, act=False) + self._conv(f"{p}.rep{i}.c1", y, act=False))
        return y + self._conv(p + ".conv2", x)

    @torch.no_grad()
    def forward(self, x: np.ndarray):
        """x: f32 [B,3,H,W] (H, W multiples of 32).  Returns (logits [B,300,C], boxes [B,300,4] cxcywh in [0,1])."""
        source, shapes = self.encode(x)
        return self.decode(source, shapes)

    @torch.no_grad()
    def encode(self, x: np.ndarray):
        """backbone + hybrid encoder + decoder-input projections: the decoder memory [B, sum(H_l W_l), 256] and the
        three map sizes.  models.build_layout_encoder states the same computation as an OARG layer list."""
        cap = {}
        self.backbone.forward(x, capture=cap)
        feats = [self._conv(f"input_proj{l}", cap[self.taps[l + 1]], act=False) for l in range(3)]
        # AIFI on the stride-32 map (encoder.rs:131-176)
        B, _, h, w = feats[2].shape
        t = feats[2].flatten(2).transpose(1, 2)
        pos = sine_position_embedding(h, w)
        t = self._ln("aifi.ln1", t + self._mha("aifi", t, pos))
        t = self._ln("aifi.ln2", t + self._lin("aifi.fc2", F.gelu(self._lin("aifi.fc1", t))))
        feats[2] = t.transpose(1, 2).reshape(B, D, h, w)
        # CCFM (encoder.rs:404-455)
        fpn = [feats[2]]
        for i in range(2):
            top = self._conv(f"lateral{i}", fpn[-1])
            fpn[-1] = top
            up = F.interpolate(top, scale_factor=2, mode="nearest")
            fpn.append(self._csp(f"fpn{i}", torch.cat([up, feats[1 - i]], 1)))
        fpn.reverse()
        pan = [fpn[0]]
        for i in range(2):
            down = self._conv(f"down{i}", pan[-1], stride=2)
            pan.append(self._csp(f"pan{i}", torch.cat([down, fpn[i + 1]], 1)))
        # decoder input (model.rs:296-345)
        shapes, flat = [], []
        for l, f in enumerate(pan):
            s = self._conv(f"dec_input_proj{l}", f, act=False)
            shapes.append((s.shape[2], s.shape[3]))
            flat.append(s.flatten(2).transpose(1, 2))
        return torch.cat(flat, 1), shapes

    @torch.no_grad()
    def decode(self, source, shapes):
        B = source.shape[0]
        anchors, valid = generate_anchors(shapes)
        memory = self._ln("enc_output_ln", self._lin("enc_output", source * valid))
        enc_class = self._lin("enc_score", memory)
        z = memory
        for i in range(3):
            z = self._lin(f"enc_bbox{i}", z)
            z = F.relu(z) if i < 2 else z
        enc_coords = z + anchors
        # top-300 on the max class logit, ties by index (model.rs:459-476)
        scores = enc_class.max(-1).values
        idx = torch.stack([torch.tensor(sorted(range(scores.shape[1]), key=lambda a: (-float(scores[b, a]), a))[:QUERIES])
                           for b in range(B)])
        gather = idx[:, :, None]
        hidden = torch.gather(memory, 1, gather.expand(-1, -1, D))
        reference = torch.sigmoid(torch.gather(enc_coords, 1, gather.expand(-1, -1, 4)))
        value_src = source
        logits = None
        for i in range(DEC_LAYERS):  # decoder.rs:696-765
            p = f"dec{i}"
            qpos = self._lin("query_pos1", F.relu(self._lin("query_pos0", reference)))
            hidden = self._ln(p + ".ln1", hidden + self._mha(p + ".sa", hidden, qpos))
            q = hidden + qpos
            Q = q.shape[1]
            offsets = self._lin(p + ".ca.offsets", q).reshape(B, Q, HEADS, LEVELS, POINTS, 2)
            wts = torch.softmax(self._lin(p + ".ca.weights", q).reshape(B, Q, HEADS, LEVELS * POINTS), -1)
            wts = wts.reshape(B, Q, HEADS, LEVELS, POINTS)
            ref = reference.reshape(B, Q, 1, 1, 1, 4)
            loc = ref[..., :2] + offsets / POINTS * ref[..., 2:] * 0.5
            value = self._lin(p + ".ca.value", value_src).reshape(B, -1, HEADS, D // HEADS)
            cross = self._lin(p + ".ca.out", deformable_attention(value, shapes, loc, wts))
            hidden = self._ln(p + ".ln2", hidden + cross)
            hidden = self._ln(p + ".ln3", hidden + self._lin(p + ".fc2", F.relu(self._lin(p + ".fc1", hidden))))
            z = hidden
            for j in range(3):
                z = self._lin(f"{p}.bbox{j}", z)
                z = F.relu(z) if j < 2 else z
            reference = torch.sigmoid(z + inverse_sigmoid(reference))
            logits = self._lin(p + ".score", hidden)
        return logits.numpy(), reference.numpy()

    def rows(self, x: np.ndarray, src_wh):
        """the exported detector's own post-process: [B,300,6] rows [class_id, score, x1, y1, x2, y2] in source pixels"""
        logits, boxes = self.forward(x)
        B, Q, C = logits.shape
        out = np.zeros((B, QUERIES, 6), np.float32)
        for b in range(B):
            s = 1.0 / (1.0 + np.exp(-logits[b].astype(np.float64))).astype(np.float32)
            flat = s.reshape(-1)
            top = np.argsort(-flat, kind="stable")[:QUERIES]
            qi, ci = top // C, top % C
            cx, cy, bw, bh = boxes[b, qi].T
            w, h = np.float32(src_wh[b][0]), np.float32(src_wh[b][1])
            out[b] = np.stack([ci.astype(np.float32), flat[top], (cx - bw / 2) * w, (cy - bh / 2) * h, (cx + bw / 2) * w,
                               (cy + bh / 2) * h], 1)
        return out

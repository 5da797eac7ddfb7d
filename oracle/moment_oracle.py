"""CPU oracle for the MomentModel inference path (fp32 torch restatement; TEST INFRASTRUCTURE ONLY).

Pinned against the unmodified reference MomentModel run in the build container (oracle/make_golden_moment.py ->
tests/golden/moment_*.pt).  Reference lines cited per function.
"""
from __future__ import annotations

import math
from copy import deepcopy

import torch
import torch.nn.functional as F


def tf_layernorm(x, w, b, eps=1e-12):
    """clip4caption/modules/until_module.py:49-53 (epsilon inside the square root, biased variance)."""
    u = x.mean(-1, keepdim=True)
    s = (x - u).pow(2).mean(-1, keepdim=True)
    return w * ((x - u) / torch.sqrt(s + eps)) + b


def gelu_erf(x):
    """until_module.py:28-33."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def time_grid(video_mask):
    """modeling.py:178-193: per sample linspace(0,1,n_b) -> [-1,1], zero padded to max_b n_b."""
    n = video_mask.sum(dim=-1).long()
    T = int(n.max())
    rows = []
    for nb in n.tolist():
        t = (torch.linspace(0, 1, nb) - 0.5) * 2
        rows.append(torch.cat([t, torch.zeros(T - nb)]).view(1, T, 1))
    return torch.cat(rows, dim=0)


def visual_encoder(sd, feats, n_layers=2, heads=12, prefix="clip4cap_model.visual."):
    """VisualModel.forward with the all-zeros mask passed at modeling.py:208 (module_visual.py:396-424): embeddings
    (:118-130), post-LN layers (:154-247); every attention logit gets -10000.0 added (:406-414)."""
    B, T, _ = feats.shape
    p = prefix + "embeddings."
    x = F.linear(feats, sd[p + "word_embeddings.weight"], sd[p + "word_embeddings.bias"]) + sd[p + "position_embeddings.weight"][:T]
    x = tf_layernorm(x, sd[p + "LayerNorm.weight"], sd[p + "LayerNorm.bias"])
    Hd = x.shape[-1]
    dh = Hd // heads
    for i in range(n_layers):
        q_ = f"{prefix}encoder.layer.{i}."
        q = F.linear(x, sd[q_ + "attention.self.query.weight"], sd[q_ + "attention.self.query.bias"])
        k = F.linear(x, sd[q_ + "attention.self.key.weight"], sd[q_ + "attention.self.key.bias"])
        v = F.linear(x, sd[q_ + "attention.self.value.weight"], sd[q_ + "attention.self.value.bias"])
        q, k, v = (t.view(B, T, heads, dh).permute(0, 2, 1, 3) for t in (q, k, v))
        s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(dh)
        s = s + (-10000.0)                                    # (1 - 0) * -10000 on every key
        a = torch.matmul(s.softmax(dim=-1), v).permute(0, 2, 1, 3).contiguous().view(B, T, Hd)
        h = tf_layernorm(F.linear(a, sd[q_ + "attention.output.dense.weight"], sd[q_ + "attention.output.dense.bias"]) + x,
                         sd[q_ + "attention.output.LayerNorm.weight"], sd[q_ + "attention.output.LayerNorm.bias"])
        m = gelu_erf(F.linear(h, sd[q_ + "intermediate.dense.weight"], sd[q_ + "intermediate.dense.bias"]))
        x = tf_layernorm(F.linear(m, sd[q_ + "output.dense.weight"], sd[q_ + "output.dense.bias"]) + h,
                         sd[q_ + "output.LayerNorm.weight"], sd[q_ + "output.LayerNorm.bias"])
    return x


def moment_shared(sd, video_feats, text_feat, video_mask, moment_mask, asr_feats, boundary_mask=None):
    """MomentModel.foward_moment_shared (modeling.py:155-210)."""
    v = F.linear(video_feats, sd["clip_g_map.weight"], sd["clip_g_map.bias"])                       # :158
    t = F.linear(text_feat, sd["clip_g_map_text.weight"], sd["clip_g_map_text.bias"])                # :159
    v = tf_layernorm(v.float(), sd["clip4cap_model.normalize_video.visual_norm2d.weight"],
                     sd["clip4cap_model.normalize_video.visual_norm2d.bias"])                        # :161
    t = t / t.norm(dim=-1, keepdim=True)                                                             # :163
    feats = v * t.unsqueeze(1)                                                                       # :165
    a = F.layer_norm(asr_feats, (asr_feats.shape[-1],), sd["asr_enc_layer.0.weight"], sd["asr_enc_layer.0.bias"], 1e-5)
    feats = feats + F.linear(a, sd["asr_enc_layer.1.weight"], sd["asr_enc_layer.1.bias"])            # :167-169
    if boundary_mask is not None:
        feats = feats + sd["boundary_embed.weight"][boundary_mask]                                   # :171-173
    g = time_grid(video_mask)                                                                        # :178-193
    te = torch.tanh(F.linear(g, sd["temporal_embed.0.weight"], sd["temporal_embed.0.bias"]))
    feats = feats + F.linear(te, sd["temporal_embed.2.weight"], sd["temporal_embed.2.bias"])         # :195-196
    feats = feats + sd["mask_embed.weight"][moment_mask]                                             # :198-199
    return visual_encoder(sd, feats)                                                                 # :208


def head(sd, feats, name):
    return F.linear(feats, sd[name + ".0.weight"], sd[name + ".0.bias"]).squeeze(2)


def test_moment_retrieval(sd, batch, text_feat):
    """modeling.py:272-310 -> [[start, end]] per sample."""
    feats = moment_shared(sd, batch["vis_feats"], text_feat, batch["vis_mask"], batch["moment_mask"], batch["asr_feats"])
    s, e = head(sd, feats, "start_predictor"), head(sd, feats, "end_predictor")
    s[batch["vis_mask"] == 0] = -1e10
    e[batch["vis_mask"] == 0] = -1e10
    return torch.stack([s.argmax(dim=1), e.argmax(dim=1)], dim=-1).tolist(), s, e


def grow_region(scores, max_idx, threshold):
    """modeling.py:409-423: widen [l, r] while score/max > threshold (host-side control flow, restated verbatim)."""
    max_score = scores[max_idx]
    left = right = max_idx
    while (scores[left] / max_score) > threshold:
        if left == 0:
            break
        left -= 1
    while (scores[right] / max_score) > threshold:
        if right == (len(scores) - 1):
            break
        right += 1
    return left, right


def postprocess_steps(steps, last_bound):
    """modeling.py:435-463."""
    steps = sorted(steps + [[last_bound, last_bound]], key=lambda x: x[0])
    flat = [v for pair in steps for v in pair]
    while flat[-1] > last_bound:
        flat.pop(-1)
    flat = sorted(set(flat))
    out = [flat[0]]
    cur = flat[0]
    for i in range(1, len(flat) - 1):
        if flat[i] - cur >= 5:
            out.append(flat[i])
            cur = flat[i]
    return out


def test_moment_segmentation(sd, batch, text_feat, threshold=0.5, max_iterations=20, return_probs=False):
    """modeling.py:353-474."""
    vis, vmask, asr = batch["vis_feats"], batch["vis_mask"], batch["asr_feats"]
    B, T = vmask.shape
    starts = batch["moment_bound_frames"][:, 0].tolist()
    lasts = batch["moment_bound_frames"][:, 1].tolist()
    moment_mask = torch.zeros(B, T, dtype=torch.long)
    prev_boundary = torch.zeros(B, T, dtype=torch.long)
    for b in range(B):
        moment_mask[b, starts[b]:lasts[b] + 1] = 1
        prev_boundary[b, starts[b]] = 1
    steps = [[[starts[b], starts[b]]] for b in range(B)]
    probs_log = []
    for _ in range(max_iterations):
        feats = moment_shared(sd, vis, text_feat, vmask, moment_mask, asr, boundary_mask=prev_boundary)
        logits = head(sd, feats, "segment_predictor")
        logits[moment_mask == 0] = -torch.finfo(logits.dtype).max
        probs = logits.softmax(dim=1)
        probs_log.append(probs.clone())
        arg = probs.argmax(dim=1)
        for b in range(B):
            sc = probs[b].tolist()
            mi = arg[b].item()
            if sc[mi] < 0.00001:
                continue
            l, r = grow_region(sc, mi, threshold)
            if l == 0 or r == 0:
                continue
            moment_mask[b, l:r + 1] = 0
            prev_boundary[b, l] = 1
            prev_boundary[b, r] = 1
            steps[b].append([l, r])
    pred = [postprocess_steps(steps[b], lasts[b]) for b in range(B)]
    return (pred, probs_log) if return_probs else pred


def trim_feats(x, moment_mask, max_frames=20):
    """modeling.py:529-554: frames inside the moment, truncated to max_frames or repeat-padded."""
    B = x.shape[0]
    out = torch.zeros(B, max_frames, x.shape[-1], dtype=x.dtype)
    for b in range(B):
        z = x[b][moment_mask[b] == 1]
        n = z.shape[0]
        if n > max_frames:
            out[b] = z[:max_frames]
        else:
            for k in range(n):
                out[b, (k * max_frames) // n:((k + 1) * max_frames) // n] = z[k]
    return out

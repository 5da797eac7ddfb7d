"""CPU oracle: fp32 restatement of the reference's EVA-CLIP inference math and retrieval scoring.

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs — never by the product path (hirest_b200/), which has no CPU fallback.

Parity status: the reference ships no tests / golden vectors; this oracle is pinned
instead against OUTPUTS OF THE REFERENCE ITSELF, imported in the build container with the shims in
oracle/ref_shims.py — see oracle/make_golden.py and tests/golden/*.pt (tests/test_oracle_golden.py).

Every function cites the reference lines it restates.  Floating point, fp32 throughout (the reference's
default precision, EVA_clip/eva_clip.py:90).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _ln(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


# ---------------------------------------------------------------------------------------------------
# EVA ViT  (EVA_clip/vit_model.py)
# ---------------------------------------------------------------------------------------------------
def vit_patch_embed(sd, img, cfg, prefix="visual."):
    """PatchEmbed.forward (vit_model.py:200-206) + cls/pos (vit_model.py:330-333) -> [B, T, D]."""
    P = cfg["vision_cfg"]["patch_size"]
    x = F.conv2d(img, sd[prefix + "patch_embed.proj.weight"], sd[prefix + "patch_embed.proj.bias"], stride=P)
    x = x.flatten(2).transpose(1, 2)
    cls = sd[prefix + "cls_token"].expand(x.shape[0], -1, -1)
    return torch.cat((cls, x), dim=1) + sd[prefix + "pos_embed"]


def vit_attention(sd, h, i, cfg, prefix="visual."):
    """Attention.forward (vit_model.py:120-150): q/v-only bias, q*scale, softmax(qk^T)v, proj."""
    v = cfg["vision_cfg"]
    D = v["width"]
    H = D // v["head_width"]
    p = f"{prefix}blocks.{i}.attn."
    B, N, _ = h.shape
    bias = torch.cat((sd[p + "q_bias"], torch.zeros_like(sd[p + "v_bias"]), sd[p + "v_bias"]))  # :124
    qkv = F.linear(h, sd[p + "qkv.weight"], bias).reshape(B, N, 3, H, -1).permute(2, 0, 3, 1, 4)  # :126-127
    q, k, vv = qkv[0], qkv[1], qkv[2]
    q = q * (v["head_width"] ** -0.5)                                                            # :130
    attn = (q @ k.transpose(-2, -1)).softmax(dim=-1)                                             # :131,144
    o = (attn @ vv).transpose(1, 2).reshape(B, N, -1)                                            # :147
    return F.linear(o, sd[p + "proj.weight"], sd[p + "proj.bias"])                               # :148


def vit_block(sd, x, i, cfg, prefix="visual."):
    """Block.forward with gamma=None, DropPath=identity in eval (vit_model.py:175-178)."""
    p = f"{prefix}blocks.{i}."
    x = x + vit_attention(sd, _ln(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-6), i, cfg, prefix)
    h = _ln(x, sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-6)
    h = F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
    h = F.gelu(h)                                                                                # nn.GELU (erf), :47,59
    return x + F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])


def encode_image(sd, img, cfg, prefix="visual.", taps=None):
    """EVA_CLIP.encode_image (eva_model.py:317-318) = VisionTransformer.forward (vit_model.py:326-351).
    taps: optional dict {layer_index: None}; filled with the residual stream after that many blocks."""
    x = vit_patch_embed(sd, img, cfg, prefix)
    if taps is not None and 0 in taps:
        taps[0] = x.clone()
    for i in range(cfg["vision_cfg"]["layers"]):
        x = vit_block(sd, x, i, cfg, prefix)
        if taps is not None and (i + 1) in taps:
            taps[i + 1] = x.clone()
    x = _ln(x, sd[prefix + "norm.weight"], sd[prefix + "norm.bias"], 1e-6)                       # :340
    return F.linear(x[:, 0], sd[prefix + "head.weight"], sd[prefix + "head.bias"])               # :346,350


# ---------------------------------------------------------------------------------------------------
# EVA-CLIP text tower  (EVA_clip/eva_model.py:177-250)
# ---------------------------------------------------------------------------------------------------
def encode_text(sd, text, cfg, prefix="text."):
    """EVA_CLIP.encode_text (eva_model.py:320-321): TextTransformer.forward_features/forward (:232-250)."""
    t = cfg["text_cfg"]
    W, H, L = t["width"], t["heads"], t["layers"]
    dh = W // H
    x = sd[prefix + "token_embedding.weight"][text] + sd[prefix + "positional_embedding"]        # :233-235
    Q, Cn, _ = x.shape
    mask = torch.full((Cn, Cn), float("-inf")).triu_(1)                                          # :224-230
    for i in range(L):
        p = f"{prefix}transformer.resblocks.{i}."
        h = _ln(x, sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], 1e-5)
        # nn.MultiheadAttention(768, 12): packed in_proj (q,k,v all biased), q scaled by dh^-0.5, additive mask
        qkv = F.linear(h, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"])
        q, k, v = qkv.split(W, dim=-1)
        q = q.reshape(Q, Cn, H, dh).transpose(1, 2) * (dh ** -0.5)
        k = k.reshape(Q, Cn, H, dh).transpose(1, 2)
        v = v.reshape(Q, Cn, H, dh).transpose(1, 2)
        a = (q @ k.transpose(-2, -1) + mask).softmax(dim=-1)
        o = (a @ v).transpose(1, 2).reshape(Q, Cn, W)
        x = x + F.linear(o, sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"])        # :157
        h = _ln(x, sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], 1e-5)
        h = F.gelu(F.linear(h, sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"]))
        x = x + F.linear(h, sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])              # :158
    x = _ln(x, sd[prefix + "ln_final.weight"], sd[prefix + "ln_final.bias"], 1e-5)               # :239
    x = x[torch.arange(Q), text.argmax(dim=-1)]                                                  # :243
    return x @ sd[prefix + "text_projection"]                                                    # :249


# ---------------------------------------------------------------------------------------------------
# Retrieval scoring  (inference_video_retrieval.py) and ranking (evaluate.py)
# ---------------------------------------------------------------------------------------------------
def normalize_text(text_embeds):
    """inference_video_retrieval.py:210-212."""
    t = text_embeds.float()
    return t / t.norm(dim=-1, keepdim=True)


def pool_normalize_video(frame_embeds, n_frames):
    """Raw-frame path: view(B, F, E).mean(1) then /= norm (inference_video_retrieval.py:273,283-285)."""
    v = frame_embeds.float().view(-1, n_frames, frame_embeds.shape[-1]).mean(dim=1, keepdim=False)
    return v / v.norm(dim=-1, keepdim=True)


def subsample_cached_features(video_embeds, n_model_frames):
    """Cached-feature path: linspace(0, n-1, F).astype(int) subsample (inference_video_retrieval.py:311-317)."""
    import numpy as np

    n = video_embeds.shape[0]
    ids = torch.from_numpy(np.linspace(0, n - 1, n_model_frames).astype(int))
    return video_embeds[ids]


def cached_video_embedding(video_embeds, n_model_frames):
    """Loop body of the cached-feature path, inference_video_retrieval.py:306-326: subsample (if n_model_frames > 0),
    .float(), mean(dim=0, keepdim=True), /= norm.  Returns [1, E]."""
    if n_model_frames > 0:
        video_embeds = subsample_cached_features(video_embeds, n_model_frames)
    v = video_embeds.float().mean(dim=0, keepdim=True)
    return v / v.norm(dim=-1, keepdim=True)


def similarity(text_hat, video_hat):
    """inference_video_retrieval.py:334 — fp32 matmul, no logit scale."""
    return torch.matmul(text_hat, video_hat.T)


def rank_videos(scores_row, video_names):
    """evaluate.py:58-60: sort (score, name) ascending, then reverse. Returns indices into video_names, best first."""
    order = sorted(range(len(video_names)), key=lambda j: (float(scores_row[j]), video_names[j]))
    return order[::-1]


def recall_at_k(scores, video_names, gt_sets, ks=(1, 5, 10, 50)):
    """evaluate.py:33-81 restricted to the 'all' category: R@k in percent."""
    hits = {k: 0 for k in ks}
    for i in range(scores.shape[0]):
        order = rank_videos(scores[i].tolist(), video_names)
        for k in ks:
            if any(video_names[j] in gt_sets[i] for j in order[:k]):
                hits[k] += 1
    n = max(1, scores.shape[0])
    return {f"R@{k}": hits[k] / n * 100 for k in ks}

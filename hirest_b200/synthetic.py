"""Seeded synthetic weights / frames / token ids in the reference's state_dict layout.

Used by bench.py, __graft_entry__.smoke(), the tests and (re-exported as oracle/weights.py) the CPU oracle.

The reference ships no weights offline (README.md:272-346 are download links), so parity is defined on
seeded random weights.  Key names / shapes follow the reference modules exactly:
  EVA visual : EVA_clip/vit_model.py:256-300 (VisionTransformer.__init__), :66-119 (Attention), :46-55 (Mlp)
  EVA text   : EVA_clip/eva_model.py:177-222 (TextTransformer), :110-141 (ResidualAttentionBlock)
The distribution is OUR choice (documented here, not the reference's init): N(0, 0.02) linear weights with the
reference's depth rescale of proj/fc2 (vit_model.py:302-308), non-trivial biases and LN affine terms so that
every bias / affine code path is exercised.  Generation uses one CPU torch.Generator walked in a fixed key
order, so the same (cfg, seed) gives bit-identical tensors on any box with this torch build.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

EVA_G14 = {
    "embed_dim": 1024,
    "vision_cfg": {"image_size": 224, "layers": 40, "width": 1408, "head_width": 88, "mlp_ratio": 4.3637,
                   "patch_size": 14},
    "text_cfg": {"context_length": 77, "vocab_size": 49408, "width": 768, "heads": 12, "layers": 12},
}

# Small config with the same awkward shapes (head_width 88, mlp_ratio 4.3637, 257 tokens, 77 ctx) that the CPU
# oracle finishes in well under a second; used by the fast parity tests and the committed golden vectors.
EVA_TINY = {
    "embed_dim": 256,
    "vision_cfg": {"image_size": 224, "layers": 3, "width": 352, "head_width": 88, "mlp_ratio": 4.3637,
                   "patch_size": 14},
    "text_cfg": {"context_length": 77, "vocab_size": 1024, "width": 128, "heads": 2, "layers": 2},
}


def _normal(gen, shape, std):
    return torch.randn(shape, generator=gen, dtype=torch.float32, device=gen.device) * std


def _generator(seed: int, device="cpu"):
    """CPU generators give tensors that are bit-identical across boxes (used by the golden vectors); a cuda
    generator is only for benchmarks, where the values need the right distribution but no oracle match."""
    return torch.Generator(device=device).manual_seed(seed)


def make_visual_state_dict(cfg: dict, seed: int = 0, device="cpu") -> "OrderedDict[str, torch.Tensor]":
    v = cfg["vision_cfg"]
    D, L, P = v["width"], v["layers"], v["patch_size"]
    F = int(D * v["mlp_ratio"])
    n_tok = (v["image_size"] // P) ** 2 + 1
    E = cfg["embed_dim"]
    g = _generator(seed, device)
    sd = OrderedDict()
    sd["cls_token"] = _normal(g, (1, 1, D), 0.02)
    sd["pos_embed"] = _normal(g, (1, n_tok, D), 0.02)
    sd["patch_embed.proj.weight"] = _normal(g, (D, 3, P, P), 0.02)
    sd["patch_embed.proj.bias"] = _normal(g, (D,), 0.02)
    for i in range(L):
        p = f"blocks.{i}."
        sd[p + "norm1.weight"] = 1.0 + _normal(g, (D,), 0.05)
        sd[p + "norm1.bias"] = _normal(g, (D,), 0.02)
        sd[p + "attn.q_bias"] = _normal(g, (D,), 0.02)
        sd[p + "attn.v_bias"] = _normal(g, (D,), 0.02)
        sd[p + "attn.qkv.weight"] = _normal(g, (3 * D, D), 0.02)
        sd[p + "attn.proj.weight"] = _normal(g, (D, D), 0.02) / math.sqrt(2.0 * (i + 1))
        sd[p + "attn.proj.bias"] = _normal(g, (D,), 0.02)
        sd[p + "norm2.weight"] = 1.0 + _normal(g, (D,), 0.05)
        sd[p + "norm2.bias"] = _normal(g, (D,), 0.02)
        sd[p + "mlp.fc1.weight"] = _normal(g, (F, D), 0.02)
        sd[p + "mlp.fc1.bias"] = _normal(g, (F,), 0.02)
        sd[p + "mlp.fc2.weight"] = _normal(g, (D, F), 0.02) / math.sqrt(2.0 * (i + 1))
        sd[p + "mlp.fc2.bias"] = _normal(g, (D,), 0.02)
    sd["norm.weight"] = 1.0 + _normal(g, (D,), 0.05)
    sd["norm.bias"] = _normal(g, (D,), 0.02)
    sd["head.weight"] = _normal(g, (E, D), 0.02)
    sd["head.bias"] = _normal(g, (E,), 0.02)
    return sd


def make_text_state_dict(cfg: dict, seed: int = 1, device="cpu") -> "OrderedDict[str, torch.Tensor]":
    t = cfg["text_cfg"]
    W, L, V, C = t["width"], t["layers"], t["vocab_size"], t["context_length"]
    E = cfg["embed_dim"]
    g = _generator(seed, device)
    sd = OrderedDict()
    sd["positional_embedding"] = _normal(g, (C, W), 0.01)
    sd["text_projection"] = _normal(g, (W, E), W ** -0.5)
    sd["logit_scale"] = torch.tensor(math.log(1 / 0.07), dtype=torch.float32, device=device)
    sd["token_embedding.weight"] = _normal(g, (V, W), 0.02)
    proj_std = (W ** -0.5) * ((2 * L) ** -0.5)
    for i in range(L):
        p = f"transformer.resblocks.{i}."
        sd[p + "ln_1.weight"] = 1.0 + _normal(g, (W,), 0.05)
        sd[p + "ln_1.bias"] = _normal(g, (W,), 0.02)
        sd[p + "attn.in_proj_weight"] = _normal(g, (3 * W, W), W ** -0.5)
        sd[p + "attn.in_proj_bias"] = _normal(g, (3 * W,), 0.02)
        sd[p + "attn.out_proj.weight"] = _normal(g, (W, W), proj_std)
        sd[p + "attn.out_proj.bias"] = _normal(g, (W,), 0.02)
        sd[p + "ln_2.weight"] = 1.0 + _normal(g, (W,), 0.05)
        sd[p + "ln_2.bias"] = _normal(g, (W,), 0.02)
        sd[p + "mlp.c_fc.weight"] = _normal(g, (4 * W, W), (2 * W) ** -0.5)
        sd[p + "mlp.c_fc.bias"] = _normal(g, (4 * W,), 0.02)
        sd[p + "mlp.c_proj.weight"] = _normal(g, (W, 4 * W), proj_std)
        sd[p + "mlp.c_proj.bias"] = _normal(g, (W,), 0.02)
    sd["ln_final.weight"] = 1.0 + _normal(g, (W,), 0.05)
    sd["ln_final.bias"] = _normal(g, (W,), 0.02)
    return sd


def make_eva_state_dict(cfg: dict, seed: int = 0, device="cpu") -> "OrderedDict[str, torch.Tensor]":
    """Full EVA_CLIP state dict: 'visual.*' + 'text.*' (EVA_clip/eva_model.py:292-315)."""
    sd = OrderedDict()
    for k, v in make_visual_state_dict(cfg, seed, device).items():
        sd["visual." + k] = v
    for k, v in make_text_state_dict(cfg, seed + 1, device).items():
        sd["text." + k] = v
    return sd


def make_frames(n: int, image_size: int = 224, seed: int = 1, device="cpu") -> torch.Tensor:
    """Synthetic normalised frames [n, 3, S, S] fp32."""
    g = _generator(seed, device)
    return torch.randn((n, 3, image_size, image_size), generator=g, dtype=torch.float32, device=device)


def make_tokens(n: int, cfg: dict, seed: int = 2) -> torch.Tensor:
    """Synthetic CLIP token rows [n, ctx] int64: SOT, 3..20 word ids, EOT (= the row max,
    EVA_clip/eva_model.py:243), zero padding.  SOT/EOT are the two highest ids of the vocabulary, as in
    the real tokenizer (49406/49407 for vocab 49408, EVA_clip/clip.py:218-219)."""
    t = cfg["text_cfg"]
    V, C = t["vocab_size"], t["context_length"]
    sot, eot = V - 2, V - 1
    g = torch.Generator().manual_seed(seed)
    out = torch.zeros((n, C), dtype=torch.int64)
    for i in range(n):
        L = int(torch.randint(3, 21, (1,), generator=g))
        ids = torch.randint(1, V - 2, (L,), generator=g)
        out[i, 0] = sot
        out[i, 1:1 + L] = ids
        out[i, 1 + L] = eot
    return out


def encode_image_flops(cfg) -> float:
    """Algorithmic FLOPs per frame (2*M*N*K), SURVEY.md §8(d): 534.063 GFLOP for EVA-CLIP-g/14."""
    v = cfg["vision_cfg"]
    D, L, P = v["width"], v["layers"], v["patch_size"]
    Fh = int(D * v["mlp_ratio"])
    n = (v["image_size"] // P) ** 2
    T = n + 1
    patch = 2.0 * n * D * (3 * P * P)
    per_layer = 2.0 * T * D * 3 * D + 2.0 * 2 * T * T * D + 2.0 * T * D * D + 2.0 * 2 * T * D * Fh
    head = 2.0 * D * cfg["embed_dim"]
    return patch + L * per_layer + head


# ---------------------------------------------------------------------------------------------------
# MomentModel (modeling.py:18-123) non-CLIP parameters, reference state_dict layout
# ---------------------------------------------------------------------------------------------------
MOMENT_CFG = {"embed_dim": 512, "hidden": 768, "heads": 12, "ffn": 3072, "visual_layers": 2, "decoder_layers": 2,
              "max_pos_visual": 2048, "max_pos_decoder": 512, "vocab": 30522, "asr_dim": 384, "clip_dim": 1024,
              "max_frames_caption": 20, "max_words": 48}


def make_moment_state_dict(seed: int = 3, device="cpu", cfg: dict = MOMENT_CFG) -> "OrderedDict[str, torch.Tensor]":
    """Seeded weights for every non-CLIP parameter of the reference MomentModel (same keys / shapes; the tied decoder
    classifier weight is the word-embedding tensor itself, clip4caption/modules/modeling.py:121-123,137)."""
    g = _generator(seed, device)
    E, Hd, Ff, A, Cd, V = cfg["embed_dim"], cfg["hidden"], cfg["ffn"], cfg["asr_dim"], cfg["clip_dim"], cfg["vocab"]
    sd = OrderedDict()

    def lin(name, out_f, in_f, std=0.03, bias=True):
        sd[name + ".weight"] = _normal(g, (out_f, in_f), std)
        if bias:
            sd[name + ".bias"] = _normal(g, (out_f,), 0.02)

    def ln(name, dim):
        sd[name + ".weight"] = 1.0 + _normal(g, (dim,), 0.05)
        sd[name + ".bias"] = _normal(g, (dim,), 0.02)

    ln("asr_enc_layer.0", A)
    lin("asr_enc_layer.1", E, A)
    lin("temporal_embed.0", E, 1, std=1.0)
    lin("temporal_embed.2", E, E)
    sd["mask_embed.weight"] = _normal(g, (2, E), 0.5)
    sd["boundary_embed.weight"] = _normal(g, (2, E), 0.5)
    sd["moment_conv.0.weight"] = _normal(g, (E, E, 3), 0.02)   # dead parameters (modeling.py:60-74), kept for layout
    sd["moment_conv.0.bias"] = _normal(g, (E,), 0.02)
    sd["moment_conv.2.weight"] = _normal(g, (E, E, 3), 0.02)
    sd["moment_conv.2.bias"] = _normal(g, (E,), 0.02)
    for h in ("start_predictor.0", "end_predictor.0", "segment_predictor.0"):
        lin(h, 1, Hd, std=0.08)
    v = "clip4cap_model.visual."
    lin(v + "embeddings.word_embeddings", Hd, E)
    sd[v + "embeddings.position_embeddings.weight"] = _normal(g, (cfg["max_pos_visual"], Hd), 0.02)
    ln(v + "embeddings.LayerNorm", Hd)
    for i in range(cfg["visual_layers"]):
        p = f"{v}encoder.layer.{i}."
        for n in ("query", "key", "value"):
            lin(p + "attention.self." + n, Hd, Hd)
        lin(p + "attention.output.dense", Hd, Hd)
        ln(p + "attention.output.LayerNorm", Hd)
        lin(p + "intermediate.dense", Ff, Hd)
        lin(p + "output.dense", Hd, Ff)
        ln(p + "output.LayerNorm", Hd)
    lin(v + "pooler.dense", Hd, Hd)   # dead (module_visual.py:421), kept for layout
    d = "clip4cap_model.decoder."
    sd[d + "embeddings.word_embeddings.weight"] = _normal(g, (V, Hd), 0.03)
    sd[d + "embeddings.position_embeddings.weight"] = _normal(g, (cfg["max_pos_decoder"], Hd), 0.02)
    ln(d + "embeddings.LayerNorm", Hd)
    for i in range(cfg["decoder_layers"]):
        p = f"{d}decoder.layer.{i}."
        for att in ("slf_attn", "enc_attn"):
            for n in ("query", "key", "value"):
                lin(f"{p}{att}.att.{n}", Hd, Hd)
            lin(f"{p}{att}.output.dense", Hd, Hd)
            ln(f"{p}{att}.output.LayerNorm", Hd)
        lin(p + "intermediate.dense", Ff, Hd)
        lin(p + "output.dense", Hd, Ff)
        ln(p + "output.LayerNorm", Hd)
    sd[d + "classifier.cls.predictions.bias"] = _normal(g, (V,), 0.02)
    lin(d + "classifier.cls.predictions.transform.dense", Hd, Hd)
    ln(d + "classifier.cls.predictions.transform.LayerNorm", Hd)
    sd[d + "classifier.cls.predictions.decoder.weight"] = sd[d + "embeddings.word_embeddings.weight"]
    ln("clip4cap_model.normalize_video.visual_norm2d", E)
    lin("clip_g_map", E, Cd)
    lin("clip_g_map_text", E, Cd)
    return sd


def _placeholder_ids(B: int):
    """Token rows for batches whose text features come from a stand-in encoder: zeros except a distinct value per sample in slot 1,
    so the samples count as different prompts (MomentModel encodes repeated prompts of a batch once)."""
    ids = torch.zeros((B, 77), dtype=torch.int64)
    ids[:, 1] = torch.arange(B)
    return ids


def make_moment_batch(B: int, T: int, seed: int = 5, ragged: bool = True, cfg: dict = MOMENT_CFG):
    """Synthetic collate output (hirest_dataset.py:409-531): unit-norm frame features (extract_features.py:64), ASR
    features, masks, moment bounds; plus fixed stand-in text features [B, 1024] (what encode_text would return)."""
    g = torch.Generator().manual_seed(seed)
    vis = torch.randn((B, T, cfg["clip_dim"]), generator=g)
    vis = vis / vis.norm(dim=-1, keepdim=True)
    asr = torch.randn((B, T, cfg["asr_dim"]), generator=g)
    n = torch.full((B,), T, dtype=torch.int64)
    if ragged and B > 1:
        n[1:] = torch.randint(max(8, T // 2), T + 1, (B - 1,), generator=g)
    vis_mask = (torch.arange(T)[None, :] < n[:, None]).long()
    vis = vis * vis_mask[..., None]
    asr = asr * vis_mask[..., None]
    bounds = torch.stack([torch.div(n, 8, rounding_mode="floor") + 1, n - torch.div(n, 8, rounding_mode="floor") - 2], dim=1)
    moment_mask = ((torch.arange(T)[None, :] >= bounds[:, :1]) & (torch.arange(T)[None, :] <= bounds[:, 1:])).long()
    text_feat = torch.randn((B, cfg["clip_dim"]), generator=g)
    return {"vis_feats": vis, "vis_mask": vis_mask, "asr_feats": asr, "moment_mask": moment_mask,
            "moment_bound_frames": bounds, "text_feat": text_feat, "n_frames": n,
            "clip_text_ids": _placeholder_ids(B)}


# ---------------------------------------------------------------------------------------------------
# encode_text -> MomentModel chain (tests/golden/chain.pt, oracle/make_golden_chain.py)
# ---------------------------------------------------------------------------------------------------
CHAIN_CLIP = {"embed_dim": 1024, "vision_cfg": dict(EVA_TINY["vision_cfg"]), "text_cfg": dict(EVA_G14["text_cfg"])}


def make_chain_clip_state_dict(device="cpu"):
    """EVA_CLIP state dict for the chain tests: tiny visual tower (MomentModel never calls encode_image), EVA-CLIP-g/14 text
    tower with the same seeded weights as make_eva_state_dict(EVA_G14)['text.*']."""
    sd = OrderedDict()
    for k, v in make_visual_state_dict(CHAIN_CLIP, 0, device).items():
        sd["visual." + k] = v
    for k, v in make_text_state_dict(EVA_G14, 1, device).items():
        sd["text." + k] = v
    return sd


def make_chain_batch(B: int, T: int, seed: int):
    """make_moment_batch with real CLIP token rows (the prompt of each clip) instead of the zero placeholder."""
    batch = make_moment_batch(B, T, seed=seed)
    batch["clip_text_ids"] = make_tokens(B, EVA_G14, seed=seed + 100)
    del batch["text_feat"]
    return batch

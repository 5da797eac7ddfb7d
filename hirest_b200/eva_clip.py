"""Host-side mirror of the reference's EVA-CLIP interface, backed by libhirest_b200.so.

Same names, argument meaning and error behaviour as the reference for this path:
  * ``EVA_CLIP(embed_dim, vision_cfg, text_cfg)``             EVA_clip/eva_model.py:273-315
  * ``.encode_image(image)`` / ``.encode_text(text)``          EVA_clip/eva_model.py:317-321
    (encode_image additionally accepts raw uint8 frames [B,3,224,224]; ToTensor + Normalize then run on the GPU)
  * ``.visual.image_size / image_mean / image_std``            EVA_clip/eva_clip.py:117-118,170
  * ``state_dict()`` keys ``visual.*`` / ``text.*``            SURVEY.md Appendix B (strict load works)
  * ``build_eva_model_and_transforms(model_name, pretrained, precision, device, ...)``  eva_clip.py:155-171

The nn.Module owns the fp32 parameters (so ``.to()``, ``.parameters()``, ``.float()``, ``load_state_dict`` and
``freeze`` loops of modeling.py:120-129 behave as before); the extension holds repacked bf16 copies that are rebuilt
lazily whenever the parameters change.  Inference only (the reference always calls these under
``torch.no_grad()`` in eval mode); there is no CPU / eager fallback.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from copy import deepcopy
from typing import Optional, Tuple

import torch
from torch import nn

from . import _lib

OPENAI_DATASET_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_DATASET_STD = (0.26862954, 0.26130258, 0.27577711)

_MODEL_CONFIGS = {
    # EVA_clip/model_configs/EVA_CLIP_g_14.json
    "EVA_CLIP_g_14": {
        "embed_dim": 1024,
        "vision_cfg": {"image_size": 224, "layers": 40, "width": 1408, "head_width": 88, "mlp_ratio": 4.3637,
                       "patch_size": 14, "drop_path_rate": 0.4},
        "text_cfg": {"context_length": 77, "vocab_size": 49408, "width": 768, "heads": 12, "layers": 12},
    },
}


def list_models():
    return list(_MODEL_CONFIGS.keys())


def get_model_config(model_name):
    return deepcopy(_MODEL_CONFIGS.get(model_name))


def add_model_config(name_or_path, cfg: Optional[dict] = None):
    """Register a config dict, or a json file path as the reference's add_model_config does (eva_clip.py:55-60)."""
    if cfg is None:
        with open(name_or_path) as f:
            cfg = json.load(f)
        name_or_path = os.path.splitext(os.path.basename(str(name_or_path)))[0]
    _MODEL_CONFIGS[name_or_path] = deepcopy(cfg)


class _ParamTree(nn.Module):
    """Holds parameters under dotted names so that state_dict() reproduces the reference key layout."""

    def __init__(self, shapes: dict):
        super().__init__()
        children = {}
        for name, shape in shapes.items():
            head, _, rest = name.partition(".")
            if rest:
                children.setdefault(head, {})[rest] = shape
            else:
                self.register_parameter(head, nn.Parameter(torch.zeros(shape), requires_grad=False))
        for head, sub in children.items():
            self.add_module(head, _ParamTree(sub))


def _visual_shapes(embed_dim: int, v: dict) -> dict:
    D, L, P = v["width"], v["layers"], v["patch_size"]
    F = int(D * v["mlp_ratio"])
    T = (v["image_size"] // P) ** 2 + 1
    s = {"cls_token": (1, 1, D), "pos_embed": (1, T, D), "patch_embed.proj.weight": (D, 3, P, P),
         "patch_embed.proj.bias": (D,)}
    for i in range(L):
        p = f"blocks.{i}."
        s.update({p + "norm1.weight": (D,), p + "norm1.bias": (D,), p + "attn.q_bias": (D,), p + "attn.v_bias": (D,),
                  p + "attn.qkv.weight": (3 * D, D), p + "attn.proj.weight": (D, D), p + "attn.proj.bias": (D,),
                  p + "norm2.weight": (D,), p + "norm2.bias": (D,), p + "mlp.fc1.weight": (F, D), p + "mlp.fc1.bias": (F,),
                  p + "mlp.fc2.weight": (D, F), p + "mlp.fc2.bias": (D,)})
    s.update({"norm.weight": (D,), "norm.bias": (D,), "head.weight": (embed_dim, D), "head.bias": (embed_dim,)})
    return s


def _text_shapes(embed_dim: int, t: dict) -> dict:
    W, L, V, Cn = t["width"], t["layers"], t["vocab_size"], t["context_length"]
    s = {"positional_embedding": (Cn, W), "text_projection": (W, embed_dim), "logit_scale": (),
         "token_embedding.weight": (V, W)}
    for i in range(L):
        p = f"transformer.resblocks.{i}."
        s.update({p + "ln_1.weight": (W,), p + "ln_1.bias": (W,), p + "attn.in_proj_weight": (3 * W, W),
                  p + "attn.in_proj_bias": (3 * W,), p + "attn.out_proj.weight": (W, W), p + "attn.out_proj.bias": (W,),
                  p + "ln_2.weight": (W,), p + "ln_2.bias": (W,), p + "mlp.c_fc.weight": (4 * W, W),
                  p + "mlp.c_fc.bias": (4 * W,), p + "mlp.c_proj.weight": (W, 4 * W), p + "mlp.c_proj.bias": (W,)})
    s.update({"ln_final.weight": (W,), "ln_final.bias": (W,)})
    return s


def _param_key(params):
    """Cache key of an engine's repacked weights: device, storage address and version counter of EVERY parameter.  In-place
    updates bump ``_version``; ``p.data = other`` / ``.to()`` change ``data_ptr()``.  Writes through ``p.data.copy_()`` are invisible
    to both: call ``invalidate()`` on the module after such a write."""
    params = list(params)
    return (params[0].device,) + tuple((q.data_ptr(), q._version) for q in params)


class _Engine:
    """Owns one native handle; destroys it when collected."""

    def __init__(self, handle, destroy):
        self.handle = handle
        self._destroy = destroy

    def __del__(self):
        try:
            if self.handle:
                self._destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class VisionTransformer(_ParamTree):
    """EVA ViT (EVA_clip/vit_model.py:248-351) — parameters here, compute in libhirest_b200.so."""

    def __init__(self, embed_dim: int, cfg: dict, max_batch: int = 1024):
        super().__init__(_visual_shapes(embed_dim, cfg))
        self.cfg = dict(cfg)
        self.embed_dim = embed_dim
        self.image_size = cfg["image_size"]
        self.num_features = cfg["width"]
        self.image_mean = OPENAI_DATASET_MEAN
        self.image_std = OPENAI_DATASET_STD
        self.max_batch = max_batch
        self._engine = None
        self._engine_key = None

    def _key(self):
        return _param_key(self.parameters())

    def invalidate(self):
        """Drop the native handle (repacked bf16 weights); the next forward rebuilds it from the current parameters."""
        self._engine = None
        self._engine_key = None

    def _get_engine(self):
        key = self._key()
        if self._engine is not None and key == self._engine_key:
            return self._engine
        dev = self.cls_token.device
        if dev.type != "cuda":
            raise RuntimeError("hirest_b200: the frame encoder runs on a B200 only; move the model to a cuda device "
                               "(there is no CPU fallback)")
        for q in self.parameters():
            if q.dtype != torch.float32:
                raise RuntimeError("hirest_b200: parameters must be fp32 (the engine keeps its own bf16 copies)")
        lib = _lib.init(dev.index or 0)
        v = self.cfg
        D, L = v["width"], v["layers"]
        cfg = _lib.HbVitConfig(v["image_size"], v["patch_size"], D, L, D // v["head_width"], int(D * v["mlp_ratio"]),
                               self.embed_dim, 1e-6)
        sd = {k: t.detach().contiguous() for k, t in self.state_dict().items()}
        keep = []

        def per_layer(fmt):
            arr = _lib.ptr_array([sd[fmt.format(i)] for i in range(L)])
            keep.append(arr)
            return C.cast(arr, C.c_void_p)

        w = _lib.HbVitWeights(
            sd["cls_token"].data_ptr(), sd["pos_embed"].data_ptr(), sd["patch_embed.proj.weight"].data_ptr(),
            sd["patch_embed.proj.bias"].data_ptr(),
            per_layer("blocks.{}.norm1.weight"), per_layer("blocks.{}.norm1.bias"),
            per_layer("blocks.{}.attn.q_bias"), per_layer("blocks.{}.attn.v_bias"), per_layer("blocks.{}.attn.qkv.weight"),
            per_layer("blocks.{}.attn.proj.weight"), per_layer("blocks.{}.attn.proj.bias"),
            per_layer("blocks.{}.norm2.weight"), per_layer("blocks.{}.norm2.bias"),
            per_layer("blocks.{}.mlp.fc1.weight"), per_layer("blocks.{}.mlp.fc1.bias"),
            per_layer("blocks.{}.mlp.fc2.weight"), per_layer("blocks.{}.mlp.fc2.bias"),
            sd["norm.weight"].data_ptr(), sd["norm.bias"].data_ptr(), sd["head.weight"].data_ptr(),
            sd["head.bias"].data_ptr())
        handle = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(lib.hb_vit_create(C.byref(cfg), C.byref(w), int(self.max_batch), _lib.stream_ptr(dev),
                                         C.byref(handle)), "hb_vit_create")
        self._engine = _Engine(handle, lib.hb_vit_destroy)
        self._engine_key = key
        return self._engine

    @torch.no_grad()
    def forward(self, x: torch.Tensor, tap: Optional[Tuple[int, torch.Tensor]] = None) -> torch.Tensor:
        B, Cc, H, W = x.shape
        # same shape contract as PatchEmbed.forward (vit_model.py:202-204)
        assert H == self.image_size and W == self.image_size, \
            f"Input image size ({H}*{W}) doesn't match model ({self.image_size}*{self.image_size})."
        assert Cc == 3
        eng = self._get_engine()
        dev = self.cls_token.device
        if x.device != dev:
            raise RuntimeError(f"input is on {x.device}, model on {dev}")
        out = torch.empty((B, self.embed_dim), dtype=torch.float32, device=dev)
        lib = _lib.load()
        with torch.cuda.device(dev):
            if tap is not None:
                _lib.check(lib.hb_vit_set_tap(eng.handle, int(tap[0]), tap[1].data_ptr()), "hb_vit_set_tap")
            if x.dtype == torch.uint8:
                # raw frames: ToTensor + Normalize(image_mean, image_std) happen inside the patch-gather kernel
                x = x.contiguous()
                mean = (C.c_float * 3)(*self.image_mean)
                std = (C.c_float * 3)(*self.image_std)
                _lib.check(lib.hb_vit_encode_u8(eng.handle, x.data_ptr(), B, mean, std, out.data_ptr(), _lib.stream_ptr(dev)),
                           "hb_vit_encode_u8")
            else:
                x = x.float().contiguous()
                _lib.check(lib.hb_vit_encode(eng.handle, x.data_ptr(), B, out.data_ptr(), _lib.stream_ptr(dev)), "hb_vit_encode")
            if tap is not None:
                lib.hb_vit_set_tap(eng.handle, -1, None)
        return out


class TextTransformer(_ParamTree):
    """EVA-CLIP text tower (EVA_clip/eva_model.py:177-250)."""

    def __init__(self, embed_dim: int, cfg: dict, max_batch: int = 512, precise: bool = True):
        super().__init__(_text_shapes(embed_dim, cfg))
        self.cfg = dict(cfg)
        self.embed_dim = embed_dim
        self.context_length = cfg["context_length"]
        self.vocab_size = cfg["vocab_size"]
        self.max_batch = max_batch
        # precise=True (default): 3-term split-bf16 GEMMs + fp32 LayerNorm / attention, ~1e-5 of the fp32 reference.  The text
        # feature multiplies every frame feature inside MomentModel (modeling.py:159-165) and decides argmax / 0.5-ratio /
        # beam top-k outcomes there, and the tower is tiny (13.3 GFLOP per query), so accuracy is worth 3x its GEMM work.
        # precise=False: plain bf16 GEMMs + bf16 attention (rel. error ~8e-3), the north_star's bf16 wording.
        self.precise = bool(precise)
        self._engine = None
        self._engine_key = None

    def _key(self):
        return _param_key(self.parameters()) + (self.precise,)

    def invalidate(self):
        self._engine = None
        self._engine_key = None

    def _get_engine(self):
        key = self._key()
        if self._engine is not None and key == self._engine_key:
            return self._engine
        dev = self.positional_embedding.device
        if dev.type != "cuda":
            raise RuntimeError("hirest_b200: the text tower runs on a B200 only (no CPU fallback)")
        for q in self.parameters():
            if q.dtype != torch.float32:
                raise RuntimeError("hirest_b200: parameters must be fp32 (the engine keeps its own bf16 copies)")
        lib = _lib.init(dev.index or 0)
        t = self.cfg
        W, L = t["width"], t["layers"]
        cfg = _lib.HbTextConfig(t["context_length"], t["vocab_size"], W, t["heads"], L, self.embed_dim, 1e-5, 1 if self.precise else 0)
        sd = {k: v.detach().contiguous() for k, v in self.state_dict().items()}
        keep = []

        def per_layer(fmt):
            arr = _lib.ptr_array([sd[fmt.format(i)] for i in range(L)])
            keep.append(arr)
            return C.cast(arr, C.c_void_p)

        r = "transformer.resblocks.{}."
        w = _lib.HbTextWeights(
            sd["token_embedding.weight"].data_ptr(), sd["positional_embedding"].data_ptr(),
            per_layer(r + "ln_1.weight"), per_layer(r + "ln_1.bias"), per_layer(r + "attn.in_proj_weight"),
            per_layer(r + "attn.in_proj_bias"), per_layer(r + "attn.out_proj.weight"), per_layer(r + "attn.out_proj.bias"),
            per_layer(r + "ln_2.weight"), per_layer(r + "ln_2.bias"), per_layer(r + "mlp.c_fc.weight"),
            per_layer(r + "mlp.c_fc.bias"), per_layer(r + "mlp.c_proj.weight"), per_layer(r + "mlp.c_proj.bias"),
            sd["ln_final.weight"].data_ptr(), sd["ln_final.bias"].data_ptr(), sd["text_projection"].data_ptr())
        handle = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(lib.hb_text_create(C.byref(cfg), C.byref(w), int(self.max_batch), _lib.stream_ptr(dev),
                                          C.byref(handle)), "hb_text_create")
        self._engine = _Engine(handle, lib.hb_text_destroy)
        self._engine_key = key
        return self._engine

    def validate_ids(self, text: torch.Tensor) -> None:
        """nn.Embedding raises on out-of-range ids (eva_model.py:233); the device kernel clamps them.  Callers that hold the ids on
        the HOST (MomentModel._inputs, collate output) check them here before the copy -- checking a device tensor would force a
        host-device synchronisation into every encode_text call."""
        if text.numel() and (int(text.min()) < 0 or int(text.max()) >= self.vocab_size):
            raise IndexError(f"token id out of range [0, {self.vocab_size})")

    @torch.no_grad()
    def forward(self, text: torch.Tensor) -> torch.Tensor:
        assert text.dim() == 2 and text.shape[1] == self.context_length, \
            f"expected token ids [n, {self.context_length}], got {tuple(text.shape)}"
        eng = self._get_engine()
        dev = self.positional_embedding.device
        if text.device != dev:
            raise RuntimeError(f"input is on {text.device}, model on {dev}")
        ids = text.to(torch.int64).contiguous()
        out = torch.empty((ids.shape[0], self.embed_dim), dtype=torch.float32, device=dev)
        lib = _lib.load()
        with torch.cuda.device(dev):
            _lib.check(lib.hb_text_encode(eng.handle, ids.data_ptr(), ids.shape[0], out.data_ptr(), _lib.stream_ptr(dev)),
                       "hb_text_encode")
        return out


class EVA_CLIP(nn.Module):
    """Drop-in for EVA_clip/eva_model.py:273-334 (inference surface)."""

    def __init__(self, embed_dim: int, vision_cfg: dict, text_cfg: dict, quick_gelu: bool = False,
                 max_image_batch: int = 1024, max_text_batch: int = 512, precise_text: bool = True):
        super().__init__()
        if quick_gelu:
            raise NotImplementedError("quick_gelu is not used by EVA_CLIP_g_14 (eva_model.py:289)")
        vision_cfg = {k: v for k, v in dict(vision_cfg).items()}
        self.visual = VisionTransformer(embed_dim, vision_cfg, max_batch=max_image_batch)
        self.text = TextTransformer(embed_dim, dict(text_cfg), max_batch=max_text_batch, precise=precise_text)

    def encode_image(self, image):
        return self.visual(image)

    def encode_text(self, text):
        return self.text(text)

    def forward(self, image, text):
        if image is None:
            return self.encode_text(text)
        elif text is None:
            return self.encode_image(image)
        image_features = torch.nn.functional.normalize(self.encode_image(image), dim=-1)
        text_features = torch.nn.functional.normalize(self.encode_text(text), dim=-1)
        return image_features, text_features, self.text.logit_scale.exp()


def load_state_dict(checkpoint_path: str, map_location: str = "cpu", model_key="model|module|state_dict"):
    """Checkpoint unwrapping rules of eva_clip.py:68-79."""
    checkpoint = torch.load(checkpoint_path, map_location=map_location)
    state_dict = checkpoint
    for mk in model_key.split("|"):
        if isinstance(checkpoint, dict) and mk in checkpoint:
            state_dict = checkpoint[mk]
            break
    if next(iter(state_dict.items()))[0].startswith("module"):
        state_dict = {k[7:]: v for k, v in state_dict.items()}
    return state_dict


def create_model(model_name: str, pretrained: str = "", precision: str = "fp32",
                 device: torch.device = torch.device("cpu"), force_quick_gelu: bool = False):
    model_name = model_name.replace("/", "-")
    cfg = get_model_config(model_name)
    if cfg is None:
        raise RuntimeError(f"Model config for {model_name} not found.")
    if precision != "fp32":
        raise RuntimeError("hirest_b200 keeps fp32 master weights and computes in bf16; pass precision='fp32'")
    model = EVA_CLIP(**cfg)
    # the reference always loads a checkpoint here, strictly (eva_clip.py:81-85,109)
    incompatible = model.load_state_dict(load_state_dict(pretrained), strict=True)
    print(incompatible)
    model.to(device=device)
    model.visual.image_mean = OPENAI_DATASET_MEAN
    model.visual.image_std = OPENAI_DATASET_STD
    return model


def image_transform(image_size: int, mean=None, std=None):
    """Same torchvision pipeline as eva_clip.py:125-153 (CPU preprocessing is outside the hot path)."""
    from torchvision.transforms import CenterCrop, Compose, InterpolationMode, Normalize, Resize, ToTensor

    mean = mean or OPENAI_DATASET_MEAN
    std = std or OPENAI_DATASET_STD
    if not isinstance(mean, (list, tuple)):
        mean = (mean,) * 3
    if not isinstance(std, (list, tuple)):
        std = (std,) * 3
    if isinstance(image_size, (list, tuple)) and image_size[0] == image_size[1]:
        image_size = image_size[0]
    return Compose([Resize(image_size, interpolation=InterpolationMode.BICUBIC), CenterCrop(image_size),
                    lambda im: im.convert("RGB"), ToTensor(), Normalize(mean=mean, std=std)])


def build_eva_model_and_transforms(model_name: str, pretrained: str = "", precision: str = "fp32",
                                   device: torch.device = torch.device("cpu"), force_quick_gelu: bool = False,
                                   image_mean: Optional[Tuple[float, ...]] = None,
                                   image_std: Optional[Tuple[float, ...]] = None):
    model = create_model(model_name, pretrained, precision, device, force_quick_gelu=force_quick_gelu)
    image_mean = image_mean or getattr(model.visual, "image_mean", None)
    image_std = image_std or getattr(model.visual, "image_std", None)
    preprocess_val = image_transform(model.visual.image_size, mean=image_mean, std=image_std)
    return model, preprocess_val

"""Import shims that let the UNMODIFIED reference (/root/reference) be imported in the build container.

TEST INFRASTRUCTURE ONLY; used by oracle/make_golden.py.  /root/reference does not exist on the GPU box, so
nothing here is imported by tests marked gpu, smoke() or bench.py.

Missing third-party packages carry no arithmetic on the hot path:
  timm.models.layers.{drop_path (identity in eval), to_2tuple, trunc_normal_}, timm.models.registry.register_model,
  tkinter.E (stray import, EVA_clip/eva_clip.py:8).
"""
import sys
import types

import torch

REFERENCE_ROOT = "/root/reference"


def install():
    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    if "timm" not in sys.modules:
        timm = mod("timm")
        models = mod("timm.models")
        layers = mod("timm.models.layers")
        registry = mod("timm.models.registry")
        timm.models = models
        models.layers = layers
        models.registry = registry

        def drop_path(x, drop_prob: float = 0.0, training: bool = False):
            if drop_prob == 0.0 or not training:
                return x
            raise RuntimeError("drop_path shim is inference-only")

        layers.drop_path = drop_path
        layers.to_2tuple = lambda x: x if isinstance(x, (tuple, list)) else (x, x)
        layers.trunc_normal_ = torch.nn.init.trunc_normal_
        registry.register_model = lambda f: f
    try:
        import tkinter  # noqa: F401
    except Exception:
        tk = mod("tkinter")
        tk.E = "e"
    eva_dir = REFERENCE_ROOT + "/EVA_clip"
    if eva_dir not in sys.path:
        sys.path.insert(0, eva_dir)


def reference_eva_clip(cfg):
    """Instantiate the reference's EVA_CLIP (EVA_clip/eva_model.py:273) for a config dict."""
    install()
    import eva_model  # noqa: E402  (from /root/reference/EVA_clip)

    return eva_model.EVA_CLIP(**cfg).eval()

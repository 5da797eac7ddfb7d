"""Bring-up of the UNMODIFIED reference MomentModel in the build container.

TEST INFRASTRUCTURE ONLY; used by oracle/make_golden_moment.py.  Works on a writable copy of /root/reference (the model
resolves ./pretrained_weights, ./clip4caption, ./EVA_clip relative to the CWD, modeling.py:10,102,115-117), with:
  * stubs for packages that carry no arithmetic on the path (timm, tkinter, kornia, boto3, pycocoevalcap, ...),
  * synthetic fixtures: BERT vocab (30522 lines, real special-token ids) + bert_config.json, an empty clip4caption
    checkpoint (=> the model's own random init), and a stand-in for the 4.5 GB EVA-CLIP build (MomentModel only calls
    clip_model.encode_text, which the golden script replaces by fixed text features).
"""
import json
import os
import shutil
import sys
import types

import torch

REFERENCE_ROOT = "/root/reference"


def _mod(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def prepare_copy(dst="/tmp/hirest_ref_copy"):
    if not os.path.exists(os.path.join(dst, "modeling.py")):
        shutil.copytree(REFERENCE_ROOT, dst, dirs_exist_ok=True)
    bert_dir = os.path.join(dst, "clip4caption", "modules", "bert-base-uncased")
    os.makedirs(bert_dir, exist_ok=True)
    vocab = [f"[unused{i}]" for i in range(30522)]
    vocab[0], vocab[100], vocab[101], vocab[102], vocab[103] = "[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"
    for i in range(1000, 30522):
        vocab[i] = f"w{i}"
    with open(os.path.join(bert_dir, "vocab.txt"), "w") as f:
        f.write("\n".join(vocab) + "\n")
    with open(os.path.join(bert_dir, "bert_config.json"), "w") as f:
        json.dump({"attention_probs_dropout_prob": 0.1, "hidden_act": "gelu", "hidden_dropout_prob": 0.1, "hidden_size": 768,
                   "initializer_range": 0.02, "intermediate_size": 3072, "max_position_embeddings": 512,
                   "num_attention_heads": 12, "num_hidden_layers": 12, "type_vocab_size": 2, "vocab_size": 30522}, f)
    os.makedirs(os.path.join(dst, "pretrained_weights"), exist_ok=True)
    torch.save({}, os.path.join(dst, "pretrained_weights", "clip4caption_vit-b-32_model.bin"))
    return dst


def install_stubs():
    from oracle import ref_shims

    ref_shims.install()
    _mod("kornia")
    _mod("boto3")
    bc = _mod("botocore")
    _mod("botocore.exceptions", ClientError=type("ClientError", (Exception,), {}))
    bc.exceptions = sys.modules["botocore.exceptions"]
    for pkg, cls in (("bleu", "Bleu"), ("rouge", "Rouge"), ("cider", "Cider"), ("meteor", "Meteor")):
        _mod("pycocoevalcap") if "pycocoevalcap" not in sys.modules else None
        _mod(f"pycocoevalcap.{pkg}")
        _mod(f"pycocoevalcap.{pkg}.{pkg}", **{cls: type(cls, (), {})})


class _StubClip(torch.nn.Module):
    def encode_text(self, ids):
        raise RuntimeError("replace encode_text before use")


def build_reference_moment_model(num_beams=3, clip_cfg=None, asr_dim=384):
    """clip_cfg=None: clip_model is a stub whose encode_text the caller replaces by fixed features.  clip_cfg=dict: clip_model
    is the reference's own EVA_CLIP(**clip_cfg) (EVA_clip/eva_model.py:273, unmodified code, random init -- the caller loads
    seeded weights), so test_step runs the reference text tower on real clip_text_ids (modeling.py:286,364,568)."""
    dst = prepare_copy()
    install_stubs()
    os.chdir(dst)
    for p in (dst, os.path.join(dst, "clip4caption"), os.path.join(dst, "EVA_clip")):
        if p not in sys.path:
            sys.path.insert(0, p)
    # stand-in for the EVA build (eva_clip.build_eva_model_and_transforms, modeling.py:117): the checkpoint file it would read
    # (pretrained_weights/eva_clip_psz14.pt, 4.5 GB) does not exist offline
    if clip_cfg is None:
        build = lambda *a, **k: (_StubClip(), None)  # noqa: E731
    else:
        import eva_model  # /root/reference/EVA_clip/eva_model.py

        build = lambda *a, **k: (eva_model.EVA_CLIP(**clip_cfg), None)  # noqa: E731
    _mod("eva_clip", build_eva_model_and_transforms=build)
    import args as ref_args
    import modeling as ref_modeling

    a = ref_args.get_parser().parse_args(["--data_dir", "x", "--video_feature_dir", "x", "--num_beams", str(num_beams)])
    torch.manual_seed(0)
    model = ref_modeling.MomentModel(n_frames=-1, asr_dim=asr_dim, args=a).eval()
    return model, a

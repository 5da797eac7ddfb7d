"""Host-side mirror of the reference's joint model for the inference path (modeling.py:18-632), backed by libhirest_b200.so.

Same surface as the reference for this path: ``MomentModel(n_frames, asr_dim, args)``, ``.test_step(batch, **kwargs)``
returning ``{'prediction': list}``, ``.freeze_clip()``, ``state_dict()`` with the reference's key layout (``clip_model.*`` plus
the non-CLIP keys of SURVEY.md Appendix B, so ``HiREST_BEST.pth`` loads with ``strict=False`` exactly as trainer_base.py:128-147
does).  Batches are the collate dicts of hirest_dataset.py:409-531 (CPU tensors; moved to the model's device here, as
modeling.py:275-286 does).

Covers moment retrieval, moment segmentation and step captioning (KV-cached beam decoder on the device); training is out of
scope and raises.  No CPU / eager fallback.
"""
from __future__ import annotations

import ctypes as C
from copy import deepcopy
from types import SimpleNamespace

import torch
from torch import nn

from . import _lib
from .eva_clip import _Engine, _ParamTree, _param_key

_DEFAULTS = dict(embed_dim=512, hidden=768, heads=12, ffn=3072, max_pos_visual=2048, max_pos_decoder=512, vocab=30522, clip_dim=1024)


def _moment_shapes(asr_dim: int, visual_layers: int, decoder_layers: int, d=_DEFAULTS) -> dict:
    E, Hd, Ff, V, Cd = d["embed_dim"], d["hidden"], d["ffn"], d["vocab"], d["clip_dim"]
    s = {}

    def lin(n, o, i):
        s[n + ".weight"] = (o, i)
        s[n + ".bias"] = (o,)

    def ln(n, k):
        s[n + ".weight"] = (k,)
        s[n + ".bias"] = (k,)

    if asr_dim > 0:
        ln("asr_enc_layer.0", asr_dim)
        lin("asr_enc_layer.1", E, asr_dim)
    lin("temporal_embed.0", E, 1)
    lin("temporal_embed.2", E, E)
    s["mask_embed.weight"] = (2, E)
    s["boundary_embed.weight"] = (2, E)
    for c in ("0", "2"):
        s[f"moment_conv.{c}.weight"] = (E, E, 3)
        s[f"moment_conv.{c}.bias"] = (E,)
    for h in ("start_predictor.0", "end_predictor.0", "segment_predictor.0"):
        lin(h, 1, Hd)
    v = "clip4cap_model.visual."
    lin(v + "embeddings.word_embeddings", Hd, E)
    s[v + "embeddings.position_embeddings.weight"] = (d["max_pos_visual"], Hd)
    ln(v + "embeddings.LayerNorm", Hd)
    for i in range(visual_layers):
        p = f"{v}encoder.layer.{i}."
        for n in ("query", "key", "value"):
            lin(p + "attention.self." + n, Hd, Hd)
        lin(p + "attention.output.dense", Hd, Hd)
        ln(p + "attention.output.LayerNorm", Hd)
        lin(p + "intermediate.dense", Ff, Hd)
        lin(p + "output.dense", Hd, Ff)
        ln(p + "output.LayerNorm", Hd)
    lin(v + "pooler.dense", Hd, Hd)
    dd = "clip4cap_model.decoder."
    s[dd + "embeddings.word_embeddings.weight"] = (V, Hd)
    s[dd + "embeddings.position_embeddings.weight"] = (d["max_pos_decoder"], Hd)
    ln(dd + "embeddings.LayerNorm", Hd)
    for i in range(decoder_layers):
        p = f"{dd}decoder.layer.{i}."
        for att in ("slf_attn", "enc_attn"):
            for n in ("query", "key", "value"):
                lin(f"{p}{att}.att.{n}", Hd, Hd)
            lin(f"{p}{att}.output.dense", Hd, Hd)
            ln(f"{p}{att}.output.LayerNorm", Hd)
        lin(p + "intermediate.dense", Ff, Hd)
        lin(p + "output.dense", Hd, Ff)
        ln(p + "output.LayerNorm", Hd)
    s[dd + "classifier.cls.predictions.bias"] = (V,)
    lin(dd + "classifier.cls.predictions.transform.dense", Hd, Hd)
    ln(dd + "classifier.cls.predictions.transform.LayerNorm", Hd)
    s[dd + "classifier.cls.predictions.decoder.weight"] = (V, Hd)
    ln("clip4cap_model.normalize_video.visual_norm2d", E)
    lin("clip_g_map", E, Cd)
    lin("clip_g_map_text", E, Cd)
    return s


def default_args(**over):
    """The reference's argparse defaults that this path reads (args.py:51-61)."""
    a = dict(max_frames_step_captioning=20, max_words=48, visual_num_hidden_layers=2, decoder_num_hidden_layers=2,
             moment_segmentation_difference_threshold=0.50, moment_segmentation_max_iterations=20, num_beams=5,
             eva_clip_path="./pretrained_weights/eva_clip_psz14.pt")
    a.update(over)
    return SimpleNamespace(**a)


class MomentModel(nn.Module):
    def __init__(self, n_frames=-1, asr_dim=-1, args=None, clip_model=None, max_rows: int = 64 * 300, max_batch: int = 64):
        super().__init__()
        self.args = args if args is not None else default_args()
        self.n_frames = n_frames
        self.asr_dim = asr_dim
        self.use_asr = asr_dim > 0   # modeling.py:28-35: asr_dim <= 0 builds no asr_enc_layer and ignores batch['asr_feats']
        vl = getattr(self.args, "visual_num_hidden_layers", 2)
        dl = getattr(self.args, "decoder_num_hidden_layers", 2)
        self.visual_layers = vl
        for head, sub in _split_top(_moment_shapes(asr_dim, vl, dl)).items():
            self.add_module(head, _ParamTree(sub))
        # mutate args like the reference does (modeling.py:103-105)
        self.args.d_model = 512
        self.args.video_dim = 512
        self.args.max_frames = getattr(self.args, "max_frames_step_captioning", 20)
        if clip_model is None:
            from .eva_clip import build_eva_model_and_transforms

            clip_model, self.clip_preprocess = build_eva_model_and_transforms(
                "EVA_CLIP_g_14", pretrained=getattr(self.args, "eva_clip_path", "./pretrained_weights/eva_clip_psz14.pt"))
        self.clip_model = clip_model
        self.max_rows, self.max_batch = max_rows, max_batch
        self.dedup_prompts = True     # encode each distinct prompt of a batch once (same features)
        self.ms_early_exit = True     # stop the segmentation loop after an iteration that accepted no step (same predictions)
        self.ms_iterations_run = 0    # forwards of the last test_moment_segmentation call
        self._engine = None
        self._engine_key = None
        self.freeze_clip()
        self.eval()

    # ------------------------------------------------------------------ reference surface
    def freeze_clip(self):
        if isinstance(self.clip_model, nn.Module):
            for p in self.clip_model.parameters():
                p.requires_grad = False
            self.clip_model.eval()

    def train_step(self, batch):
        raise NotImplementedError("hirest_b200 implements the inference path only (training is out of scope, SURVEY.md §2)")

    def test_step(self, batch, **kwargs):
        task = batch["tasks"][0]
        if task == "moment_retrieval":
            return self.test_moment_retrieval(batch, **kwargs)
        elif task == "moment_segmentation":
            return self.test_moment_segmentation(batch, **kwargs)
        elif task == "step_captioning":
            return self.test_step_captioning(batch, **kwargs)
        else:
            raise NotImplementedError

    # ------------------------------------------------------------------ engine
    def _own_params(self):
        return [p for n, p in self.named_parameters() if not n.startswith("clip_model.")]

    def _device(self):
        return self.clip_g_map.weight.device

    def _get_engine(self, rows: int, batch: int):
        ps = self._own_params()
        key = _param_key(ps) + (self.max_rows, self.max_batch)
        if self._engine is not None and key == self._engine_key and rows <= self.max_rows and batch <= self.max_batch:
            return self._engine
        self.max_rows, self.max_batch = max(self.max_rows, rows), max(self.max_batch, batch)
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("hirest_b200: MomentModel runs on a B200 only; move it to a cuda device (no CPU fallback)")
        lib = _lib.init(dev.index or 0)
        sd = {k: v.detach().float().contiguous() for k, v in self.state_dict().items() if not k.startswith("clip_model.")}
        L = self.visual_layers
        keep = []
        head_w = torch.cat([sd["start_predictor.0.weight"], sd["end_predictor.0.weight"], sd["segment_predictor.0.weight"]], 0).contiguous()
        head_b = torch.cat([sd["start_predictor.0.bias"], sd["end_predictor.0.bias"], sd["segment_predictor.0.bias"]], 0).contiguous()
        keep += [head_w, head_b]

        def per_layer(fmt):
            arr = _lib.ptr_array([sd[fmt.format(i)] for i in range(L)])
            keep.append(arr)
            return C.cast(arr, C.c_void_p)

        v = "clip4cap_model.visual."
        e = v + "encoder.layer.{}."
        asr_ptrs = [sd[k].data_ptr() if self.use_asr else None
                    for k in ("asr_enc_layer.0.weight", "asr_enc_layer.0.bias", "asr_enc_layer.1.weight", "asr_enc_layer.1.bias")]
        w = _lib.HbMomentWeights(
            *asr_ptrs, sd["temporal_embed.0.weight"].data_ptr(), sd["temporal_embed.0.bias"].data_ptr(),
            sd["temporal_embed.2.weight"].data_ptr(), sd["temporal_embed.2.bias"].data_ptr(), sd["mask_embed.weight"].data_ptr(),
            sd["boundary_embed.weight"].data_ptr(), head_w.data_ptr(), head_b.data_ptr(),
            sd["clip4cap_model.normalize_video.visual_norm2d.weight"].data_ptr(),
            sd["clip4cap_model.normalize_video.visual_norm2d.bias"].data_ptr(), sd["clip_g_map.weight"].data_ptr(),
            sd["clip_g_map.bias"].data_ptr(), sd["clip_g_map_text.weight"].data_ptr(), sd["clip_g_map_text.bias"].data_ptr(),
            sd[v + "embeddings.word_embeddings.weight"].data_ptr(), sd[v + "embeddings.word_embeddings.bias"].data_ptr(),
            sd[v + "embeddings.position_embeddings.weight"].data_ptr(), sd[v + "embeddings.LayerNorm.weight"].data_ptr(),
            sd[v + "embeddings.LayerNorm.bias"].data_ptr(),
            per_layer(e + "attention.self.query.weight"), per_layer(e + "attention.self.query.bias"),
            per_layer(e + "attention.self.key.weight"), per_layer(e + "attention.self.key.bias"),
            per_layer(e + "attention.self.value.weight"), per_layer(e + "attention.self.value.bias"),
            per_layer(e + "attention.output.dense.weight"), per_layer(e + "attention.output.dense.bias"),
            per_layer(e + "attention.output.LayerNorm.weight"), per_layer(e + "attention.output.LayerNorm.bias"),
            per_layer(e + "intermediate.dense.weight"), per_layer(e + "intermediate.dense.bias"),
            per_layer(e + "output.dense.weight"), per_layer(e + "output.dense.bias"),
            per_layer(e + "output.LayerNorm.weight"), per_layer(e + "output.LayerNorm.bias"))
        d = _DEFAULTS
        cfg = _lib.HbMomentConfig(d["embed_dim"], d["hidden"], d["heads"], d["ffn"], L, max(0, self.asr_dim), d["clip_dim"], d["max_pos_visual"])
        handle = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(lib.hb_moment_create(C.byref(cfg), C.byref(w), int(self.max_rows), int(self.max_batch), _lib.stream_ptr(dev),
                                            C.byref(handle)), "hb_moment_create")
        self._engine = _Engine(handle, lib.hb_moment_destroy)
        self._engine_key = _param_key(ps) + (self.max_rows, self.max_batch)
        return self._engine

    @torch.no_grad()
    def _forward(self, video, text_feat, asr, vmask, mmask, bmask=None, reuse_base=False, want_feats=False):
        """foward_moment_shared (modeling.py:155-210) + heads -> (logits [B,T,3] = start/end/segment, feats [B,T,768] or None)."""
        B, T, _ = video.shape
        eng = self._get_engine(B * T, B)
        dev = self._device()
        logits = torch.empty((B, T, 3), dtype=torch.float32, device=dev)
        feats = torch.empty((B, T, _DEFAULTS["hidden"]), dtype=torch.float32, device=dev) if want_feats else None
        lib = _lib.load()
        with torch.cuda.device(dev):
            _lib.check(lib.hb_moment_forward(eng.handle, video.data_ptr(), text_feat.data_ptr(),
                                             asr.data_ptr() if asr is not None else None, vmask.data_ptr(),
                                             mmask.data_ptr(), bmask.data_ptr() if bmask is not None else None, B, T,
                                             1 if reuse_base else 0, feats.data_ptr() if want_feats else None, logits.data_ptr(),
                                             _lib.stream_ptr(dev)), "hb_moment_forward")
        return logits, feats

    def _inputs(self, batch):
        dev = self._device()
        video = batch["vis_feats"].to(dev).float().contiguous()
        vmask = batch["vis_mask"].to(dev).long().contiguous()
        asr = batch["asr_feats"].to(dev).float().contiguous() if self.use_asr else None
        ids = batch["clip_text_ids"]
        if ids.device.type == "cpu" and hasattr(getattr(self.clip_model, "text", None), "validate_ids"):
            self.clip_model.text.validate_ids(ids)   # host-side range check, as nn.Embedding would raise (no device sync)
        # Rows of one batch often repeat a prompt (every step of a video is captioned under the video's prompt, several videos answer
        # one query): each distinct token row goes through the text tower once and the features are gathered back.  A row's embedding
        # does not depend on what else is in the batch (bit-identical, tests/test_gpu_chain.py), so this only removes recomputation.
        inv = None
        if self.dedup_prompts and ids.device.type == "cpu" and ids.dim() == 2 and ids.shape[0] > 1:
            uniq, inverse = torch.unique(ids, dim=0, return_inverse=True)
            if uniq.shape[0] < ids.shape[0]:
                ids, inv = uniq, inverse.to(dev)
        text_feat = self.clip_model.encode_text(ids.to(dev)).float()
        if inv is not None:
            text_feat = text_feat.index_select(0, inv)
        return video, vmask, asr, text_feat.contiguous()

    def foward_moment_shared(self, video_feats, text_feat, video_mask=None, moment_mask=None, asr_feats=None, boundary_mask=None):
        """Same (misspelt) name and argument order as modeling.py:155."""
        dev = self._device()
        B, T, _ = video_feats.shape
        if video_mask is None:
            video_mask = torch.ones((B, T), dtype=torch.long, device=dev)
        _, feats = self._forward(video_feats.to(dev).float().contiguous(), text_feat.to(dev).float().contiguous(),
                                 asr_feats.to(dev).float().contiguous() if self.use_asr else None, video_mask.to(dev).long().contiguous(),
                                 moment_mask.to(dev).long().contiguous(),
                                 boundary_mask.to(dev).long().contiguous() if boundary_mask is not None else None, want_feats=True)
        return feats

    @torch.no_grad()
    def test_moment_retrieval(self, batch, **kwargs):
        """modeling.py:272-310."""
        dev = self._device()
        video, vmask, asr, text_feat = self._inputs(batch)
        mmask = batch["moment_mask"].to(dev).long().contiguous()
        logits, _ = self._forward(video, text_feat, asr, vmask, mmask)
        B, T = vmask.shape
        pred = torch.empty((B, 2), dtype=torch.int64, device=dev)
        lib = _lib.load()
        with torch.cuda.device(dev):
            _lib.check(lib.hb_moment_mr_decode(logits.data_ptr(), vmask.data_ptr(), pred.data_ptr(), B, T, _lib.stream_ptr(dev)),
                       "hb_moment_mr_decode")
        return {"prediction": pred.tolist()}

    @torch.no_grad()
    def test_moment_segmentation(self, batch, threshold=0.15, **kwargs):
        """modeling.py:353-474: <= 20 shared forwards, each followed by the on-device region-growing step; one D2H at the end."""
        dev = self._device()
        video, vmask, asr, text_feat = self._inputs(batch)
        B, T = vmask.shape
        starts = batch["moment_bound_frames"][:, 0].tolist()
        lasts = batch["moment_bound_frames"][:, 1].tolist()
        idx = torch.arange(T)[None, :]
        b0 = batch["moment_bound_frames"][:, :1].cpu()
        b1 = batch["moment_bound_frames"][:, 1:2].cpu()
        mmask = ((idx >= b0) & (idx <= b1)).long().to(dev).contiguous()       # :376-378
        bmask = (idx == b0).long().to(dev).contiguous()                       # :380-382
        n_iter = self.args.moment_segmentation_max_iterations
        thr = float(self.args.moment_segmentation_difference_threshold)
        steps = torch.zeros((B, n_iter, 2), dtype=torch.int32, device=dev)
        nsteps = torch.zeros((B,), dtype=torch.int32, device=dev)
        lib = _lib.load()
        # An iteration that accepts no step for any clip leaves both masks as they were, so every later iteration would recompute
        # exactly the same logits and accept nothing again (the reference still runs all n_iter forwards, modeling.py:391): the loop
        # stops there.  The accepted-step total of iteration i is read back one iteration late, while iteration i + 1 is already
        # running, so the host never waits on an idle GPU; at most one no-op iteration is enqueued.
        totals = torch.zeros(n_iter, dtype=torch.int32).pin_memory() if self.ms_early_exit else None
        events = []
        for it in range(n_iter):
            logits, _ = self._forward(video, text_feat, asr, vmask, mmask, bmask, reuse_base=(it > 0))
            with torch.cuda.device(dev):
                _lib.check(lib.hb_moment_ms_step(logits.data_ptr(), mmask.data_ptr(), bmask.data_ptr(), steps.data_ptr(),
                                                 nsteps.data_ptr(), n_iter, B, T, thr, None, _lib.stream_ptr(dev)), "hb_moment_ms_step")
                if totals is not None:
                    totals[it:it + 1].copy_(nsteps.sum(dim=0, keepdim=True, dtype=torch.int32), non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record()
                    events.append(ev)
                    if it >= 1:
                        events[it - 1].synchronize()
                        if int(totals[it - 1]) == (int(totals[it - 2]) if it >= 2 else 0):
                            break
        self.ms_iterations_run = it + 1
        steps_h, n_h = steps.cpu().tolist(), nsteps.cpu().tolist()
        preds = []
        for b in range(B):
            sp = [[starts[b], starts[b]]] + [list(x) for x in steps_h[b][:n_h[b]]]
            preds.append(_postprocess_steps(sp, lasts[b]))
        return {"raw_predictions": deepcopy(preds), "prediction": preds}

    # ------------------------------------------------------------------ step captioning
    def _get_decoder(self, n_inst: int, beam: int):
        ps = self._own_params()
        key = _param_key(ps)
        cap = getattr(self, "_dec_cap", (0, 0))
        if getattr(self, "_dec_engine", None) is not None and key == self._dec_key and n_inst <= cap[0] and beam <= cap[1]:
            return self._dec_engine
        cap = (max(cap[0], n_inst, 16), max(cap[1], beam, 5))
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("hirest_b200: the caption decoder runs on a B200 only (no CPU fallback)")
        lib = _lib.init(dev.index or 0)
        sd = {k: v.detach().float().contiguous() for k, v in self.state_dict().items() if k.startswith("clip4cap_model.decoder.")}
        L = getattr(self.args, "decoder_num_hidden_layers", 2)
        keep = []

        def per_layer(fmt):
            arr = _lib.ptr_array([sd[fmt.format(i)] for i in range(L)])
            keep.append(arr)
            return C.cast(arr, C.c_void_p)

        d = "clip4cap_model.decoder."
        l_ = d + "decoder.layer.{}."
        c_ = d + "classifier.cls.predictions."
        fields = [sd[d + "embeddings.word_embeddings.weight"].data_ptr(), sd[d + "embeddings.position_embeddings.weight"].data_ptr(),
                  sd[d + "embeddings.LayerNorm.weight"].data_ptr(), sd[d + "embeddings.LayerNorm.bias"].data_ptr()]
        for att in ("slf_attn", "enc_attn"):
            for n in ("query", "key", "value"):
                fields += [per_layer(f"{l_}{att}.att.{n}.weight"), per_layer(f"{l_}{att}.att.{n}.bias")]
            fields += [per_layer(f"{l_}{att}.output.dense.weight"), per_layer(f"{l_}{att}.output.dense.bias"),
                       per_layer(f"{l_}{att}.output.LayerNorm.weight"), per_layer(f"{l_}{att}.output.LayerNorm.bias")]
        fields += [per_layer(l_ + "intermediate.dense.weight"), per_layer(l_ + "intermediate.dense.bias"),
                   per_layer(l_ + "output.dense.weight"), per_layer(l_ + "output.dense.bias"),
                   per_layer(l_ + "output.LayerNorm.weight"), per_layer(l_ + "output.LayerNorm.bias"),
                   sd[c_ + "transform.dense.weight"].data_ptr(), sd[c_ + "transform.dense.bias"].data_ptr(),
                   sd[c_ + "transform.LayerNorm.weight"].data_ptr(), sd[c_ + "transform.LayerNorm.bias"].data_ptr(), sd[c_ + "bias"].data_ptr()]
        w = _lib.HbDecoderWeights(*fields)
        dd = _DEFAULTS
        cfg = _lib.HbDecoderConfig(dd["hidden"], dd["heads"], dd["ffn"], L, dd["vocab"], dd["max_pos_decoder"], int(self.args.max_words),
                                   self.BOS, self.EOS)
        handle = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(lib.hb_decoder_create(C.byref(cfg), C.byref(w), cap[0], cap[1], int(self.args.max_frames), _lib.stream_ptr(dev),
                                             C.byref(handle)), "hb_decoder_create")
        self._dec_engine = _Engine(handle, lib.hb_decoder_destroy)
        self._dec_key, self._dec_cap = key, cap
        return self._dec_engine

    BOS, EOS, PAD = 101, 102, 0  # [CLS] / [SEP] / [PAD] of the BERT vocabulary (beam.py:22-29 via the tokenizer)

    @torch.no_grad()
    def generate_caption_ids(self, feats: torch.Tensor, num_beams: int):
        """Beam search of modeling.py:575-613 over encoder features [B, F, 768]; returns the best hypothesis' token ids per sample.
        The whole loop stays on the device (finished instances are frozen instead of compacted); one D2H at the end."""
        dev = self._device()
        B, Fr, _ = feats.shape
        eng = self._get_decoder(B, num_beams)
        lib = _lib.load()
        W = int(self.args.max_words)
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            _lib.check(lib.hb_decoder_begin(eng.handle, feats.float().contiguous().data_ptr(), B, Fr, num_beams, st), "hb_decoder_begin")
            done = torch.zeros((B,), dtype=torch.int32, device=dev)
            for step in range(W):
                _lib.check(lib.hb_decoder_step(eng.handle, st), "hb_decoder_step")
                if step % 8 == 7 and step + 1 < W:   # cheap early exit: all instances finished?
                    _lib.check(lib.hb_decoder_read(eng.handle, None, None, None, done.data_ptr(), None, st), "hb_decoder_read")
                    if bool(done.all()):
                        break
            pk = torch.empty((W, B, num_beams), dtype=torch.int32, device=dev)
            ys = torch.empty((W, B, num_beams), dtype=torch.int32, device=dev)
            ns = torch.empty((B,), dtype=torch.int32, device=dev)
            _lib.check(lib.hb_decoder_read(eng.handle, pk.data_ptr(), ys.data_ptr(), ns.data_ptr(), None, None, st), "hb_decoder_read")
        pk, ys, ns = pk.cpu().tolist(), ys.cpu().tolist(), ns.cpu().tolist()
        out = []
        for b in range(B):   # Beam.get_hypothesis(0): scores are kept sorted, so the best tail is beam 0 (beam.py:103-123)
            k, hyp = 0, []
            for j in range(ns[b] - 1, -1, -1):
                hyp.append(ys[j][b][k])
                k = pk[j][b][k]
            out.append(hyp[::-1])
        return out

    def _vocab(self):
        v = getattr(self, "_vocab_list", None)
        if v is None:
            import os

            path = getattr(self.args, "bert_vocab_path", None)
            cands = [path] if path else []
            cands += ["./clip4caption/modules/bert-base-uncased/vocab.txt", os.path.expanduser("~/.pytorch_pretrained_bert/vocab.txt")]
            for c in cands:
                if c and os.path.exists(c):
                    with open(c, encoding="utf-8") as f:
                        v = [ln.rstrip("\n") for ln in f]
                    break
            if v is None:
                raise RuntimeError("BERT vocab.txt not found (set args.bert_vocab_path); needed to turn token ids into text")
            self._vocab_list = v
        return v

    def ids_to_text(self, ids):
        """modeling.py:615-626."""
        vocab = self._vocab()
        toks = [vocab[i] for i in ids]
        if "[SEP]" in toks:
            toks = toks[:toks.index("[SEP]")]
        if "[PAD]" in toks:
            toks = toks[:toks.index("[PAD]")]
        return str(" ".join(toks).replace(" ##", "").strip("##").strip())

    @torch.no_grad()
    def caption_token_ids(self, batch, num_beams: int = 5):
        """The device part of modeling.py:556-613: trim to the moment, shared encoder, beam search -> best-hypothesis token ids."""
        dev = self._device()
        video, vmask, asr, text_feat = self._inputs(batch)
        mmask = batch["moment_mask"].to(dev).long().contiguous()
        B = video.shape[0]
        video = self.trim_feats(video, mmask)
        asr = self.trim_feats(asr, mmask) if self.use_asr else None
        ones = torch.ones((B, self.args.max_frames), dtype=torch.long, device=dev)
        _, feats = self._forward(video, text_feat, asr, ones, ones, want_feats=True)
        return self.generate_caption_ids(feats, num_beams)

    @torch.no_grad()
    def test_step_captioning(self, batch, **kwargs):
        """modeling.py:556-632."""
        ids = self.caption_token_ids(batch, kwargs.get("num_beams", 5))
        return {"prediction": [self.ids_to_text(x) for x in ids], "token_ids": ids}

    @torch.no_grad()
    def trim_feats(self, visual_output, moment_mask, B=None, device=None):
        """modeling.py:529-554 on the device."""
        dev = self._device()
        x = visual_output.to(dev).float().contiguous()
        mk = moment_mask.to(dev).long().contiguous()
        Bx, T, Cc = x.shape
        F = self.args.max_frames
        out = torch.empty((Bx, F, Cc), dtype=torch.float32, device=dev)
        lib = _lib.init(dev.index or 0)
        with torch.cuda.device(dev):
            _lib.check(lib.hb_trim_feats(x.data_ptr(), mk.data_ptr(), out.data_ptr(), Bx, T, Cc, F, _lib.stream_ptr(dev)), "hb_trim_feats")
        return out


def _split_top(shapes: dict) -> dict:
    out = {}
    for name, shape in shapes.items():
        head, _, rest = name.partition(".")
        out.setdefault(head, {})[rest] = shape
    return out


def _postprocess_steps(steps, last_bound):
    """Host post-processing of modeling.py:435-463 (sort, flatten, drop values past the moment end, unique, min gap of 5)."""
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

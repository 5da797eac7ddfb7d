"""CPU oracle for step captioning: clip4caption decoder + beam search (fp32 torch restatement; TEST INFRASTRUCTURE ONLY).

Restates, with the reference's quirks kept: full-prefix recompute each step, -10000 on every cross-attention key
(all-zeros video mask, modeling.py:591), soft-causal -10000 self-attention mask, tied classifier, per-instance flat
top-k with `//` back-pointers, "done when the BEST beam emits [SEP]".  Pinned against the unmodified reference by
oracle/make_golden_moment.py (tests/golden/moment.pt, key 'caption').
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from oracle.moment_oracle import gelu_erf, moment_shared, tf_layernorm, trim_feats

BOS, EOS, PAD = 101, 102, 0   # [CLS], [SEP], [PAD] ids of the BERT vocabulary (beam.py:22-29)


def _mha(sd, p, q_in, kv_in, mask, heads=12):
    """MultiHeadAttention.forward (module_decoder.py:220-247)."""
    B, Tq, Hd = q_in.shape
    Tk = kv_in.shape[1]
    dh = Hd // heads
    q = F.linear(q_in, sd[p + "query.weight"], sd[p + "query.bias"]).view(B, Tq, heads, dh).permute(0, 2, 1, 3)
    k = F.linear(kv_in, sd[p + "key.weight"], sd[p + "key.bias"]).view(B, Tk, heads, dh).permute(0, 2, 1, 3)
    v = F.linear(kv_in, sd[p + "value.weight"], sd[p + "value.bias"]).view(B, Tk, heads, dh).permute(0, 2, 1, 3)
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(dh) + mask
    return torch.matmul(s.softmax(dim=-1), v).permute(0, 2, 1, 3).contiguous().view(B, Tq, Hd)


def decoder_logits(sd, ids, enc, n_layers=2, prefix="clip4cap_model.decoder."):
    """DecoderModel.forward (module_decoder.py:372-406) with answer_mask = ones and encoder_mask = zeros."""
    B, t = ids.shape
    e = prefix + "embeddings."
    x = sd[e + "word_embeddings.weight"][ids] + sd[e + "position_embeddings.weight"][:t]            # :309-317
    x = tf_layernorm(x, sd[e + "LayerNorm.weight"], sd[e + "LayerNorm.bias"])
    enc_mask = torch.full((1, 1, 1, enc.shape[1]), -10000.0)                                           # (1 - 0) * -10000, :385-387
    self_mask = torch.triu(torch.ones(t, t), diagonal=1).gt(0).float()[None, None] * -10000.0           # :393-396
    for i in range(n_layers):
        p = f"{prefix}decoder.layer.{i}."
        a = _mha(sd, p + "slf_attn.att.", x, x, self_mask)
        s = tf_layernorm(F.linear(a, sd[p + "slf_attn.output.dense.weight"], sd[p + "slf_attn.output.dense.bias"]) + x,
                         sd[p + "slf_attn.output.LayerNorm.weight"], sd[p + "slf_attn.output.LayerNorm.bias"])
        a = _mha(sd, p + "enc_attn.att.", s, enc, enc_mask)
        c = tf_layernorm(F.linear(a, sd[p + "enc_attn.output.dense.weight"], sd[p + "enc_attn.output.dense.bias"]) + s,
                         sd[p + "enc_attn.output.LayerNorm.weight"], sd[p + "enc_attn.output.LayerNorm.bias"])
        m = gelu_erf(F.linear(c, sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"]))
        x = tf_layernorm(F.linear(m, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"]) + c,
                         sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"])
    c_ = prefix + "classifier.cls.predictions."
    h = tf_layernorm(gelu_erf(F.linear(x, sd[c_ + "transform.dense.weight"], sd[c_ + "transform.dense.bias"])),
                     sd[c_ + "transform.LayerNorm.weight"], sd[c_ + "transform.LayerNorm.bias"])          # :160-166
    return F.linear(h, sd[c_ + "decoder.weight"]) + sd[c_ + "bias"]                                    # :180-183


def beam_search(sd, enc, beam=3, max_words=48):
    """beam_decode_step / Beam.advance / collect_hypothesis (train.py:547-599, beam.py:70-123), one instance at a time
    (instances are independent in the reference: finished ones are merely dropped from the batch)."""
    out = []
    for b in range(enc.shape[0]):
        e = enc[b:b + 1].expand(beam, -1, -1)
        scores = torch.zeros(beam)
        prev_ks, next_ys = [], [torch.full((beam,), BOS, dtype=torch.long)]

        def hypothesis(k):
            hyp = []
            for j in range(len(prev_ks) - 1, -1, -1):
                hyp.append(int(next_ys[j + 1][k]))
                k = int(prev_ks[j][k])
            return hyp[::-1]

        for step in range(1, max_words + 1):
            if len(next_ys) == 1:
                seq = next_ys[0].unsqueeze(1)
            else:
                keys = torch.sort(scores, 0, True)[1]
                seq = torch.LongTensor([[BOS] + hypothesis(int(k)) for k in keys])
            logp = F.log_softmax(decoder_logits(sd, seq, e)[:, -1, :], dim=1)
            V = logp.shape[1]
            lk = logp + scores.unsqueeze(1) if prev_ks else logp[0]
            best, idx = lk.view(-1).topk(beam, 0, True, True)
            scores = best
            pk = idx // V
            prev_ks.append(pk)
            next_ys.append(idx - pk * V)
            if int(next_ys[-1][0]) == EOS:
                break
        tail = torch.sort(scores, 0, True)[1]
        out.append(hypothesis(int(tail[0])))
    return out


def ids_to_text(ids, vocab):
    """modeling.py:615-626."""
    toks = [vocab[i] for i in ids]
    if "[SEP]" in toks:
        toks = toks[:toks.index("[SEP]")]
    if "[PAD]" in toks:
        toks = toks[:toks.index("[PAD]")]
    return " ".join(toks).replace(" ##", "").strip("##").strip()


def test_step_captioning(sd, batch, text_feat, beam=3, max_frames=20, max_words=48):
    """modeling.py:556-613 -> token id lists of the best hypothesis per sample."""
    B = batch["vis_feats"].shape[0]
    vis = trim_feats(batch["vis_feats"], batch["moment_mask"], max_frames)
    asr = trim_feats(batch["asr_feats"], batch["moment_mask"], max_frames)
    ones = torch.ones((B, max_frames), dtype=torch.long)
    enc = moment_shared(sd, vis, text_feat, ones, ones, asr)
    return beam_search(sd, enc, beam, max_words), enc

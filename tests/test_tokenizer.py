"""hirest_b200.tokenizer against token ids produced by the reference's own tokenizer
(oracle/make_golden_tokenizer.py -> tests/golden/tokenizer.json).  The BPE merge table is user-supplied data (it ships with CLIP
and with the reference); the tests run where it can be found: HIREST_BPE_PATH or the build container's /root/reference."""
import json
import os

import pytest
import torch

from hirest_b200 import tokenizer

BPE = os.environ.get("HIREST_BPE_PATH") or "/root/reference/EVA_clip/bpe_simple_vocab_16e6.txt.gz"
needs_real_table = pytest.mark.skipif(not os.path.exists(BPE), reason="CLIP's own BPE merge table is not available on this machine")


@pytest.fixture(scope="module")
def tok():
    if not os.path.exists(BPE):
        pytest.skip("CLIP's own BPE merge table is not available on this machine")
    return tokenizer.ClipBpeTokenizer(BPE)


def test_synthetic_merge_table_matches_reference_everywhere(golden, golden_dir):
    """Runs on every machine (GPU box included): a small merge table committed under tests/golden/ and the ids the REFERENCE
    tokenizer produced with it (oracle/make_golden_tokenizer.py) -- same algorithm, table-independent."""
    path = os.path.join(golden_dir, "bpe_synthetic.txt.gz")
    t = tokenizer.ClipBpeTokenizer(path)
    g = golden["synthetic"]
    assert len(t.encoder) == g["vocab_size"] and t.sot_token == g["sot"] and t.eot_token == g["eot"]
    for prompt, ids, dec in zip(golden["prompts"], g["ids"], g["decoded"]):
        assert t.encode(prompt) == ids, prompt
        assert t.decode(ids) == dec, prompt
    assert t.encode(golden["long_prompt"]) == g["long_ids"]
    assert tokenizer.get_tokenizer(path) is tokenizer.get_tokenizer(path)            # cached per path
    row = tokenizer.tokenize("make tea", bpe_path=path)[0].tolist()
    assert row[0] == g["sot"] and g["eot"] in row and tokenizer.get_tokenizer(path) is tokenizer.get_tokenizer(path)


@pytest.fixture(scope="module")
def golden(golden_dir):
    return json.load(open(os.path.join(golden_dir, "tokenizer.json")))


def test_vocabulary_layout(tok, golden):
    assert len(tok.encoder) == golden["vocab_size"] == 49408
    assert tok.sot_token == golden["sot"] == 49406 and tok.eot_token == golden["eot"] == 49407
    assert tok.encoder["!"] == 0 and tok.encoder["!</w>"] == 256


def test_encode_matches_reference(tok, golden):
    for prompt, ids in zip(golden["prompts"], golden["ids"]):
        assert tok.encode(prompt) == ids, prompt
    assert tok.encode(golden["long_prompt"]) == golden["long_ids"]


def test_decode_matches_reference(tok, golden):
    for prompt, ids, dec in zip(golden["prompts"], golden["ids"], golden["decoded"]):
        assert tok.decode(ids) == dec, prompt


def test_tokenize_layout_padding_and_length_rules(tok, golden):
    t = tok.tokenize(golden["prompts"][:4])
    assert t.shape == (4, 77) and t.dtype == torch.int64
    for row, ids in zip(t.tolist(), golden["ids"][:4]):
        assert row[:len(ids) + 2] == [49406] + ids + [49407] and set(row[len(ids) + 2:]) <= {0}
        assert int(torch.tensor(row).argmax()) == len(ids) + 1     # EOT is the row maximum (eva_model.py:243 relies on it)
    assert tok.tokenize("single string").shape == (1, 77)
    with pytest.raises(RuntimeError, match="too long"):
        tok.tokenize(golden["long_prompt"])
    tr = tok.tokenize(golden["long_prompt"], truncate=True)[0].tolist()
    assert tr[0] == 49406 and tr[-1] == 49407 and tr[1:76] == golden["long_ids"][:75]


def test_missing_table_is_an_error(tmp_path):
    with pytest.raises(FileNotFoundError):
        tokenizer.ClipBpeTokenizer(str(tmp_path / "nope.txt.gz"))

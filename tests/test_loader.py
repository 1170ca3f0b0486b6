"""Model-file formats: CTranslate2 model.bin + config.json + tokenizer.json -- what the reference
loads at /root/reference/asr_components.py:232-265 -- and Hugging Face safetensors, round-tripped through synthetic
files written by the independent test-side writer tests/ct2_format.py (no real checkpoint exists offline)."""
import os

import pytest
import torch

from realtime_whisper_asr_b200 import loader as L
from realtime_whisper_asr_b200 import weights as W
from realtime_whisper_asr_b200.tokenizer import Tokenizer
from tests import ct2_format as C

MD = W.MODEL_DIMS["micro"]
HEADS = [(1, 0), (1, 1)]


@pytest.fixture(scope="module")
def sd():
    return W.synthetic_state_dict(MD, seed=3)


def _same(a, b, tol=0.0):
    assert set(a) == set(b), set(a) ^ set(b)
    for k in a:
        assert a[k].shape == b[k].shape, k
        err = (a[k].float() - b[k].float()).abs().max().item()
        assert err <= tol * max(1.0, b[k].abs().max().item()), (k, err)


def test_ct2_float32_roundtrip_is_exact(sd, tmp_path):
    d = C.write_ct2_dir(str(tmp_path / "m"), sd, MD, dtype="float32", alignment_heads=HEADS,
                        suppress_ids=[1, 2, 7], suppress_ids_begin=[220, 50257])
    m = L.load_model_dir(d)
    _same(m.state_dict, sd)
    assert m.dims == MD and m.alignment_heads == HEADS
    assert m.suppress_ids == [1, 2, 7] and m.suppress_ids_begin == [220, 50257] and len(m.lang_ids) == 99
    assert m.tokenizer_json.endswith("tokenizer.json") and "ctranslate2" in m.source
    # the key projection has no bias in Whisper: the fused in-projection's zero block must not come back as one
    assert "encoder.blocks.0.attn.key.bias" not in m.state_dict


@pytest.mark.parametrize("dtype,tol", [("float16", 1e-3), ("bfloat16", 0.0), ("int8", 1e-2)])
def test_ct2_storage_types(sd, tmp_path, dtype, tol):
    """float16 (faster-whisper's published conversions), bfloat16 (exact here: the synthetic matrices are
    bf16-representable... vectors are not, hence the tolerance on them), int8 + weight_scale (de-quantised)."""
    d = C.write_ct2_dir(str(tmp_path / dtype), sd, MD, dtype=dtype, alignment_heads=HEADS)
    m = L.load_model_dir(d)
    _same(m.state_dict, sd, tol=max(tol, 4e-3 if dtype == "bfloat16" else tol))
    if dtype == "bfloat16":
        for k in ("decoder.token_embedding.weight", "encoder.blocks.0.mlp.0.weight"):
            assert torch.equal(m.state_dict[k], sd[k]), k


def test_ct2_header_and_aliases(sd, tmp_path):
    d = C.write_ct2_dir(str(tmp_path / "m"), sd, MD, dtype="float16")
    header, variables, aliases = L.read_ct2_model_bin(os.path.join(d, "model.bin"))
    assert header == {"binary_version": 6, "spec": "WhisperSpec", "revision": 3}
    assert aliases == {"decoder/projection/weight": "decoder/embeddings/weight"}
    assert variables["decoder/layer_0/attention/linear_1/weight"].shape == (2 * MD.d_model, MD.d_model)
    assert variables["encoder/num_heads"].item() == MD.n_heads
    with open(os.path.join(d, "model.bin"), "r+b") as f:  # wrong version is refused, truncated files too
        f.write(b"\x09\0\0\0")
    with pytest.raises(ValueError):
        L.read_ct2_model_bin(os.path.join(d, "model.bin"))


def test_hf_safetensors_roundtrip(sd, tmp_path):
    d = C.write_hf_safetensors_dir(str(tmp_path / "hf"), sd, alignment_heads=HEADS)
    m = L.load_model_dir(d)
    _same(m.state_dict, sd)
    assert m.dims == MD and m.alignment_heads == HEADS and "transformers" in m.source


def test_real_checkpoint_without_tokenizer_is_refused(sd, tmp_path):
    d = C.write_ct2_dir(str(tmp_path / "m"), sd, MD, with_tokenizer=False)
    with pytest.raises(FileNotFoundError, match="tokenizer.json"):
        L.load_model_dir(d)


def test_resolve_model_dir_layouts(sd, tmp_path):
    root = tmp_path / "models_fast"
    plain = C.write_ct2_dir(str(root / "micro"), sd, MD)
    assert L.resolve_model_dir("micro", str(root)) == plain
    assert L.resolve_model_dir(plain) == plain
    hub = root / "models--Systran--faster-whisper-tiny" / "snapshots" / "abc123"
    C.write_ct2_dir(str(hub), sd, MD)
    assert L.resolve_model_dir("tiny", str(root)) == str(hub)
    assert L.resolve_model_dir("large-v3", str(root)) is None


def test_hf_vocabulary_and_shared_vocab(tmp_path):
    p = str(tmp_path / "tokenizer.json")
    C.write_tokenizer_json(p)
    tk = Tokenizer(51865, True, language="en", tokenizer_json=p)
    text = " the quick brown fox, hello world."
    assert tk.decode(tk.encode(text)) == text
    tk2 = Tokenizer(51865, True, language="de", tokenizer_json=p, vocab=tk.vocab)
    assert tk2.vocab is tk.vocab and tk2.sot_sequence[1] == tk2.lang0 + 2
    words, toks = tk.split_to_word_tokens(tk.encode(" hello world") + [tk.eot])
    assert [w for w in words[:2]] == [" hello", " world"] and sum(len(t) for t in toks) == len(tk.encode(" hello world")) + 1

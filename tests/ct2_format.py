"""Test-side WRITER of the CTranslate2 model directory format (independent of the product's reader in
realtime_whisper_asr_b200/loader.py): model.bin in binary spec version 6 as ctranslate2's ModelSpec._serialize
lays it out, config.json with the Whisper fields, and a small byte-level BPE tokenizer.json.

Layout of model.bin (little endian): uint32 version (6); string spec name; uint32 revision; uint32 #variables;
per variable (sorted by name): string name, uint8 rank, uint32 dims[rank], uint8 dtype id (0 f32, 1 i8, 2 i16,
3 i32, 4 f16, 5 bf16), uint32 byte count, bytes; uint32 #aliases; per alias: string alias, string target.
A string is uint16 (length + 1), the utf-8 bytes, a NUL.
Variable names follow ctranslate2.specs.WhisperSpec: fused self-attention in-projection `linear_0` ([3d, d]:
query | key | value, the key's bias zero), cross attention `linear_0` = query, `linear_1` = fused key | value,
last linear = output projection; LayerNorm = gamma / beta; decoder/projection/weight is an ALIAS of
decoder/embeddings/weight (tied)."""
import json
import os
import struct

import numpy as np
import torch

DTYPE_ID = {"float32": 0, "int8": 1, "int16": 2, "int32": 3, "float16": 4, "bfloat16": 5}


def whisper_spec_variables(sd, dims):
    """OpenAI-keyed state dict -> {ct2 variable name: fp32 tensor}."""
    v = {}

    def norm(dst, src):
        v[dst + "/gamma"] = sd[src + ".weight"]
        v[dst + "/beta"] = sd[src + ".bias"]

    def lin(dst, src):
        v[dst + "/weight"] = sd[src + ".weight"]
        v[dst + "/bias"] = sd[src + ".bias"]

    def fused(dst, src, names):
        ws, bs = [], []
        for n in names:
            w = sd[f"{src}.{n}.weight"]
            ws.append(w)
            bs.append(sd.get(f"{src}.{n}.bias", torch.zeros(w.shape[0])))
        v[dst + "/weight"] = torch.cat(ws, 0)
        v[dst + "/bias"] = torch.cat(bs, 0)

    for c in ("conv1", "conv2"):
        v[f"encoder/{c}/weight"] = sd[f"encoder.{c}.weight"]
        v[f"encoder/{c}/bias"] = sd[f"encoder.{c}.bias"]
    v["encoder/position_encodings/encodings"] = sd["encoder.positional_embedding"]
    for i in range(dims.n_enc_layers):
        s, t = f"encoder.blocks.{i}", f"encoder/layer_{i}"
        norm(t + "/self_attention/layer_norm", s + ".attn_ln")
        fused(t + "/self_attention/linear_0", s + ".attn", ("query", "key", "value"))
        lin(t + "/self_attention/linear_1", s + ".attn.out")
        norm(t + "/ffn/layer_norm", s + ".mlp_ln")
        lin(t + "/ffn/linear_0", s + ".mlp.0")
        lin(t + "/ffn/linear_1", s + ".mlp.2")
    norm("encoder/layer_norm", "encoder.ln_post")
    v["encoder/num_heads"] = torch.tensor(dims.n_heads, dtype=torch.int16)
    v["decoder/embeddings/weight"] = sd["decoder.token_embedding.weight"]
    v["decoder/position_encodings/encodings"] = sd["decoder.positional_embedding"]
    for i in range(dims.n_dec_layers):
        s, t = f"decoder.blocks.{i}", f"decoder/layer_{i}"
        norm(t + "/self_attention/layer_norm", s + ".attn_ln")
        fused(t + "/self_attention/linear_0", s + ".attn", ("query", "key", "value"))
        lin(t + "/self_attention/linear_1", s + ".attn.out")
        norm(t + "/attention/layer_norm", s + ".cross_attn_ln")
        lin(t + "/attention/linear_0", s + ".cross_attn.query")
        fused(t + "/attention/linear_1", s + ".cross_attn", ("key", "value"))
        lin(t + "/attention/linear_2", s + ".cross_attn.out")
        norm(t + "/ffn/layer_norm", s + ".mlp_ln")
        lin(t + "/ffn/linear_0", s + ".mlp.0")
        lin(t + "/ffn/linear_1", s + ".mlp.2")
    norm("decoder/layer_norm", "decoder.ln")
    v["decoder/num_heads"] = torch.tensor(dims.n_heads, dtype=torch.int16)
    v["decoder/scale_embeddings"] = torch.tensor(0, dtype=torch.int8)
    return v


def _wstr(f, s):
    b = s.encode("utf-8")
    f.write(struct.pack("<H", len(b) + 1))
    f.write(b)
    f.write(b"\0")


def write_model_bin(path, variables, dtype="float16", aliases=None):
    """dtype: storage type of the float variables ("float32" | "float16" | "bfloat16" | "int8": 2-D weights are
    quantised per row with a weight_scale variable, everything else stays float32)."""
    out = {}
    for name, t in variables.items():
        if not t.dtype.is_floating_point:
            out[name] = (t, {torch.int8: "int8", torch.int16: "int16", torch.int32: "int32"}[t.dtype])
        elif dtype == "int8" and name.endswith("/weight") and t.ndim == 2:
            scale = 127.0 / t.abs().amax(dim=1).clamp_min(1e-12)
            out[name] = ((t * scale[:, None]).round().clamp(-127, 127).to(torch.int8), "int8")
            out[name + "_scale"] = (scale.to(torch.float32), "float32")
        elif dtype in ("float16", "bfloat16"):
            out[name] = (t.to(getattr(torch, dtype)), dtype)
        else:
            out[name] = (t.to(torch.float32), "float32")
    with open(path, "wb") as f:
        f.write(struct.pack("<I", 6))
        _wstr(f, "WhisperSpec")
        f.write(struct.pack("<I", 3))
        f.write(struct.pack("<I", len(out)))
        for name in sorted(out):
            t, dt = out[name]
            _wstr(f, name)
            f.write(struct.pack("<B", t.ndim))
            for dim in t.shape:
                f.write(struct.pack("<I", dim))
            f.write(struct.pack("<B", DTYPE_ID[dt]))
            raw = t.contiguous().view(torch.uint8).numpy().tobytes() if t.ndim else \
                t.reshape(1).view(torch.uint8).numpy().tobytes()
            f.write(struct.pack("<I", len(raw)))
            f.write(raw)
        aliases = aliases or {}
        f.write(struct.pack("<I", len(aliases)))
        for a in sorted(aliases):
            _wstr(f, a)
            _wstr(f, aliases[a])


def write_tokenizer_json(path, vocab_size=600):
    from tokenizers import Tokenizer, decoders, models, pre_tokenizers, trainers
    tok = Tokenizer(models.BPE())
    tok.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False)
    tok.decoder = decoders.ByteLevel()
    trainer = trainers.BpeTrainer(vocab_size=vocab_size, initial_alphabet=pre_tokenizers.ByteLevel.alphabet(),
                                  special_tokens=[], show_progress=False)
    corpus = ["the quick brown fox jumps over the lazy dog", "seven wizards quietly mix potions before dawn",
              "bright stars above the old harbour town", "hello world, this is a test of word timing."] * 4
    tok.train_from_iterator(corpus, trainer)
    tok.save(path)


def write_ct2_dir(path, sd, dims, dtype="float16", alignment_heads=None, suppress_ids=None,
                  suppress_ids_begin=None, with_tokenizer=True):
    os.makedirs(path, exist_ok=True)
    write_model_bin(os.path.join(path, "model.bin"), whisper_spec_variables(sd, dims), dtype,
                    aliases={"decoder/projection/weight": "decoder/embeddings/weight"})
    cfg = {"alignment_heads": [list(x) for x in (alignment_heads or [])],
           "lang_ids": list(range(50259, 50259 + (100 if dims.n_vocab >= 51866 else 99))),
           "suppress_ids": list(suppress_ids or []), "suppress_ids_begin": list(suppress_ids_begin or [])}
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(cfg, f)
    if with_tokenizer:
        write_tokenizer_json(os.path.join(path, "tokenizer.json"))
    return path


def write_hf_safetensors_dir(path, sd, alignment_heads=None, with_tokenizer=True):
    """transformers' WhisperForConditionalGeneration key names."""
    from safetensors.torch import save_file
    ren = [("encoder.blocks.", "model.encoder.layers."), ("decoder.blocks.", "model.decoder.layers."),
           (".cross_attn_ln", ".encoder_attn_layer_norm"), (".attn_ln", ".self_attn_layer_norm"),
           (".mlp_ln", ".final_layer_norm"), (".cross_attn.query", ".encoder_attn.q_proj"),
           (".cross_attn.key", ".encoder_attn.k_proj"), (".cross_attn.value", ".encoder_attn.v_proj"),
           (".cross_attn.out", ".encoder_attn.out_proj"), (".attn.query", ".self_attn.q_proj"),
           (".attn.key", ".self_attn.k_proj"), (".attn.value", ".self_attn.v_proj"), (".attn.out", ".self_attn.out_proj"),
           (".mlp.0", ".fc1"), (".mlp.2", ".fc2")]
    out = {}
    for k, t in sd.items():
        nk = k
        for a, b in ren:
            nk = nk.replace(a, b)
        if nk == "encoder.ln_post.weight" or nk == "encoder.ln_post.bias":
            nk = "model.encoder.layer_norm." + nk.split(".")[-1]
        elif nk.startswith("decoder.ln."):
            nk = "model.decoder.layer_norm." + nk.split(".")[-1]
        elif nk == "decoder.token_embedding.weight":
            nk = "model.decoder.embed_tokens.weight"
        elif nk.endswith("positional_embedding"):
            nk = "model." + nk.replace("positional_embedding", "embed_positions.weight")
        elif not nk.startswith("model."):
            nk = "model." + nk
        out[nk] = t.contiguous()
    os.makedirs(path, exist_ok=True)
    save_file(out, os.path.join(path, "model.safetensors"))
    with open(os.path.join(path, "generation_config.json"), "w") as f:
        json.dump({"alignment_heads": [list(x) for x in (alignment_heads or [])]}, f)
    if with_tokenizer:
        write_tokenizer_json(os.path.join(path, "tokenizer.json"))
    return path

"""Checkpoint loading: a model directory -> (OpenAI-keyed state dict, dims, tokenizer.json, model config).

What the reference loads (/root/reference/asr_components.py:232-265): ``WhisperModel(size_or_dir, download_root=...)``,
i.e. a CTranslate2 model directory -- ``model.bin`` + ``config.json`` (``alignment_heads``, ``suppress_ids``,
``suppress_ids_begin``, ``lang_ids``) + ``tokenizer.json`` / ``vocabulary.json|txt`` -- either given as a path or found
under ``download_root`` (Hugging Face hub cache layout ``models--Systran--faster-whisper-<size>/snapshots/<rev>/`` or a
plain ``<download_root>/<size>/`` directory).  Supported layouts:

  * CTranslate2 ``model.bin`` (binary spec version 6, what ``ct2-transformers-converter`` / faster-whisper's
    published conversions contain): float32 / float16 / bfloat16 variables, or int8 / int16 weights with their
    ``weight_scale`` (de-quantised on load: this backend computes in bf16).  The format is restated from the
    published CTranslate2 serialisation (``ModelSpec._serialize``: uint32 version, name, uint32 revision, the
    variables sorted by name -- each name, rank, dims, dtype id, byte count, bytes -- then the aliases); no real
    file exists offline, tests round-trip synthetic files written by an independent writer (tests/ct2_format.py).
  * Hugging Face ``model.safetensors`` (transformers' Whisper key names) + ``config.json`` / ``generation_config.json``;
  * an OpenAI-keyed ``model.safetensors`` (keys start with ``encoder.`` / ``decoder.``).

A real checkpoint without ``tokenizer.json`` is refused: decoding real weights with the synthetic byte vocabulary
would print well-formed gibberish.
"""
from __future__ import annotations

import glob
import json
import os
import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from .weights import ModelDims, dims_from_state_dict, sinusoids

_HF_RENAMES = [
    ("model.", ""), ("layers.", "blocks."), ("self_attn_layer_norm", "attn_ln"),
    ("encoder_attn_layer_norm", "cross_attn_ln"), ("final_layer_norm", "mlp_ln"),
    ("self_attn.q_proj", "attn.query"), ("self_attn.k_proj", "attn.key"), ("self_attn.v_proj", "attn.value"),
    ("self_attn.out_proj", "attn.out"), ("encoder_attn.q_proj", "cross_attn.query"),
    ("encoder_attn.k_proj", "cross_attn.key"), ("encoder_attn.v_proj", "cross_attn.value"),
    ("encoder_attn.out_proj", "cross_attn.out"), ("fc1", "mlp.0"), ("fc2", "mlp.2"),
    ("embed_positions.weight", "positional_embedding"), ("embed_tokens", "token_embedding"),
    ("encoder.layer_norm", "encoder.ln_post"), ("decoder.layer_norm", "decoder.ln"),
]


@dataclass
class LoadedModel:
    state_dict: Dict[str, torch.Tensor]
    dims: ModelDims
    tokenizer_json: Optional[str]
    alignment_heads: Optional[List[Tuple[int, int]]] = None
    suppress_ids: Optional[List[int]] = None        # CT2 config.json: what suppress_tokens=[-1] expands to
    suppress_ids_begin: Optional[List[int]] = None  # CT2 config.json: barred at the first sampled position
    lang_ids: Optional[List[int]] = None
    source: str = ""
    extra: dict = field(default_factory=dict)


def hf_to_openai_keys(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in sd.items():
        if k.startswith("proj_out"):
            continue  # tied to the token embedding
        nk = k
        for a, b in _HF_RENAMES:
            nk = nk.replace(a, b)
        out[nk] = v.to(torch.float32)
    return out


# ------------------------------------------------------------------------------- CTranslate2 model.bin
CT2_DTYPES = {0: np.dtype("<f4"), 1: np.dtype("i1"), 2: np.dtype("<i2"), 3: np.dtype("<i4"), 4: np.dtype("<f2"),
              5: "bfloat16"}


def read_ct2_model_bin(path: str) -> Tuple[dict, Dict[str, torch.Tensor], Dict[str, str]]:
    """Parse a CTranslate2 model.bin -> (header, {variable name: tensor}, {alias: variable name})."""
    with open(path, "rb") as f:
        data = f.read()
    pos = 0

    def take(fmt):
        nonlocal pos
        v = struct.unpack_from("<" + fmt, data, pos)
        pos += struct.calcsize("<" + fmt)
        return v[0]

    def take_str():
        nonlocal pos
        n = take("H")
        s = data[pos:pos + n - 1].decode("utf-8")
        if data[pos + n - 1] != 0:
            raise ValueError(f"{path}: malformed string at byte {pos}")
        pos += n
        return s

    version = take("I")
    if version < 5 or version > 6:
        raise ValueError(f"{path}: CTranslate2 binary version {version} is not supported (5 and 6 are)")
    header = {"binary_version": version, "spec": take_str(), "revision": take("I")}
    variables: Dict[str, torch.Tensor] = {}
    for _ in range(take("I")):
        name = take_str()
        rank = take("B")
        shape = [take("I") for _ in range(rank)]
        dtype_id = take("B")
        nbytes = take("I")
        if dtype_id not in CT2_DTYPES:
            raise ValueError(f"{path}: variable {name} has unknown dtype id {dtype_id}")
        raw = data[pos:pos + nbytes]
        if len(raw) != nbytes:
            raise ValueError(f"{path}: truncated at variable {name}")
        pos += nbytes
        dt = CT2_DTYPES[dtype_id]
        if dt == "bfloat16":
            t = torch.frombuffer(bytearray(raw), dtype=torch.bfloat16)
        else:
            t = torch.from_numpy(np.frombuffer(raw, dtype=dt).copy())
        n_el = int(np.prod(shape)) if shape else 1
        if t.numel() != n_el:
            raise ValueError(f"{path}: variable {name}: {t.numel()} elements for shape {shape}")
        variables[name] = t.reshape(shape)
    aliases = {}
    if pos < len(data):
        for _ in range(take("I")):
            alias = take_str()
            aliases[alias] = take_str()
    return header, variables, aliases


def _ct2_weight(v: Dict[str, torch.Tensor], aliases: Dict[str, str], name: str) -> Optional[torch.Tensor]:
    """Variable `name` as fp32; int8 / int16 weights are divided by their per-row `weight_scale`."""
    name = aliases.get(name, name)
    if name not in v:
        return None
    t = v[name]
    if t.dtype in (torch.int8, torch.int16):
        sname = name[: -len("weight")] + "weight_scale" if name.endswith("weight") else None
        sname = aliases.get(sname, sname) if sname else None
        if not sname or sname not in v:
            raise ValueError(f"quantised variable {name} has no weight_scale")
        scale = v[sname].to(torch.float32)
        t = t.to(torch.float32) / (scale[:, None] if scale.ndim == 1 and t.ndim == 2 else scale)
    return t.to(torch.float32)


def ct2_to_openai_keys(v: Dict[str, torch.Tensor], aliases: Dict[str, str]) -> Dict[str, torch.Tensor]:
    """CTranslate2 WhisperSpec variable names -> OpenAI Whisper state-dict keys (fused projections split)."""
    def w(name):
        t = _ct2_weight(v, aliases, name)
        if t is None:
            raise KeyError(f"model.bin: variable {name} is missing")
        return t

    def opt(name):
        return _ct2_weight(v, aliases, name)

    sd: Dict[str, torch.Tensor] = {}

    def norm(dst, src):
        sd[dst + ".weight"] = w(src + "/gamma")
        sd[dst + ".bias"] = w(src + "/beta")

    def lin(dst, src, bias=True):
        sd[dst + ".weight"] = w(src + "/weight")
        if bias:
            b = opt(src + "/bias")
            sd[dst + ".bias"] = b if b is not None else torch.zeros(sd[dst + ".weight"].shape[0])

    def split_fused(dst, src, names, key_has_bias=False):
        W_, b_ = w(src + "/weight"), opt(src + "/bias")
        d = W_.shape[0] // len(names)
        for i, n in enumerate(names):
            sd[f"{dst}.{n}.weight"] = W_[i * d:(i + 1) * d].contiguous()
            if n != "key" or key_has_bias:
                sd[f"{dst}.{n}.bias"] = (b_[i * d:(i + 1) * d].contiguous() if b_ is not None else torch.zeros(d))

    for c in ("conv1", "conv2"):
        sd[f"encoder.{c}.weight"] = w(f"encoder/{c}/weight")
        sd[f"encoder.{c}.bias"] = w(f"encoder/{c}/bias")
    d = sd["encoder.conv1.weight"].shape[0]
    pe = opt("encoder/position_encodings/encodings")
    sd["encoder.positional_embedding"] = pe if pe is not None else sinusoids(1500, d)
    n_enc = 0
    while f"encoder/layer_{n_enc}/self_attention/linear_0/weight" in v:
        n_enc += 1
    for i in range(n_enc):
        s, t = f"encoder/layer_{i}", f"encoder.blocks.{i}"
        norm(t + ".attn_ln", s + "/self_attention/layer_norm")
        split_fused(t + ".attn", s + "/self_attention/linear_0", ("query", "key", "value"))
        lin(t + ".attn.out", s + "/self_attention/linear_1")
        norm(t + ".mlp_ln", s + "/ffn/layer_norm")
        lin(t + ".mlp.0", s + "/ffn/linear_0")
        lin(t + ".mlp.2", s + "/ffn/linear_1")
    norm("encoder.ln_post", "encoder/layer_norm")
    sd["decoder.token_embedding.weight"] = w("decoder/embeddings/weight")
    sd["decoder.positional_embedding"] = w("decoder/position_encodings/encodings")
    n_dec = 0
    while f"decoder/layer_{n_dec}/self_attention/linear_0/weight" in v:
        n_dec += 1
    for i in range(n_dec):
        s, t = f"decoder/layer_{i}", f"decoder.blocks.{i}"
        norm(t + ".attn_ln", s + "/self_attention/layer_norm")
        split_fused(t + ".attn", s + "/self_attention/linear_0", ("query", "key", "value"))
        lin(t + ".attn.out", s + "/self_attention/linear_1")
        norm(t + ".cross_attn_ln", s + "/attention/layer_norm")
        lin(t + ".cross_attn.query", s + "/attention/linear_0")
        split_fused(t + ".cross_attn", s + "/attention/linear_1", ("key", "value"))
        lin(t + ".cross_attn.out", s + "/attention/linear_2")
        norm(t + ".mlp_ln", s + "/ffn/layer_norm")
        lin(t + ".mlp.0", s + "/ffn/linear_0")
        lin(t + ".mlp.2", s + "/ffn/linear_1")
    norm("decoder.ln", "decoder/layer_norm")
    if n_enc == 0 or n_dec == 0:
        raise ValueError("model.bin holds no encoder / decoder layers: not a Whisper model")
    return sd


# ------------------------------------------------------------------------------- directory handling
def resolve_model_dir(size_or_path: str, download_root: Optional[str] = None) -> Optional[str]:
    """Directory holding the checkpoint for `size_or_path`, or None.  Mirrors where faster-whisper looks: the path
    itself, then `download_root` (plain sub-directory or Hugging Face hub cache layout), then the default hub cache."""
    if size_or_path and os.path.isdir(size_or_path):
        return size_or_path
    roots = [r for r in (download_root, os.environ.get("HF_HUB_CACHE"),
                         os.path.join(os.environ.get("HF_HOME", os.path.expanduser("~/.cache/huggingface")), "hub")) if r]
    names = [size_or_path, f"faster-whisper-{size_or_path}", f"whisper-{size_or_path}"]
    for root in roots:
        for n in names:
            p = os.path.join(root, n)
            if os.path.isdir(p) and (os.path.exists(os.path.join(p, "model.bin")) or
                                     os.path.exists(os.path.join(p, "model.safetensors"))):
                return p
        for org in ("Systran", "guillaumekln", "openai"):
            for n in names[1:]:
                hits = sorted(glob.glob(os.path.join(root, f"models--{org}--{n}", "snapshots", "*")))
                hits = [h for h in hits if os.path.exists(os.path.join(h, "model.bin")) or
                        os.path.exists(os.path.join(h, "model.safetensors"))]
                if hits:
                    return hits[-1]
    return None


def load_model_dir(path: str) -> Optional[LoadedModel]:
    if not path or not os.path.isdir(path):
        return None
    cfg = {}
    for name in ("config.json", "generation_config.json"):
        p = os.path.join(path, name)
        if os.path.exists(p):
            with open(p) as f:
                for k, val in json.load(f).items():
                    cfg.setdefault(k, val)
    bin_path, st_path = os.path.join(path, "model.bin"), os.path.join(path, "model.safetensors")
    if os.path.exists(bin_path):
        header, variables, aliases = read_ct2_model_bin(bin_path)
        if "Whisper" not in header["spec"]:
            raise ValueError(f"{bin_path}: spec {header['spec']!r} is not a Whisper model")
        sd = ct2_to_openai_keys(variables, aliases)
        source = f"ctranslate2 model.bin v{header['binary_version']} ({header['spec']} r{header['revision']})"
    elif os.path.exists(st_path):
        from safetensors.torch import load_file
        raw = load_file(st_path)
        hf = any(k.startswith("model.") for k in raw)
        sd = hf_to_openai_keys(raw) if hf else {k: v.to(torch.float32) for k, v in raw.items()}
        source = "safetensors (" + ("transformers" if hf else "openai") + " keys)"
    else:
        raise FileNotFoundError(f"{path}: neither model.bin (CTranslate2) nor model.safetensors")
    if "encoder.positional_embedding" not in sd:
        sd["encoder.positional_embedding"] = sinusoids(1500, sd["encoder.conv1.weight"].shape[0])
    dims = dims_from_state_dict(sd)
    tj = os.path.join(path, "tokenizer.json")
    if not os.path.exists(tj):
        raise FileNotFoundError(
            f"{path}: tokenizer.json is missing.  A real checkpoint cannot be decoded with the synthetic byte "
            "vocabulary; put the model's Hugging Face tokenizer.json next to the weights.")
    heads = [tuple(x) for x in cfg["alignment_heads"]] if cfg.get("alignment_heads") else None
    return LoadedModel(sd, dims, tj, heads, cfg.get("suppress_ids"), cfg.get("suppress_ids_begin"),
                       cfg.get("lang_ids"), source)

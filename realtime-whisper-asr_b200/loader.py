"""Checkpoint loading: a model directory -> (OpenAI-keyed state dict, dims, tokenizer.json, heads).

Supported layouts (none ships with this repo or the image; tests write synthetic files):
  * Hugging Face ``model.safetensors`` (+ ``config.json``, ``tokenizer.json``,
    ``generation_config.json`` with ``alignment_heads``): transformers' Whisper key names;
  * an OpenAI-keyed ``model.safetensors`` (keys start with ``encoder.`` / ``decoder.``).
A plain size name ("tiny", "large-v3", ...) is not a directory: ``load_model_dir`` returns None and
the caller falls back to seeded synthetic weights (there is no network and no checkpoint cache).
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional, Tuple

import torch

from .weights import ModelDims, dims_from_state_dict

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


def load_model_dir(path: str) -> Optional[Tuple[Dict[str, torch.Tensor], ModelDims, Optional[str], Optional[List[Tuple[int, int]]]]]:
    if not path or not os.path.isdir(path):
        return None
    st_path = os.path.join(path, "model.safetensors")
    if not os.path.exists(st_path):
        if os.path.exists(os.path.join(path, "model.bin")):
            raise NotImplementedError(
                "CTranslate2 model.bin checkpoints are not supported yet; convert to safetensors "
                "(Hugging Face or OpenAI key names)")
        raise FileNotFoundError(f"{path}: no model.safetensors")
    from safetensors.torch import load_file
    raw = load_file(st_path)
    sd = hf_to_openai_keys(raw) if any(k.startswith("model.") for k in raw) else {k: v.to(torch.float32) for k, v in raw.items()}
    if "encoder.positional_embedding" not in sd:
        from .weights import sinusoids
        d = sd["encoder.conv1.weight"].shape[0]
        sd["encoder.positional_embedding"] = sinusoids(1500, d)
    dims = dims_from_state_dict(sd)
    tj = os.path.join(path, "tokenizer.json")
    heads = None
    for name in ("generation_config.json", "config.json"):
        p = os.path.join(path, name)
        if os.path.exists(p):
            with open(p) as f:
                cfg = json.load(f)
            if cfg.get("alignment_heads"):
                heads = [tuple(x) for x in cfg["alignment_heads"]]
                break
    return sd, dims, tj if os.path.exists(tj) else None, heads

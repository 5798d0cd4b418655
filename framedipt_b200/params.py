"""Parameter inventory of the score network (the `state_dict` contract).

Key names and shapes are those of the reference's ``ScoreNetwork.state_dict()``
(framedipt/model/score_network.py:67-112, framedipt/model/ipa_pytorch.py:105-168, 416-459) so a
checkpoint written for the reference loads unchanged.  ``param_specs`` is used by
``ScoreNetwork.load_state_dict`` for validation and by ``synthetic_state_dict`` to build
deterministic random weights when no checkpoint is available (benchmarks / tests: the published
weights are a HuggingFace download and there is no network).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch


@dataclass(frozen=True)
class ModelDims:
    """config/base.yaml:55-79."""

    c_s: int = 256
    c_z: int = 128
    c_hidden: int = 256
    c_skip: int = 64
    no_heads: int = 8
    no_qk_points: int = 8
    no_v_points: int = 12
    num_blocks: int = 4
    index_embed_size: int = 32
    num_bins: int = 22
    min_bin: float = 1e-5
    max_bin: float = 20.0
    seq_tfmr_num_heads: int = 4
    seq_tfmr_num_layers: int = 2
    coordinate_scaling: float = 0.1
    embed_self_conditioning: bool = True  # model.embed.embed_self_conditioning (score_network.py:95-96, 185)

    @property
    def concat_dim(self) -> int:
        return self.no_heads * (self.c_z // 4 + self.c_hidden + self.no_v_points * 4)


def node_feat_dim(d: ModelDims, with_aatype: bool) -> int:
    return d.index_embed_size + 1 + (21 if with_aatype else 0)


def param_specs(d: ModelDims = ModelDims(), with_aatype: bool = True):
    """Yields (key, shape, kind); kind in {w, relu, final, bias, bias_default, bias_final, ln_w, ln_b, head, unused}."""
    f1 = node_feat_dim(d, with_aatype)
    node_in = f1 + d.index_embed_size
    edge_in = 2 * f1 + d.index_embed_size + (d.num_bins if d.embed_self_conditioning else 0)
    out = []

    def lin(name, o, i, kind="w", bias="bias"):
        if kind == "final" and bias == "bias":
            bias = "bias_final"
        out.append((name + ".weight", (o, i), kind))
        out.append((name + ".bias", (o,), bias))

    def ln(name, c):
        out.append((name + ".weight", (c,), "ln_w"))
        out.append((name + ".bias", (c,), "ln_b"))

    for pre, cin, c in (("embedding_layer.node_embedder", node_in, d.c_s), ("embedding_layer.edge_embedder", edge_in, d.c_z)):
        lin(pre + ".0", c, cin, "w", "bias_default")
        lin(pre + ".2", c, c, "w", "bias_default")
        lin(pre + ".4", c, c, "w", "bias_default")
        ln(pre + ".5", c)
    t = "score_model.trunk."
    hc = d.no_heads * d.c_hidden
    dt = d.c_s + d.c_skip
    for b in range(d.num_blocks):
        p = f"{t}ipa_{b}"
        out.append((p + ".head_weights", (d.no_heads,), "head"))
        lin(p + ".linear_q", hc, d.c_s)
        lin(p + ".linear_kv", 2 * hc, d.c_s)
        lin(p + ".linear_q_points", d.no_heads * d.no_qk_points * 3, d.c_s)
        lin(p + ".linear_kv_points", d.no_heads * (d.no_qk_points + d.no_v_points) * 3, d.c_s)
        lin(p + ".linear_b", d.no_heads, d.c_z)
        lin(p + ".down_z", d.c_z // 4, d.c_z)
        lin(p + ".linear_out", d.c_s, d.concat_dim, "final")
        lin(p + ".linear_rbf", 1, 20, "unused", "unused")
        ln(f"{t}ipa_ln_{b}", d.c_s)
        lin(f"{t}skip_embed_{b}", d.c_skip, d.c_s, "final")
        for l in range(d.seq_tfmr_num_layers):
            q = f"{t}seq_tfmr_{b}.layers.{l}"
            out.append((q + ".self_attn.in_proj_weight", (3 * dt, dt), "w"))
            out.append((q + ".self_attn.in_proj_bias", (3 * dt,), "bias"))
            lin(q + ".self_attn.out_proj", dt, dt)
            lin(q + ".linear1", dt, dt, "w", "bias_default")
            lin(q + ".linear2", dt, dt, "w", "bias_default")
            ln(q + ".norm1", dt)
            ln(q + ".norm2", dt)
        lin(f"{t}post_tfmr_{b}", d.c_s, dt, "final")
        p = f"{t}node_transition_{b}"
        lin(p + ".linear_1", d.c_s, d.c_s, "relu")
        lin(p + ".linear_2", d.c_s, d.c_s, "relu")
        lin(p + ".linear_3", d.c_s, d.c_s, "final")
        ln(p + ".ln", d.c_s)
        lin(f"{t}bb_update_{b}.linear", 6, d.c_s, "final")
        if b < d.num_blocks - 1:
            p = f"{t}edge_transition_{b}"
            hid = d.c_z + d.c_s  # c_z + 2 * (c_s // 2)
            lin(p + ".initial_embed", d.c_s // 2, d.c_s, "relu")
            lin(p + ".trunk.0", hid, hid, "relu")
            lin(p + ".trunk.2", hid, hid, "relu")
            lin(p + ".final_layer", d.c_z, hid, "final")
            ln(p + ".layer_norm", d.c_z)
    p = "score_model.torsion_pred"
    lin(p + ".linear_1", d.c_s, d.c_s, "relu")
    lin(p + ".linear_2", d.c_s, d.c_s, "relu")
    lin(p + ".linear_3", d.c_s, d.c_s, "unused", "unused")
    lin(p + ".linear_final", 2, d.c_s, "final")
    return out


def synthetic_state_dict(seed: int = 0, d: ModelDims = ModelDims(), with_aatype: bool = True,
                         final_std: float = 0.002, bias_std: float = 0.01) -> dict[str, torch.Tensor]:
    """Deterministic random weights of realistic scale (float32, CPU).

    Weight scales follow the reference's initialisers (framedipt/model/layers.py:231-337: LeCun normal
    for default layers, He normal for ``relu``) and, as in SURVEY.md §8d, every layer the reference
    zero-initialises (``init="final"``) receives N(0, final_std) so that the network is not a no-op.
    Biases get a small N(0, bias_std) so bias paths are exercised; LayerNorm affine = 1/0 + small noise.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    sd = {}
    for key, shape, kind in param_specs(d, with_aatype):
        if kind in ("w", "relu", "unused"):
            fan_in = shape[-1]
            std = math.sqrt((2.0 if kind == "relu" else 1.0) / fan_in)
            v = torch.randn(shape, generator=g) * std
        elif kind == "final":
            v = torch.randn(shape, generator=g) * final_std
        elif kind in ("bias", "bias_default"):
            v = torch.randn(shape, generator=g) * bias_std
        elif kind == "bias_final":
            v = torch.randn(shape, generator=g) * final_std
        elif kind == "ln_w":
            v = 1.0 + 0.05 * torch.randn(shape, generator=g)
        elif kind == "ln_b":
            v = 0.05 * torch.randn(shape, generator=g)
        elif kind == "head":
            v = 0.541324854612918 + 0.1 * torch.randn(shape, generator=g)  # softplus^-1(1), layers.py:199-203
        else:  # pragma: no cover
            raise ValueError(kind)
        sd[key] = v.float().contiguous()
    return sd

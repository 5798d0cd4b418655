"""Default configuration values (config/base.yaml:33-79, config/inference.yaml:29-35 of the reference) as an
attribute-style dict that the shims accept in place of an OmegaConf DictConfig."""
from __future__ import annotations


class AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(d):
    return AttrDict({k: to_attr(v) for k, v in d.items()}) if isinstance(d, dict) else d


def default_conf(input_aatype: bool = True, seed: int | None = 123):
    return to_attr({
        "diffuser": {
            "diffuse_trans": True, "diffuse_rot": True,
            "r3": {"min_b": 0.1, "max_b": 20.0, "coordinate_scaling": 0.1, "seed": seed},
            "so3": {"num_omega": 1000, "num_sigma": 1000, "min_sigma": 0.1, "max_sigma": 1.5, "schedule": "logarithmic",
                    "cache_dir": ".cache/", "use_cached_score": False, "seed": seed},
        },
        "model": {
            "input_aatype": input_aatype, "node_embed_size": 256, "edge_embed_size": 128, "dropout": 0.0,
            "embed": {"index_embed_size": 32, "aatype_embed_size": 64, "embed_self_conditioning": True, "num_bins": 22,
                      "min_bin": 1e-5, "max_bin": 20.0},
            "ipa": {"c_s": 256, "c_z": 128, "c_hidden": 256, "c_skip": 64, "no_heads": 8, "no_qk_points": 8, "no_v_points": 12,
                    "seq_tfmr_num_heads": 4, "seq_tfmr_num_layers": 2, "num_blocks": 4, "coordinate_scaling": 0.1},
        },
        "inference": {"diffusion": {"num_t": 100, "noise_scale": 0.1, "min_t": 0.01}, "seed": seed},
    })

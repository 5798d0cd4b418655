"""framedipt.data.utils: the reference's own module when its dependencies import (everything then comes from it); the few functions the
sampler path needs (preprocess_aatype, pad_feats, pad_rigid, pad, move_to_np, read_pkl / write_pkl) otherwise."""
from framedipt_b200 import dropin as _d

_ref = _d.load_reference_module("framedipt/data/utils.py", "_framedipt_ref_data_utils")
if _ref is not None:
    globals().update({k: v for k, v in vars(_ref).items() if not k.startswith("__")})
else:
    import pickle

    import numpy as np
    import torch

    from framedipt_b200.sampler import PAIR_FEATS, RIGID_FEATS, UNPADDED_FEATS, pad, pad_feats, pad_rigid  # noqa: F401
    from framedipt_b200.score_network import preprocess_aatype  # noqa: F401

    def move_to_np(x):
        return x.cpu().detach().numpy()

    def read_pkl(read_path, verbose=True, use_torch=False, map_location=None):
        if use_torch:
            return torch.load(read_path, map_location=map_location, weights_only=False)
        with open(read_path, "rb") as handle:
            return pickle.load(handle)

    def write_pkl(save_path, pkl_data, create_dir=False, use_torch=False):
        import os

        if create_dir:
            os.makedirs(os.path.dirname(save_path), exist_ok=True)
        if use_torch:
            torch.save(pkl_data, save_path, pickle_protocol=pickle.HIGHEST_PROTOCOL)
        else:
            with open(save_path, "wb") as handle:
                pickle.dump(pkl_data, handle, protocol=pickle.HIGHEST_PROTOCOL)

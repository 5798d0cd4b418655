"""PDB output of sampled backbones with the reference's call surface (a "next" row).

``write_prot_to_pdb`` mirrors ``framedipt/analysis/utils.py:78-156`` (same arguments, same file-indexing rules, same text) and
``to_pdb_text`` the string produced by ``framedipt/protein/protein.py:165-279`` for one model; the formatting itself runs in
``libfdpt.so`` (``fdpt_to_pdb``, a host function).  Only the 5 backbone slots the sampler produces are supported: positions may
be given as atom37 ``[..., N, 37, 3]`` (slots >= 5 must be zero, as they are in the sampler's output) or compact ``[..., N, 5, 3]``.
"""
from __future__ import annotations

import ctypes as C
import os
import pathlib
import re

import numpy as np

from . import runtime


def _chain_layout(n: int, residue_index, chain_index):
    """create_full_prot's re-indexing (analysis/utils.py:45-60): chains renumbered 0.., residues 0.. within each chain."""
    final_res = np.arange(n)
    final_chain = np.zeros(n)
    if residue_index is not None and chain_index is not None:
        chain_index = np.asarray(chain_index)
        prev = 0
        for i, index in enumerate(np.unique(chain_index)):
            ln = int((chain_index == index).sum())
            final_chain[prev:prev + ln] = i
            final_res[prev:prev + ln] = np.arange(ln)
            prev += ln
    return final_res.astype(np.int32), final_chain.astype(np.int32)


def to_pdb_text(pos: np.ndarray, aatype=None, b_factors=None, residue_index=None, chain_index=None, model: int = 1,
                add_end: bool = False) -> str:
    pos = np.asarray(pos, np.float32)
    if pos.ndim != 3 or pos.shape[-1] != 3 or pos.shape[-2] not in (5, 37):
        raise ValueError(f"atom37 should have shape [..., 37, 3], got {pos.shape}.")
    if pos.shape[-2] == 37:
        if np.any(pos[:, 5:] != 0):
            raise ValueError("only backbone atoms (atom37 slots 0..4) are supported by the fast PDB writer")
        pos = pos[:, :5]
    n = pos.shape[0]
    pos = np.ascontiguousarray(pos)
    res_i, chain_i = _chain_layout(n, residue_index, chain_index)
    aa = None if aatype is None else np.ascontiguousarray(np.asarray(aatype), np.int32)
    if aa is not None and np.any(aa > 20):
        raise ValueError("Invalid aatypes.")
    bf = None
    if b_factors is not None:
        bf = np.asarray(b_factors, np.float32)
        bf = np.ascontiguousarray(bf[:, :5])
    cap = 128 * (5 * n + 8 + 64)
    buf = C.create_string_buffer(cap)
    ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
    nbytes = runtime.lib().fdpt_to_pdb(ptr(pos), ptr(aa), ptr(res_i), ptr(chain_i), ptr(bf), n, int(model), int(add_end), buf, cap)
    if nbytes < 0:
        raise ValueError(f"fdpt_to_pdb failed ({nbytes}): invalid aatype / more than 62 chains / buffer too small")
    return buf.raw[:nbytes].decode("ascii")


def write_prot_to_pdb(prot_pos: np.ndarray, file_path, aatype=None, overwrite: bool = False, no_indexing: bool = False,
                      b_factors=None, residue_index=None, chain_index=None) -> pathlib.Path:
    if isinstance(file_path, str):
        file_path = pathlib.Path(file_path)
    if overwrite:
        max_existing_idx = 0
    else:
        file_dir = os.path.dirname(file_path)
        file_name = os.path.basename(file_path).strip(".pdb")
        existing = [x for x in os.listdir(file_dir) if file_name in x]
        max_existing_idx = max([int(re.findall(r"_(\d+).pdb", x)[0]) for x in existing if re.findall(r"_(\d+).pdb", x)] + [0])
    save_path = file_path if no_indexing else file_path.with_name(f"{file_path.stem}_{max_existing_idx + 1}.pdb")
    prot_pos = np.asarray(prot_pos)
    with open(save_path, "w", encoding="utf-8") as f:
        if prot_pos.ndim == 4:
            for t, pos in enumerate(prot_pos):
                f.write(to_pdb_text(pos, aatype, b_factors, residue_index, chain_index, model=t + 1, add_end=False))
        elif prot_pos.ndim == 3:
            f.write(to_pdb_text(prot_pos, aatype, b_factors, residue_index, chain_index, model=1, add_end=False))
        else:
            raise ValueError(f"Invalid positions shape {prot_pos.shape}")
        f.write("END")
    return save_path

"""PDB writer: byte-exact against text written by the unmodified reference (oracle/make_golden_pdb.py).
Host-only code in libfdpt.so: runs without a GPU."""
import os

import numpy as np
import pytest


@pytest.fixture(scope="module")
def g(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "pdb_text.npz")))


def test_trajectory_multichain_text_is_byte_exact(g, tmp_path):
    from framedipt_b200.pdb import write_prot_to_pdb

    p = write_prot_to_pdb(g["pos"], tmp_path / "a.pdb", aatype=g["aatype"], no_indexing=True, b_factors=g["b_factors"],
                          residue_index=g["residue_index"], chain_index=g["chain_index"])
    got, ref = open(p).read(), str(g["traj_multichain"])
    assert len(got) == len(ref)
    assert got == ref


def test_single_model_defaults_and_compact_input(g, tmp_path):
    from framedipt_b200.pdb import write_prot_to_pdb

    ref = str(g["single_default"])
    assert open(write_prot_to_pdb(g["pos"][2], tmp_path / "b.pdb", no_indexing=True)).read() == ref
    assert open(write_prot_to_pdb(g["pos"][2][:, :5], tmp_path / "c.pdb", no_indexing=True)).read() == ref  # [N,5,3] sampler layout


def test_file_indexing_rules(g, tmp_path):
    from framedipt_b200.pdb import write_prot_to_pdb

    a = write_prot_to_pdb(g["pos"][0], tmp_path / "s.pdb")
    b = write_prot_to_pdb(g["pos"][0], tmp_path / "s.pdb")
    c = write_prot_to_pdb(g["pos"][0], tmp_path / "s.pdb", overwrite=True)
    assert (a.name, b.name, c.name) == ("s_1.pdb", "s_2.pdb", "s_1.pdb")


def test_errors_like_the_reference(g, tmp_path):
    from framedipt_b200.pdb import to_pdb_text, write_prot_to_pdb

    with pytest.raises(ValueError):
        to_pdb_text(g["pos"][0], aatype=np.full(23, 21))          # "Invalid aatypes."
    with pytest.raises(ValueError):
        write_prot_to_pdb(g["pos"][0][0], tmp_path / "x.pdb")     # "Invalid positions shape"
    full = g["pos"][0].copy()
    full[0, 7] = 1.0
    with pytest.raises(ValueError):
        to_pdb_text(full)                                         # side-chain atoms are outside the fast writer's scope

"""TEST INFRASTRUCTURE ONLY — golden PDB text from the UNMODIFIED reference (framedipt.analysis.utils.write_prot_to_pdb).

Run in the build container (needs /root/reference):   python oracle/make_golden_pdb.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402

rh.install_stubs()
sys.path.insert(0, rh.ref_root())
from framedipt.analysis import utils as au  # noqa: E402

rs = np.random.RandomState(7)
n = 23
pos = np.zeros((3, n, 37, 3), np.float32)
pos[:, :, :5] = (rs.normal(size=(3, n, 5, 3)) * 40).astype(np.float32)
pos[0, 3, 3] = 0  # a GLY-like missing CB
pos[1, 0, :5] = [[1234.5678, -999.9996, 0.0005], [0.0015, 0.0025, -0.0035], [9.9995, 99.9995, -99.9995], [1, 2, 3], [-0.0004, 7, 8]]
aatype = rs.randint(0, 21, size=n)
chain_index = np.array([5] * 9 + [2] * 6 + [9] * 8)       # unsorted chain ids, re-indexed by create_full_prot
residue_index = rs.randint(0, 500, size=n)
b = (rs.rand(n, 37) * 100).astype(np.float32)
out = {}
with tempfile.TemporaryDirectory() as d:
    p = au.write_prot_to_pdb(pos, os.path.join(d, "a.pdb"), aatype=aatype, no_indexing=True, b_factors=b, residue_index=residue_index,
                             chain_index=chain_index)
    out["traj_multichain"] = open(p).read()
    p = au.write_prot_to_pdb(pos[2], os.path.join(d, "b.pdb"), no_indexing=True)
    out["single_default"] = open(p).read()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pdb_text.npz"), pos=pos, aatype=aatype, chain_index=chain_index,
                    residue_index=residue_index, b_factors=b, **{k: np.array(v) for k, v in out.items()})
print({k: len(v) for k, v in out.items()})

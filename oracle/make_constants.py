"""TEST INFRASTRUCTURE ONLY — provenance of framedipt_b200/backbone_constants.py.

Derives the three backbone tables the CUDA `backbone_kernel` uses from the reference's own constants
(framedipt/protein/residue_constants.py via framedipt/protein/all_atom.py:10-16) and
  * writes them to tests/golden/backbone_tables.npz (the CPU test tests/test_host.py::test_backbone_constants_match_reference_tables
    pins framedipt_b200/backbone_constants.py to it for all 20 residue types), and
  * with --write regenerates framedipt_b200/backbone_constants.py itself.

Run in the build container (needs /root/reference):   python oracle/make_constants.py [--write]
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402


def tables():
    rh.load_reference()
    from framedipt.protein import all_atom  # type: ignore

    ideal = all_atom.IDEALIZED_POS.numpy()[:20, :5].astype(np.float32)          # [20, 5, 3] atom14 order N, CA, C, O, CB
    psi_frame = all_atom.DEFAULT_FRAMES.numpy()[:20, 3].astype(np.float32)      # [20, 4, 4] rigid group 3 (psi)
    mask = all_atom.ATOM_MASK.numpy()[:20, :5].astype(np.float32)               # [20, 5]
    group = all_atom.GROUP_IDX.numpy()[:20, :5]
    assert (group[:, [0, 1, 2, 4]] == 0).all() and (group[:, 3] == 3).all()     # N, CA, C, CB in the backbone frame; O in the psi frame
    return ideal, psi_frame, mask


def main():
    ideal, psi_frame, mask = tables()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "backbone_tables.npz"), ideal=ideal, psi_frame=psi_frame, mask=mask)
    if "--write" in sys.argv:
        np.set_printoptions(threshold=100000, precision=8, suppress=True)
        with open(os.path.join(ROOT, "framedipt_b200", "backbone_constants.py")) as f:
            head = f.read().split("import numpy as np")[0]
        with open(os.path.join(ROOT, "framedipt_b200", "backbone_constants.py"), "w") as f:
            f.write(head + "import numpy as np\n\n")
            f.write("# [20, 5, 3] atom14 order N, CA, C, O, CB; N/CA/C/CB live in the backbone frame (group 0), O in the psi frame (group 3)\n")
            f.write("IDEAL_BB_POS = np." + repr(ideal).replace("dtype=float32", "dtype=np.float32") + "\n\n")
            f.write("# [20, 4, 4] default frame of rigid group 3 (psi) relative to the backbone frame\n")
            f.write("PSI_DEFAULT_FRAME = np." + repr(psi_frame).replace("dtype=float32", "dtype=np.float32") + "\n\n")
            f.write("# [20, 5] atom14 existence mask (GLY has no CB)\n")
            f.write("BB_ATOM_MASK = np." + repr(mask).replace("dtype=float32", "dtype=np.float32") + "\n")
    print("ok", ideal.shape, psi_frame.shape, mask.shape)


if __name__ == "__main__":
    main()

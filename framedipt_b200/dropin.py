"""Drop-in wiring: makes the reference's own module paths resolve to this package for the sampler hot path, so that
`experiments/inference.py` of instadeepai/FrameDiPT runs UNCHANGED on the B200 path (BASELINE.json north star).

The directory `shim/` at the repository root holds overlay packages named like the reference's (`framedipt`, `experiments`,
`openfold`).  Put it on `sys.path` AHEAD of the reference checkout:

    PYTHONPATH=/path/to/framedipt_b200_repo/shim:/path/to/framedipt_b200_repo:/path/to/FrameDiPT  python experiments/inference.py ...

Each overlay package extends its `__path__` with the same-named directory of the reference checkout (found on `sys.path` or through
`$FRAMEDIPT_REF`), so every module the overlay does not provide — analysis, protein constants, PDB IO, ProteinMPNN glue ... — is still the
reference's own file, untouched.  What the overlay provides, under the reference's names:

    framedipt.model.score_network      ScoreNetwork (+ Embedder helpers)           -> framedipt_b200.score_network / runtime
    framedipt.diffusion.se3_diffuser   SE3Diffuser                                 -> framedipt_b200.se3_diffuser
    experiments.utils                  everything of the reference's module, with  inference_fn, logp_confidence_score -> framedipt_b200.inference
    experiments.sampler                the reference's module when its dependencies import, plus UnconditionalSampler / synthetic samplers
    openfold.utils.rigid_utils         the reference's module when present (pure torch), else framedipt_b200.rigid

Nothing here runs on the hot path; it is import plumbing.
"""
from __future__ import annotations

import importlib.util
import os
import sys


def reference_root() -> str | None:
    """The reference checkout: $FRAMEDIPT_REF, or the first sys.path entry that holds framedipt/model/ipa_pytorch.py."""
    cands = [os.environ.get("FRAMEDIPT_REF")] + list(sys.path)
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "framedipt", "model", "ipa_pytorch.py")):
            return os.path.abspath(c)
    return None


def overlay_path(pkg_file: str, rel: str) -> list[str]:
    """`__path__` of an overlay package: its own directory first, then the reference's directory `rel` (e.g. "framedipt/model")."""
    here = os.path.dirname(os.path.abspath(pkg_file))
    ref = reference_root()
    out = [here]
    if ref:
        d = os.path.join(ref, *rel.split("/"))
        if os.path.isdir(d) and os.path.abspath(d) != here:
            out.append(d)
    return out


def load_reference_module(rel_file: str, alias: str):
    """Executes the reference's own source file `rel_file` (e.g. "experiments/utils.py") as module `alias`; None when the reference
    checkout (or one of the file's imports) is unavailable."""
    ref = reference_root()
    if not ref:
        return None
    path = os.path.join(ref, *rel_file.split("/"))
    if not os.path.isfile(path):
        return None
    if alias in sys.modules:
        return sys.modules[alias]
    spec = importlib.util.spec_from_file_location(alias, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    try:
        spec.loader.exec_module(mod)
    except Exception as e:
        sys.modules.pop(alias, None)
        if os.environ.get("FDPT_DROPIN_DEBUG"):
            print(f"[framedipt_b200.dropin] {rel_file} of the reference did not import ({type(e).__name__}: {e}); using the built-in subset", file=sys.stderr)
        return None
    return mod

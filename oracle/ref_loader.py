"""Locate and import the UNMODIFIED reference (thomas-schillaci/SimPLe).  TEST INFRASTRUCTURE ONLY.

Search order: $SIMPLE_REF_ROOT, /root/reference (build container), <repo>/baseline/_ref
(git-ignored copy made by `python -m oracle.build_ref`, which travels to the GPU box).
`gym` / `baselines` are not installable offline; on-disk shims under oracle/shims are put on
sys.path and PYTHONPATH (forkserver workers re-import them).
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)


def find_reference_root():
    for cand in (os.environ.get('SIMPLE_REF_ROOT'), '/root/reference', os.path.join(_REPO, 'baseline', '_ref')):
        if cand and os.path.isfile(os.path.join(cand, 'simple', 'next_frame_predictor.py')):
            return cand
    return None


def load_reference():
    """Returns a namespace of reference modules, or None if the reference is not reachable."""
    root = find_reference_root()
    if root is None:
        return None
    shims = os.path.join(_HERE, 'shims')
    paths = [shims, root, os.path.join(root, 'atari_utils'), os.path.join(root, 'a2c_ppo_acktr')]
    for p in paths:
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ['PYTHONPATH'] = os.pathsep.join(paths + [os.environ.get('PYTHONPATH', '')])
    from types import SimpleNamespace
    import simple.next_frame_predictor as nfp
    import simple.adafactor as adafactor
    import simple.trainer as trainer
    import simple.utils as utils
    return SimpleNamespace(root=root, nfp=nfp, adafactor=adafactor, trainer=trainer, utils=utils)

"""TEST / BENCH INFRASTRUCTURE — stages and imports the UNMODIFIED reference loss modules for the CPU baseline.

The reference (jizongFox/Contrast-You) is pure Python; its loss modules run on torch-CPU.  ``/root/reference`` only
exists in the build container, so ``stage()`` (called by ``__graft_entry__.build()`` there) copies the Python
packages the modules import from (``contrastyou/``, ``semi_seg/``, ``script/``; ~1.3 MB, .py files only) to ``baseline/_ref/`` — a
git-ignored directory that still travels to the GPU box with the working-tree snapshot.  Nothing of it is tracked,
imported by the product path, or modified: ``bench.py --impl reference`` and the ``cpu_baseline`` leg time the
reference's own ``SupConLoss1`` / ``IIDSegmentationLoss`` through ``load()``; when ``baseline/_ref`` is absent the
callers fall back to the C/OpenMP port (oracle/oracle.c) and say so (``kind: "port"``).

Import shims: ``contrastyou.losses.discreteMI`` imports plotting / medical-imaging packages that are
not installed in this image (termcolor, matplotlib, medpy) and ``semi_seg.hooks.midl`` (which would pull the whole hook
tree); they are replaced by empty stand-ins BEFORE import.  No arithmetic is shimmed.
"""
import os
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "baseline", "_ref")
SOURCE = os.environ.get("CY_REFERENCE", "/root/reference")
_PACKAGES = ("contrastyou", "semi_seg", "script")


def available() -> bool:
    return os.path.isfile(os.path.join(STAGED, "contrastyou", "losses", "contrastive.py"))


def stage(force: bool = False) -> bool:
    """copy the reference's python packages to baseline/_ref (build container only).  Returns available()."""
    if not os.path.isdir(os.path.join(SOURCE, "contrastyou")):
        return available()
    if available() and not force:
        return True
    os.makedirs(STAGED, exist_ok=True)
    for pkg in _PACKAGES:
        dst = os.path.join(STAGED, pkg)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(SOURCE, pkg), dst,
                        ignore=lambda d, names: [n for n in names if not (n.endswith(".py") or os.path.isdir(os.path.join(d, n)))])
    return available()


def load():
    """import the staged reference and return (SupConLoss1, SelfPacedSupConLoss, IIDSegmentationLoss) classes"""
    if not available():
        raise RuntimeError("baseline/_ref is not staged (run __graft_entry__.build() in the build container)")
    if STAGED not in sys.path:
        sys.path.insert(0, STAGED)
    os.environ.setdefault("LOGURU_LEVEL", "ERROR")

    def shim(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        m.__dict__.update(attrs)
        return m

    shim("termcolor", colored=lambda s, *a, **k: s)
    mpl = shim("matplotlib", use=lambda *a, **k: None, get_backend=lambda: "agg")
    mpl.pyplot = shim("matplotlib.pyplot", switch_backend=lambda *a, **k: None)
    medpy = shim("medpy")
    medpy.metric = shim("medpy.metric", assd=None)
    medpy.metric.binary = shim("medpy.metric.binary", __surface_distances=None)
    try:
        from loguru import logger
        logger.disable("contrastyou")
    except Exception:  # noqa
        pass
    import contrastyou  # noqa: F401  (mkdirs .data/ runs/ config/ opt/ inside baseline/_ref; logs a git warning)
    from contrastyou.losses.kl import Entropy
    pkg = shim("semi_seg")
    pkg.__path__ = [os.path.join(STAGED, "semi_seg")]
    hooks = shim("semi_seg.hooks")
    hooks.__path__ = [os.path.join(STAGED, "semi_seg", "hooks")]
    shim("semi_seg.hooks.midl", entropy_criterion=Entropy(reduction="none", eps=1e-8))      # semi_seg/hooks/midl.py:13
    from contrastyou.losses.contrastive import SupConLoss1, SelfPacedSupConLoss
    from contrastyou.losses.discreteMI import IIDSegmentationLoss
    return SupConLoss1, SelfPacedSupConLoss, IIDSegmentationLoss

"""Recipe for oracle/_ref: the UNMODIFIED reference modules of the hot path, byte-compiled where they lie.

The reference is pure Python, so "compiling it from its own sources" means py_compile: every module the timed
reference arm needs (model/*.py, trainer/trainer.py, evaluator/evaluator.py, sampler/sampler.py under /root/reference) is
compiled to sourceless bytecode (``<module>.bin``: a .pyc under a neutral extension, because snapshot tools skip ``*.pyc``)
under ``oracle/_ref/`` -- outputs only, no reference source is copied into the repo.
``oracle/_ref/`` is git-ignored but travels to the GPU box (like our own ``.so``), where ``/root/reference`` does not
exist; the box runs the same image, hence the same CPython bytecode magic.

Used by ``bench.py --impl reference`` / ``cpu_baseline`` (kind "reference") and by the "eager reference on cuda" numbers.
TEST / MEASUREMENT INFRASTRUCTURE ONLY: nothing in the product package imports it.

    python oracle/build_ref.py            # no-op (exit 0) when /root/reference is absent
"""
import glob
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("RS_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
PACKAGES = ("model", "trainer", "evaluator", "sampler")


def build_ref(verbose=False):
    if not os.path.isdir(REF):
        return None
    n = 0
    for pkg in PACKAGES:
        src_dir = os.path.join(REF, pkg)
        if not os.path.isdir(src_dir):
            continue
        dst_dir = os.path.join(OUT, pkg)
        os.makedirs(dst_dir, exist_ok=True)
        for src in sorted(glob.glob(os.path.join(src_dir, "*.py"))):
            name = os.path.splitext(os.path.basename(src))[0]
            if name == "__init__":
                continue
            py_compile.compile(src, cfile=os.path.join(dst_dir, name + ".bin"), dfile=f"<reference>/{pkg}/{name}.py", doraise=True)
            n += 1
    with open(os.path.join(OUT, "MANIFEST.txt"), "w") as f:
        f.write(f"byte-compiled from {REF} by oracle/build_ref.py with CPython {sys.version.split()[0]}: {n} modules of {PACKAGES}\n")
    if verbose:
        print(f"oracle/_ref: {n} reference modules byte-compiled")
    return OUT


def import_ref():
    """-> dict of the reference's hot-path modules loaded from oracle/_ref (sourceless bytecode), or None when it was never
    built.  The reference's top-level package names (model, trainer, evaluator, sampler) are registered only while
    loading, so they never shadow anything of this repo."""
    if not os.path.exists(os.path.join(OUT, "MANIFEST.txt")):
        return None
    import importlib.machinery
    import importlib.util
    import types
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split(".")[0] in PACKAGES}
    try:
        for pkg in PACKAGES:
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m
        mods = {}
        for name in ("evaluator.evaluator", "model.lr", "model.mf", "model.ffm", "model.deepfm", "model.afm", "model.nfm", "model.pnn",
                     "model.din", "model.dien", "model.neuralcf", "trainer.trainer"):
            pkg, mod = name.split(".")
            path = os.path.join(OUT, pkg, mod + ".bin")
            loader = importlib.machinery.SourcelessFileLoader(name, path)
            spec = importlib.util.spec_from_loader(name, loader)
            module = importlib.util.module_from_spec(spec)
            sys.modules[name] = module
            loader.exec_module(module)
            setattr(sys.modules[pkg], mod, module)
            mods[name] = module
        return mods
    finally:
        for k in list(sys.modules):
            if k.split(".")[0] in PACKAGES:
                del sys.modules[k]
        sys.modules.update(saved)


if __name__ == "__main__":
    out = build_ref(verbose=True)
    if out is None:
        print(f"{REF} not present: oracle/_ref left as it is")

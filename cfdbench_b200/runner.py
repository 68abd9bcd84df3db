"""Run the reference's own scripts (train_auto.py / test_multistep.py) UNCHANGED on top of this package.

    python -m cfdbench_b200.runner /path/to/CFDBench/src train_auto.py --model fno --loss_name nmse ...

What it does (no file of the reference is edited; SURVEY.md 3.1 lists why each step is needed):
  1. puts `<src>` first on sys.path so `models.base_model.AutoCfdModel` is the reference's class
     (test_multistep.py:109 checks isinstance against it) -- `cfdbench_b200.base_model` picks it up;
  2. rebinds `models.fno.fno2d.Fno2d` to `cfdbench_b200.Fno2d` BEFORE `utils.autoregressive` imports it
     (utils/autoregressive.py:10 is the plug-in seam);
  3. optionally (`--stub-missing`) provides stand-ins for packages that are NOT installed: `tap.Tap` (a minimal
     typed-argument-parser, cfdbench_b200/_stubs.py) and no-op modules for what only the fork's plotting / diffusion /
     VAE code needs (matplotlib, diffusers, sklearn, seaborn, accelerate, ...), and injects `Args.lr_step_size` which
     train_auto.py:357 reads but args.py never declares (defect 1);
  4. runs the script as `__main__`.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import os
import runpy
import sys
import types

STUBBABLE = ("diffusers", "sklearn", "seaborn", "accelerate", "matplotlib", "h5py", "diffsci")


class _StubLoader(importlib.abc.Loader):
    def create_module(self, spec):
        from ._stubs import make_stub_module
        return make_stub_module(spec.name)

    def exec_module(self, module):
        pass


class _StubFinder(importlib.abc.MetaPathFinder):
    def __init__(self, roots):
        self.roots = tuple(roots)

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in self.roots:
            return importlib.machinery.ModuleSpec(fullname, _StubLoader(), is_package=True)
        return None


def missing(roots=STUBBABLE):
    out = []
    for r in roots:
        try:
            importlib.import_module(r)
        except Exception:  # noqa: BLE001
            out.append(r)
    return out


def install(src_dir: str, stub_missing: bool = False, act_dtype: str | None = None) -> None:
    """Steps 1-3 above.  Safe to call from tests; does not import any reference script."""
    src_dir = os.path.abspath(src_dir)
    if not os.path.isdir(os.path.join(src_dir, "models")):
        raise FileNotFoundError(f"{src_dir} does not look like CFDBench/src")
    if src_dir not in sys.path:
        sys.path.insert(0, src_dir)
    sys.dont_write_bytecode = True
    if stub_missing:
        from ._stubs import install_tap
        install_tap()   # `tap.Tap` (reference args.py:1) when typed-argument-parser is not installed
        gone = missing()
        if gone:
            sys.meta_path.append(_StubFinder(gone))
    import cfdbench_b200.base_model as bm
    if not bm.USING_REFERENCE_BASE:
        importlib.reload(bm)  # now that <src> is importable, bind to the reference's AutoCfdModel
    import cfdbench_b200.fno2d as ours
    importlib.reload(ours)
    ref = importlib.import_module("models.fno.fno2d")
    new_cls = ours.Fno2d
    if act_dtype is not None:
        base = ours.Fno2d

        class Fno2d(base):  # type: ignore[misc,valid-type]
            def __init__(self, *a, **k):
                k.setdefault("act_dtype", act_dtype)
                super().__init__(*a, **k)
        new_cls = Fno2d
    ref.Fno2d = new_cls
    try:  # defect 1: train_auto.py:357 / train.py:329 read args.lr_step_size
        args_mod = importlib.import_module("args")
        if not hasattr(args_mod.Args, "lr_step_size"):
            args_mod.Args.lr_step_size = 20
            args_mod.Args.__annotations__["lr_step_size"] = int
    except Exception:  # noqa: BLE001  (tap not installed: the scripts cannot run anyway)
        pass


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    stub = "--stub-missing" in argv
    if stub:
        argv.remove("--stub-missing")
    act = None
    if "--act-dtype" in argv:
        i = argv.index("--act-dtype")
        act = argv[i + 1]
        del argv[i:i + 2]
    if len(argv) < 2:
        raise SystemExit(__doc__)
    src, script = argv[0], argv[1]
    install(src, stub_missing=stub, act_dtype=act)
    sys.argv = [os.path.join(src, script)] + argv[2:]
    os.chdir(src)
    runpy.run_path(os.path.join(src, script), run_name="__main__")


if __name__ == "__main__":
    main()

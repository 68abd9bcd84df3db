"""Import stand-ins the runner installs ONLY for packages that are not installed (cfdbench_b200/runner.py).

`tap` (typed-argument-parser) is what the reference's `args.py:1` builds its CLI on; `matplotlib` and the
diffusion / VAE stacks are imported at module scope by files the FNO path never executes.
These stand-ins implement exactly what `train_auto.py` / `test_multistep.py` touch:
  * `Tap`: class annotations + defaults -> argparse (`--name value`, `--flag` for bool False defaults, `List[int]`
    as nargs="*"), `parse_args()`, `save(path)` (JSON), `as_dict()`, printable.
  * everything else: `Anything`, an object that absorbs attribute access, calls, indexing, iteration (as a pair:
    `fig, axs = plt.subplots(...)`) and context management, so plotting calls become no-ops.
"""
from __future__ import annotations

import argparse
import json
import sys
import types
import typing


class Tap:
    """Minimal typed-argument-parser: enough for reference src/args.py (plain int/float/str/bool/List[...] fields)."""

    def __init__(self, *args, **kwargs):
        self._fields = self._collect()

    @classmethod
    def _collect(cls):
        fields = {}
        for klass in reversed(cls.__mro__):
            ann = klass.__dict__.get("__annotations__", {})
            for name, tp in ann.items():
                if name.startswith("_"):
                    continue
                fields[name] = tp
        return fields

    @staticmethod
    def _resolve(tp):
        if isinstance(tp, str):
            tp = {"int": int, "float": float, "str": str, "bool": bool}.get(tp, tp)
            if isinstance(tp, str):
                inner = tp.strip()
                if inner.startswith("List[") and inner.endswith("]"):
                    return typing.List[Tap._resolve(inner[5:-1])]
                return str
        return tp

    def parse_args(self, argv=None):
        ap = argparse.ArgumentParser()
        for name, tp in self._fields.items():
            tp = self._resolve(tp)
            default = getattr(type(self), name, None)
            origin = typing.get_origin(tp)
            if tp is bool:
                if default:
                    ap.add_argument(f"--{name}", type=lambda s: s.lower() in ("1", "true", "yes"), nargs="?",
                                    const=True, default=True)
                else:
                    ap.add_argument(f"--{name}", action="store_true", default=False)
            elif origin in (list, typing.List):
                (elt,) = typing.get_args(tp) or (str,)
                ap.add_argument(f"--{name}", type=elt, nargs="*", default=default)
            else:
                ap.add_argument(f"--{name}", type=tp if callable(tp) else str, default=default,
                                required=not hasattr(type(self), name))
        ns = ap.parse_args(argv)
        for k, v in vars(ns).items():
            setattr(self, k, v)
        return self

    def as_dict(self):
        return {k: getattr(self, k) for k in self._fields if hasattr(self, k)}

    def save(self, path, **_):
        with open(path, "w", encoding="utf8") as f:
            json.dump(self.as_dict(), f, indent=2, default=str)

    def __repr__(self):
        return f"{type(self).__name__}({self.as_dict()})"

    __str__ = __repr__


class Anything:
    """Absorbs whatever plotting / optional-dependency code does with it."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return Anything()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return Anything()

    def __getitem__(self, key):
        return Anything()

    def __setitem__(self, key, value):
        pass

    def __iter__(self):
        return iter((Anything(), Anything()))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def __mro_entries__(self, bases):  # `class X(stub.A, stub.B)` in files the FNO path never runs: distinct dummy bases
        return (type("StubBase", (), {"__init__": lambda self, *a, **k: None}),)

    def __bool__(self):
        return False


def make_stub_module(name: str) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__path__ = []  # a package: submodule imports resolve to further stubs

    def _getattr(attr):
        if attr.startswith("__") and attr.endswith("__"):
            raise AttributeError(attr)
        return Anything()

    m.__getattr__ = _getattr  # type: ignore[attr-defined]
    return m


def install_tap() -> bool:
    """Provide `tap.Tap` if typed-argument-parser is not installed.  Returns True when the stand-in is used."""
    try:
        import tap  # noqa: F401
        return False
    except ImportError:
        m = types.ModuleType("tap")
        m.Tap = Tap
        sys.modules["tap"] = m
        return True

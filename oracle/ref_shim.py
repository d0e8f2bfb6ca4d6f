"""Headless import shim for the UNMODIFIED reference at /root/reference.

TEST INFRASTRUCTURE ONLY (container-side): used by oracle/gen_golden.py to
produce the committed fixtures under tests/golden/ and by the container-only
tests that pin the C restatement against the live reference.  Nothing in the
product package imports this, and nothing that runs on the GPU box does
(/root/reference does not exist there).

What is shimmed -- no reference file is edited or copied:
  * cv2.imshow / cv2.waitKey -> no-ops (headless OpenCV raises otherwise);
  * stub ``matplotlib`` / ``matplotlib.pyplot`` (not installed; imported by
    UavFntsmcParam/collector.py:3);
  * sys.path entries for the bare ``from uav import ...`` style imports;
  * sys.dont_write_bytecode (the reference mount is read-only);
  * stdout silenced around reference calls (terminal prints on the hot path).
UavFntsmcParam and UavRobust define clashing top-level module names (uav,
FNTSMC, ...): ``use_family`` purges them and re-orders sys.path.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
# The live source tree in the build container; on the GPU box (no /root/reference) the byte-compiled staging of the same
# files made by oracle/stage_reference.py (sourceless .pyb bytecode modules under oracle/_ref/, git-ignored).
REF = os.environ.get("RLP_REFERENCE") or ("/root/reference" if os.path.isdir("/root/reference/environment")
                                          else os.path.join(HERE, "_ref"))
_CLASH = ("uav", "FNTSMC", "collector", "ref_cmd", "uav_att_ctrl", "uav_pos_ctrl", "uav_att_ctrl_RL",
          "uav_pos_ctrl_RL", "UavHover", "UavHoverOuterLoop", "UavInnerLoop", "UavTrackingOuterLoop", "Color")
_installed = False


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "environment"))


def install() -> None:
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference not found at {REF}")
    sys.dont_write_bytecode = True
    import cv2
    cv2.imshow = lambda *a, **k: None
    cv2.waitKey = lambda *a, **k: 0
    cv2.destroyAllWindows = lambda *a, **k: None
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            m = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            for name in ("figure", "plot", "show", "legend", "grid", "xlabel", "ylabel", "title", "subplot",
                         "ylim", "xlim", "yticks", "xticks", "pause", "ion", "ioff", "savefig", "close"):
                setattr(plt, name, lambda *a, **k: None)
            m.pyplot = plt
            sys.modules["matplotlib"] = m
            sys.modules["matplotlib.pyplot"] = plt
    if os.path.exists(os.path.join(REF, "STAGED")):
        # byte-compiled staging (oracle/stage_reference.py): modules are <name>.pyb files; teach the path finder about them
        import importlib.machinery as M
        hook = M.FileFinder.path_hook((M.SourcelessFileLoader, [".pyb"]), (M.SourceFileLoader, M.SOURCE_SUFFIXES),
                                      (M.ExtensionFileLoader, M.EXTENSION_SUFFIXES))
        sys.path_hooks.insert(0, hook)
        sys.path_importer_cache.clear()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    _installed = True


def use_family(subdir: str) -> None:
    """Put environment/<subdir> first on sys.path and purge clashing bare modules."""
    install()
    d = os.path.join(REF, "environment", subdir)
    for name in _CLASH:
        sys.modules.pop(name, None)
    for other in ("UavFntsmcParam", "UavRobust", "UGVForwardObstacleAvoidance"):
        p = os.path.join(REF, "environment", other)
        while p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, d)
    importlib.invalidate_caches()


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def load(modname: str):
    install()
    with quiet():
        return importlib.import_module(modname)


def load_file(path: str, name: str):
    """Import a demo-copy env file (e.g. demonstration/PPO2/.../cartpole_angleonly.py) by path."""
    install()
    import importlib.machinery
    import importlib.util
    full = os.path.join(REF, path)
    staged = full[:-3] + ".pyb"
    if not os.path.exists(full) and os.path.exists(staged):            # staged bytecode (oracle/_ref)
        loader = importlib.machinery.SourcelessFileLoader(name, staged)
        spec = importlib.util.spec_from_loader(name, loader, origin=staged)
    else:
        spec = importlib.util.spec_from_file_location(name, full)
    mod = importlib.util.module_from_spec(spec)
    with quiet():
        spec.loader.exec_module(mod)
    return mod

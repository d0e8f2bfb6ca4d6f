"""Stage the UNMODIFIED Python reference for the GPU box: compile the modules of the hot path to bytecode.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, and reference SOURCES must not be
copied into the repo.  The reference is Python, so its "build" is `py_compile`: every module the timed loops import is
compiled FROM WHERE IT LIES under /root/reference INTO oracle/_ref/<same relative path>.pyb -- ordinary CPython bytecode
files under an extension of their own, because the GPU-box snapshot drops *.pyc; oracle/ref_shim.py installs an import
hook that loads them (importlib SourcelessFileLoader).
oracle/_ref/ is git-ignored (never in history) but not gpurun-ignored, so it travels to the GPU box like the built .so
files, where bench.py's CPU legs run it (ref_shim with RLP_REFERENCE=oracle/_ref).  Called by __graft_entry__.build()
when /root/reference is present.

    python oracle/stage_reference.py   ->  oracle/_ref/{environment,algorithm,utils,demonstration/...}/**/*.pyb
"""
import os
import py_compile
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("RLP_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")

# directories whose .py files the adapters of oracle/ref_adapters.py (and bench_reference_python.config1) import
TREES = ["environment", "algorithm", "utils",
         "demonstration/PPO2/PPO2-4-CartPoleAngleOnly",
         "demonstration/PPO2/PPO2-4-FlightAttitudeSimulator",
         "demonstration/DPPO2/DPPO2-4-SecondOrderIntegration",
         "demonstration/DPPO2/DPPO2-4-UGVForwardObstacleAvoidance",
         "demonstration/PPO2/PPO2-4-UGVForwardObstacleAvoidance"]


def stage() -> int:
    if not os.path.isdir(os.path.join(SRC, "environment")):
        raise SystemExit(f"reference not found at {SRC}")
    n = 0
    for tree in TREES:
        top = os.path.join(SRC, tree)
        if not os.path.isdir(top):
            continue
        for d, _, files in os.walk(top):
            rel = os.path.relpath(d, SRC)
            for f in files:
                if not f.endswith(".py"):
                    continue
                out = os.path.join(DST, rel, f[:-3] + ".pyb")
                os.makedirs(os.path.dirname(out), exist_ok=True)
                try:
                    warnings.simplefilter("ignore", SyntaxWarning)   # the reference's own '\d' docstrings
                    py_compile.compile(os.path.join(d, f), cfile=out, dfile=os.path.join(rel, f), doraise=True)
                    n += 1
                except py_compile.PyCompileError:
                    pass  # a demo script that does not parse is not on the timed path
    with open(os.path.join(DST, "STAGED"), "w") as fh:
        fh.write(f"{n} modules compiled from {SRC} with python {sys.version.split()[0]}\n")
    return n


if __name__ == "__main__":
    print(f"staged {stage()} modules -> {DST}")

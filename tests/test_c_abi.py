"""The C ABI from plain C (examples/c_abi_cartpole.c): compiles against include/b200env.h and links against
libb200env.so + the CUDA runtime only (CPU check); on a GPU the binary's final state equals the C oracle's on the same
inputs (no Python, no torch on the product side of this test)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

PKG = os.path.join(ROOT, "reinforcementlearningplatform_b200")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def build(tmp_path):
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(PKG, "libb200env.so")):
        g.build()
    exe = str(tmp_path / "c_abi_cartpole")
    cmd = ["gcc", "-O2", "-std=c99", os.path.join(ROOT, "examples", "c_abi_cartpole.c"), "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(CUDA, "include"), "-L" + PKG, "-lb200env", "-L" + os.path.join(CUDA, "lib64"), "-lcudart", "-lm",
           "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def env_vars():
    e = dict(os.environ)
    e["LD_LIBRARY_PATH"] = PKG + ":" + os.path.join(CUDA, "lib64") + ":" + e.get("LD_LIBRARY_PATH", "")
    return e


def test_c_example_compiles_and_links_against_the_header(tmp_path):
    exe = build(tmp_path)
    out = subprocess.run(["ldd", exe], capture_output=True, text=True, env=env_vars()).stdout
    assert "libb200env.so" in out and "libcudart" in out and "libtorch" not in out and "python" not in out.lower()


@pytest.mark.gpu
def test_c_example_matches_the_oracle(tmp_path, oracle_lib):
    from oracle import oracle
    import reinforcementlearningplatform_b200 as rlp
    from reinforcementlearningplatform_b200 import _lib
    exe = build(tmp_path)
    n, steps = 4096, 120
    dump = str(tmp_path / "state.bin")
    r = subprocess.run([exe, str(n), str(steps), dump], capture_output=True, text=True, env=env_vars(), timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    line = r.stdout.strip().splitlines()[-1]
    episodes = int(line.split("episodes=")[1].split()[0])
    reward_sum = float(line.split("reward_sum=")[1])
    raw = np.fromfile(dump, dtype=np.uint8)
    f64 = raw[: 8 * 5 * n].view(np.float64).reshape(5, n)
    ep = raw[8 * 5 * n:].view(np.uint32)
    host = rlp.CartPole(n_envs=n, host_only=True)
    orc = oracle.OracleEnv(_lib.CARTPOLE, host._params, n, 4, 4, 1, 0, seed=2024, auto_reset=True, nthreads=8)
    orc.reset()
    i = np.arange(n, dtype=np.float64)
    rs, eps = 0.0, 0
    for t in range(steps):
        orc.step((8.0 * np.sin(0.37 * i + 0.11 * t)).reshape(1, n))
        rs += float(orc.reward.sum())
        eps += int(orc.done.sum())
    assert episodes == eps and eps > 0
    assert np.array_equal(ep, orc.episode) and np.array_equal(f64[4], orc.time)
    np.testing.assert_allclose(f64[:4], orc.state, rtol=0, atol=1e-9)
    assert abs(reward_sum - rs) <= 1e-8 * abs(rs)

"""CPU: the C restatement (oracle/) against the fixtures recorded from the live reference.
This is the pin of the oracle: the reference ships no golden vectors, so the
fixtures under tests/golden/ -- outputs of the unmodified reference classes -- are the authority."""
import pytest

from helpers import OracleBackend, load_golden, replay

# name -> tolerance of the mixed metric.  0.0 = bit-exact (same glibc sin/cos as numpy, same operation order).
CASES = {
    "cartpole": 0.0,
    "cartpole_gentle": 0.0,
    "cartpole_angleonly_env": 0.0,
    "cartpole_angleonly_ppo2": 0.0,
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("resync", [False, True])
def test_oracle_matches_reference_fixture(name, resync, oracle_lib):
    g = load_golden(name)
    res = replay(g, OracleBackend(name, g["reward"].shape[1]), resync=resync)
    assert res["flag_mismatch"] == 0 and res["done_mismatch"] == 0, res
    assert res["worst"]["time"] == 0.0, res
    for k, v in res["worst"].items():
        assert v <= CASES[name], (k, res)

"""Free-running drift of a backend on UAV fixtures, raw (no chaos cut-off): worst mixed error of the persistent state per
lane over the whole fixture, state re-injected only after the reference's own resets.

    python tests/drift_report.py engine|oracle <fixture | path.npz:family> ...     (engine: the library B200ENV_LIB selects)

Prints one JSON object.  Used by tests/test_drift_gpu.py to compare, on the same fixtures: the reference's own twin drift
(fixture `twin_err`: a second reference instance nudged by one ulp per step), the C oracle (reference operation order,
glibc), the strict engine build (`make strict`) and the shipped engine."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from helpers import EngineBackend, OracleBackend, load_golden, replay  # noqa: E402


def main():
    backend = sys.argv[1]
    out = {}
    for spec in sys.argv[2:]:
        if ":" in spec:
            path, family = spec.rsplit(":", 1)
            with np.load(path) as z:
                g = {k: z[k] for k in z.files}
            label = os.path.basename(path).replace(".npz", "")
        else:
            g, family, label = load_golden(spec), spec, spec
        L = g["reward"].shape[1]
        b = EngineBackend(family, L) if backend == "engine" else OracleBackend(family, L)
        r = replay(g, b, resync=False, name=family, chaos_cut=1e300)
        out[label] = {"state": r["worst"]["state"], "next_obs": r["worst"]["next_obs"], "reward": r["worst"]["reward"],
                      "flag_mismatch": r["flag_mismatch"], "steps": r["steps"], "lanes": int(L),
                      "lane_state": [float(v) for v in r["lane_state"]],
                      "lane_twin": [float(v) for v in np.max(g["twin_err"], axis=0)] if "twin_err" in g else None}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

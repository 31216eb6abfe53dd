"""Regenerates tests/golden/oracle_golden.json from the CPU oracle.

PARITY UNPINNED: the reference ships no golden vectors and PCL is not installable here, so these
vectors pin the ORACLE against silent regressions (and are cross-checked against scipy / numpy
in tests/test_oracle.py); they are not outputs of the reference.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from icpslam_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    O.build()
    O.set_threads(1)
    g = {}
    tgt = synth.integer_cloud(11 + 4096, 4096)
    q = synth.integer_cloud(12 + 4096, 4096, unique=False)
    idx, d2 = O.nn_brute(tgt, q)
    g["nn_integer_4096"] = {"idx_sha256": sha(idx), "d2_sha256": sha(d2), "idx_head": idx[:16].tolist()}
    tgt = synth.integer_cloud(3, 3000, -8, 8)
    q = synth.integer_cloud(4, 20000, -10, 10, unique=False)
    idx, d2 = O.nn_brute(tgt, q)
    g["nn_lattice_ties"] = {"idx_sha256": sha(idx), "d2_sha256": sha(d2)}
    _, _, sw = synth.sweep_sequence(1, 2, n_beams=64, n_az=64)
    r = O.align(O.default_params("odometer"), sw[1], sw[0], record_iter=-1)
    g["c1_p2p"] = {"T": r["T"].tolist(), "iterations": r["iterations"], "n_corr": r["n_corr"],
                   "corr_idx_sha256": sha(r["corr_idx"]), "cloud_sha256": sha(sw[0])}
    r = O.align(O.default_params("odometer", O.MODE_GICP_BFGS), sw[1], sw[0])
    g["c1_gicp"] = {"T": r["T"].tolist(), "iterations": r["iterations"], "n_corr": r["n_corr"]}
    _, _, sc = synth.planar_stream(3, 2)
    r = O.align(O.default_params("odometer"), sc[1], sc[0])
    g["c3_planar_p2p"] = {"T": r["T"].tolist(), "iterations": r["iterations"]}
    v = O.voxel_filter(sw[0], 0.2)
    g["voxel_0p2"] = {"n": int(len(v)), "sha256": sha(v), "first": v[0].tolist()}
    m = O.map_insert(None, sw[0], 0.2)
    a = O.map_insert(m, sw[1], 0.2)
    g["map_insert_0p2"] = {"n_first": int(len(m)), "n_added": int(len(a)), "sha256": sha(np.concatenate([m, a]))}
    C = O.covariances(sw[0][:512])
    g["cov_512"] = {"trace_sum": float(np.trace(C, axis1=1, axis2=2).sum()), "first": C[0].tolist()}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.json"), "w") as f:
        json.dump(g, f, indent=1)
    print("wrote oracle_golden.json")


if __name__ == "__main__":
    main()

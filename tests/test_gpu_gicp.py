"""GPU parity of the GICP mode — the class the reference instantiates (reference
src/icpslam/icp_odometer.cpp:188, src/icpslam/octree_mapper.cpp:104) — against the oracle, stage by stage.
PCL's GICP ends its line search on a round-off test, so end-to-end parity needs every stage to be
bit-identical to the same arithmetic definition; the stage tests below pin exactly that."""
import numpy as np
import pytest

from icpslam_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R(b2lib):
    import torch
    assert torch.cuda.is_available()
    return b2lib


def close(Ta, Tb, tol_t=1e-4, tol_r=1e-4):
    Rr = Ta[:3, :3].T @ Tb[:3, :3]
    ang = np.linalg.norm(0.5 * np.array([Rr[2, 1] - Rr[1, 2], Rr[0, 2] - Rr[2, 0], Rr[1, 0] - Rr[0, 1]]))
    return np.abs(Ta[:3, 3] - Tb[:3, 3]).max() <= tol_t and ang <= tol_r


def test_covariances_bit_identical(R, oracle):
    """K5: k = 20 neighbourhoods in (d2, index) order, double covariance from float products, SVD,
    eigenvalues -> (1, 1, 1e-3): every one of the 9 doubles equal to the oracle's."""
    _, _, sw = synth.sweep_sequence(1, 1, n_beams=64, n_az=256)
    reg = R.Registration(mode=R.MODE_GICP_BFGS)
    C = reg.computeCovariances(sw[0])
    Co = oracle.covariances(sw[0])
    assert np.array_equal(C, Co)
    # planar 2-D scan: rank-2 neighbourhoods (u2 completed by a cross product), far / sparse points
    _, _, sc = synth.planar_stream(3, 1)
    assert np.array_equal(reg.computeCovariances(sc[0]), oracle.covariances(sc[0]))
    rng = np.random.default_rng(0)
    sparse = synth.as_xyzw(np.concatenate([rng.uniform(-3, 3, (300, 3)), rng.uniform(50, 400, (40, 3))]))
    assert np.array_equal(reg.computeCovariances(sparse), oracle.covariances(sparse))
    with pytest.raises(R.B2icpError) as e:
        reg.computeCovariances(sparse[:10])
    assert e.value.code == -3   # TOO_FEW_POINTS


@pytest.mark.parametrize("n_az,preset", [(64, "odometer"), (256, "mapper")])
def test_gicp_align_matches_oracle(R, oracle, n_az, preset):
    _, _, sw = synth.sweep_sequence(1, 3, n_beams=64, n_az=n_az)
    pre = R.PRESET_ODOMETER if preset == "odometer" else R.PRESET_MAPPER
    for i in (1, 2):
        reg = R.Registration(preset=pre, mode=R.MODE_GICP_BFGS)
        reg.setInputSource(sw[i])
        reg.setInputTarget(sw[i - 1])
        aligned = reg.align(want_aligned=True)
        o = oracle.align(oracle.default_params(preset, oracle.MODE_GICP_BFGS), sw[i], sw[i - 1], want_aligned=True,
                         record_iter=-1)
        assert o["rc"] == 0
        assert reg.iterations == o["iterations"] and reg.hasConverged() == bool(o["converged"])
        assert reg.result.n_corr_last == o["n_corr"]
        assert close(reg.getFinalTransformation(), o["T"])
        idx, d2 = reg.getCorrespondences()
        assert np.array_equal(idx, o["corr_idx"])
        assert np.abs(aligned - o["aligned"]).max() <= 2e-5
        fit = reg.getFitnessScore()
        assert abs(fit - oracle.fitness(sw[i], sw[i - 1], reg.getFinalTransformation().astype(np.float32))) <= 1e-9


def test_gicp_kat_and_errors(R, oracle):
    rng = np.random.default_rng(3)
    P = synth.as_xyzw(rng.uniform(-5, 5, (3000, 3)) * [1, 1, 0.2])
    T = synth.random_rigid(rng, 0.05, 0.01)
    Q = oracle.transform_cloud(P, T)
    reg = R.Registration(preset=R.PRESET_MAPPER, mode=R.MODE_GICP_BFGS)
    reg.setInputSource(P)
    reg.setInputTarget(Q)
    reg.align()
    assert np.abs(reg.getFinalTransformation() - T).max() < 2e-5
    o = oracle.align(oracle.default_params("mapper", oracle.MODE_GICP_BFGS), P, Q)
    assert close(reg.getFinalTransformation(), o["T"]) and reg.iterations == o["iterations"]
    reg.setInputSource(P[:10])          # N < k_correspondences: PCL's computeCovariances bails out
    with pytest.raises(R.B2icpError) as e:
        reg.align()
    assert e.value.code == -3
    far = P.copy()
    far[:, :3] += 100
    reg.setInputSource(far)
    with pytest.raises(R.B2icpError) as e:
        reg.align()
    assert e.value.code == -4 and not reg.hasConverged()


def test_gicp_batch_equals_single_calls(R, oracle):
    """b2icp_align_batch in GICP mode (consecutive sweeps: pair i registers against sweep i-1, and an explicit
    target for pair 0) gives exactly what b2icp_align gives pair by pair, fitness included."""
    _, _, sw = synth.sweep_sequence(2, 4, n_beams=64, n_az=128)
    reg = R.Registration(preset=R.PRESET_ODOMETER, mode=R.MODE_GICP_BFGS)
    rc, res = reg.alignBatch(sw[1:], [sw[0], None, None], with_fitness=True)
    assert rc == 0 and len(res) == 3
    for i in range(3):
        one = R.Registration(preset=R.PRESET_ODOMETER, mode=R.MODE_GICP_BFGS)
        one.setInputSource(sw[i + 1])
        one.setInputTarget(sw[i])
        one.align()
        assert np.array_equal(res[i].matrix(), one.getFinalTransformation())
        assert res[i].iterations == one.iterations and res[i].converged == int(one.hasConverged())
        assert abs(res[i].fitness - one.getFitnessScore()) <= 1e-12
    # shared target
    reg.setInputTarget(sw[0])
    rc, res2 = reg.alignBatch([sw[1], sw[1]], None)
    assert rc == 0 and np.array_equal(res2[0].matrix(), res[0].matrix()) and np.array_equal(res2[1].matrix(), res[0].matrix())


def test_concurrent_handles_are_independent(R):
    """Four host threads, each with its own handle, run batches at once (what bench.py's concurrent GICP leg does, and
    what a node with several registration clients would): every result equals the one the same call gives alone —
    GICP (fibers, grouped rounds on private streams, pinned read-backs) and point-to-point alike."""
    import threading
    _, _, sw = synth.sweep_sequence(3, 6, n_beams=64, n_az=128)
    jobs = [(R.MODE_GICP_BFGS, sw[1:4]), (R.MODE_P2P_SVD, sw[2:5]), (R.MODE_GICP_BFGS, sw[3:6]), (R.MODE_P2P_SVD, sw[1:6])]

    def run(mode, clouds):
        reg = R.Registration(preset=R.PRESET_ODOMETER, mode=mode)
        reg.setInputTarget(sw[0])
        rc, res = reg.alignBatch(list(clouds), None, with_fitness=True)
        return rc, [(r.matrix().copy(), r.iterations, r.converged, r.fitness) for r in res]

    alone = [run(m, c) for m, c in jobs]
    for _ in range(2):
        got = [None] * len(jobs)

        def worker(k):
            got[k] = run(*jobs[k])

        ths = [threading.Thread(target=worker, args=(k,)) for k in range(len(jobs))]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        for a, g in zip(alone, got):
            assert g is not None and a[0] == g[0] == 0 and len(a[1]) == len(g[1])
            for (Ta, ia, ca, fa), (Tg, ig, cg, fg) in zip(a[1], g[1]):
                assert np.array_equal(Ta, Tg) and ia == ig and ca == cg and fa == fg

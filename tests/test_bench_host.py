"""Host-side pieces of bench.py that need no GPU: the CPU arm (spawned oracle workers with pinned numerical threads),
the parity bookkeeping over the CPU sample, and the algorithmic flop counts."""
import numpy as np

import bench


def test_cpu_arm_runs_the_oracle_with_pinned_threads():
    batch, probs = bench.sample_problems("cfg2", 1)
    assert batch.B == 6 and len(probs) == 6
    arm = bench.CpuArm(2)
    try:
        wall, values, busy = arm.run(probs[:2])
    finally:
        arm.close()
    assert arm.blas_threads == 1 and wall > 0 and busy > 0
    # member 0 of cfg 2 at 6.925 / 10.65 GHz: the reference values of SURVEY.md 8(d)
    np.testing.assert_allclose(values[0].ravel(), (248.695458986944, 226.167355602153), rtol=1e-9)
    np.testing.assert_allclose(values[1].ravel(), (244.567319165792, 222.495966384025), rtol=1e-9)
    # bookkeeping: GPU problem (f, s) of an S-member run sits at f * S + s
    S = 5
    gpu = np.zeros((6 * S, 2, 1))
    gpu[0 * S + 0] = values[0].reshape(2, 1) * (1 + 1e-12)
    gpu[1 * S + 0] = values[1].reshape(2, 1)
    err, n = bench.max_rel_err(gpu, values + [None] * 4, S, 1, 6)
    assert n == 2 and 0.5e-12 < err < 2e-12


def test_sample_sizes_and_flop_counts():
    assert bench.cpu_sample_size("cfg2", 16, 20.0) == 200  # capped at 200 snowpacks
    assert bench.cpu_sample_size("cfg2", 32, 7.0) >= 43    # >= 8 solves per core
    fe, fb = bench.f_alg_upper("cfg2")
    assert abs(fe + fb - 10.54 * 20 * 128**3) / (fe + fb) < 2e-3  # 10.54 L N^3 = 0.442 GFLOP
    # with every layer keeping all its streams the actual-stream count equals the upper bound
    eps = np.full((3, 20), 1.5 + 1e-4j)
    ea, ba, mean = bench.f_alg_actual("cfg2", eps, np.full(3, 20))
    assert mean == 32 and abs(ea - fe) / fe < 1e-12 and abs(ba - fb) / fb < 1e-12
    # a layer of lower permittivity than the most refringent one keeps fewer streams
    eps[:, 0] = 1.2 + 1e-4j
    ea2, _, mean2 = bench.f_alg_actual("cfg2", eps, np.full(3, 20))
    assert mean2 < 32 and ea2 < ea

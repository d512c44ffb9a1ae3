"""Experiment: kernel times of a cfg-2-like ensemble at a given stream count for several boundary-kernel block sizes
(SMRT_B200_BOUNDARY_THREADS), i.e. for 1 / 2 / ... resident boundary CTAs per SM.  usage: exp_threads.py <streams> <S>"""
import os
import subprocess
import sys

if len(sys.argv) > 3:  # child: one configuration
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, ROOT)
    import bench
    from smrt_b200 import capi

    streams, S = int(sys.argv[1]), int(sys.argv[2])
    batch = bench.make_batch("cfg2", S)
    plan = capi.Plan(capi.make_options(batch, n_max_stream=streams, serialize=True))
    for _ in range(3):
        out = plan.solve_host(batch)
        t = plan.last_timing()
    print(f"threads={os.environ.get('SMRT_B200_BOUNDARY_THREADS')} stream_fg={os.environ.get('SMRT_B200_STREAM_FG')} "
          f"streams={streams} solves={batch.B} eigen {1e3 * t['eigen_ms'] / batch.B:.3f} us boundary "
          f"{1e3 * t['boundary_ms'] / batch.B:.3f} us per solve ({t['chunks']} chunks) Tb0={out.values[0].ravel()[:2]}")
else:
    for thr, stream in (("256", "0"), ("128", "0"), ("128", "1")):
        env = dict(os.environ, SMRT_B200_BOUNDARY_THREADS=thr, SMRT_B200_STREAM_FG=stream)
        subprocess.run([sys.executable, __file__, sys.argv[1], sys.argv[2], "child"], env=env)

"""CPU suite, part 3: host-side logic that needs no device — pulse generators against the REFERENCE's
own spin14.py output, iso helpers, sharding, multi-rank statistics exchange over gloo (world_size 2)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import rbqoc_b200 as rbq
from rbqoc_b200 import dist as rdist
from rbqoc_b200 import spin
from _cases import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pulses_equal_reference_generated_fixture():
    """tests/golden/spin14_pulses.npz was produced by calling the reference's spin14.py generators."""
    z = np.load(os.path.join(GOLDEN, "spin14_pulses.npz"))
    assert float(z["dt_inv"]) == 100.0
    y, x = spin.controls_ypiby2(100.0), spin.controls_xpiby2(100.0)
    assert y.shape == z["ypiby2"].shape == (1947,) and np.array_equal(y, z["ypiby2"])
    assert x.shape == z["xpiby2"].shape == (5680,) and np.array_equal(x, z["xpiby2"])
    assert float(z["ypiby2_evolution_time"]) == spin.TTOT_YPIBY2
    assert float(z["xpiby2_evolution_time"]) == spin.TTOT_XPIBY2
    assert y[0] == 0 and y[-1] == 0 and x[0] == 0 and x[-1] == 0


def test_iso_helpers_and_operators():
    v = np.array([1 + 2j, 3 - 1j])
    assert np.array_equal(spin.get_vec_uniso(spin.get_vec_iso(v)), v)
    g = spin.GATES[spin.GateType.xpiby2]
    G = spin.get_mat_iso(g)
    assert np.allclose(G @ spin.get_vec_iso(v), spin.get_vec_iso(g @ v))
    # spin.jl:234-236
    assert np.allclose(spin.NEGI_H0_ISO / np.pi, [[0, 0, 1, 0], [0, 0, 0, -1], [-1, 0, 0, 0], [0, 1, 0, 0]])
    assert np.allclose(spin.NEGI_H1_ISO / np.pi, [[0, 0, 0, 1], [0, 0, 1, 0], [0, -1, 0, 0], [-1, 0, 0, 0]])
    assert np.allclose(spin.IS3_ISO, spin.get_vec_iso(np.array([1, 1j]) / np.sqrt(2)))
    assert spin.NAMP_PREFACTOR == pytest.approx(2.574e-5, rel=1e-3)


def test_initial_states_and_fixed_seeds():
    assert spin.spin12_x0().shape == (43,) and spin.spin13_x0().shape == (11,)
    x0 = spin.spin30_x0()
    assert x0.shape == (56,)
    assert np.allclose(np.linalg.norm(x0[15:55].reshape(10, 4), axis=1), 1.0)
    assert np.allclose(spin.gen_rand_state_iso(-3), spin.IS3_ISO)
    assert np.allclose(spin.gen_rand_state_iso(-4), spin.IS4_ISO)
    with pytest.raises(ValueError):
        spin.gen_rand_state_iso(-9)
    p = spin.pulse_sin(1001)
    assert p.shape == (1001,) and p[0] == 0 and abs(p[-1]) < 1e-15 and np.max(np.abs(p)) <= 0.5


def test_shard_ranges_tile_the_sweep():
    for S in (0, 1, 7, 1000, 10**6 + 3):
        for world in (1, 2, 3, 8):
            spans = [rdist.shard_range(S, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == S
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_stats_words_roundtrip():
    st = rbq.Stats()
    st.count, st.sum, st.min, st.max = 5, 1.25, 0.1, 0.9
    st.hist[17] = 5
    back = rbq.stats_from_words(rdist.stats_to_words(st))
    assert (back.count, back.sum, back.min, back.max, back.hist[17]) == (5, 1.25, 0.1, 0.9, 5)


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import numpy as np, torch, torch.distributed as dist
    import rbqoc_b200 as rbq
    from rbqoc_b200 import dist as rdist
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    S = 1001
    first, count = rdist.shard_range(S, rank, world)
    e = np.random.default_rng(3).random(S)          # the whole sweep's infidelities, same on every rank
    st = rbq.Stats()
    mine = e[first:first + count]
    st.count, st.sum, st.sumsq, st.min, st.max = count, float(mine.sum()), float((mine**2).sum()), float(mine.min()), float(mine.max())
    for v in mine:
        st.hist[min(255, max(0, int(np.floor((np.log10(v) + 16) * 16))))] += 1
    total = rdist.allgather_merge_stats(torch.from_numpy(rdist.stats_to_words(st)))
    assert total.count == S and sum(total.hist[:]) == S
    assert abs(total.sum - e.sum()) < 1e-12 * e.sum()
    assert total.min == e.min() and total.max == e.max()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def test_two_rank_statistics_exchange_over_gloo(tmp_path):
    """The N>1 path: shard by contiguous index range, ONE all-gather of the 2.1 KB statistics, host merge."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2

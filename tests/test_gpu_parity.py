"""GPU parity tests (run on the B200 box with `-m gpu`).  Everything goes through the C ABI
(libtip.so) via the drop-in Model / EMEngine and is compared with
  * the golden vectors produced by the unmodified reference (tests/golden), and
  * the CPU oracle on the same seeded inputs,
then, at BASELINE.json's full sizes, through size-independent properties.
Tolerances: ids / folds / packing bit-exact; theta, p, log-likelihood 1e-9 relative (fp64 mode);
AUC 1e-6 (north_star)."""
import ctypes
import json
import os
import random

import numpy as np
import pytest

from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu

BASE = os.path.join(GOLDEN, "base")
DUPS = os.path.join(GOLDEN, "dups")
RTOL = 1e-9


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _model(case, train, test, **kw):
    from trigenicinteractionpredictor_b200 import Model
    m = Model(**kw)
    m.get_traintest(os.path.join(case, train), os.path.join(case, test))
    return m


def _relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


@pytest.mark.parametrize("flags", [0, 1], ids=["specialised", "anyK"])
@pytest.mark.parametrize("K", [1, 2, 3, 10])
def test_em_trace_matches_reference(torch_cuda, K, flags):
    tr = np.load(os.path.join(BASE, "trace_K%d.npz" % K))
    m = _model(BASE, "train1.dat", "test1.dat", flags=flags)
    random.seed(1000)
    m.initialize_parameters(K)
    assert np.array_equal(np.array(m.theta), tr["theta0"])
    assert m.compute_likelihood() == pytest.approx(tr["loglik"][0], rel=RTOL)
    for it in range(5):
        m.make_iteration()
        assert _relerr(m.theta, tr["theta%d" % (it + 1)]) < RTOL, "theta iteration %d" % (it + 1)
        assert _relerr(m.pr, tr["pr%d" % (it + 1)]) < RTOL, "p iteration %d" % (it + 1)
        assert m.compute_likelihood() == pytest.approx(tr["loglik"][it + 1], rel=RTOL)
        assert m.likelihood == pytest.approx(tr["loglik"][it + 1], rel=RTOL)
    assert m.compute_likelihood("test") == pytest.approx(tr["heldout"][0], rel=RTOL)
    # scoring, table order and metrics
    m.calculate_test_set_results()
    sc = m._scores.cpu().numpy()
    assert _relerr(sc, tr["scores_test_order"]) < RTOL
    res = m.results
    assert len(res) == len(tr["result_keys"])
    assert _relerr([r[0] for r in res], tr["result_scores"]) < RTOL
    met = m.calculate_metrics()
    if K > 1:
        # the table is the reference's row for row; a row may only sit elsewhere if the reference's own scores of the two
        # neighbours are within 1e-9 of each other (the order of such a pair is decided below the tolerance of the scores)
        gs, gk = tr["result_scores"], tr["result_keys"].tolist()
        near_tie = False
        for i, (r, k) in enumerate(zip(res, gk)):
            if r[1] != k:
                gap = min(abs(gs[i] - gs[j]) for j in (i - 1, i + 1) if 0 <= j < len(gs))
                assert gap <= 1e-9 * abs(gs[i]), "row %d of the sorted test table differs from the reference's" % i
                near_tie = True
        assert met[3] == pytest.approx(tr["metrics"][3], abs=1e-6)          # AUC
        if not near_tie:
            # no near-tie anywhere: the rank cut falls on the same row, precision / recall / fallout are the same integers
            np.testing.assert_allclose(met[:3], tr["metrics"][:3], rtol=1e-12, atol=0)
            assert [r[2] for r in res] == [int(x) for x in tr["result_labels"]]
        else:
            np.testing.assert_allclose(met[:3], tr["metrics"][:3], atol=1e-2)   # the rank cut may move by one row on a tie
    else:
        # K=1 is degenerate: theta == 1 - O(eps), every score equals p[0][0][0][1] to ~1e-10, so the order of
        # the table, the rank cut and the AUC are decided at rounding level (and by the order in which the
        # atomics land); the 1e-6 AUC gate applies to K >= 2
        assert met[3] == pytest.approx(tr["metrics"][3], abs=5e-3)
    # single-triplet prediction by id strings and by gene names (TIP.py:541-545)
    key = next(iter(m.test_links)).split("_")
    names = [m.id_gene[int(t)] for t in key]
    assert m.do_prediction(*key) == pytest.approx(tr["predict_by_name"][1], rel=RTOL)
    assert m.do_prediction(*names) == pytest.approx(tr["predict_by_name"][0], rel=RTOL)


@pytest.mark.parametrize("flags", [0, 1], ids=["specialised", "anyK"])
def test_duplicates_conflicts_and_string_sorted_slots(torch_cuda, flags):
    tr = np.load(os.path.join(DUPS, "trace_K3.npz"))
    m = _model(DUPS, "train.dat", "test.dat", flags=flags)
    random.seed(1001)
    m.initialize_parameters(3)
    for it in range(3):
        m.make_iteration()
        assert _relerr(m.theta, tr["theta%d" % (it + 1)]) < RTOL
        assert _relerr(m.pr, tr["pr%d" % (it + 1)]) < RTOL
        assert m.compute_likelihood() == pytest.approx(tr["loglik"][it + 1], rel=RTOL)
    assert m.compute_likelihood("test") == pytest.approx(tr["heldout"][0], rel=RTOL)
    m.calculate_test_set_results()
    assert _relerr(m._scores.cpu().numpy(), tr["scores_test_order"]) < RTOL
    assert m.calculate_metrics()[3] == pytest.approx(tr["metrics"][3], abs=1e-6)


@pytest.mark.parametrize("K", [2, 3, 10])
def test_fp32_compute_mode_within_1e5(torch_cuda, K):
    """TIP_EM_FP32_COMPUTE: fp32 contractions, fp64 accumulation - theta, p, log-likelihood per iteration within
    1e-5 relative of the reference (north_star tolerance for this mode), AUC within 1e-6."""
    tr = np.load(os.path.join(BASE, "trace_K%d.npz" % K))
    m = _model(BASE, "train1.dat", "test1.dat", flags=2)
    random.seed(1000)
    m.initialize_parameters(K)
    for it in range(5):
        m.make_iteration()
        assert _relerr(m.theta, tr["theta%d" % (it + 1)]) < 1e-5, "theta iteration %d" % (it + 1)
        assert _relerr(m.pr, tr["pr%d" % (it + 1)]) < 1e-5, "p iteration %d" % (it + 1)
        assert m.compute_likelihood() == pytest.approx(tr["loglik"][it + 1], rel=1e-5)
    m.calculate_test_set_results()
    assert _relerr(m._scores.cpu().numpy(), tr["scores_test_order"]) < 1e-5
    assert m.calculate_metrics()[3] == pytest.approx(tr["metrics"][3], abs=1e-6)


@pytest.mark.parametrize("K", [5, 8, 9, 10, 13, 16])
def test_gene_segmented_mode_matches_oracle(torch_cuda, K):
    """TIP_EM_GENE_SEGMENTED (flag 8): p contracted with theta per gene first - same statistics to rounding."""
    from oracle import mmsbm_oracle as orc
    from trigenicinteractionpredictor_b200.engine import EMEngine
    P, L = 300, 6000 if K <= 10 else 1500
    g, n0, n1, theta, pr = _random_problem(P, L, K, 40 + K)
    cnt = np.stack([n0, n1], axis=1).astype(np.int64)
    ent, enp, deg = orc.em_step_np(theta, pr, g.astype(np.int64), cnt, return_stats=True)
    th1, pr1 = orc.normalise_np(ent, enp, deg)
    eng = EMEngine(P, K, flags=8)
    eng.set_train_links(g[:, 0], g[:, 1], g[:, 2], n0, n1)
    eng.set_params(theta, pr)
    eng.em_iteration()
    th, p = eng.get_params()
    assert _relerr(th, th1) < 1e-11 and _relerr(p, pr1) < 1e-11


def test_gene_segmented_mode_reference_trace(torch_cuda):
    tr = np.load(os.path.join(BASE, "trace_K10.npz"))
    m = _model(BASE, "train1.dat", "test1.dat", flags=8)
    random.seed(1000)
    m.initialize_parameters(10)
    for it in range(5):
        m.make_iteration()
        assert _relerr(m.theta, tr["theta%d" % (it + 1)]) < RTOL
        assert _relerr(m.pr, tr["pr%d" % (it + 1)]) < RTOL
        assert m.compute_likelihood() == pytest.approx(tr["loglik"][it + 1], rel=RTOL)


KUZMIN = os.path.join(GOLDEN, "kuzmin")


@pytest.mark.parametrize("flags", [None, 32, 0, 8, 1], ids=["default", "slots", "k3", "gene_seg", "anyK"])
@pytest.mark.parametrize("K,iters", [(3, 5), (10, 3)])
def test_em_trace_matches_reference_on_hub_shaped_links(torch_cuda, K, iters, flags):
    """Kuzmin shape: 20 query genes carry every link (hubs in all three key slots).  theta, p, log-likelihood per
    iteration within 1e-9 of the unmodified reference, scores 1e-9, AUC 1e-6 - for every E-step formulation."""
    tr = np.load(os.path.join(KUZMIN, "trace_K%d.npz" % K))
    m = _model(KUZMIN, "train1.dat", "test1.dat", flags=flags)
    random.seed(1000)
    m.initialize_parameters(K)
    assert np.array_equal(np.array(m.theta), tr["theta0"])
    assert m.compute_likelihood() == pytest.approx(tr["loglik"][0], rel=RTOL)
    for it in range(iters):
        m.make_iteration()
        assert _relerr(m.theta, tr["theta%d" % (it + 1)]) < RTOL, "theta iteration %d" % (it + 1)
        assert _relerr(m.pr, tr["pr%d" % (it + 1)]) < RTOL, "p iteration %d" % (it + 1)
        assert m.compute_likelihood() == pytest.approx(tr["loglik"][it + 1], rel=RTOL)
    assert m.compute_likelihood("test") == pytest.approx(tr["heldout"][0], rel=RTOL)
    m.calculate_test_set_results()
    assert _relerr(m._scores.cpu().numpy(), tr["scores_test_order"]) < RTOL
    assert m.calculate_metrics()[3] == pytest.approx(tr["metrics"][3], abs=1e-6)


@pytest.mark.parametrize("K", [1, 2, 3, 10])
def test_slot_segmented_mode_reference_trace(torch_cuda, K):
    """TIP_EM_SLOT_SEGMENTED (flag 32) on the reference trace of the base case, every K the golden set has."""
    tr = np.load(os.path.join(BASE, "trace_K%d.npz" % K))
    m = _model(BASE, "train1.dat", "test1.dat", flags=32)
    random.seed(1000)
    m.initialize_parameters(K)
    for it in range(5):
        m.make_iteration()
        assert _relerr(m.theta, tr["theta%d" % (it + 1)]) < RTOL, "theta iteration %d" % (it + 1)
        assert _relerr(m.pr, tr["pr%d" % (it + 1)]) < RTOL, "p iteration %d" % (it + 1)
        assert m.compute_likelihood() == pytest.approx(tr["loglik"][it + 1], rel=RTOL)


def test_slot_segmented_duplicates_and_conflicts(torch_cuda):
    tr = np.load(os.path.join(DUPS, "trace_K3.npz"))
    m = _model(DUPS, "train.dat", "test.dat", flags=32)
    random.seed(1001)
    m.initialize_parameters(3)
    for it in range(3):
        m.make_iteration()
        assert _relerr(m.theta, tr["theta%d" % (it + 1)]) < RTOL
        assert _relerr(m.pr, tr["pr%d" % (it + 1)]) < RTOL


@pytest.mark.parametrize("K", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 16, 17, 20, 24, 25, 32])
@pytest.mark.parametrize("shape", ["uniform", "hubs"])
def test_slot_segmented_mode_matches_oracle(torch_cuda, K, shape):
    """Statistics of the slot-segmented E-step against the NumPy oracle (1e-11), uniform links and links whose three
    slots are all dominated by a handful of hub genes; every accumulator-block count (K <= 8, 16, 24, 32)."""
    from oracle import mmsbm_oracle as orc
    from trigenicinteractionpredictor_b200.engine import EMEngine
    from trigenicinteractionpredictor_b200 import synth
    P, L = 300, 6000 if K <= 10 else (2000 if K <= 16 else 700)
    g, n0, n1, theta, pr = _random_problem(P, L, K, 500 + K)
    if shape == "hubs":
        a, b, c, lab = synth.kuzmin_links_soa(P, L, seed=K, n_query=6)
        g = np.stack([a, b, c], axis=1).astype(np.int32)
        g[:P, 0] = np.arange(P)
        n0, n1 = (1 - lab).astype(np.int32), lab.astype(np.int32)
        n0[5:40:7] += 2                             # repeated sightings: counts above one
    cnt = np.stack([n0, n1], axis=1).astype(np.int64)
    ent, enp, deg = orc.em_step_np(theta, pr, g.astype(np.int64), cnt, return_stats=True)
    th1, pr1 = orc.normalise_np(ent, enp, deg)
    eng = EMEngine(P, K, flags=32)
    eng.set_train_links(g[:, 0], g[:, 1], g[:, 2], n0, n1)
    eng.set_params(theta, pr)
    eng.em_step()
    st = eng.stats.cpu().numpy()
    nth = st[: P * K].reshape(P, K)
    S = st[P * K: P * K + 2 * K ** 3].reshape(2, K, K, K)
    assert _relerr(nth, ent) < 1e-11, "Ntheta K=%d" % K
    assert _relerr(pr * np.moveaxis(S, 0, -1), np.maximum(enp, 1e-300)) < 1e-11, "Np K=%d" % K
    eng.normalise()
    th, p = eng.get_params()
    assert _relerr(th, th1) < 1e-11 and _relerr(p, pr1) < 1e-11
    eng.set_params(theta, pr)
    eng.em_iterations(5)                            # CUDA-graph replay of the four-kernel E-step
    tho, pro = theta, pr
    for _ in range(5):
        tho, pro = orc.em_step_np(tho, pro, g.astype(np.int64), cnt)
    th, p = eng.get_params()
    assert _relerr(th, tho) < 1e-10 and _relerr(p, pro) < 1e-10


def test_order_rows_bit_exact(torch_cuda):
    """tip_order_rows against NumPy: the slot-b / slot-c orders are stable sorts of the packed rows by (rating, gene),
    rating blocks keep their padded sizes, every row carries the position of its link in the slot-a order."""
    from trigenicinteractionpredictor_b200.engine import EMEngine
    rng = np.random.default_rng(9)
    P, L = 200, 5003
    g = rng.integers(0, P, size=(L, 3)).astype(np.int32)
    n0 = rng.integers(0, 2, size=L).astype(np.int32)
    n1 = (1 - n0) * rng.integers(0, 2, size=L).astype(np.int32)
    eng = EMEngine(P, 5, flags=32)
    eng.set_train_links(g[:, 0], g[:, 1], g[:, 2], n0, n1)
    t = eng.train
    rows3 = t.rows3.cpu().numpy()
    n, n_r0 = t.n_rows, t.n_rows_r0
    ra = rows3[:n]
    assert np.array_equal(ra, t.rows.cpu().numpy())
    for slot, blk in ((1, rows3[n: 2 * n]), (2, rows3[2 * n:])):
        exp = np.zeros_like(ra)
        exp[:, 3] = -1
        for lo, hi in ((0, n_r0), (n_r0, n)):
            idx = np.arange(lo, hi)
            live = idx[(ra[lo:hi, 3] >> 1) > 0]
            order = live[np.argsort(ra[live, slot], kind="stable")]
            other = 2 if slot == 1 else 1
            exp[lo: lo + len(order)] = np.stack([ra[order, slot], ra[order, 0], ra[order, other], order], axis=1)
        assert np.array_equal(blk, exp)
    # the tile schedules behind the orders: every tile of a launch exactly once, chunks of <= 4 consecutive tiles,
    # costs (1 + run ends per tile) never increase along the list except where single tiles close it
    T, T0 = n // 32, n_r0 // 32
    sched = t.rows3_buf.cpu().numpy()[12 * n: 12 * n + 3 * T]
    for name, ent, nt, rows_l in (("a", sched[:T], T, rows3[:n]), ("bc", sched[T:], 2 * T, rows3[n:])):
        tt = np.arange(nt * 32) // 32
        so = (tt >= T).astype(np.int64)
        key = ((so * 2 + ((tt - so * T) >= T0)) << 32) | rows_l[:, 0].astype(np.int64)
        ends = np.zeros(nt * 32, dtype=np.int64)
        ends[1:] = key[1:] != key[:-1]
        cost = 1 + ends.reshape(nt, 32).sum(axis=1)
        live = ent[ent != 0]
        assert np.all(ent[len(live):] == 0), name
        seen = np.zeros(nt, dtype=np.int64)
        flagged = [int(e) < 0 for e in live.tolist()]           # bit 31: end-game entries, the last ones of the list
        assert flagged == sorted(flagged) and sum(flagged) <= 148 * 16
        for e in live.tolist():
            e &= 0x7FFFFFFF
            t0, cnt = e >> 3, e & 7
            assert 1 <= cnt <= 4 and t0 + cnt <= nt
            seen[t0: t0 + cnt] += 1
        assert np.all(seen == 1), "schedule %s does not cover every tile exactly once" % name
        first = int(live[0]) & 0x7FFFFFFF
        assert cost[first >> 3: (first >> 3) + (first & 7)].max() == max(cost.max(), 0) or cost.max() < 7


def test_fp32_mode_is_rejected_where_it_does_not_exist(torch_cuda):
    from trigenicinteractionpredictor_b200._cabi import TipLibraryError
    from trigenicinteractionpredictor_b200.engine import EMEngine
    g, n0, n1, theta, pr = _random_problem(100, 640, 12, 1)
    eng = EMEngine(100, 12, flags=2)
    eng.set_train_links(g[:, 0], g[:, 1], g[:, 2], n0, n1)
    eng.set_params(theta, pr)
    with pytest.raises(TipLibraryError):
        eng.em_step()


def test_gene_seen_only_in_test_raises_like_reference(torch_cuda):
    m = _model(os.path.join(GOLDEN, "testonly"), "train.dat", "test.dat")
    random.seed(5)
    m.initialize_parameters(2)
    with pytest.raises(ZeroDivisionError):
        m.make_iteration()


def test_pack_rows_bit_exact(torch_cuda):
    from trigenicinteractionpredictor_b200.engine import EMEngine
    rng = np.random.default_rng(3)
    P, L = 500, 7001
    g = rng.integers(0, P, size=(L, 3)).astype(np.int32)
    n0 = rng.integers(0, 3, size=L).astype(np.int32)
    n1 = rng.integers(0, 3, size=L).astype(np.int32)
    eng = EMEngine(P, 2)
    pk = eng.pack(g[:, 0], g[:, 1], g[:, 2], n0, n1)
    rows = pk.rows.cpu().numpy()
    exp = []
    for r, n in ((0, n0), (1, n1)):
        sel = np.nonzero(n > 0)[0]
        order = np.lexsort((g[sel, 2], g[sel, 1], g[sel, 0]))
        blk = np.stack([g[sel, 0], g[sel, 1], g[sel, 2], (n[sel] << 1) | r], axis=1)[order]
        pad = (-len(blk)) % 32
        blk = np.concatenate([blk, np.tile(np.array([[0, 0, 0, r]], dtype=np.int32), (pad, 1))])
        exp.append(blk)
    assert pk.n_rows_r0 == len(exp[0]) and pk.n_rows == len(exp[0]) + len(exp[1])
    assert pk.n_real == int((n0 > 0).sum() + (n1 > 0).sum())
    # rows with identical (a,b,c) may be permuted among themselves; compare as sorted tuples per block
    for blk, lo in ((exp[0], 0), (exp[1], pk.n_rows_r0)):
        got = rows[lo: lo + len(blk)]
        assert np.array_equal(got[:, :3], blk[:, :3])
        assert sorted(map(tuple, got.tolist())) == sorted(map(tuple, blk.tolist()))
    live = (n0 > 0) | (n1 > 0)
    deg = np.bincount(g[live].ravel(), minlength=P)
    assert np.array_equal(pk.deg.cpu().numpy(), deg)


def _random_problem(P, L, K, seed):
    rng = np.random.default_rng(seed)
    g = rng.integers(0, P, size=(L, 3)).astype(np.int32)
    g[:P, 0] = np.arange(P)                       # every gene has a training link
    lab = (rng.random(L) < 0.15).astype(np.int32)
    theta = rng.dirichlet(np.ones(K), size=P)
    pr = rng.random((K, K, K, 2))
    pr /= pr.sum(axis=3, keepdims=True)
    return g, 1 - lab, lab, theta, pr


@pytest.mark.parametrize("K", [4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 20, 23, 24, 32])
def test_em_step_vs_oracle_all_k(torch_cuda, K):
    from oracle import mmsbm_oracle as orc
    from trigenicinteractionpredictor_b200.engine import EMEngine
    P, L = 300, 6000 if K <= 10 else (1500 if K <= 16 else 400)
    g, n0, n1, theta, pr = _random_problem(P, L, K, 10 + K)
    cnt = np.stack([n0, n1], axis=1).astype(np.int64)
    ent, enp, deg = orc.em_step_np(theta, pr, g.astype(np.int64), cnt, return_stats=True)
    ll = orc.loglik_np(theta, pr, g.astype(np.int64), cnt)
    th1, pr1 = orc.normalise_np(ent, enp, deg)
    # 4 = TIP_EM_WITH_LOGLIK, 1 = TIP_EM_FORCE_GENERIC; K = 11..16 have a specialised kernel without the by-product
    for flags in ([4, 1, 0] if K <= 10 else [0, 1]):
        eng = EMEngine(P, K, flags=flags)
        eng.set_train_links(g[:, 0], g[:, 1], g[:, 2], n0, n1)
        eng.set_params(theta, pr)
        eng.em_step()
        st = eng.stats.cpu().numpy()
        nth = st[: P * K].reshape(P, K)
        S = st[P * K: P * K + 2 * K ** 3].reshape(2, K, K, K)
        npr = pr * np.moveaxis(S, 0, -1)
        assert _relerr(nth, ent) < 1e-11, "Ntheta K=%d flags=%d" % (K, flags)
        assert _relerr(npr, np.maximum(enp, 1e-300)) < 1e-11, "Np K=%d flags=%d" % (K, flags)
        if flags != 0:
            assert st[-1] == pytest.approx(ll, rel=1e-12)      # by-product of the E-step
        else:
            assert st[-1] == 0.0                                # not requested on the specialised path
        assert eng.loglik("train") == pytest.approx(ll, rel=1e-12)
        eng.normalise()
        th, p = eng.get_params()
        assert _relerr(th, th1) < 1e-11 and _relerr(p, pr1) < 1e-11
        assert np.array_equal(eng.degrees(), deg.astype(np.int32))


def test_graph_replay_and_host_entry_equal_stepwise(torch_cuda):
    from trigenicinteractionpredictor_b200 import _cabi
    from trigenicinteractionpredictor_b200.engine import EMEngine
    P, L, K = 400, 9000, 10
    g, n0, n1, theta, pr = _random_problem(P, L, K, 77)
    a = EMEngine(P, K)
    a.set_train_links(g[:, 0], g[:, 1], g[:, 2], n0, n1)
    a.set_params(theta, pr)
    for _ in range(6):
        a.em_iteration()
    th_a, p_a = a.get_params()
    b = EMEngine(P, K)
    b.set_train_links(g[:, 0], g[:, 1], g[:, 2], n0, n1)
    b.set_params(theta, pr)
    b.em_iterations(6, use_graph=True)
    th_b, p_b = b.get_params()
    # atomics make the summation order run-dependent: agreement is to rounding, not bitwise
    assert _relerr(th_b, th_a) < 1e-11 and _relerr(p_b, p_a) < 1e-11
    # host-buffer entry point
    lib = _cabi.load()
    rows = np.ascontiguousarray(a.train.rows.cpu().numpy())
    deg = np.ascontiguousarray(a.degrees().astype(np.int32))
    th_h, p_h = np.ascontiguousarray(theta.copy()), np.ascontiguousarray(pr.copy())
    rc = lib.tip_em_iterations_host(P, K, rows.ctypes.data, a.train.n_rows, a.train.n_rows_r0, deg.ctypes.data,
                                    th_h.ctypes.data, p_h.ctypes.data, 6, 0)
    assert rc == 0, lib.tip_last_error()
    assert _relerr(th_h, th_a) < 1e-11 and _relerr(p_h, p_a) < 1e-11
    # the same entry with 8-byte host rows (TIP_ROWS_COMPACT8), and the device-side expansion on its own
    rows8 = np.zeros(a.train.n_rows, dtype=np.uint64)
    assert lib.tip_rows_compact_host(rows.ctypes.data, a.train.n_rows, rows8.ctypes.data) == 0
    th_c, p_c = np.ascontiguousarray(theta.copy()), np.ascontiguousarray(pr.copy())
    rc = lib.tip_em_iterations_host(P, K, rows8.ctypes.data, a.train.n_rows, a.train.n_rows_r0, deg.ctypes.data,
                                    th_c.ctypes.data, p_c.ctypes.data, 6, _cabi.TIP_ROWS_COMPACT8)
    assert rc == 0, lib.tip_last_error()
    assert _relerr(th_c, th_a) < 1e-11 and _relerr(p_c, p_a) < 1e-11
    torch = torch_cuda
    d8 = torch.from_numpy(rows8.view(np.int64)).to(a.device)
    d16 = torch.zeros_like(a.train.rows)
    assert lib.tip_rows_expand(d8.data_ptr(), d16.data_ptr(), a.train.n_rows, 0) == 0
    torch.cuda.synchronize()
    assert torch.equal(d16, a.train.rows)                          # bit-exact round trip


def test_em_step_host_rows_equals_em_step(torch_cuda):
    """The streamed E-step (rows polled as the host-to-device copy lands) gives the statistics of tip_em_step, for both
    host row formats, also when calls follow each other without a synchronisation in between."""
    torch = torch_cuda
    from trigenicinteractionpredictor_b200 import _cabi
    from trigenicinteractionpredictor_b200.engine import EMEngine
    lib = _cabi.load()
    for K, L in ((10, 50000), (4, 20000), (7, 20000)):
        P = 500
        g, n0, n1, theta, pr = _random_problem(P, L, K, 300 + K)
        eng = EMEngine(P, K)
        eng.set_train_links(g[:, 0], g[:, 1], g[:, 2], n0, n1)
        eng.set_params(theta, pr)
        eng.em_step()
        ref = eng.stats.clone()
        rows_h = eng.train.rows.cpu().pin_memory()
        rows8_h = torch.empty(eng.train.n_rows, dtype=torch.int64).pin_memory()
        assert lib.tip_rows_compact_host(rows_h.data_ptr(), eng.train.n_rows, rows8_h.data_ptr()) == 0
        for compact, src in ((False, rows_h), (True, rows8_h)):
            outs = [torch.zeros_like(ref) for _ in range(3)]
            for o in outs:                                       # back to back: the copy stream must wait for the reader
                eng.em_step_host_rows(src, compact, o)
            torch.cuda.synchronize()
            assert eng.host_rows_arrived()
            for o in outs:
                assert _relerr(o.cpu().numpy(), ref.cpu().numpy()) < 1e-11     # (0 against 0 counts as equal)
        # whole iterations through the host-rows path equal the resident path
        a = EMEngine(P, K)
        a.set_train_links(g[:, 0], g[:, 1], g[:, 2], n0, n1)
        a.set_params(theta, pr)
        for _ in range(3):
            eng.em_iteration_host_rows(rows8_h, True)
            a.em_iteration()
        th_e, p_e = eng.get_params()
        th_a, p_a = a.get_params()
        assert _relerr(th_e, th_a) < 1e-11 and _relerr(p_e, p_a) < 1e-11
    big = EMEngine(300, 12)
    g, n0, n1, theta, pr = _random_problem(300, 2000, 12, 5)
    big.set_train_links(g[:, 0], g[:, 1], g[:, 2], n0, n1)
    big.set_params(theta, pr)
    with pytest.raises(_cabi.TipLibraryError):                   # no streamed kernel for K > 10: refuses, no fallback
        big.em_step_host_rows(big.train.rows.cpu().pin_memory(), False)


def test_metrics_kernel_vs_reference_loops(torch_cuda):
    torch = torch_cuda
    from oracle import mmsbm_oracle as orc
    from trigenicinteractionpredictor_b200.engine import EMEngine
    rng = np.random.default_rng(5)
    T = 3000
    scores = np.round(rng.random(T), 2)            # many exact ties
    n0 = (rng.random(T) < 0.8).astype(np.int32)
    g = rng.integers(0, 50, size=(T, 3)).astype(np.int32)
    eng = EMEngine(50, 2)
    eng.set_test_links(g[:, 0], g[:, 1], g[:, 2], n0, 1 - n0)
    test_links = {"%d" % i: [int(n0[i]), int(1 - n0[i])] for i in range(T)}
    res = orc.test_results(scores, test_links)
    train_links = {"a": [0, 1], "b": [1, 0], "c": [1, 0], "d": [1, 0], "e": [0, 1]}
    exp = orc.metrics_quadratic(res, train_links, T)
    npos_rank = int(2 / 5 * T)
    c = eng.metric_counts(torch.from_numpy(scores).to(eng.device), npos_rank)
    got = [c["tp"] / (c["tp"] + c["fp"]), c["tp"] / (c["tp"] + c["fn"]), c["fp"] / (c["fp"] + c["tn"]),
           c["wins"] / (c["n_pos"] * c["n_neg"])]
    assert got == exp                               # integers -> identical floats
    assert c["cut_value"] == res[npos_rank][0]
    big = eng.metric_counts(torch.from_numpy(scores).to(eng.device), T + 5)   # index past the end: cut stays 0
    assert big["cut_value"] == 0.0 and big["fn"] == 0 and big["tn"] == 0


def test_training_loop_and_report_match_reference(torch_cuda, tmp_path):
    from trigenicinteractionpredictor_b200 import train_sample
    man = json.load(open(os.path.join(GOLDEN, "manifest.json")))["base"]["loop"]
    m = _model(BASE, "train1.dat", "test1.dat")
    random.seed(1000)
    out = str(tmp_path / "Sample_0_K2.csv")
    conv, done, checks = train_sample(m, man["K"], man["iterations"], man["fcheck"], man["bcheck"], outfile=out,
                                      verbose=False)
    assert conv and done - 1 == man["converged_at_iteration"]
    np.testing.assert_allclose(checks, man["checks"], rtol=RTOL)
    got = open(out, encoding="utf-8").read().split("\n")
    exp = open(os.path.join(BASE, "Sample_0_K2.csv"), encoding="utf-8").read().split("\n")
    assert len(got) == len(exp)
    swaps = 0
    for lg, le in zip(got, exp):
        fg, fe = lg.split("\t"), le.split("\t")
        assert len(fg) == len(fe)
        for xg, xe in zip(fg, fe):
            if xg == xe:
                continue
            try:
                assert float(xg) == pytest.approx(float(xe), rel=1e-8, abs=1e-12), (lg, le)
            except ValueError:
                swaps += 1                          # a key column differing: near-tie swap in the table
    assert swaps <= 4


def test_full_size_properties_cfg2(torch_cuda):
    """BASELINE config 2 shape (6,000 genes, 800k training links, K=10): invariants of one EM step."""
    torch = torch_cuda
    from trigenicinteractionpredictor_b200 import synth
    from trigenicinteractionpredictor_b200.engine import EMEngine
    P, L, K = 6000, 800_000, 10
    g1, g2, g3, lab = synth.planted_links_soa(P, L, seed=11, device="cuda")
    g1[:P] = torch.arange(P, dtype=torch.int32, device="cuda")
    rng = np.random.default_rng(0)
    theta = rng.dirichlet(np.ones(K), size=P)
    pr = rng.random((K, K, K, 2))
    pr /= pr.sum(axis=3, keepdims=True)
    # the chunked NumPy oracle on the same 800,000 links (a few seconds): statistics and parameters to 1e-10
    from oracle import mmsbm_oracle as orc
    ids_h = torch.stack([g1, g2, g3], dim=1).cpu().numpy().astype(np.int64)
    lab_h = lab.cpu().numpy().astype(np.int64)
    th_o, p_o = orc.em_step_np(theta, pr, ids_h, np.stack([1 - lab_h, lab_h], axis=1))
    outs = []
    for flags in (4, 1, 32, 8):                        # K^3 kernel (+loglik by-product), any-K, slot-segmented, gene-segmented
        eng = EMEngine(P, K, flags=flags)
        eng.set_train_links(g1, g2, g3, 1 - lab, lab)
        assert eng.train.n_real == L
        eng.set_params(theta, pr)
        ll0 = eng.loglik("train")
        eng.em_step()
        st = eng.stats.cpu().numpy()
        if flags in (4, 1):
            assert st[-1] == pytest.approx(ll0, rel=1e-12)      # by-product == dedicated reduction
        S = st[P * K: P * K + 2 * K ** 3].reshape(2, K, K, K)
        npr = pr * np.moveaxis(S, 0, -1)
        # every link distributes (d-eps)/d ~ 1 unit of responsibility: over cells and over each slot
        assert npr.sum() == pytest.approx(L, rel=1e-6)
        assert st[: P * K].sum() == pytest.approx(3 * L, rel=1e-6)
        eng.normalise()
        th, p = eng.get_params()
        np.testing.assert_allclose(th.sum(axis=1), 1.0, atol=1e-6)      # trap 2: count-1 data keeps rows on the simplex
        np.testing.assert_allclose(p.sum(axis=3), 1.0, atol=1e-6)
        assert eng.loglik("train") > ll0                                 # EM ascent
        assert _relerr(th, th_o) < 1e-10 and _relerr(p, p_o) < 1e-10, "flags %d against the oracle at cfg2 size" % flags
        outs.append((th, p))
    assert _relerr(outs[0][0], outs[1][0]) < 1e-10 and _relerr(outs[0][1], outs[1][1]) < 1e-10


def test_full_size_hub_shaped_cfg2_against_oracle(torch_cuda):
    """BASELINE config 2 in the Kuzmin shape (77 query genes of degree ~20,000 among 6,000, 800k links, K=10): the default
    E-step (slot-segmented, chunk schedule) against the chunked NumPy oracle, and its schedule covers every tile."""
    torch = torch_cuda
    from oracle import mmsbm_oracle as orc
    from trigenicinteractionpredictor_b200 import synth
    from trigenicinteractionpredictor_b200.engine import EMEngine
    P, L, K = 6000, 800_000, 10
    g1, g2, g3, lab = synth.kuzmin_links_soa(P, L, seed=12, device="cuda")
    g1[:P] = torch.arange(P, dtype=torch.int32, device="cuda")
    rng = np.random.default_rng(1)
    theta = rng.dirichlet(np.ones(K), size=P)
    pr = rng.random((K, K, K, 2))
    pr /= pr.sum(axis=3, keepdims=True)
    ids_h = torch.stack([g1, g2, g3], dim=1).cpu().numpy().astype(np.int64)
    lab_h = lab.cpu().numpy().astype(np.int64)
    cnt = np.stack([1 - lab_h, lab_h], axis=1)
    th_o, p_o = theta, pr
    for _ in range(2):
        th_o, p_o = orc.em_step_np(th_o, p_o, ids_h, cnt)
    eng = EMEngine(P, K)
    eng.set_train_links(g1, g2, g3, 1 - lab, lab)
    assert eng.flags & 32, "slot-segmented E-step is the default for K >= 4"
    eng.set_params(theta, pr)
    eng.em_iteration()
    eng.em_iteration()
    th, p = eng.get_params()
    assert _relerr(th, th_o) < 1e-10 and _relerr(p, p_o) < 1e-10
    assert eng.loglik("train") == pytest.approx(orc.loglik_np(th_o, p_o, ids_h, cnt), rel=1e-11)


def test_results_table_order_with_exact_ties(torch_cuda):
    """calculate_test_set_results: descending (score, key string, label) like list.sort(); reverse() (TIP.py:568-569).
    Test triplets over genes with IDENTICAL theta rows give bit-equal scores, so the key-string order decides."""
    from trigenicinteractionpredictor_b200 import Model
    rng = np.random.default_rng(8)
    P, K, T = 40, 3, 300
    m = Model()
    g = np.stack([rng.permutation(P)[:3] for _ in range(200)]).astype(np.int32)
    gt = np.stack([rng.permutation(P)[:3] for _ in range(T)]).astype(np.int32)
    gt = np.unique(gt, axis=0)
    labt = (rng.random(len(gt)) < 0.3).astype(np.int32)
    g[:P // 3, :] = np.arange(P // 3 * 3).reshape(-1, 3)
    lab = (rng.random(len(g)) < 0.3).astype(np.int32)
    m.set_links_soa((g[:, 0], g[:, 1], g[:, 2], 1 - lab, lab), (gt[:, 0], gt[:, 1], gt[:, 2], 1 - labt, labt), P=P)
    random.seed(3)
    m.initialize_parameters(K)
    th = np.array(m.theta)
    th[:] = th[rng.integers(0, 4, P)]                  # four distinct theta rows only: many exactly equal scores
    m.theta = th.tolist()
    m.calculate_test_set_results()
    sc = m._scores.cpu().numpy().tolist()
    exp = [[s, "%d_%d_%d" % tuple(k), int(l)] for s, k, l in zip(sc, gt.tolist(), labt.tolist())]
    exp.sort()
    exp.reverse()
    assert len({r[0] for r in exp}) < len(exp) // 2, "the case must contain exact ties"
    assert m.results == exp


def test_links_handed_over_as_arrays_train_like_the_files(torch_cuda):
    """Model.set_links_soa against Model.get_traintest on the same links: iterations, likelihoods, scores, metrics."""
    from trigenicinteractionpredictor_b200 import Model
    ref = _model(BASE, "train1.dat", "test1.dat")
    tr_arr = Model._soa(ref.links)
    te_arr = Model._soa(ref.test_links)
    m = Model()
    m.set_links_soa(tr_arr, te_arr, P=ref.P)
    for mod in (ref, m):
        random.seed(1000)
        mod.initialize_parameters(10)
        mod.make_iterations(6)
    assert np.array_equal(np.array(m.theta), np.array(ref.theta)) or _relerr(m.theta, ref.theta) < 1e-12
    assert m.compute_likelihood() == pytest.approx(ref.compute_likelihood(), rel=1e-12)
    assert m.compute_likelihood("test") == pytest.approx(ref.compute_likelihood("test"), rel=1e-12)
    m.calculate_test_set_results()
    ref.calculate_test_set_results()
    assert [r[1:] for r in m.results] == [r[1:] for r in ref.results]
    np.testing.assert_allclose(m.calculate_metrics(), ref.calculate_metrics(), rtol=1e-12)
    assert m.to_string().split("\n")[2:6] == ref.to_string().split("\n")[2:6]


def test_one_device_per_process_is_enforced(torch_cuda):
    """libtip keeps per-process device state: a second device in one process is refused, not silently mis-launched."""
    torch = torch_cuda
    from trigenicinteractionpredictor_b200._cabi import TipLibraryError
    from trigenicinteractionpredictor_b200.engine import EMEngine
    EMEngine(10, 2, device="cuda:0")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs to name a second device")
    with pytest.raises(TipLibraryError):
        EMEngine(10, 2, device="cuda:1")


def test_eager_iterations_and_graph_replays_can_be_mixed(torch_cuda):
    """em_iterations (CUDA-graph replays) interleaved with em_iteration (eager) = the same number of eager iterations."""
    from trigenicinteractionpredictor_b200.engine import EMEngine
    g, n0, n1, theta, pr = _random_problem(200, 3000, 6, 77)
    out = []
    for plan in ((5, 1, 4, 1, 1), (12,)):
        eng = EMEngine(200, 6)
        eng.set_train_links(g[:, 0], g[:, 1], g[:, 2], n0, n1)
        eng.set_params(theta, pr)
        for n in plan:
            if n == 1:
                eng.em_iteration()
            else:
                eng.em_iterations(n)
        out.append(eng.get_params())
    assert _relerr(out[0][0], out[1][0]) < 1e-11 and _relerr(out[0][1], out[1][1]) < 1e-11


def test_cli_writes_reference_shaped_sample_files(torch_cuda, tmp_path, capsys):
    """The command line of TIP.py:1148-1279: same flags, Sample_{s}_K{k}.csv naming, skip-if-exists."""
    from trigenicinteractionpredictor_b200 import main
    out = str(tmp_path) + os.sep
    argv = ["-t", os.path.join(BASE, "train1.dat"), "-e", os.path.join(BASE, "test1.dat"), "-k", "2", "-i", "300",
            "-f", "5", "-b", "10", "-n", "2", "-o", out, "--seed", "1000"]
    assert main(argv) == 0
    files = sorted(os.listdir(out))
    assert files == ["Sample_0_K2.csv", "Sample_1_K2.csv"]
    got = open(os.path.join(out, "Sample_0_K2.csv"), encoding="utf-8").read().split("\n")
    exp = open(os.path.join(BASE, "Sample_0_K2.csv"), encoding="utf-8").read().split("\n")
    assert got[0].split("\t")[0] == "Max Likelihood:" and len(got) == len(exp)
    assert float(got[0].split("\t")[1]) == pytest.approx(float(exp[0].split("\t")[1]), rel=1e-9)
    assert got[2:8] == exp[2:8]                      # P, #links, K, R lines are identical text
    stamp = os.path.getmtime(os.path.join(out, "Sample_0_K2.csv"))
    assert main(argv) == 0                           # second run: both samples exist and are skipped
    assert os.path.getmtime(os.path.join(out, "Sample_0_K2.csv")) == stamp
    assert "Likelihood has converged" in capsys.readouterr().out
    # --mode segmented: the gene-segmented E-step through the same command line, same report to 1e-9
    out2 = str(tmp_path / "seg") + os.sep
    os.makedirs(out2)
    assert main(argv[:-4] + ["-n", "1", "-o", out2, "--seed", "1000", "--mode", "segmented"]) == 0
    got2 = open(os.path.join(out2, "Sample_0_K2.csv"), encoding="utf-8").read().split("\n")
    assert float(got2[0].split("\t")[1]) == pytest.approx(float(exp[0].split("\t")[1]), rel=1e-9)
    # --reducible: the sample file carries the gene list and goes straight through the package's own reducer
    from trigenicinteractionpredictor_b200 import testResultsReducer as trr
    out3 = str(tmp_path / "red") + os.sep
    os.makedirs(out3 + "fold1")
    assert main(argv[:-4] + ["-n", "2", "-o", out3 + "fold1" + os.sep, "--seed", "1000", "--reducible"]) == 0
    text = open(os.path.join(out3, "fold1", "Sample_0_K2.csv"), encoding="utf-8").read()
    lines = text.split("\n")
    assert "LIST OF REGISTERED GENES" in text and "LIST OF LINKS BETWEEN GENE IDS" in text and lines[2:8] == got[2:8]
    assert float(lines[0].split("\t")[1]) == pytest.approx(float(got[0].split("\t")[1]), rel=1e-9)
    names = trr._read_gene_names(os.path.join(out3, "fold1", "Sample_0_K2.csv"))
    assert len(names) == int(got[2].split("\t")[1])       # the reducer's parser finds every gene of the model


def test_cfg1_full_run_matches_reference(torch_cuda, tmp_path):
    """BASELINE.json configs[0] end to end: 1,000 genes x 100,000 triplets, get_input -> fold -> fold 1 -> K=2 ->
    100 EM iterations with the likelihood after every one, against the record the unmodified reference produced
    (oracle/gen_golden_cfg1.py, ~2.5 min of CPython there).  Inputs are regenerated here and pinned by sha256."""
    import hashlib
    from trigenicinteractionpredictor_b200 import Model, synth
    with open(os.path.join(GOLDEN, "cfg1", "record.json")) as fh:
        rec = json.load(fh)

    def sha(path):
        with open(path, "rb") as f:
            return hashlib.sha256(f.read()).hexdigest()

    g, lab = synth.planted_triplets(1000, 100_000, seed=1, shape="uniform")
    raw = str(tmp_path / "input_s2.tsv")
    synth.write_raw_s2(raw, g, lab, synth.gene_names(1000))
    assert sha(raw) == rec["sha256"]["raw"]
    m = Model()
    m.get_input(raw)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        np.random.seed(2)
        m.fold()
    finally:
        os.chdir(cwd)
    train, test = str(tmp_path / "train1.dat"), str(tmp_path / "test1.dat")
    assert sha(train) == rec["sha256"]["train"] and sha(test) == rec["sha256"]["test"]      # fold: bit-exact
    mm = Model()
    mm.get_traintest(train, test)
    assert (mm.P, len(mm.links), len(mm.test_links)) == (rec["P"], rec["train_links"], rec["test_links"])
    random.seed(1000)
    mm.initialize_parameters(2)
    like = [mm.compute_likelihood()]
    for _ in range(100):
        mm.make_iteration()
        like.append(mm.compute_likelihood())
    np.testing.assert_allclose(like, rec["loglik"], rtol=RTOL)                               # every iteration
    theta = np.array(mm.theta)
    assert _relerr(theta[rec["theta_rows"]], rec["theta_final_rows"]) < RTOL
    assert theta.sum() == pytest.approx(rec["theta_final_sum"], rel=RTOL)
    assert (theta ** 2).sum() == pytest.approx(rec["theta_final_sq"], rel=RTOL)
    assert _relerr(np.array(mm.pr).reshape(-1), rec["pr_final"]) < RTOL
    assert mm.compute_likelihood("test") == pytest.approx(rec["heldout"], rel=RTOL)
    mm.calculate_test_set_results()
    met = mm.calculate_metrics()
    assert met[3] == pytest.approx(rec["metrics"][3], abs=1e-6)                              # AUC
    np.testing.assert_allclose(met[:3], rec["metrics"][:3], atol=1e-3)
    head = mm.results[:50]
    assert _relerr([r[0] for r in head], [r[0] for r in rec["results_head"]]) < RTOL
    assert [r[2] for r in head] == [r[2] for r in rec["results_head"]]


def test_streamed_host_entry_gives_up_cleanly(torch_cuda, monkeypatch):
    """A streamed E-step whose rows do not arrive in time reports it (-3 / host_rows_arrived() False) instead of
    hanging, and the copy-then-compute path (TIP_HOST_NO_STREAM=1) of the same entry point still gives the result."""
    from trigenicinteractionpredictor_b200 import _cabi
    from trigenicinteractionpredictor_b200.engine import EMEngine
    lib = _cabi.load()
    P, L, K = 400, 200000, 10
    g, n0, n1, theta, pr = _random_problem(P, L, K, 91)
    a = EMEngine(P, K)
    a.set_train_links(g[:, 0], g[:, 1], g[:, 2], n0, n1)
    a.set_params(theta, pr)
    a.em_iteration()
    th_a, p_a = a.get_params()
    rows = np.ascontiguousarray(a.train.rows.cpu().numpy())
    deg = np.ascontiguousarray(a.degrees().astype(np.int32))
    monkeypatch.setenv("TIP_STREAM_TIMEOUT_US", "0")                    # every wait gives up at once
    th_h, p_h = np.ascontiguousarray(theta.copy()), np.ascontiguousarray(pr.copy())
    rcs = [lib.tip_em_iterations_host(P, K, rows.ctypes.data, a.train.n_rows, a.train.n_rows_r0, deg.ctypes.data,
                                      th_h.ctypes.data, p_h.ctypes.data, 1, 0) for _ in range(3)]
    # (pageable rows are staged by the driver before the kernel starts, so a call may also succeed: both outcomes are legal)
    assert all(rc in (0, -3) for rc in rcs)
    if -3 in rcs:
        assert b"did not arrive" in lib.tip_last_error()
    monkeypatch.delenv("TIP_STREAM_TIMEOUT_US")
    monkeypatch.setenv("TIP_HOST_NO_STREAM", "1")
    th_h, p_h = np.ascontiguousarray(theta.copy()), np.ascontiguousarray(pr.copy())
    rc = lib.tip_em_iterations_host(P, K, rows.ctypes.data, a.train.n_rows, a.train.n_rows_r0, deg.ctypes.data,
                                    th_h.ctypes.data, p_h.ctypes.data, 1, 0)
    assert rc == 0, lib.tip_last_error()
    assert _relerr(th_h, th_a) < 1e-11 and _relerr(p_h, p_a) < 1e-11
    monkeypatch.delenv("TIP_HOST_NO_STREAM")
    th_h, p_h = np.ascontiguousarray(theta.copy()), np.ascontiguousarray(pr.copy())
    rc = lib.tip_em_iterations_host(P, K, rows.ctypes.data, a.train.n_rows, a.train.n_rows_r0, deg.ctypes.data,
                                    th_h.ctypes.data, p_h.ctypes.data, 1, 0)                # streamed again, healthy
    assert rc == 0, lib.tip_last_error()
    assert _relerr(th_h, th_a) < 1e-11 and _relerr(p_h, p_a) < 1e-11


def test_streamed_host_rows_many_sizes_and_repeats(torch_cuda):
    """The streamed E-step races a kernel against the rows' DMA: it must give the resident-row statistics
    every time - sizes from a few tiles to 1.6 M rows, both row formats, pinned rows, repeated back to back - and the
    verification of what it consumed must never fire on a healthy platform."""
    torch = torch_cuda
    from trigenicinteractionpredictor_b200 import _cabi
    from trigenicinteractionpredictor_b200.engine import EMEngine
    lib = _cabi.load()
    P, K = 250, 10
    for L, reps in ((300, 4), (5000, 4), (120_000, 6), (1_600_000, 6)):
        g, n0, n1, theta, pr = _random_problem(P, L, K, 1000 + L % 97)
        eng = EMEngine(P, K, flags=0)
        eng.set_train_links(g[:, 0], g[:, 1], g[:, 2], n0, n1)
        eng.set_params(theta, pr)
        eng.em_step()
        want = eng.stats.cpu().numpy().copy()
        rows16 = eng.train.rows.cpu().pin_memory()
        rows8 = torch.empty(eng.train.n_rows, dtype=torch.int64).pin_memory()
        assert lib.tip_rows_compact_host(rows16.data_ptr(), eng.train.n_rows, rows8.data_ptr()) == 0
        for rep in range(reps):
            for compact, src in ((False, rows16), (True, rows8)):
                eng.em_step_host_rows(src, compact)
                got = eng.stats.cpu().numpy()
                assert eng.host_rows_arrived(), "L=%d rep %d: the streamed step reported a stall or a checksum mismatch" % (L, rep)
                assert _relerr(got[: P * K], want[: P * K]) < 1e-12 and _relerr(got[P * K:-1], want[P * K:-1]) < 1e-12


def test_streamed_step_checksum_mismatch_is_detected_and_repaired(torch_cuda, monkeypatch):
    """TIP_STREAM_INJECT_FAULT makes the verification disagree with what the kernel consumed: tip_em_iterations_host
    must still return the exact iteration (M-steps held back, iterations repeated from the resident rows) and
    tip_em_step_host_rows must flag the statistics (host_rows_arrived() False)."""
    torch = torch_cuda
    from trigenicinteractionpredictor_b200 import _cabi
    from trigenicinteractionpredictor_b200.engine import EMEngine
    lib = _cabi.load()
    P, L, K = 400, 60000, 10
    g, n0, n1, theta, pr = _random_problem(P, L, K, 17)
    a = EMEngine(P, K, flags=0)
    a.set_train_links(g[:, 0], g[:, 1], g[:, 2], n0, n1)
    a.set_params(theta, pr)
    for _ in range(3):
        a.em_iteration()
    th_a, p_a = a.get_params()
    rows = a.train.rows.cpu().pin_memory()
    deg = np.ascontiguousarray(a.degrees().astype(np.int32))
    monkeypatch.setenv("TIP_STREAM_INJECT_FAULT", "1")
    monkeypatch.setenv("TIP_HOST_SEG3_AFTER_STREAM", "1")    # flags 32: streamed K^3 first iteration, then the rows are ordered
    for n_iter_flags in ((3, 0), (3, 32)):
        th_h, p_h = np.ascontiguousarray(theta.copy()), np.ascontiguousarray(pr.copy())
        rc = lib.tip_em_iterations_host(P, K, rows.data_ptr(), a.train.n_rows, a.train.n_rows_r0, deg.ctypes.data,
                                        th_h.ctypes.data, p_h.ctypes.data, n_iter_flags[0], n_iter_flags[1])
        assert rc == 0, lib.tip_last_error()
        assert _relerr(th_h, th_a) < 1e-10 and _relerr(p_h, p_a) < 1e-10
    a.set_params(theta, pr)
    a.em_step_host_rows(rows, False)
    assert not a.host_rows_arrived()
    monkeypatch.delenv("TIP_STREAM_INJECT_FAULT")
    monkeypatch.delenv("TIP_HOST_SEG3_AFTER_STREAM")
    a.em_step_host_rows(rows, False)
    assert a.host_rows_arrived()
    # and without the fault: the slot-segmented kernels behind the host entry (rows ordered on the device while pass A runs),
    # 16-byte and 8-byte host rows, repeated (the scratch and the events are reused)
    rows8 = torch.empty(a.train.n_rows, dtype=torch.int64).pin_memory()
    assert lib.tip_rows_compact_host(rows.data_ptr(), a.train.n_rows, rows8.data_ptr()) == 0
    for rep in range(3):
        for src, fl in ((rows, 32), (rows8, 32 | 16), (rows8, 32 | 64 | 16)):
            th_h, p_h = np.ascontiguousarray(theta.copy()), np.ascontiguousarray(pr.copy())
            rc = lib.tip_em_iterations_host(P, K, src.data_ptr(), a.train.n_rows, a.train.n_rows_r0, deg.ctypes.data,
                                            th_h.ctypes.data, p_h.ctypes.data, 3, fl)
            assert rc == 0, lib.tip_last_error()
            assert _relerr(th_h, th_a) < 1e-10 and _relerr(p_h, p_a) < 1e-10


def test_order_rows_by_gene_holds_the_same_runs(torch_cuda):
    """tip_order_rows_by_gene (counting sort, not stable) against tip_order_rows (stable radix sort): every (rating, gene)
    run holds the same links at the same place, only their order inside a run may differ; hub-shaped and uniform rows."""
    torch = torch_cuda
    from trigenicinteractionpredictor_b200 import _cabi, synth
    from trigenicinteractionpredictor_b200.engine import EMEngine
    lib = _cabi.load()
    P = 300
    for shape, L in (("uniform", 40_000), ("kuzmin", 40_000), ("uniform", 700)):
        if shape == "kuzmin":
            a, b, c, lab = synth.kuzmin_links_soa(P, L, seed=5, n_query=6)
        else:
            a, b, c, lab = synth.planted_links_soa(P, L, seed=5)
        eng = EMEngine(P, 5, flags=32)
        eng.set_train_links(a, b, c, 1 - lab, lab)
        t = eng.train
        n = t.n_rows
        want = t.rows3_buf.cpu().numpy()
        out = torch.zeros_like(t.rows3_buf)
        out[: 4 * n].copy_(t.rows3_buf[: 4 * n])
        ws = torch.empty_like(t.order_ws)
        rc = lib.tip_order_rows_by_gene(out.data_ptr(), n, t.n_rows_r0, P, ws.data_ptr(), int(ws.numel()),
                                        out.data_ptr() + n * 16, torch.cuda.current_stream().cuda_stream)
        assert rc == 0, lib.tip_last_error()
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        for blk in (1, 2):
            w = want[4 * n * blk: 4 * n * (blk + 1)].reshape(n, 4)
            g = got[4 * n * blk: 4 * n * (blk + 1)].reshape(n, 4)
            assert np.array_equal(w[:, 0], g[:, 0]), "run layout differs (%s, order %d)" % (shape, blk)
            blockid = (np.arange(n) >= t.n_rows_r0).astype(np.int64)

            def canon(x):   # rows sorted by (rating block, gene with padding last, position in order a)
                gene = np.where(x[:, 3] < 0, 1 << 30, x[:, 0].astype(np.int64))
                return x[np.lexsort((x[:, 3], gene, blockid))]
            assert np.array_equal(canon(g), canon(w)) and np.array_equal(canon(w), w)
        # and the E-step on the re-ordered rows gives the same statistics
        g0, n0, n1, theta, pr = _random_problem(P, 640, 5, 3)
        eng.set_params(theta, pr)
        eng.em_step()
        ref = eng.stats.cpu().numpy().copy()
        t.rows3_buf.copy_(out)
        eng.em_step()
        np.testing.assert_allclose(eng.stats.cpu().numpy(), ref, rtol=1e-11, atol=1e-300)


@pytest.mark.parametrize("flags", [None, 0, 1], ids=["default", "k3", "anyK"])
@pytest.mark.parametrize("K", [2, 3, 10])
def test_digenic_extension_trace_matches_the_patched_reference(torch_cuda, K, flags):
    """SURVEY f-4: the drop-in for TrigenicInteractionPredictor_23.py (triplets + pair links with their own qr): theta, pr,
    qr and the log-likelihood per iteration within 1e-9 of the author's code (golden from oracle/gen_golden_23.py)."""
    import contextlib
    import io
    from trigenicinteractionpredictor_b200.TrigenicInteractionPredictor_23 import Model
    case = os.path.join(GOLDEN, "digenic")
    tr = np.load(os.path.join(case, "trace23_K%d.npz" % K))
    m = Model(flags=flags)
    with contextlib.redirect_stdout(io.StringIO()):
        m.get_train_test(os.path.join(case, "train_mixed.dat"), os.path.join(case, "test_mixed.dat"))
    random.seed(2300 + K)
    m.initialize_parameters(K)
    assert np.array_equal(np.array(m.qr), tr["qr0"])
    assert m.compute_likelihood() == pytest.approx(tr["loglik"][0], rel=RTOL)
    for it in range(len(tr["loglik"]) - 1):
        m.make_iteration()
        assert _relerr(m.theta, tr["theta%d" % (it + 1)]) < RTOL, "theta iteration %d" % (it + 1)
        assert _relerr(m.pr, tr["pr%d" % (it + 1)]) < RTOL, "pr iteration %d" % (it + 1)
        assert _relerr(m.qr, tr["qr%d" % (it + 1)]) < RTOL, "qr iteration %d" % (it + 1)
        assert m.compute_likelihood() == pytest.approx(tr["loglik"][it + 1], rel=RTOL)

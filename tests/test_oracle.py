"""CPU: the oracle against the golden fixtures (the reference's own FAISS-written file) and
against itself (C restatement vs numpy twins vs float64 brute force)."""
import hashlib
import os
import pickle

import numpy as np
import pytest

import oracle as orc

FLT_MAX = np.finfo(np.float32).max


def _queries(golden, name, xb):
    return xb if name == "self" else golden["perturbed"]


def test_fixture_file_layout(golden, tmp_path):
    raw = open(golden["index_path"], "rb").read()
    assert hashlib.sha256(raw).hexdigest() == golden["sha256_index"]
    x, metric = orc.np_read_index(golden["index_path"])
    xc, metric_c = orc.c_read_index(golden["index_path"])
    assert metric == metric_c == orc.METRIC_L2
    assert x.shape == (golden["ntotal"], golden["d"]) == (23, 384)
    assert np.array_equal(x, xc)
    # writers reproduce the FAISS-written bytes exactly
    p1, p2 = tmp_path / "np.bin", tmp_path / "c.bin"
    orc.np_write_index(p1, x, metric)
    orc.c_write_index(str(p2), x, metric)
    assert open(p1, "rb").read() == raw
    assert open(p2, "rb").read() == raw
    assert pickle.load(open(golden["mapping_path"], "rb")) == golden["mapping"]


def test_ip_fourcc_roundtrip(tmp_path):
    x = orc.np_synth_rows(1, 0, 7, 12)
    p = tmp_path / "ip.bin"
    orc.c_write_index(str(p), x, orc.METRIC_INNER_PRODUCT)
    assert open(p, "rb").read()[:4] == b"IxFI"
    y, m = orc.np_read_index(p)
    assert m == orc.METRIC_INNER_PRODUCT and np.array_equal(x, y)


@pytest.mark.parametrize("algo", [1, 2])
def test_fixture_known_answers(golden, algo):
    xb, _ = orc.np_read_index(golden["index_path"])
    for case in golden["cases"]:
        q = _queries(golden, case["queries"], xb)
        D, I = orc.c_search(xb, q, case["k"], case["metric"], algo=algo)
        I_ref = np.asarray(case["ids"], np.int64)
        D_ref = np.asarray(case["dists"], np.float64)
        r = orc.recall_and_errors(D, I, D_ref, I_ref, case["metric"],
                                  rel_tol=1e-5 if algo == 1 else 2e-5)
        assert r["recall"] == 1.0 and r["id_mismatch"] == 0 and r["padding_ok"], (case["metric"], case["k"], r)
        # exact-difference path: <=1e-5 relative; expanded (blas) path is allowed its cancellation error
        assert r["max_rel_err"] < (1e-5 if algo == 1 else 2e-5)
        if case["k"] > xb.shape[0]:
            assert (I[:, xb.shape[0]:] == -1).all()
            pad = FLT_MAX if case["metric"] == orc.METRIC_L2 else -FLT_MAX
            assert (D[:, xb.shape[0]:] == pad).all()


def test_synth_generator_pinned(synth_known):
    d = synth_known["d"]
    r0 = orc.c_synth_rows(synth_known["seed"], 0, 1, d)
    assert (np.round(r0[0, :16] * 65536).astype(int).tolist() == synth_known["ints_row0"])
    r1 = orc.np_synth_rows(synth_known["seed"], 1000003, 1, d)
    assert (np.round(r1[0, :16] * 65536).astype(int).tolist() == synth_known["ints_row1000003"])
    q = orc.c_synth_rows(synth_known["seed_q"], 7, 1, d)
    assert (np.round(q[0, :16] * 65536).astype(int).tolist() == synth_known["ints_q_row7"])
    # C and numpy generators agree bit for bit, plain and normalised, at an arbitrary offset
    for norm in (False, True):
        a = orc.c_synth_rows(99, 12345, 33, 100, norm)
        b = orc.np_synth_rows(99, 12345, 33, 100, norm)
        assert np.array_equal(a, b)
    big = orc.c_synth_rows(1234, 0, 2000, 384)
    assert abs(big.mean()) < 0.01 and abs(big.std() - 1.0) < 0.01
    nrm = orc.c_synth_rows(1234, 0, 50, 768, True)
    assert np.allclose(np.linalg.norm(nrm.astype(np.float64), axis=1), 1.0, atol=1e-6)


@pytest.mark.parametrize("metric", [orc.METRIC_L2, orc.METRIC_INNER_PRODUCT])
@pytest.mark.parametrize("nq,k", [(1, 1), (3, 10), (25, 10), (40, 100), (5, 128)])
def test_c_vs_f64_random(metric, nq, k):
    xb = orc.c_synth_rows(1234, 0, 3000, 96)
    xq = orc.c_synth_rows(5678, 0, nq, 96)
    D_ref, I_ref = orc.np_search_f64(xb, xq, k, metric)
    for algo in (0, 1, 2):
        D, I = orc.c_search(xb, xq, k, metric, algo=algo)
        r = orc.recall_and_errors(D, I, D_ref, I_ref, metric, rel_tol=1e-5)
        assert r["recall"] == 1.0 and r["id_mismatch"] == 0, (algo, r)
    D, I = orc.np_search_blas(xb, xq, k, metric)
    r = orc.recall_and_errors(D, I, D_ref, I_ref, metric, rel_tol=1e-5)
    assert r["recall"] == 1.0 and r["id_mismatch"] == 0, r


@pytest.mark.parametrize("metric", [orc.METRIC_L2, orc.METRIC_INNER_PRODUCT])
def test_oracle_vs_independent_brute_force(metric):
    """faiss itself cannot run here (section 1 of DESIGN.md), so the oracle's notion of "the true top-k" is checked
    against an implementation this repo did not write: scikit-learn's brute-force NearestNeighbors in float64 (squared
    euclidean; for inner product, cosine distance on normalised rows, which ranks them identically).  It pins the
    neighbour SETS and distances, not faiss's tie order."""
    sk = pytest.importorskip("sklearn.neighbors")
    n, d, nq, k = 4000, 64, 33, 10
    xb = orc.c_synth_rows(77, 0, n, d)
    xq = orc.c_synth_rows(78, 0, nq, d)
    if metric == orc.METRIC_INNER_PRODUCT:
        xb = (xb / np.linalg.norm(xb, axis=1, keepdims=True)).astype(np.float32)
        xq = (xq / np.linalg.norm(xq, axis=1, keepdims=True)).astype(np.float32)
    nn = sk.NearestNeighbors(n_neighbors=k, algorithm="brute", metric="sqeuclidean" if metric == orc.METRIC_L2 else "cosine")
    nn.fit(xb.astype(np.float64))
    d_sk, i_sk = nn.kneighbors(xq.astype(np.float64))
    d_sk = d_sk if metric == orc.METRIC_L2 else 1.0 - d_sk          # cosine distance -> inner product of unit vectors
    for name, (D, I) in {"float64 brute force": orc.np_search_f64(xb, xq, k, metric),
                         "C seq path": orc.c_search(xb, xq, k, metric, algo=1),
                         "C blas path": orc.c_search(xb, xq, k, metric, algo=2)}.items():
        assert all(set(I[q].tolist()) == set(i_sk[q].tolist()) for q in range(nq)), name
        assert np.allclose(np.sort(D, axis=1), np.sort(d_sk, axis=1), rtol=1e-5, atol=1e-5), name


def test_edge_cases():
    xb = orc.c_synth_rows(1, 0, 10, 8)
    # k > ntotal
    D, I = orc.c_search(xb, xb[:2], 16)
    assert (I[:, 10:] == -1).all() and (D[:, 10:] == FLT_MAX).all() and (I[:, 0] == [0, 1]).all()
    # empty database
    D, I = orc.c_search(np.zeros((0, 8), np.float32), xb[:2], 3)
    assert (I == -1).all() and (D == FLT_MAX).all()
    D, I = orc.c_search(np.zeros((0, 8), np.float32), xb[:2], 3, orc.METRIC_INNER_PRODUCT)
    assert (I == -1).all() and (D == -FLT_MAX).all()
    # duplicates: ties resolve to the lower id
    dup = np.repeat(xb[:1], 5, 0)
    D, I = orc.c_search(dup, xb[:1], 3)
    assert I.tolist() == [[0, 1, 2]] and (D == 0).all()
    # NaN rows never enter
    bad = xb.copy()
    bad[3] = np.nan
    D, I = orc.c_search(bad, xb[:1], 10)
    assert 3 not in I[0].tolist() and I[0, -1] == -1
    # k <= 0 is an assertion in faiss
    with pytest.raises(AssertionError):
        orc.c_search(xb, xb[:1], 0)
    # blas path clamps negative expanded distances to 0
    big = (xb * 1000).astype(np.float32)
    D, I = orc.c_search(big, big, 1, algo=2)
    assert (D >= 0).all() and (I[:, 0] == np.arange(10)).all()

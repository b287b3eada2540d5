"""Known-answer tests for the oracle.  The reference ships no golden vectors for this path (SURVEY.md
§8c: parity unpinned), so these small cases are hand-computed here; they pin the restatement of
faiss.IndexFlatIP / normalize_L2 (faissRetriever.py:18-24,34-37), bm25s scoring (bm25Retriever.py:75-79)
and the fusion semantics (ensembleRetriever.py:58-229)."""
import math

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import bm25 as obm, flat_ip, fusion as ofu, sharded as osh

FLT_MAX = np.finfo(np.float32).max


def test_normalize_l2_known_values_and_zero_row():
    x = np.array([[3.0, 4.0], [0.0, 0.0], [1.0, 0.0], [-2.0, 0.0]], dtype=np.float32)
    y = flat_ip.normalize_l2(x)
    assert y[0].tolist() == [np.float32(3.0) * (np.float32(1.0) / np.float32(5.0)), np.float32(4.0) * (np.float32(1.0) / np.float32(5.0))]
    assert y[1].tolist() == [0.0, 0.0]          # zero rows untouched
    assert y[2].tolist() == [1.0, 0.0]
    assert y[3].tolist() == [-1.0, 0.0]
    assert x[0].tolist() == [3.0, 4.0]          # returns a copy


def test_bf16_round_to_nearest_even():
    # 1 + 2^-8 is exactly halfway between bf16 neighbours 1.0 and 1 + 2^-7 -> ties to even (1.0)
    x = np.array([1.0 + 2.0 ** -8, 1.0 + 3 * 2.0 ** -8, 1.0 + 2.0 ** -8 + 2.0 ** -20, -0.3], dtype=np.float32)
    y = flat_ip.bf16_round(x)
    assert y[0] == np.float32(1.0)
    assert y[1] == np.float32(1.0 + 2.0 ** -6)  # halfway between 1+2^-7 and 1+2^-6 -> even mantissa
    assert y[2] == np.float32(1.0 + 2.0 ** -7)  # just above halfway rounds up
    assert (y.view(np.uint32) & 0xFFFF == 0).all()
    from veritasfi_b200.synth import bf16_round_np
    r = np.random.default_rng(0).standard_normal(10000).astype(np.float32)
    assert (bf16_round_np(r) == flat_ip.bf16_round(r)).all()


def test_flat_ip_hand_case_ties_duplicates_padding():
    xb = np.array([[1, 0], [0, 1], [1, 0], [0.6, 0.8], [0, 0]], dtype=np.float32)  # rows 0 and 2 identical, row 4 zero
    xq = np.array([[1, 0], [0, 1]], dtype=np.float32)
    D, I = flat_ip.search_exhaustive(xq, xb, 7)
    # q0: scores [1,0,1,0.6,0] -> 1(id0),1(id2),0.6(id3),0(id1),0(id4) then padding
    assert I[0].tolist() == [0, 2, 3, 1, 4, -1, -1]
    assert D[0, :5].tolist() == [1.0, 1.0, np.float32(0.6), 0.0, 0.0]
    assert (D[0, 5:] == -FLT_MAX).all()
    # q1: scores [0,1,0,0.8,0]
    assert I[1].tolist() == [1, 3, 0, 2, 4, -1, -1]
    D2, I2 = flat_ip.search_exhaustive(xq, xb, 2, id_base=100)
    assert I2.tolist() == [[100, 102], [101, 103]]


def test_canonical_score_is_sequential_fp64_rounded_once():
    rng = np.random.default_rng(3)
    q = rng.standard_normal(1024).astype(np.float32)
    x = rng.standard_normal((5, 1024)).astype(np.float32)
    got = flat_ip.canon_scores(q, x, np.arange(5))
    for i in range(5):
        acc = 0.0
        for j in range(1024):
            acc += float(q[j]) * float(x[i, j])   # python floats are fp64; product of two fp32 is exact
        assert got[i] == np.float32(acc)


def test_search_fast_path_equals_exhaustive():
    rng = np.random.default_rng(7)
    xb = flat_ip.normalize_l2(rng.standard_normal((30000, 96)).astype(np.float32))
    xb[100] = xb[5]
    xb[20000] = xb[5]
    xq = flat_ip.normalize_l2(rng.standard_normal((24, 96)).astype(np.float32))
    xq[0] = xb[5]
    D0, I0 = flat_ip.search_exhaustive(xq, xb, 20)
    D1, I1 = flat_ip.search(xq, xb, 20)
    assert (I0 == I1).all() and (D0 == D1).all()
    assert I0[0, :3].tolist() == [5, 100, 20000]


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 300), st.integers(1, 24), st.integers(1, 40), st.integers(0, 2 ** 31 - 1))
def test_oracle_equals_stable_argsort_bruteforce(n, d, k, seed):
    rng = np.random.default_rng(seed)
    xb = rng.integers(-3, 4, size=(n, d)).astype(np.float32)   # small integers: many exact ties, exact sums
    xq = rng.integers(-3, 4, size=(3, d)).astype(np.float32)
    D, I = flat_ip.search_exhaustive(xq, xb, k)
    S = xq.astype(np.float64) @ xb.astype(np.float64).T
    for q in range(3):
        order = np.argsort(-S[q], kind="stable")[:k]
        m = len(order)
        assert I[q, :m].tolist() == order.tolist()
        assert D[q, :m].tolist() == S[q, order].astype(np.float32).tolist()
        assert (I[q, m:] == -1).all()


def test_bm25_three_doc_hand_case():
    # docs (token ids): d0 = [0,0,1], d1 = [1,2], d2 = [2,2,2,0]; N=3, avgdl=3
    docs = [[0, 0, 1], [1, 2], [2, 2, 2, 0]]
    indptr, indices, data = obm.build_index(docs, 3)
    assert indptr.tolist() == [0, 2, 4, 6]
    assert indices.tolist() == [0, 2, 0, 1, 1, 2]
    k1, b, N, avgdl = 1.5, 0.75, 3, 3.0

    def impact(df, tf, dl):
        idf = math.log(1 + (N - df + 0.5) / (df + 0.5))
        return np.float32(idf * tf / (tf + k1 * (1 - b + b * dl / avgdl)))

    want = [impact(2, 2, 3), impact(2, 1, 4), impact(2, 1, 3), impact(2, 1, 2), impact(2, 1, 2), impact(2, 3, 4)]
    assert data.tolist() == want
    # query tokens [2, 0, 2, 7]: token 7 unknown (skipped), token 2 counted twice, order matters in fp32
    s = obm.scores(indptr, indices, data, [2, 0, 2, 7], 3)
    s0 = np.float32(0) + want[0]
    s1 = np.float32(np.float32(0) + want[4]) + want[4]
    s2 = np.float32(np.float32(np.float32(0) + want[5]) + want[1]) + want[5]
    assert s.tolist() == [s0, np.float32(s1), np.float32(s2)]
    I, S = obm.retrieve(indptr, indices, data, [[2, 0, 2, 7], []], 3, 3)
    assert I[0].tolist() == np.argsort(-s, kind="stable").tolist()
    assert I[1].tolist() == [0, 1, 2] and S[1].tolist() == [0, 0, 0]   # empty query: all zero, id order
    assert (obm.scores_numpy(indptr, indices, data, [2, 0, 2, 7], 3) == s).all()


def test_rrf_textbook_example():
    # three lists over docs {1,2,3,4}: doc 2 is 1st,2nd,1st; doc 1 is 2nd,1st,absent ...
    ids = np.array([[[2, 1, 3], [1, 2, 4], [2, 4, -1]]], dtype=np.int64)
    oi, os_ = ofu.rrf(ids, 60.0, 4)
    f = np.float32
    s2 = f(f(f(1) / f(61)) + f(1) / f(62)) + f(1) / f(61)
    s1 = f(f(1) / f(62)) + f(1) / f(61)
    s4 = f(f(1) / f(63)) + f(1) / f(62)
    s3 = f(1) / f(63)
    assert oi[0].tolist() == [2, 1, 4, 3]
    assert os_[0].tolist() == [f(s2), f(s1), f(s4), f(s3)]
    pi, ps = ofu.rrf_python(ids, 60.0, 4)
    assert (pi == oi).all() and (ps == os_).all()
    # ties: docs 8 and 5 both appear once at rank 1 -> lower id first
    ids = np.array([[[8, -1], [5, -1]]], dtype=np.int64)
    oi, _ = ofu.rrf(ids, 60.0, 3)
    assert oi[0].tolist() == [5, 8, -1]


def test_union_is_priority_ordered_first_occurrence():
    ids = np.array([[[7, 3, 9], [3, 5, 7], [1, 9, -1]]], dtype=np.int64)
    sc = np.arange(9, dtype=np.float32).reshape(1, 3, 3)
    oi, os_, op, cnt = ofu.union(ids, sc)
    assert cnt.tolist() == [5]
    assert oi[0, :5].tolist() == [7, 3, 9, 5, 1]
    assert os_[0, :5].tolist() == [0, 1, 2, 4, 6]
    assert op[0, :5].tolist() == [0, 0, 0, 1, 2]
    assert (oi[0, 5:] == -1).all()
    assert ofu.union_python(ids, sc)[0] == [(7, 0.0, 0), (3, 1.0, 0), (9, 2.0, 0), (5, 4.0, 1), (1, 6.0, 2)]


@pytest.mark.parametrize("g", [1, 2, 3, 8])
def test_shard_merge_equals_unsharded(g):
    rng = np.random.default_rng(11)
    xb = flat_ip.normalize_l2(rng.standard_normal((1003, 32)).astype(np.float32))
    xb[500] = xb[2]
    xb[900] = xb[2]
    xq = flat_ip.normalize_l2(rng.standard_normal((6, 32)).astype(np.float32))
    xq[0] = xb[2]
    D0, I0 = flat_ip.search_exhaustive(xq, xb, 25)
    I1, D1 = osh.search_sharded(xq, xb, 25, g)
    assert (I0 == I1).all() and (D0 == D1).all()

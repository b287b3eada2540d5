"""GPU parity tests for what round 2 added (run on the B200 box with `-m gpu`): the exact streaming scorer and its
multi-CTA radix select, deep-k BM25 and the lazy k = N ranking, long BM25 queries, the one-launch hybrid fusion,
concurrent searches on one index.  Same bar as test_gpu_parity.py: ids and scores bit-exact against the oracle."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    return torch


def _world(n, d, nq, seed, bf16):
    from veritasfi_b200 import synth
    xb = synth.dense_corpus_np(n, d, seed, bf16=bf16)
    xq = synth.dense_queries_np(nq, d, seed, xb, bf16=bf16)
    return xb, xq


@pytest.mark.parametrize("n,d,nq,k,store", [
    (70_000, 128, 3, 2048, "bf16"),     # radix select (n > 32768), group of 3 queries
    (70_000, 100, 11, 500, "f32"),      # two groups (8 + 3), d padded, fp32 rows
    (20_000, 192, 8, 2048, "bf16"),     # single-CTA selection from the score array, row = 384 B (1.5 pieces)
    (40_000, 1024, 2, 2048, "f32"),     # 4 KB rows
    (1_000, 64, 5, 2048, "bf16"),       # k > n: padded like faiss
])
def test_exact_streaming_scorer_matches_oracle(torch_cuda, n, d, nq, k, store):
    torch = torch_cuda
    from oracle import flat_ip
    from veritasfi_b200 import _native as N
    from veritasfi_b200.dense import DenseIndex
    xb, xq = _world(n, d, nq, 41, store == "bf16")
    idx = DenseIndex(d, store=store)
    idx.add(xb)
    idx.set_option(N.OPT_FORCE_PATH, N.PATH_EXACT)
    ids, scores = idx.search_batch(torch.from_numpy(xq).cuda(), k)
    D0, I0 = flat_ip.search(xq, xb, k) if n > k else flat_ip.search_exhaustive(xq, xb, k)
    assert (ids.cpu().numpy() == I0).all()
    assert (scores.cpu().numpy() == D0).all()
    assert idx.stats().last_path == N.PATH_EXACT
    idx.close()


def test_exact_streaming_radix_select_breaks_mass_ties_by_id(torch_cuda):
    """45 000 copies of one row all tie at the top: the threshold key is decided by the id digits of the radix walk and
    the k winners are the lowest ids; zero rows tie in bulk at score 0 below them."""
    torch = torch_cuda
    from oracle import flat_ip
    from veritasfi_b200 import _native as N
    from veritasfi_b200.dense import DenseIndex
    n, d, k = 60_000, 64, 2048
    xb, xq = _world(n, d, 3, 43, True)
    rng = np.random.default_rng(5)
    dup = rng.permutation(n)[:45_000]
    xb[dup] = xb[dup[0]]
    xb[rng.permutation(n)[:5_000]] = 0.0
    xq[0] = xb[dup[0]]
    xq[1] = -xb[dup[0]]          # the copies tie at the bottom, the zero rows in the middle
    idx = DenseIndex(d, store="bf16")
    idx.add(xb)
    idx.set_option(N.OPT_FORCE_PATH, N.PATH_EXACT)
    ids, scores = idx.search_batch(torch.from_numpy(xq).cuda(), k)
    D0, I0 = flat_ip.search_exhaustive(xq, xb, k)
    assert (ids.cpu().numpy() == I0).all() and (scores.cpu().numpy() == D0).all()
    idx.close()


def test_deep_k_on_the_fused_kernel(torch_cuda):
    """k = 2048 for a batch too large for the streaming scorers: K1 with k' = 2560 candidates per query."""
    torch = torch_cuda
    from oracle import flat_ip
    from veritasfi_b200 import _native as N
    from veritasfi_b200.dense import DenseIndex
    xb, xq = _world(90_000, 128, 20, 47, True)
    idx = DenseIndex(128, store="bf16")
    idx.add(xb)
    ids, scores = idx.search_batch(torch.from_numpy(xq).cuda(), 2048)
    D0, I0 = flat_ip.search(xq, xb, 2048)
    assert idx.stats().last_path == N.PATH_FUSED
    assert (ids.cpu().numpy() == I0).all() and (scores.cpu().numpy() == D0).all()
    idx.close()


@pytest.mark.parametrize("m", [16, 32])
def test_denser_admission_samples_do_not_change_results(torch_cuda, m):
    """The m-th best of a denser row sample (small shards with deep lists: the title path of the hybrid retriever)."""
    torch = torch_cuda
    from oracle import flat_ip
    from veritasfi_b200 import _native as N
    from veritasfi_b200.dense import DenseIndex
    xb, xq = _world(125_000, 128, 140, 53, True)
    idx = DenseIndex(128, store="bf16")
    idx.add(xb)
    idx.set_option(N.OPT_FORCE_PATH, N.PATH_FUSED)
    idx.set_option(N.OPT_TAU_M, m)
    ids, scores = idx.search_batch(torch.from_numpy(xq).cuda(), 200)
    D0, I0 = flat_ip.search(xq, xb, 200)
    assert (ids.cpu().numpy() == I0).all() and (scores.cpu().numpy() == D0).all()
    assert idx.stats().hint_retries == 0
    idx.set_option(N.OPT_TAU_M, 0)            # automatic choice: 125k rows at k' = 256 takes the dense sample
    ids, scores = idx.search_batch(torch.from_numpy(xq).cuda(), 200)
    assert (ids.cpu().numpy() == I0).all() and (scores.cpu().numpy() == D0).all()
    idx.close()


@pytest.mark.parametrize("n,d,nq,k,store,hint", [
    (150_001, 128, 9, 100, "bf16", 1),     # smallest batch past the streaming scorers, ragged last corpus tile
    (90_000, 1024, 64, 100, "bf16", 1),    # the largest resident query block (128 KB)
    (60_000, 256, 33, 10, "bf16", 0),      # no admission hint: thresholds rise by buffer compaction alone
    (70_000, 100, 20, 50, "f32", 1),       # fp32 rows: 3-term split, K' = 384
    (50_000, 64, 48, 300, "bf16", 1),      # k' = 384: the deep-list tail behind the swapped kernel
])
def test_small_batch_swapped_kernel_matches_oracle(torch_cuda, n, d, nq, k, store, hint):
    """K1s (corpus rows on the M side of the MMA, the query block resident in shared memory) against the oracle, and against
    the pair kernel it replaces for 9..64 queries."""
    torch = torch_cuda
    from oracle import flat_ip
    from veritasfi_b200 import _native as N
    from veritasfi_b200.dense import DenseIndex
    xb, xq = _world(n, d, nq, 59, store == "bf16")
    xb[n // 2: n // 2 + 40] = xb[3]                      # duplicates: ties inside one tile
    xq[0] = xb[3]
    idx = DenseIndex(d, store=store)
    idx.add(xb)
    idx.set_option(N.OPT_FORCE_PATH, N.PATH_FUSED)
    idx.set_option(N.OPT_TAU_HINT, hint)
    q = torch.from_numpy(xq).cuda()
    ids, scores = idx.search_batch(q, k)
    D0, I0 = flat_ip.search(xq, xb, k)
    assert (ids.cpu().numpy() == I0).all()
    assert (scores.cpu().numpy() == D0).all()
    idx.set_option(N.OPT_SMALL_BATCH, 1)                 # the pair kernel on the same batch
    ids2, scores2 = idx.search_batch(q, k)
    assert torch.equal(ids, ids2) and torch.equal(scores, scores2)
    idx.close()


@pytest.mark.parametrize("store", ["bf16", "f32"])
def test_bf16_query_buffers_give_the_bits_of_fp32_buffers(torch_cuda, store):
    """vfi_index_search_ex with VFI_DTYPE_BF16 queries (half the H2D bytes) == the same values passed as fp32, on all paths."""
    torch = torch_cuda
    from oracle import flat_ip
    from veritasfi_b200.dense import DenseIndex
    xb, xq = _world(40_000, 192, 70, 61, True)           # queries already bf16-representable
    idx = DenseIndex(192, store=store)
    idx.add(xb)
    D0, I0 = flat_ip.search(xq, xb, 20)
    for nq in (70, 20, 3):                               # pair kernel, small-batch kernel, streaming GEMV
        q32 = torch.from_numpy(xq[:nq]).cuda()
        ids, scores = idx.search_batch(q32.to(torch.bfloat16), 20)
        assert (ids.cpu().numpy() == I0[:nq]).all() and (scores.cpu().numpy() == D0[:nq]).all()
        t = idx.search_begin(q32.to(torch.bfloat16), 20)
        i2, s2 = idx.search_finish(t)
        assert torch.equal(i2, ids) and torch.equal(s2, scores)
    idx.close()


def test_random_shapes_on_every_dense_path(torch_cuda):
    """A seeded sweep of ragged shapes (d not a multiple of 64, n not a multiple of the tile, k from 1 to 300, both stores)
    through whatever path the library picks: exact streaming, GEMV, small-batch kernel, pair kernel."""
    torch = torch_cuda
    from oracle import flat_ip
    from veritasfi_b200.dense import DenseIndex
    rng = np.random.default_rng(2026)
    seen = set()
    fixed = {14: (30_000, 768, 3, 10, "f32"), 15: (4_100, 48, 70, 5, "bf16"), 16: (9_000, 64, 4, 2000, "bf16")}   # exact streaming scorer
    for case in range(17):
        n = int(rng.integers(4_200, 60_000))
        d = int(rng.choice([33, 48, 64, 100, 200, 384, 768]))
        nq = int(rng.choice([1, 2, 5, 8, 9, 17, 40, 64, 65, 129, 200]))
        k = int(rng.choice([1, 3, 10, 64, 100, 257]))
        store = "bf16" if case % 2 == 0 else "f32"
        if case in fixed:
            n, d, nq, k, store = fixed[case]
        xb, xq = _world(n, d, nq, 700 + case, store == "bf16")
        idx = DenseIndex(d, store=store)
        idx.add(xb)
        ids, scores = idx.search_batch(torch.from_numpy(xq).cuda(), k)
        D0, I0 = flat_ip.search(xq, xb, k)
        assert (ids.cpu().numpy() == I0).all(), (case, n, d, nq, k, store)
        assert (scores.cpu().numpy() == D0).all(), (case, n, d, nq, k, store)
        seen.add(idx.stats().last_path)
        idx.close()
    assert seen == {1, 2, 3}


def test_global_ids_must_fit_32_bits(torch_cuda):
    from veritasfi_b200 import _native as N
    from veritasfi_b200.dense import DenseIndex
    idx = DenseIndex(64, store="bf16")
    idx.add(np.ones((10, 64), np.float32))
    with pytest.raises(N.VfiError):
        idx.set_id_offset(2 ** 32 - 5)
    idx.set_id_offset(2 ** 32 - 12)          # 10 rows still fit below 2^32 - 1
    with pytest.raises(N.VfiError):
        idx.add(np.ones((5, 64), np.float32))
    idx.close()


def test_concurrent_searches_on_one_index_equal_the_serial_run(torch_cuda):
    """The reference calls retriever.invoke from concurrent request threads with no lock (vllmChatService.py:85-88,404):
    eight threads hammer one index through the faiss facade, every shape class (exact streaming, GEMV, fused) mixed."""
    from veritasfi_b200 import faiss_compat
    xb, xq = _world(60_000, 128, 64, 51, False)
    index = faiss_compat.IndexFlatIP(128)
    index.add(xb)
    shapes = [(1, 10), (4, 2048), (33, 50), (3, 100), (64, 20), (2, 7), (16, 300), (8, 2048)]
    want = [index.search(xq[:nq], k) for nq, k in shapes]
    errors = []

    def worker(t):
        try:
            for rep in range(6):
                j = (t + rep) % len(shapes)
                nq, k = shapes[j]
                D, I = index.search(xq[:nq], k)
                if not ((I == want[j][1]).all() and (D == want[j][0]).all()):
                    errors.append((t, rep, j))
        except Exception as e:   # noqa: BLE001
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


# ------------------------------------------------------------------------------------------------ sparse
def _postings(n_docs, n_vocab, seed, mean_len=30):
    from veritasfi_b200 import synth
    from veritasfi_b200.bm25_compat import build_csc
    doc_ptr, toks = synth.zipf_postings(n_docs, n_vocab, seed, mean_len=mean_len)
    return build_csc(doc_ptr, toks, n_vocab)


def test_bm25_queries_longer_than_64_tokens(torch_cuda):
    """bm25s has no token limit; the reference sends paragraph-length rewritten queries (repeats count)."""
    from oracle import bm25 as obm
    from veritasfi_b200.bm25_compat import GpuPostings
    n_docs, n_vocab = 30_000, 900
    csc = _postings(n_docs, n_vocab, 61)
    rng = np.random.default_rng(7)
    qs = [rng.integers(0, n_vocab, size=t).astype(np.int32).tolist() for t in (65, 64, 200, 129, 3, 0, 500)]
    gp = GpuPostings(*csc, n_docs)
    ids, scores = gp.search(qs, 50)
    oi, os_ = obm.retrieve(*csc, qs, n_docs, 50)
    assert (ids == oi).all() and (scores == os_).all()
    s = gp.score_all(qs[2])
    assert (s == obm.scores(*csc, qs[2], n_docs)).all()


def test_bm25_top_2048_rank_range_and_the_lazy_k_equals_n_result(torch_cuda, tmp_path):
    from oracle import bm25 as obm
    from veritasfi_b200 import bm25_compat, synth
    from veritasfi_b200.bm25_compat import GpuPostings
    n_docs, n_vocab = 40_000, 3000
    csc = _postings(n_docs, n_vocab, 67)
    qs = synth.bm25_queries(6, n_vocab, 67)
    gp = GpuPostings(*csc, n_docs)
    ids, scores = gp.search(qs, 2048)
    oi, os_ = obm.retrieve(*csc, qs, n_docs, 2048)
    assert (ids == oi).all() and (scores == os_).all()
    fi, fs = obm.retrieve(*csc, qs[:2], n_docs, n_docs)          # every doc ranked
    for j in range(2):
        ri, rs = gp.rank_all(qs[j])
        assert (ri == fi[j]).all() and (rs == fs[j]).all()
        ti, ts = gp.rank_range(qs[j], 2048, 5000)
        assert (ti == fi[j][2048:7048]).all() and (ts == fs[j][2048:7048]).all()
    # facade: k = N like ensembleRetriever.py:189, read [:bm25_k] like :190 -> the tail is never produced
    eng = bm25_compat.BM25()
    eng.scores = {"data": csc[2], "indices": csc[1], "indptr": csc[0], "num_docs": n_docs}
    eng.corpus = [{"id": i, "text": str(i)} for i in range(n_docs)]
    docs, sc = eng.retrieve([qs[0]], k=n_docs, return_as="tuple")
    docs, sc = docs[0], sc[0]
    assert len(docs) == n_docs and len(sc) == n_docs
    assert [d["id"] for d in docs[:10]] == fi[0][:10].tolist() and (np.asarray(sc[:10]) == fs[0][:10]).all()
    assert eng.last_rankings[0].tail_reads == 0
    assert docs[3000]["id"] == fi[0][3000] and sc[n_docs - 1] == fs[0][-1]      # reading past the head produces the tail once
    assert eng.last_rankings[0].tail_reads == 1
    assert [d["id"] for d in docs] == fi[0].tolist()                              # the unmodified wrapper's list comprehension


def test_bm25_postings_adopted_from_device_memory(torch_cuda):
    torch = torch_cuda
    from oracle import bm25 as obm
    from veritasfi_b200 import _native as N, synth
    from veritasfi_b200.bm25_compat import GpuPostings
    n_docs, n_vocab = 9_000, 700
    csc = _postings(n_docs, n_vocab, 71)
    qs = synth.bm25_queries(9, n_vocab, 71)
    gp = GpuPostings.from_device(torch.from_numpy(csc[0]).cuda(), torch.from_numpy(csc[1]).cuda(), torch.from_numpy(csc[2]).cuda(),
                                 n_docs, id_offset=500)
    ids, scores = gp.search(qs, 30)
    oi, os_ = obm.retrieve(*csc, qs, n_docs, 30, id_base=500)
    assert (ids == oi).all() and (scores == os_).all()
    bad = csc[1].copy()
    bad[5], bad[6] = bad[6], bad[5]
    with pytest.raises(N.VfiError):
        GpuPostings.from_device(torch.from_numpy(csc[0]).cuda(), torch.from_numpy(bad).cuda(), torch.from_numpy(csc[2]).cuda(), n_docs)


# ------------------------------------------------------------------------------------------------ fusion
@pytest.mark.parametrize("layout", ["pbl", "bpl"])
def test_hybrid_fusion_kernel_equals_the_stagewise_oracle(torch_cuda, layout):
    torch = torch_cuda
    from oracle import fusion as ofu
    from veritasfi_b200.multipath import fuse_hybrid
    rng = np.random.default_rng(9)
    B, P, L, n, n_t, k = 37, 3, 200, 5000, 900, 50
    ids = np.stack([np.stack([rng.permutation(n if p != 1 else n_t)[:L] for p in range(P)]) for _ in range(B)]).astype(np.int64)
    scores = np.sort(rng.random((B, P, L)).astype(np.float32), axis=2)[:, :, ::-1].copy()
    t2c = rng.integers(0, 300, size=n_t).astype(np.int64)        # many titles per chunk: duplicates after the mapping
    for b in range(B):                                            # queries that matched fewer than L docs: zero-score filler
        m = int(rng.integers(0, L + 1))
        scores[b, 2, m:] = 0.0
        ids[b, 0, L - int(rng.integers(0, 20)):] = -1             # short dense lists (padding)
    want_i, want_s = ofu.hybrid(ids, scores, t2c, 1, 2, 60.0, k)
    ti, ts = torch.from_numpy(ids).cuda(), torch.from_numpy(scores).cuda()
    if layout == "pbl":
        ti, ts = ti.permute(1, 0, 2).contiguous(), ts.permute(1, 0, 2).contiguous()
    gi, gs = fuse_hybrid(ti, ts, torch.from_numpy(t2c).cuda(), 1, 2, k, 60.0, layout)
    assert (gi.cpu().numpy() == want_i).all() and (gs.cpu().numpy() == want_s).all()


# ------------------------------------------------------------------------------------------------ BASELINE sizes
def test_full_size_c5_shard_equals_the_oracle(torch_cuda):
    """One 1/8 shard of BASELINE config C5 (6.25M x 768 bf16, ONE query, top-10) on the streaming GEMV, equal to the CPU
    oracle streaming the shard in 1M-row blocks; global ids through the shard's id offset."""
    torch = torch_cuda
    from oracle import flat_ip
    from veritasfi_b200 import _native as N, synth
    from veritasfi_b200.dense import DenseIndex
    dev = torch.device("cuda", 0)
    n, d, k, off = 6_250_000, 768, 10, 3 * 6_250_000
    idx = DenseIndex(d, store="bf16", id_offset=off)
    idx.reserve(n)
    for c, r0 in enumerate(range(0, n, 1 << 20)):
        idx.add(synth.dense_corpus_torch(min(1 << 20, n - r0), d, 900 + c, dev))
    q = synth.dense_queries_torch(1, d, 901, dev)
    ids, scores = idx.search_batch(q, k)
    assert idx.stats().last_path == N.PATH_GEMV

    def blocks():
        for r0 in range(0, n, 1 << 20):
            yield r0 + off, idx.read_rows(r0, min(1 << 20, n - r0))
    D0, I0 = flat_ip.search_blocks(q.cpu().numpy(), blocks(), k, kc=64)
    assert (ids.cpu().numpy() == I0).all() and (scores.cpu().numpy() == D0).all()
    idx.close()


def test_full_size_c4_shard_bm25_and_fusion_equal_the_oracle(torch_cuda):
    """One 1/8 shard of BASELINE config C4 (625k docs, V = 262144, ~96 unique terms per doc, 6e7 postings built on the
    GPU): BM25 top-200 for 16 queries and the fused top-50 equal to the CPU oracle."""
    torch = torch_cuda
    from oracle import bm25 as obm, fusion as ofu
    from veritasfi_b200 import synth
    from veritasfi_b200.bm25_compat import GpuPostings
    from veritasfi_b200.multipath import fuse_hybrid
    dev = torch.device("cuda", 0)
    n, V, L, k, B = 625_000, 262_144, 200, 50, 16
    tok, doc, tf, dl = synth.zipf_postings_torch(n, V, 1001, dev, mean_len=128)
    df = torch.bincount(tok, minlength=V)
    avgdl = float(dl.sum()) / n
    indptr, indices, data = synth.bm25_impacts_torch(tok, doc, tf, dl, V, n, df, avgdl)
    assert 5.0e7 < indices.numel() < 6.5e7
    gp = GpuPostings.from_device(indptr, indices, data, n)
    qs = synth.bm25_queries(B, V, 1001)
    bi, bs = gp.search(qs, L)
    csc = (indptr.cpu().numpy(), indices.cpu().numpy(), data.cpu().numpy())
    oi, os_ = obm.retrieve(*csc, qs, n, L)
    assert (bi == oi).all() and (bs == os_).all()
    # fusion at the C4 list shapes: two synthetic dense lists + the BM25 list
    rng = np.random.default_rng(12)
    t2c = rng.integers(0, n, size=125_000).astype(np.int64)
    i0 = np.stack([rng.permutation(n)[:L] for _ in range(B)]).astype(np.int64)
    it = np.stack([rng.permutation(125_000)[:L] for _ in range(B)]).astype(np.int64)
    sc = np.sort(rng.random((B, L)).astype(np.float32), axis=1)[:, ::-1].copy()
    ids = np.stack([i0, it, oi], axis=1)
    scores = np.stack([sc, sc, os_], axis=1)
    want_i, want_s = ofu.hybrid(ids, scores, t2c, 1, 2, 60.0, k)
    gi, gs = fuse_hybrid(torch.from_numpy(ids).cuda(), torch.from_numpy(scores).cuda(), torch.from_numpy(t2c).cuda(), 1, 2, k, 60.0, "bpl")
    assert (gi.cpu().numpy() == want_i).all() and (gs.cpu().numpy() == want_s).all()
    gp.close()

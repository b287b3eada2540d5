"""GPU parity tests (run on the B200 box with `-m gpu`): every kernel family through the C ABI against the
CPU oracle on the same seeded inputs.  Bar: ids bit-exact, scores bit-exact (the canonical score is
reproducible), and within 1e-3 relative of the plain fp32 path (tolerance stated by BASELINE.json)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FLT_MAX = np.finfo(np.float32).max


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


def _oracle_inputs(xb, xq, store):
    from oracle import flat_ip
    if store == "bf16":
        return flat_ip.bf16_round(xb), flat_ip.bf16_round(xq)
    return xb, xq


# ------------------------------------------------------------------------------------------------ dense
def test_tcgen05_scores_match_fp64_matmul(torch_cuda):
    """The raw tensor-core scores (UMMA descriptors, TMA swizzle, TMEM epilogue) against an fp64 matmul."""
    torch = torch_cuda
    from veritasfi_b200.dense import DenseIndex
    for n, d, nq, store, tol in [(5000, 64, 5, "bf16", 5e-7), (33000, 1024, 257, "bf16", 2e-6), (9000, 100, 130, "f32", 4e-6)]:
        xb, xq = _world(n, d, nq, 1, store == "bf16")
        idx = DenseIndex(d, store=store)
        idx.add(xb)
        S = idx.debug_scores(torch.from_numpy(xq).cuda())
        ref = torch.from_numpy(xq).cuda().double() @ torch.from_numpy(xb).cuda().double().T
        assert (S.double() - ref).abs().max().item() < tol
        idx.close()


@pytest.mark.parametrize("n,d,nq,k,store", [
    (40000, 1024, 70, 100, "bf16"),      # k' = 128: four full warps of candidates per query
    (30000, 100, 33, 10, "f32"),         # fp32 rows, d padded to 128, k' = 32
    (20000, 64, 5, 200, "bf16"),         # k' = 256, one 128-byte row piece
    (9000, 192, 19, 37, "bf16"),         # k' = 64, row length not a multiple of 256 bytes
])
def test_tail_matches_oracle(torch_cuda, n, d, nq, k, store):
    """The selection + canonical rescoring + certificate tail (K1c + K2) is bit-exact for every candidate-buffer shape."""
    torch = torch_cuda
    from oracle import flat_ip
    from veritasfi_b200 import _native as N
    from veritasfi_b200.dense import DenseIndex
    xb, xq = _world(n, d, nq, 11, store == "bf16")
    idx = DenseIndex(d, store=store)
    idx.add(xb)
    idx.set_option(N.OPT_FORCE_PATH, N.PATH_FUSED)
    ids, scores = idx.search_batch(torch.from_numpy(xq).cuda(), k)
    ob, oq = _oracle_inputs(xb, xq, store)
    D0, I0 = flat_ip.search(oq, ob, k)
    assert (ids.cpu().numpy() == I0).all()
    assert (scores.cpu().numpy() == D0).all()
    idx.close()


@pytest.mark.parametrize("n,d,nq,k,store,path", [
    (30000, 128, 40, 10, "bf16", 2),
    (100000, 1024, 300, 100, "bf16", 2),     # 3 query tiles -> grouped CTAs share corpus tiles
    (20000, 100, 17, 50, "f32", 2),          # d not a multiple of 64, 3-term split operand
    (60000, 768, 9, 10, "bf16", 2),
    (50000, 768, 1, 10, "bf16", 3),          # latency mode: streaming scorer
    (30000, 1024, 8, 100, "bf16", 3),
    (20000, 100, 3, 10, "f32", 3),
    (3000, 64, 5, 20, "f32", 1),             # small shard: exhaustive exact path
    (70000, 256, 130, 100, "bf16", 0),       # automatic path selection
])
def test_dense_search_matches_oracle(torch_cuda, n, d, nq, k, store, path):
    torch = torch_cuda
    from oracle import flat_ip
    from veritasfi_b200 import _native as N
    from veritasfi_b200.dense import DenseIndex
    xb, xq = _world(n, d, nq, 3, store == "bf16")
    idx = DenseIndex(d, store=store)
    idx.add(xb)
    idx.set_option(N.OPT_FORCE_PATH, path)
    ids, scores = idx.search_batch(torch.from_numpy(xq).cuda(), k)
    ob, oq = _oracle_inputs(xb, xq, store)
    D0, I0 = flat_ip.search(oq, ob, k)
    assert (ids.cpu().numpy() == I0).all()
    assert (scores.cpu().numpy() == D0).all()
    st = idx.stats()
    if path:
        assert st.last_path == path
    # fp32-path tolerance of the north star: 1e-3 relative
    S32 = (torch.from_numpy(oq) @ torch.from_numpy(ob).T).numpy()
    ref = np.take_along_axis(S32, I0, axis=1)
    assert np.max(np.abs(scores.cpu().numpy() - ref) / np.maximum(np.abs(ref), 1e-6)) < 1e-3
    # the certificate's error model holds: tensor-core error is far below epsilon (~2*K*2^-24)
    if st.last_path in (2, 3):
        assert st.max_abs_err < 2.0 * idx.d * 3 * 2.0 ** -24
    idx.close()


def test_reference_online_depth_k2048(torch_cuda):
    """The reference's own online call: depth 2048 for 1-4 query strings (ensembleRetriever.py:64-66)."""
    from oracle import flat_ip
    from veritasfi_b200 import faiss_compat
    from veritasfi_b200 import _native as N
    xb, xq = _world(120000, 96, 4, 12, False)
    index = faiss_compat.IndexFlatIP(96)
    index.add(xb)
    for nq in (4, 1):
        D, I = index.search(xq[:nq], 2048)
        D0, I0 = flat_ip.search(xq[:nq], xb, 2048)
        assert (I == I0).all() and (D == D0).all()
        # k' = 2560 candidates do not fit the buffers of K1b: the exact streaming scorer (one pass over the corpus for
        # all queries, canonical scores, multi-CTA radix select) serves this call, not a per-query exhaustive sweep
        assert index.stats().last_path == N.PATH_EXACT
    xb2 = xb[:1500].copy()                       # corpus smaller than the depth: padded like faiss
    index2 = faiss_compat.IndexFlatIP(96)
    index2.add(xb2)
    D, I = index2.search(xq, 2048)
    D0, I0 = flat_ip.search_exhaustive(xq, xb2, 2048)
    assert (I == I0).all() and (D == D0).all() and (I[:, 1500:] == -1).all()


def test_admission_hint_does_not_change_results(torch_cuda):
    torch = torch_cuda
    from oracle import flat_ip
    from veritasfi_b200 import _native as N
    from veritasfi_b200.dense import DenseIndex
    xb, xq = _world(200000, 128, 150, 5, True)
    idx = DenseIndex(128, store="bf16")
    idx.add(xb)
    q = torch.from_numpy(xq).cuda()
    idx.set_option(N.OPT_FORCE_PATH, N.PATH_FUSED)
    idx.set_option(N.OPT_TAU_HINT, 0)
    i0, s0 = idx.search_batch(q, 20)
    idx.set_option(N.OPT_TAU_HINT, 1)            # the library default: threshold from the chunk maxima of a row sample
    i1, s1 = idx.search_batch(q, 20)
    assert torch.equal(i0, i1) and torch.equal(s0, s1)
    assert idx.stats().hint_retries == 0
    idx.set_option(N.OPT_TAU_HINT, 3)            # threshold from every sampled score
    i3, s3 = idx.search_batch(q, 20)
    assert torch.equal(i0, i3) and torch.equal(s0, s3)
    D0, I0 = flat_ip.search(xq, xb, 20)
    assert (i1.cpu().numpy() == I0).all() and (s1.cpu().numpy() == D0).all()
    idx.close()


@pytest.mark.parametrize("nq,pair,hint", [
    (256, 2, 1),     # one query tile pair: both epilogue sets own a buffer per query
    (256, 1, 1),     # the same batch on the single-CTA kernel
    (512, 2, 0),     # two tile pairs, no admission hint: thresholds rise by buffer compaction alone
    (1024, 2, 1),    # four tile pairs (the benchmark batch)
    (1000, 0, 1),    # ragged last tile, automatic choice (8 tiles -> pair kernel)
    (640, 0, 1),     # five tiles (odd): padded to six, three tile pairs
    (100, 0, 1),     # one ragged tile: padded to a pair whose second tile is empty
    (300, 2, 0),     # three tiles, padded to four
])
def test_cta_pair_kernel_matches_oracle_and_the_single_cta_kernel(torch_cuda, nq, pair, hint):
    """tcgen05 cta_group::2 path (two SMs share every corpus tile) against the oracle, for every query-tile-pair count."""
    torch = torch_cuda
    from oracle import flat_ip
    from veritasfi_b200 import _native as N
    from veritasfi_b200.dense import DenseIndex
    n, d, k = 150_001, 128, 100                      # n not a multiple of 256: ragged last corpus tile
    xb, xq = _world(n, d, nq, 17, True)
    idx = DenseIndex(d, store="bf16")
    idx.add(xb)
    idx.set_option(N.OPT_FORCE_PATH, N.PATH_FUSED)
    idx.set_option(N.OPT_CTA_PAIR, pair)
    idx.set_option(N.OPT_TAU_HINT, hint)
    ids, scores = idx.search_batch(torch.from_numpy(xq).cuda(), k)
    D0, I0 = flat_ip.search(xq, xb, k)
    assert (ids.cpu().numpy() == I0).all()
    assert (scores.cpu().numpy() == D0).all()
    assert idx.stats().last_path == N.PATH_FUSED
    idx.close()


@pytest.mark.parametrize("hint", [1, 2])
def test_pipelined_search_begin_finish_equals_the_synchronous_search(torch_cuda, hint):
    """Batches enqueued back to back with search_begin and collected later give the bits of the synchronous call, also
    when every certificate fails (hint = 2 admits nothing) and the repair of batch i runs after batch i+1 was launched."""
    torch = torch_cuda
    from oracle import flat_ip
    from veritasfi_b200 import _native as N
    from veritasfi_b200.dense import DenseIndex
    n, d, k = 30000, 256, 20
    xb, _ = _world(n, d, 4, 21, True)
    idx = DenseIndex(d, store="bf16")
    idx.add(xb)
    idx.set_option(N.OPT_FORCE_PATH, N.PATH_FUSED)
    idx.set_option(N.OPT_TAU_HINT, hint)
    from veritasfi_b200 import synth
    batches = [torch.from_numpy(synth.dense_queries_np(nq, d, 100 + j, xb)).cuda() for j, nq in enumerate([33, 130, 7, 64, 256])]
    want = [flat_ip.search(flat_ip.bf16_round(b.cpu().numpy()), flat_ip.bf16_round(xb), k) for b in batches]
    tickets = [idx.search_begin(b, k) for b in batches[:4]]            # four in flight, each in a workspace of its own
    got = [idx.search_finish(t) for t in tickets[:2]]
    tickets.append(idx.search_begin(batches[4], k))
    got += [idx.search_finish(t) for t in tickets[2:]]
    for (ids, scores), (D0, I0) in zip(got, want):
        assert (ids.cpu().numpy() == I0).all() and (scores.cpu().numpy() == D0).all()
    with pytest.raises(N.VfiError):
        idx.search_finish(tickets[0])                                 # already collected
    ids, scores = idx.search_batch(batches[1], k)                     # the synchronous call still works afterwards
    assert (ids.cpu().numpy() == want[1][1]).all()
    if hint == 2:
        assert idx.stats().retried_queries > 0
    idx.close()


def test_ties_duplicates_zero_rows_and_padding(torch_cuda):
    torch = torch_cuda
    from oracle import flat_ip
    from veritasfi_b200 import _native as N
    from veritasfi_b200.dense import DenseIndex
    rng = np.random.default_rng(9)
    n, d = 9000, 64
    xb = rng.integers(-2, 3, size=(n, d)).astype(np.float32)    # tiny integers: masses of exact score ties
    xb[100] = 0                                                    # a zero row
    xb[5000:5040] = xb[17]                                         # 40 exact duplicates of one row
    xq = rng.integers(-2, 3, size=(140, d)).astype(np.float32)
    xq[0] = xb[17]
    xq[1] = 0                                                      # zero query: every score ties at 0
    xb[200, 3] = np.float32(1e-39)                                 # an fp32 denormal (a bf16 denormal after rounding)
    xb[201, 5] = np.float32(-3e-40)
    for path in (1, 2, 3):
        idx = DenseIndex(d, store="bf16")
        idx.add(xb)
        idx.set_option(N.OPT_FORCE_PATH, path)
        nq = 140 if path != 3 else 8
        ids, scores = idx.search_batch(torch.from_numpy(xq[:nq]).cuda(), 50)
        D0, I0 = flat_ip.search_exhaustive(xq[:nq], xb, 50)
        assert (ids.cpu().numpy() == I0).all(), f"path {path}"
        assert (scores.cpu().numpy() == D0).all()
        assert I0[0, :41].tolist() == [17] + list(range(5000, 5040))   # duplicates in id order behind the original
        assert I0[1].tolist() == list(range(50))                        # all-zero scores: lowest ids win
        idx.close()
    # k larger than the shard: padded with -1 / -FLT_MAX like faiss
    idx = DenseIndex(d, store="f32")
    idx.add(xb[:30])
    ids, scores = idx.search_batch(torch.from_numpy(xq[:3]).cuda(), 64)
    assert (ids[:, 30:] == -1).all() and (scores[:, 30:] == -FLT_MAX).all()
    assert (ids[:, :30].sort(dim=1).values.cpu().numpy() == np.arange(30)).all()
    idx.close()


def test_certificate_failure_falls_back_to_the_exhaustive_pass(torch_cuda):
    """With the over-fetch forced down to k' = k rounded to 32 and 300 exact duplicates straddling the cut, the
    tensor-core candidate list cannot be certified; the flagged queries must be redone exactly."""
    torch = torch_cuda
    from oracle import flat_ip
    from veritasfi_b200 import _native as N
    from veritasfi_b200.dense import DenseIndex
    xb, xq = _world(20000, 64, 130, 4, True)
    xb[7000:7300] = xb[11]
    xq[0] = xb[11]
    idx = DenseIndex(64, store="bf16")
    idx.add(xb)
    idx.set_option(N.OPT_FORCE_PATH, N.PATH_FUSED)
    idx.set_option(N.OPT_OVERFETCH, 32)
    ids, scores = idx.search_batch(torch.from_numpy(xq).cuda(), 32)
    D0, I0 = flat_ip.search(xq, xb, 32)
    assert (ids.cpu().numpy() == I0).all() and (scores.cpu().numpy() == D0).all()
    assert idx.stats().retried_queries >= 1
    idx.close()


def test_incremental_add_id_offset_and_host_api(torch_cuda):
    from oracle import flat_ip
    from veritasfi_b200 import faiss_compat
    xb, xq = _world(12000, 96, 6, 8, False)
    index = faiss_compat.IndexFlatIP(96)
    for lo in range(0, 12000, 5000):                 # add() in pieces, like repeated index.add calls
        index.add(xb[lo:lo + 5000])
    assert index.ntotal == 12000 and index.d == 96
    D, I = index.search(xq, 10)
    D0, I0 = flat_ip.search(xq, xb, 10)
    assert D.dtype == np.float32 and I.dtype == np.int64 and D.shape == (6, 10)
    assert (I == I0).all() and (D == D0).all()
    assert (index.reconstruct(123) == xb[123]).all()
    with pytest.raises(AssertionError):
        index.search(xq[:, :50].copy(), 5)
    with pytest.raises(Exception):
        index.search(xq, 5000)                        # k > VFI_MAX_K
    y = xb[:100].copy() * 3.0
    y[7] = 0
    z = y.copy()
    faiss_compat.normalize_L2(z)
    assert (z == flat_ip.normalize_l2(y)).all() and (z[7] == 0).all()


def _blocks_of(index, n, block=1 << 20):
    """(first_row, fp32 rows) blocks read back from the index: what the oracle streams over at BASELINE sizes."""
    for r0 in range(0, n, block):
        yield r0, index.read_rows(r0, min(block, n - r0))


def test_full_size_c2_equals_the_oracle(torch_cuda):
    """BASELINE config C2 at full size (1M x 1024 bf16, 256 queries, top-100): ids and scores EQUAL to the CPU oracle
    (oracle.flat_ip.search: sgemm proposes, canonical rescoring decides, certificate) for every query, plus idempotence."""
    torch = torch_cuda
    from oracle import flat_ip
    from veritasfi_b200 import _native as N, synth
    from veritasfi_b200.dense import DenseIndex
    dev = torch.device("cuda", 0)
    n, d, b, k = 1_000_000, 1024, 256, 100
    xb = synth.dense_corpus_torch(n, d, 77, dev)
    q = synth.dense_queries_torch(b, d, 77, dev)
    q[:8] = xb[torch.arange(8, device=dev) * 1000 + 5].float()          # self-retrieval probes
    idx = DenseIndex(d, store="bf16")
    idx.add(xb)
    ids, scores = idx.search_batch(q, k)
    ids2, scores2 = idx.search_batch(q, k)
    assert torch.equal(ids, ids2) and torch.equal(scores, scores2)       # idempotent
    assert idx.stats().last_path == N.PATH_FUSED
    D0, I0 = flat_ip.search(q.cpu().numpy(), xb.float().cpu().numpy(), k)
    assert (ids.cpu().numpy() == I0).all()
    assert (scores.cpu().numpy() == D0).all()
    idx.close()


def test_full_size_c3_equals_the_oracle_on_a_query_subset(torch_cuda):
    """BASELINE config C3 on one GPU (10M x 1024 bf16, 1024 queries, top-100).
    (1) 16 of the 1024 queries EQUAL to the CPU oracle streaming the corpus in 1M-row blocks (oracle.flat_ip.search_blocks);
    (2) idempotence, total order, uniqueness, planted tie for the whole batch; (3) the exact streaming scorer agrees on a
    subset; (4) two half-corpus shards merged by the merge kernel equal the unsharded result bit for bit."""
    torch = torch_cuda
    from oracle import flat_ip
    from veritasfi_b200 import _native as N, synth
    from veritasfi_b200.dense import DenseIndex, merge_topk
    dev = torch.device("cuda", 0)
    free, _ = torch.cuda.mem_get_info(dev)
    if free < 70 * (1 << 30):
        pytest.skip("needs ~65 GB of free HBM")
    n, d, b, k, chunk = 10_000_000, 1024, 1024, 100, 1 << 20
    full = DenseIndex(d, store="bf16")
    halves = [DenseIndex(d, store="bf16", id_offset=0), DenseIndex(d, store="bf16", id_offset=n // 2)]
    full.reserve(n)
    for h in halves:
        h.reserve(n // 2)
    probes = []
    for c, r0 in enumerate(range(0, n, chunk)):
        rows = synth.dense_corpus_torch(min(chunk, n - r0), d, 500 + c, dev)
        if c == 0:
            rows[2 * chunk // 3] = rows[chunk // 3]               # exact duplicate inside the first chunk: a planted tie
        full.add(rows)
        cut = max(0, min(rows.shape[0], n // 2 - r0))              # rows [0, cut) belong to the first half
        if cut > 0:
            halves[0].add(rows[:cut])
        if cut < rows.shape[0]:
            halves[1].add(rows[cut:])
        if c in (0, 4, 9):
            probes.append((r0 + chunk // 3, rows[chunk // 3].float().clone()))
    q = synth.dense_queries_torch(b, d, 501, dev)
    for j, (_, v) in enumerate(probes):
        q[j] = v
    ids, scores = full.search_batch(q, k)
    ids2, scores2 = full.search_batch(q, k)
    assert full.stats().last_path == N.PATH_FUSED
    assert torch.equal(ids, ids2) and torch.equal(scores, scores2)
    s, i = scores.cpu().numpy(), ids.cpu().numpy()
    # (1) the oracle on 16 queries (the three probes, then a spread over the batch)
    pick16 = [0, 1, 2] + list(range(67, 1024, 74))[:13]
    D0, I0 = flat_ip.search_blocks(q[pick16].cpu().numpy(), _blocks_of(full, n), k)
    assert (i[pick16] == I0).all()
    assert (s[pick16] == D0).all()
    assert (i >= 0).all() and (i < n).all()
    ds = np.diff(s, axis=1)
    assert (ds <= 0).all() and (np.diff(i, axis=1)[ds == 0] > 0).all()
    assert all(len(set(r)) == k for r in i)
    for j, (row, _) in enumerate(probes):
        assert i[j, 0] <= row and s[j, 0] >= 0.99
    assert i[0, 0] == (1 << 20) // 3 and i[0, 1] == 2 * (1 << 20) // 3 and s[0, 0] == s[0, 1]   # the planted tie, lower id first
    # (3) the exact streaming scorer (every row scored canonically) on a subset of the queries
    sub = torch.cat([q[:4], q[500:504]])
    full.set_option(N.OPT_FORCE_PATH, N.PATH_EXACT)
    ie, se = full.search_batch(sub, k)
    full.set_option(N.OPT_FORCE_PATH, 0)
    pick = list(range(4)) + list(range(500, 504))
    assert torch.equal(ie, ids[pick]) and torch.equal(se, scores[pick])
    # (4) shard-merge == unsharded
    parts = [h.search_batch(q, k) for h in halves]
    gi = torch.stack([p[0] for p in parts])
    gs = torch.stack([p[1] for p in parts])
    mi, ms = merge_topk(gs, gi, k)
    assert torch.equal(mi, ids) and torch.equal(ms, scores)
    for x in [full] + halves:
        x.close()


# ------------------------------------------------------------------------------------------------ sparse
@pytest.mark.parametrize("n_docs,n_vocab,nq,k,mean_len", [(3000, 500, 20, 10, 40), (50000, 5000, 64, 50, 40), (2048 * 3 + 5, 300, 9, 100, 12)])
def test_bm25_matches_oracle(n_docs, n_vocab, nq, k, mean_len):
    from oracle import bm25 as obm, flat_ip
    from veritasfi_b200 import synth
    from veritasfi_b200.bm25_compat import GpuPostings, build_csc
    doc_ptr, toks = synth.zipf_postings(n_docs, n_vocab, 11, mean_len=mean_len)
    indptr, indices, data = build_csc(doc_ptr, toks, n_vocab)
    gp = GpuPostings(indptr, indices, data, n_docs)
    qs = synth.bm25_queries(nq, n_vocab, 11)
    qs[0] = []                                     # no known token: all scores 0 -> lowest ids
    qs[1] = [qs[1][0], qs[1][0]]                   # a repeated token adds twice
    qs[2] = [n_vocab - 1]                          # a rare token: fewer matches than k -> zero-score fill
    qs[3] = list(range(min(64, n_vocab)))          # the maximum query length
    I, S = gp.search(qs, k)
    I0, S0 = obm.retrieve(indptr, indices, data, qs, n_docs, k)
    assert (I == I0).all() and (S == S0).all()
    sa = gp.score_all(qs[4])
    ref = obm.scores(indptr, indices, data, qs[4], n_docs)
    assert (sa == ref).all()
    ri, rs = gp.rank_all(qs[4])
    s0, i0 = flat_ip.topk(ref, n_docs)
    assert (ri == i0).all() and (rs == s0).all()
    gp.close()


def test_bm25_general_impacts_and_doc_shards():
    """Non-positive impacts (other bm25s variants) disable tile skipping; doc-range shards with id offsets merge
    to the unsharded answer."""
    from oracle import bm25 as obm, sharded as osh
    from veritasfi_b200 import synth
    from veritasfi_b200.bm25_compat import GpuPostings, build_csc
    n_docs, n_vocab = 9000, 400
    doc_ptr, toks = synth.zipf_postings(n_docs, n_vocab, 2, mean_len=20)
    indptr, indices, data = build_csc(doc_ptr, toks, n_vocab)
    data2 = data.copy()
    data2[::7] *= -1
    gp = GpuPostings(indptr, indices, data2, n_docs)
    qs = synth.bm25_queries(12, n_vocab, 3)
    I, S = gp.search(qs, 30)
    I0, S0 = obm.retrieve(indptr, indices, data2, qs, n_docs, 30)
    assert (I == I0).all() and (S == S0).all()
    gp.close()
    parts_i, parts_s = [], []
    for lo, hi in osh.shard_bounds(n_docs, 3):
        keep = (indices >= lo) & (indices < hi)
        tok_of = np.repeat(np.arange(n_vocab), np.diff(indptr))
        ip = np.zeros(n_vocab + 1, np.int64)
        np.cumsum(np.bincount(tok_of[keep], minlength=n_vocab), out=ip[1:])
        g = GpuPostings(ip, indices[keep] - lo, data[keep], hi - lo, id_offset=lo)
        i, s = g.search(qs, 30)
        parts_i.append(i)
        parts_s.append(s)
        g.close()
    mi, ms = osh.merge(np.stack(parts_s), np.stack(parts_i), 30)
    I0, S0 = obm.retrieve(indptr, indices, data, qs, n_docs, 30)
    assert (mi == I0).all() and (ms == S0).all()


@pytest.mark.parametrize("k,n_docs,id_offset", [(64, 5000, 0), (200, 5000, 1000), (1000, 1500, 7), (32, 20, 0)])
def test_bm25_queries_matching_fewer_than_k_docs_are_filled_with_the_lowest_zero_score_ids(k, n_docs, id_offset, torch_cuda):
    """bm25s scores every doc, so a query that matches m < k docs gets the k - m lowest-id docs of score exactly 0 after
    them (total order: score desc, id asc).  Rare tokens + large k exercise that fill, host and device outputs."""
    torch = torch_cuda
    from oracle import bm25 as obm
    from veritasfi_b200.bm25_compat import GpuPostings
    rng = np.random.default_rng(k)
    n_vocab = 60
    # token t occurs in t+1 random docs (token 0 in one doc ... token 59 in sixty), positive impacts
    docs = [np.sort(rng.choice(n_docs, size=min(n_docs, t + 1), replace=False)) for t in range(n_vocab)]
    docs[3] = np.unique(np.concatenate([[0, 1], docs[3]]))               # low ids among the matches: the fill must skip them
    indptr = np.zeros(n_vocab + 1, np.int64)
    np.cumsum([len(x) for x in docs], out=indptr[1:])
    indices = np.concatenate(docs).astype(np.int32)
    data = rng.uniform(0.1, 3.0, size=len(indices)).astype(np.float32)
    qs = [[0], [3, 5], [59, 58, 57], [10, 10, 2], [], [1, 30, 44, 7]]
    kk = min(k, n_docs)
    I0, S0 = obm.retrieve(indptr, indices, data, qs, n_docs, kk)
    gp = GpuPostings(indptr, indices, data, n_docs, id_offset=id_offset)
    I, S = gp.search(qs, kk)
    assert (I == I0 + id_offset).all() and (S == S0).all()
    toks, qptr = GpuPostings.pack_tokens(qs)
    Id, Sd = gp.search_csr_device(toks, qptr, kk)
    assert (Id.cpu().numpy() == I0 + id_offset).all() and (Sd.cpu().numpy() == S0).all()
    assert (S0[:, -1] == 0).sum() >= 3                                   # several queries really needed the fill
    gp.close()


def test_bm25_rejects_malformed_postings():
    from veritasfi_b200 import _native as N
    from veritasfi_b200.bm25_compat import GpuPostings
    with pytest.raises(N.VfiError):
        GpuPostings(np.array([0, 2]), np.array([3, 1], np.int32), np.ones(2, np.float32), 5)      # not ascending
    with pytest.raises(N.VfiError):
        GpuPostings(np.array([0, 1]), np.array([9], np.int32), np.ones(1, np.float32), 5)         # out of range


def test_bm25_facade_retrieve_like_bm25s(tmp_path):
    from oracle import bm25 as obm
    from veritasfi_b200 import bm25_compat
    corpus = [f"doc {i} " + " ".join(w for w in ["alpha", "beta", "gamma", "delta", "epsilon"] if (i >> "abgde".index(w[0])) & 1) for i in range(1, 32)]
    eng = bm25_compat.BM25()
    eng.index(bm25_compat.tokenize(corpus))
    eng.save(str(tmp_path), corpus=[f"id{i}" for i in range(31)])
    eng = bm25_compat.BM25.load(str(tmp_path), load_corpus=True)
    docs, scores = eng.retrieve(bm25_compat.tokenize(["alpha gamma unknownword"]), k=31, return_as="tuple")
    ids = [d["id"] for d in docs[0]]
    s = eng.scores
    toks = eng.get_tokens_ids(["alpha", "gamma"])
    I0, S0 = obm.retrieve(s["indptr"], s["indices"], s["data"], [toks], 31, 31)
    assert ids == I0[0].tolist() and (scores[0] == S0[0]).all()
    with pytest.raises(ValueError):
        eng.retrieve(bm25_compat.tokenize(["alpha"]), k=32)


# ------------------------------------------------------------------------------------------------ fusion / merge
def test_rrf_union_merge_match_oracle(torch_cuda):
    torch = torch_cuda
    from oracle import fusion as ofu, sharded as osh
    from veritasfi_b200 import fusion as F
    from veritasfi_b200.dense import merge_topk
    rng = np.random.default_rng(5)
    for B, P, L, k in [(33, 3, 50, 20), (5, 3, 200, 50), (2, 1, 7, 10), (1, 4, 1000, 100)]:
        ids = np.stack([np.stack([rng.permutation(3 * L)[:L] for _ in range(P)]) for _ in range(B)]).astype(np.int64)
        ids[0, P - 1, L // 2:] = -1
        sc = rng.standard_normal((B, P, L)).astype(np.float32)
        fi, fs = F.rrf(ids, k)
        oi, os_ = ofu.rrf(ids, 60.0, k)
        assert (fi == oi).all() and (fs == os_).all()
        ti, ts = F.rrf(torch.from_numpy(ids).cuda(), k)              # device tensors in, device tensors out
        assert (ti.cpu().numpy() == oi).all() and (ts.cpu().numpy() == os_).all()
        ui, us, up, uc = F.union(ids, sc)
        vi, vs, vp, vc = ofu.union(ids, sc)
        assert (ui == vi).all() and (us == vs).all() and (up == vp).all() and (uc == vc).all()
    for G, B, k_in, k_out in [(8, 33, 100, 100), (2, 4, 10, 10), (4, 3, 2048, 100), (8, 2, 100, 2048)]:
        s = -np.sort(-rng.standard_normal((G, B, k_in)).astype(np.float32), axis=2)
        i = rng.permutation(G * B * k_in).reshape(G, B, k_in).astype(np.int64)
        i[G - 1, :, k_in - 3:] = -1
        s[0, 0, :4] = s[1, 0, 0]                                     # cross-shard score ties
        mi, ms = merge_topk(torch.from_numpy(s).cuda(), torch.from_numpy(i).cuda(), k_out)
        ri, rs = osh.merge(s, i, k_out)
        assert (mi.cpu().numpy() == ri).all() and (ms.cpu().numpy() == rs).all()


def test_cosine_topk_matches_the_experiment_scripts_semantics():
    """select_top_chunks (step3_mul.py:255-289): cosine_similarity then argsort()[-k:][::-1]."""
    from veritasfi_b200 import fusion as F
    rng = np.random.default_rng(1)
    c = rng.standard_normal((60, 32)).astype(np.float32)
    c[40] = c[3]                       # tie: argsort()[::-1] puts the HIGHER index first
    e = rng.standard_normal((4, 32)).astype(np.float32)
    e[0] = c[3] * 2.5
    ids, sims = F.cosine_topk(e, c, 5)
    cn = c / np.linalg.norm(c, axis=1, keepdims=True)
    en = e / np.linalg.norm(e, axis=1, keepdims=True)
    ref = en.astype(np.float64) @ cn.astype(np.float64).T
    assert ids[0, :2].tolist() == [40, 3]
    for j in range(4):
        want = [x for x in np.argsort(ref[j])[-5:][::-1]]
        if j:
            assert ids[j].tolist() == want
        assert np.allclose(sims[j], ref[j, ids[j]], atol=1e-6)


def test_multipath_batch_equals_stagewise_oracle(torch_cuda):
    torch = torch_cuda
    from oracle import bm25 as obm, flat_ip, fusion as ofu
    from veritasfi_b200 import synth
    from veritasfi_b200.bm25_compat import GpuPostings, build_csc
    from veritasfi_b200.dense import DenseIndex
    from veritasfi_b200.multipath import MultiPathRetriever
    n, n_ts, d, B, L, k = 20000, 6000, 128, 24, 40, 15
    xb, xq = _world(n, d, B, 31, True)
    xt = synth.dense_corpus_np(n_ts, d, 32)
    t2c = np.random.default_rng(3).integers(0, n, size=n_ts).astype(np.int64)
    doc_ptr, toks = synth.zipf_postings(n, 2000, 4, mean_len=30)
    csc = build_csc(doc_ptr, toks, 2000)
    qs = synth.bm25_queries(B, 2000, 4)
    chunks, titles = DenseIndex(d), DenseIndex(d)
    chunks.add(xb)
    titles.add(xt)
    mp = MultiPathRetriever(chunks, titles, torch.from_numpy(t2c).cuda(), GpuPostings(*csc, n), depth=L)
    q = torch.from_numpy(xq).cuda()
    fi, fs, _ = mp.multipath_batch(q, None, qs, k, fusion="rrf")
    # oracle, stage by stage
    D0, I0 = flat_ip.search(xq, xb, L)
    Dt, It = flat_ip.search(xq, xt, L)
    mapped = t2c[It]
    mi, ms, _, _ = ofu.union(mapped[:, None, :], Dt[:, None, :])
    Ib, Sb = obm.retrieve(*csc, qs, n, L)
    lists = np.stack([I0, mi, Ib], axis=1)
    oi, os_ = ofu.hybrid(np.stack([I0, It, Ib], axis=1), np.stack([D0, Dt, Sb], axis=1), t2c, 1, 2, 60.0, k)
    assert (fi.cpu().numpy() == oi).all() and (fs.cpu().numpy() == os_).all()
    ui, us, up = mp.multipath_batch(q, None, qs, k, fusion="union")
    vi, vs, vp, vc = ofu.union(lists, np.stack([D0, ms, Sb], axis=1))
    assert (ui.cpu().numpy() == vi[:, :k]).all() and (up.cpu().numpy() == vp[:, :k]).all()


# ------------------------------------------------------------------------------------------------ next rows
def test_shard_save_load_roundtrip_is_bit_exact(torch_cuda, tmp_path):
    torch = torch_cuda
    from veritasfi_b200.dense import DenseIndex
    for store in ("bf16", "f32"):
        xb, xq = _world(9000, 100, 6, 41, store == "bf16")
        a = DenseIndex(100, store=store)
        a.add(xb)
        ob, _ = _oracle_inputs(xb, xq, store)
        assert (a.read_rows(10, 500) == ob[10:510]).all()
        a.save(str(tmp_path / store))
        b = DenseIndex.load(str(tmp_path / store), id_offset=7)
        assert b.ntotal == 9000 and b.store == store
        q = torch.from_numpy(xq).cuda()
        ia, sa = a.search_batch(q, 25)
        ib, sb = b.search_batch(q, 25)
        assert torch.equal(ia + 7, ib) and torch.equal(sa, sb)


def test_pairwise_similarity_by_id_matches_oracle(torch_cuda):
    from oracle import flat_ip
    from veritasfi_b200.dense import DenseIndex
    rng = np.random.default_rng(3)
    x = flat_ip.normalize_l2(rng.standard_normal((500, 96)).astype(np.float32))
    idx = DenseIndex(96, store="f32")
    idx.add(x)
    ids = [7, 7, 499, 0, 123, 250]
    m = idx.pairwise(ids).cpu().numpy()
    for i, a in enumerate(ids):
        want = flat_ip.canon_scores(x[a], x, np.array(ids))
        assert (m[i] == want).all()
    assert np.allclose(np.diag(m), 1.0, atol=1e-6)


def test_mirror_similarity_helpers_match_the_reference_formula(tmp_path):
    """compute_similarity_mtx / compute_similarity (ensembleRetriever.py:235-281): normalise, then dot."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import fixture_world as fw
    import oracle_doubles as od
    from oracle import flat_ip
    from veritasfi_b200 import retrievers
    world = fw.make_world()
    od.write_bm25_dir(world, str(tmp_path))
    chroma, ts = fw.make_collections(world)
    emb = fw.FakeEmbeddings(world)
    r = retrievers.EnsembleRetriever(str(tmp_path), chroma, ts, 5, emb)
    chunks = ["near:7:1", "near:8:2", "near:7:3", "revenue profit"]
    m = r.compute_similarity_mtx(chunks).cpu().numpy()
    vecs = flat_ip.normalize_l2(np.array([emb.embed_query(c) for c in chunks], dtype=np.float32))
    ref = vecs.astype(np.float64) @ vecs.astype(np.float64).T
    assert m.shape == (4, 4) and np.abs(m - ref).max() < 1e-6
    for i in range(4):
        assert (m[i] == flat_ip.canon_scores(vecs[i], vecs, np.arange(4))).all()
    s = r.compute_similarity(chunks, [0, 1], 2).cpu().numpy()
    assert (s == m[[0, 1], 2]).all()
    by_row = r.similarity_mtx_by_row([3, 4, 5]).numpy()
    xb = flat_ip.normalize_l2(world["emb"])
    assert (by_row[0] == flat_ip.canon_scores(xb[3], xb, np.array([3, 4, 5]))).all()

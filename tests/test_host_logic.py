"""Host-side logic that needs no GPU: tokenisation, bm25s index construction and on-disk format,
result packing for the all-gather, shard bounds, synthetic generators, and the N>1 exchange path on gloo."""
import json
import os
import socket
import sys

import numpy as np
import torch

from oracle import bm25 as obm, sharded as osh
from veritasfi_b200 import bm25_compat, sharded, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tokenize_matches_bm25s_rules():
    t = bm25_compat.tokenize(["The quick brown fox, a fox!  It's 3 km to X-ray.", "fox AND the hound"], stopwords="english")
    rev = {i: w for w, i in t.vocab.items()}
    assert [rev[i] for i in t.ids[0]] == ["quick", "brown", "fox", "fox", "km", "ray"]   # 1-char tokens and stop words dropped
    assert [rev[i] for i in t.ids[1]] == ["fox", "hound"]
    assert t.vocab["quick"] == 0 and t.vocab["fox"] == 2                                  # first-seen order

    class Stem:
        def stemWords(self, ws):
            return [w.rstrip("s") for w in ws]
    t2 = bm25_compat.tokenize(["cats cat dogs"], stopwords=None, stemmer=Stem())
    assert t2.ids == [[0, 0, 1]] and t2.vocab == {"cat": 0, "dog": 1}
    assert bm25_compat.tokenize("hello world", return_ids=False) == [["hello", "world"]]


def test_build_csc_equals_plain_loop_oracle():
    doc_ptr, toks = synth.zipf_postings(400, 60, 5, mean_len=12)
    a = bm25_compat.build_csc(doc_ptr, toks, 60)
    docs = [toks[doc_ptr[i]:doc_ptr[i + 1]].tolist() for i in range(400)]
    b = obm.build_index(docs, 60)
    for x, y in zip(a, b):
        assert x.dtype == y.dtype and (x == y).all()
    assert (np.diff(a[0]) >= 0).all()
    for t in range(60):                      # doc ids ascending inside every posting list
        seg = a[1][a[0][t]:a[0][t + 1]]
        assert (np.diff(seg) > 0).all()


def test_bm25_save_load_roundtrip_uses_bm25s_file_layout(tmp_path):
    corpus = ["alpha beta gamma", "beta beta delta", "gamma epsilon", "zeta"]
    eng = bm25_compat.BM25()
    eng.index(bm25_compat.tokenize(corpus))
    eng.save(str(tmp_path), corpus=["id-a", "id-b", "id-c", "id-d"])
    names = sorted(os.listdir(tmp_path))
    assert names == ["corpus.jsonl", "corpus.mmindex.json", "data.csc.index.npy", "indices.csc.index.npy",
                     "indptr.csc.index.npy", "params.index.json", "vocab.index.json"]
    eng2 = bm25_compat.BM25.load(str(tmp_path), load_corpus=True)
    assert eng2.corpus == [{"id": i, "text": t} for i, t in enumerate(["id-a", "id-b", "id-c", "id-d"])]
    for k in ("data", "indices", "indptr"):
        assert (eng2.scores[k] == eng.scores[k]).all()
    assert eng2.vocab_dict == eng.vocab_dict
    assert json.load(open(tmp_path / "params.index.json"))["num_docs"] == 4
    assert eng2.get_tokens_ids(["beta", "nope", "zeta"]) == [eng.vocab_dict["beta"], eng.vocab_dict["zeta"]]


def test_pack_unpack_roundtrip_odd_and_even_k():
    for k in (1, 7, 100):
        s = torch.randn(5, k)
        i = torch.randint(-1, 10 ** 9, (5, k), dtype=torch.int64)
        buf = sharded.pack(s, i)
        assert buf.dtype == torch.int64 and buf.shape == (5, k + (k + 1) // 2)
        s2, i2 = sharded.unpack(torch.stack([buf, buf]), k)
        assert torch.equal(s2[1], s) and torch.equal(i2[0], i)


def test_shard_bounds_cover_rows_exactly_and_match_oracle():
    for n, g in [(10, 3), (10_000_000, 8), (7, 8), (0, 2)]:
        bounds = [sharded.shard_bounds(n, g, r) for r in range(g)]
        assert bounds == osh.shard_bounds(n, g)
        assert bounds[0][0] == 0 and bounds[-1][1] == n
        assert all(b[1] == c[0] for b, c in zip(bounds, bounds[1:]))


def test_synthetic_generators_are_seeded_and_plant_duplicates():
    a = synth.dense_corpus_np(5000, 32, 9)
    b = synth.dense_corpus_np(5000, 32, 9)
    assert (a == b).all()
    assert (a.view(np.uint32) & 0xFFFF == 0).all()        # values are bf16-representable
    assert len(np.unique(a, axis=0)) < 5000                # exact duplicate rows exist
    q = synth.dense_queries_np(16, 32, 9, a)
    assert abs(np.linalg.norm(q, axis=1) - 1).max() < 2e-2
    qs = synth.bm25_queries(50, 1000, 1)
    assert all(4 <= len(t) <= 16 for t in qs)


def _gloo_worker(rank, world, port, n, d, k, out_path):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from oracle import flat_ip as fi, sharded as so
    from veritasfi_b200 import sharded as sh
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(21)
    xb = fi.normalize_l2(rng.standard_normal((n, d)).astype(np.float32))
    xb[n - 1] = xb[3]                                   # a duplicate living on another shard
    xq = fi.normalize_l2(rng.standard_normal((5, d)).astype(np.float32))
    xq[0] = xb[3]
    lo, hi = sh.shard_bounds(n, world, rank)

    def local(q, kk):                                   # the CPU oracle stands in for the shard's CUDA searcher
        D, I = fi.search_exhaustive(q.numpy(), xb[lo:hi], kk, id_base=lo)
        return torch.from_numpy(I), torch.from_numpy(D)

    def merge(s, i, kk):
        oi, os_ = so.merge(s.numpy(), i.numpy(), kk)
        return torch.from_numpy(oi), torch.from_numpy(os_)

    searcher = sh.ShardedSearcher(local, merge)
    ids, scores = searcher.search(torch.from_numpy(xq), k)
    D0, I0 = fi.search_exhaustive(xq, xb, k)
    ok = bool((ids.numpy() == I0).all() and (scores.numpy() == D0).all())
    # the hybrid retriever's shape: three per-rank lists travel as 3*B independent rows in ONE exchange (exchange_rows)
    B, L = 4, 6
    rows_s, rows_i = [], []
    for p in range(3):
        D, I = fi.search_exhaustive(np.roll(xq[:B], p, axis=1), xb[lo:hi], L, id_base=lo)
        rows_s.append(D)
        rows_i.append(I)
    mi, ms = searcher.exchange_rows(torch.from_numpy(np.concatenate(rows_s)), torch.from_numpy(np.concatenate(rows_i)), L)
    for p in range(3):
        D, I = fi.search_exhaustive(np.roll(xq[:B], p, axis=1), xb, L)
        ok &= bool((mi.numpy()[p * B:(p + 1) * B] == I).all() and (ms.numpy()[p * B:(p + 1) * B] == D).all())
    with open(f"{out_path}.{rank}", "w") as f:
        f.write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


def _gloo_fused_worker(rank, world, port, out_path):
    """The control flow of the pipelined sharded search whose local search pushes its rows (ShardedSearcher.search_begin /
    search_finish on the fused route) with CPU stand-ins for the index and the peer exchange: tickets, the any-rank-failed bit,
    the repeat exchange after a repair on ONE rank."""
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from oracle import flat_ip as fi, sharded as so
    from veritasfi_b200 import sharded as sh
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, d, k = 257, 16, 7
    rng = np.random.default_rng(5)
    xb = fi.normalize_l2(rng.standard_normal((n, d)).astype(np.float32))
    xq = fi.normalize_l2(rng.standard_normal((6, d)).astype(np.float32))
    lo, hi = sh.shard_bounds(n, world, rank)

    class Ticket:
        def __init__(self, ids, scores, broken):
            self.ids, self.scores, self.broken = ids, scores, broken

    class FakeIndex:                      # DenseIndex's pipelined surface; batch 1 of rank 0 "fails its certificate"
        def __init__(self):
            self.begun = 0

        def search_begin(self, q, kk, out=None, push=None):
            D, I = fi.search_exhaustive(q.numpy(), xb[lo:hi], kk, id_base=lo)
            broken = rank == 0 and self.begun == 1
            self.begun += 1
            if broken:                    # rows that are not final yet: what finish() will repair
                D = D.copy()
                D[:, 0] = -1.0
            t = Ticket(torch.from_numpy(I), torch.from_numpy(D), broken)
            if push is not None:
                push.pushed_rows = (t.scores.clone(), t.ids.clone(), broken)
            return t

        def search_finish(self, t):
            if t.broken:                  # the repair
                D, I = fi.search_exhaustive(xq, xb[lo:hi], k, id_base=lo)
                t.ids, t.scores = torch.from_numpy(I), torch.from_numpy(D)
            return t.ids, t.scores

        def ticket_flag_ptr(self, t):
            return 0

    class FakeExchange:                   # PeerExchange's surface over gloo collectives
        max_nq, max_k = 64, 64

        def __init__(self):
            self.flags, self.pushed_rows = [], None

        def _gather(self, scores, ids, fail):
            gs = [torch.empty_like(scores) for _ in range(world)]
            gi = [torch.empty_like(ids) for _ in range(world)]
            dist.all_gather(gs, scores.contiguous())
            dist.all_gather(gi, ids.contiguous())
            f = torch.tensor([1 if fail else 0])
            dist.all_reduce(f, op=dist.ReduceOp.MAX)
            return torch.stack(gs), torch.stack(gi), bool(f.item())

        def merge_pushed(self, B, kk, k_out, fail_ptr=None):
            scores, ids, broken = self.pushed_rows
            gs, gi, any_fail = self._gather(scores, ids, broken)
            oi, os_ = so.merge(gs.numpy(), gi.numpy(), k_out)
            self.flags.append(any_fail)
            return torch.from_numpy(oi), torch.from_numpy(os_), len(self.flags) - 1

        def any_fail(self, slot):
            return self.flags[slot]

        def merge(self, scores, ids, k_out, fail_ptrs=(None, None), any_fail=None, out=None):
            gs, gi, _ = self._gather(scores, ids, False)
            oi, os_ = so.merge(gs.numpy(), gi.numpy(), k_out)
            return torch.from_numpy(oi), torch.from_numpy(os_)

    s = sh.ShardedSearcher(None, None, exchange=FakeExchange(), index=FakeIndex())
    q = torch.from_numpy(xq)
    D0, I0 = fi.search_exhaustive(xq, xb, k)
    tickets = [s.search_begin(q, k) for _ in range(3)]
    ok = all(t.pushed for t in tickets)
    for t in tickets:
        ids, scores = s.search_finish(t)
        ok &= bool((ids.numpy() == I0).all() and (scores.numpy() == D0).all())
    ok &= s.re_exchanges == 1             # on BOTH ranks, though only rank 0 repaired
    with open(f"{out_path}.{rank}", "w") as f:
        f.write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_fused_push_control_flow_and_repeat_exchange(tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "res")
    mp.spawn(_gloo_fused_worker, args=(2, port, out), nprocs=2, join=True)
    assert open(out + ".0").read() == "ok" and open(out + ".1").read() == "ok"


def test_world_size_2_gloo_allgather_merge_equals_unsharded(tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "res")
    mp.spawn(_gloo_worker, args=(2, port, 301, 24, 9, out), nprocs=2, join=True)
    assert open(out + ".0").read() == "ok" and open(out + ".1").read() == "ok"


def test_bench_reference_arm_prints_the_contract_line_and_the_gpu_arm_refuses_to_run_without_a_device():
    """`bench.py --impl reference` times the CPU port of the reference's path (no GPU needed) and prints ONE JSON line with
    the keys the driver reads; the default arm has no CPU fallback."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "small", "--steps", "2",
                        "--warmup", "0", "--cpu-rows", "20000"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "queries/sec" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 0 and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    import torch
    if not torch.cuda.is_available():
        p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--workload", "small", "--steps", "1"], capture_output=True,
                           text=True, timeout=300)
        assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)

"""Staged bring-up checks for the GPU box: each stage runs in its own process under a timeout so that a
trap or a hang in one kernel does not hide the others.  Usage: python tools/gpu_first_light.py [stage ...]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STAGES = {}


def stage(fn):
    STAGES[fn.__name__] = fn
    return fn


def _mk(n, d, nq, seed=0, bf16=True):
    import numpy as np
    from veritasfi_b200 import synth
    xb = synth.dense_corpus_np(n, d, seed, bf16=bf16)
    xq = synth.dense_queries_np(nq, d, seed, xb, bf16=bf16)
    return xb, xq


def _cmp(tag, I, D, I0, D0):
    import numpy as np
    ok_i = bool((I == I0).all())
    rel = float(np.max(np.abs(D - D0) / np.maximum(np.abs(D0), 1e-30))) if D.size else 0.0
    print(f"[{tag}] ids_equal={ok_i} mismatches={int((I != I0).sum())}/{I.size} max_rel_score_err={rel:.3e} scores_bitexact={bool((D == D0).all())}", flush=True)
    return ok_i


@stage
def simple():
    import numpy as np
    from oracle import flat_ip
    from veritasfi_b200 import faiss_compat, _native as N
    x = np.random.default_rng(1).standard_normal((3000, 100)).astype(np.float32)
    y = x.copy()
    faiss_compat.normalize_L2(y)
    y0 = flat_ip.normalize_l2(x)
    print("[normalize] bitexact", bool((y == y0).all()), flush=True)
    for store in ("f32", "bf16"):
        idx = faiss_compat.IndexFlatIP(100, store=store)
        idx.add(y)
        q = y[:7].copy()
        D, I = idx.search(q, 10)
        ref_x = y if store == "f32" else flat_ip.bf16_round(y)
        ref_q = q if store == "f32" else flat_ip.bf16_round(q)
        D0, I0 = flat_ip.search_exhaustive(ref_q, ref_x, 10)
        _cmp(f"exhaustive {store}", I, D, I0, D0)
        D, I = idx.search(q[:2], 2048)  # k > n' -> padding only beyond n (n=3000 > 2048: none) 
        print("[k=2048] path", idx.stats().last_path, bool((I >= 0).all()), flush=True)


@stage
def gemm():
    import numpy as np
    import torch
    from veritasfi_b200.dense import DenseIndex
    for (n, d, nq, store) in [(5000, 64, 5, "bf16"), (70000, 1024, 200, "bf16"), (9000, 128, 130, "f32"), (4100, 100, 3, "f32")]:
        xb, xq = _mk(n, d, nq, bf16=(store == "bf16"))
        idx = DenseIndex(d, store=store)
        idx.add(xb)
        q = torch.from_numpy(xq).cuda()
        t0 = time.time()
        S = idx.debug_scores(q)
        torch.cuda.synchronize()
        ref = (torch.from_numpy(xq).cuda().double() @ torch.from_numpy(xb).cuda().double().T).float()
        err = (S - ref).abs().max().item()
        print(f"[gemm n={n} d={d} nq={nq} {store}] max_abs_err={err:.3e} ({time.time()-t0:.2f}s)", flush=True)
        idx.close()


@stage
def fused():
    import numpy as np
    import torch
    from oracle import flat_ip
    from veritasfi_b200.dense import DenseIndex
    from veritasfi_b200 import _native as N
    for (n, d, nq, k, store) in [(30000, 128, 40, 10, "bf16"), (100000, 1024, 300, 100, "bf16"), (20000, 100, 17, 50, "f32"), (60000, 768, 9, 10, "bf16")]:
        xb, xq = _mk(n, d, nq, seed=3, bf16=(store == "bf16"))
        idx = DenseIndex(d, store=store)
        idx.add(xb)
        idx.set_option(N.OPT_FORCE_PATH, N.PATH_FUSED)
        q = torch.from_numpy(xq).cuda()
        t0 = time.time()
        I, D = idx.search_batch(q, k)
        torch.cuda.synchronize()
        dt = time.time() - t0
        D0, I0 = flat_ip.search(xq, xb, k)
        _cmp(f"fused n={n} d={d} nq={nq} k={k} {store}", I.cpu().numpy(), D.cpu().numpy(), I0, D0)
        st = idx.stats()
        print(f"   path={st.last_path} keep={st.last_overfetch} retried={st.retried_queries} max_abs_err={st.max_abs_err:.3e} t={dt:.3f}s", flush=True)
        idx.close()


@stage
def gemv():
    import torch
    from oracle import flat_ip
    from veritasfi_b200.dense import DenseIndex
    from veritasfi_b200 import _native as N
    for (n, d, nq, k, store) in [(50000, 768, 1, 10, "bf16"), (30000, 1024, 8, 100, "bf16"), (20000, 100, 3, 10, "f32")]:
        xb, xq = _mk(n, d, nq, seed=5, bf16=(store == "bf16"))
        idx = DenseIndex(d, store=store)
        idx.add(xb)
        idx.set_option(N.OPT_FORCE_PATH, N.PATH_GEMV)
        I, D = idx.search_batch(torch.from_numpy(xq).cuda(), k)
        D0, I0 = flat_ip.search(xq, xb, k)
        _cmp(f"gemv n={n} d={d} nq={nq} k={k} {store}", I.cpu().numpy(), D.cpu().numpy(), I0, D0)
        st = idx.stats()
        print(f"   path={st.last_path} keep={st.last_overfetch} retried={st.retried_queries} max_abs_err={st.max_abs_err:.3e}", flush=True)
        idx.close()


@stage
def bm25():
    import numpy as np
    from oracle import bm25 as obm
    from veritasfi_b200 import synth
    from veritasfi_b200.bm25_compat import GpuPostings, build_csc
    for (n_docs, n_vocab, nq, k) in [(3000, 500, 20, 10), (50000, 5000, 64, 50)]:
        doc_ptr, toks = synth.zipf_postings(n_docs, n_vocab, 11, mean_len=40)
        indptr, indices, data = build_csc(doc_ptr, toks, n_vocab)
        gp = GpuPostings(indptr, indices, data, n_docs)
        qs = synth.bm25_queries(nq, n_vocab, 11)
        qs[0] = []          # empty query
        qs[1] = [qs[1][0], qs[1][0]]  # repeated token
        I, S = gp.search(qs, k)
        I0, S0 = obm.retrieve(indptr, indices, data, qs, n_docs, k)
        _cmp(f"bm25 n={n_docs} V={n_vocab} nq={nq} k={k}", I, S, I0, S0)
        sa = gp.score_all(qs[2])
        print("   score_all bitexact", bool((sa == obm.scores(indptr, indices, data, qs[2], n_docs)).all()), flush=True)
        ri, rs = gp.rank_all(qs[2])
        s0, i0 = __import__("oracle").flat_ip.topk(obm.scores(indptr, indices, data, qs[2], n_docs), n_docs)
        print("   rank_all ids", bool((ri == i0).all()), "scores", bool((rs == s0).all()), flush=True)
        gp.close()


@stage
def fuse():
    import numpy as np
    from oracle import fusion as ofu, sharded as osh
    from veritasfi_b200 import fusion as F
    import torch
    from veritasfi_b200.dense import merge_topk
    rng = np.random.default_rng(5)
    B, P, L = 33, 3, 50
    ids = rng.integers(0, 300, size=(B, P, L)).astype(np.int64)
    for b in range(B):          # ranked lists have distinct ids per path
        for p in range(P):
            ids[b, p] = rng.permutation(400)[:L]
    ids[0, 1, 40:] = -1
    sc = rng.standard_normal((B, P, L)).astype(np.float32)
    fi, fs = F.rrf(ids, 20)
    oi, os_ = ofu.rrf(ids, 60.0, 20)
    _cmp("rrf", fi, fs, oi, os_)
    ui, us, up, uc = F.union(ids, sc)
    vi, vs, vp, vc = ofu.union(ids, sc)
    print("[union] ids", bool((ui == vi).all()), "scores", bool((us == vs).all()), "path", bool((up == vp).all()), "count", bool((uc == vc).all()), flush=True)
    G, k = 8, 100
    s = np.sort(rng.standard_normal((G, B, k)).astype(np.float32), axis=2)[:, :, ::-1].copy()
    i = rng.permutation(G * B * k).reshape(G, B, k).astype(np.int64)
    i[3, :, 90:] = -1
    mi, ms = merge_topk(torch.from_numpy(s).cuda(), torch.from_numpy(i).cuda(), k)
    ri, rs = osh.merge(s, i, k)
    _cmp("merge", mi.cpu().numpy(), ms.cpu().numpy(), ri, rs)


@stage
def cluster():
    """TMA-multicast clusters: same results, timing per cluster size."""
    import torch
    from veritasfi_b200 import synth, _native as N
    from veritasfi_b200.dense import DenseIndex
    dev = torch.device("cuda", 0)
    for (n, d, nq, k) in [(1_000_000, 1024, 256, 100), (4_000_000, 1024, 1024, 100)]:
        xb = synth.dense_corpus_torch(n, d, 1235, dev)
        idx = DenseIndex(d, store="bf16")
        idx.add(xb)
        del xb
        q = synth.dense_queries_torch(nq, d, 1235, dev)
        idx.set_option(N.OPT_PROFILE, 1)
        idx.set_option(N.OPT_TAU_HINT, 1)
        ref = None
        for c in (1, 2, 4, 8):
            if c > max(1, nq // 128):
                continue
            idx.set_option(N.OPT_CLUSTER, c)
            try:
                for _ in range(3):
                    I, D = idx.search_batch(q, k)
                torch.cuda.synchronize()
            except Exception as e:
                print(f"[cluster n={n} nq={nq} C={c}] FAILED: {e}", flush=True)
                break
            idx.stats(reset=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(8):
                I, D = idx.search_batch(q, k)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 8
            st = idx.stats()
            kms = st.fused_ms_total / max(1, st.fused_ms_samples)
            tf = 2.0 * nq * n * d / (kms * 1e-3) / 1e12
            if ref is None:
                ref = (I.clone(), D.clone())
            same = bool((I == ref[0]).all() and (D == ref[1]).all())
            print(f"[cluster n={n} nq={nq} C={c}] step={ms:.3f} ms fused_kernel={kms:.3f} ms ({tf:.0f} TFLOP/s) same_as_C1={same} retried={st.retried_queries}", flush=True)
        idx.close()


@stage
def perf():
    import torch
    from veritasfi_b200 import synth, _native as N
    from veritasfi_b200.dense import DenseIndex
    dev = torch.device("cuda", 0)
    for (n, d, nq, k) in [(1_000_000, 1024, 256, 100), (2_000_000, 1024, 1024, 100)]:
        xb = synth.dense_corpus_torch(n, d, 1235, dev)
        idx = DenseIndex(d, store="bf16")
        idx.add(xb)
        del xb
        q = synth.dense_queries_torch(nq, d, 1235, dev)
        idx.set_option(N.OPT_PROFILE, 1)
        ref = None
        for hint in (0, 1, 2):
            idx.set_option(N.OPT_TAU_HINT, hint)
            for _ in range(3):
                I, D = idx.search_batch(q, k)
            torch.cuda.synchronize()
            idx.stats(reset=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                I, D = idx.search_batch(q, k)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            st = idx.stats()
            kms = st.fused_ms_total / max(1, st.fused_ms_samples)
            tf = 2.0 * nq * n * d / (kms * 1e-3) / 1e12
            same = None
            if hint == 0:
                ref = (I.clone(), D.clone())
            elif hint == 1:
                same = bool((I == ref[0]).all() and (D == ref[1]).all())
            print(f"[perf n={n} nq={nq} hint={hint}] step={ms:.3f} ms  qps={nq/ms*1e3:.0f}  fused_kernel={kms:.3f} ms ({tf:.0f} TFLOP/s) "
                  f"retried={st.retried_queries} hint_retries={st.hint_retries} same_as_nohint={same}", flush=True)
        idx.close()


def main():
    names = sys.argv[1:] or list(STAGES)
    if os.environ.get("VFI_STAGE"):
        STAGES[os.environ["VFI_STAGE"]]()
        return
    rc = 0
    for name in names:
        print(f"===== stage {name} =====", flush=True)
        env = dict(os.environ, VFI_STAGE=name)
        try:
            p = subprocess.run([sys.executable, __file__], env=env, timeout=240)
            print(f"===== stage {name} exit {p.returncode} =====", flush=True)
            rc |= p.returncode != 0
        except subprocess.TimeoutExpired:
            print(f"===== stage {name} TIMEOUT =====", flush=True)
            rc = 1
    sys.exit(rc)


if __name__ == "__main__":
    main()

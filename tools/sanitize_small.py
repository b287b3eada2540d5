"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck): tiny shapes, all paths."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from veritasfi_b200 import _native as N, synth, fusion as F
from veritasfi_b200.dense import DenseIndex, merge_topk
from veritasfi_b200.bm25_compat import GpuPostings, build_csc

for (n, d, nq, k, store, path, hint) in [(6000, 128, 130, 20, "bf16", 2, 0), (6000, 100, 9, 20, "f32", 2, 0), (40000, 64, 40, 10, "bf16", 2, 1),
                                         (6000, 128, 3, 10, "bf16", 3, 0), (3000, 64, 5, 40, "f32", 1, 0)]:
    xb = synth.dense_corpus_np(n, d, 1, bf16=(store == "bf16"))
    xq = synth.dense_queries_np(nq, d, 1, xb, bf16=(store == "bf16"))
    idx = DenseIndex(d, store=store)
    idx.add(xb)
    idx.set_option(N.OPT_FORCE_PATH, path)
    idx.set_option(N.OPT_TAU_HINT, hint)
    ids, sc = idx.search_batch(torch.from_numpy(xq).cuda(), k)
    torch.cuda.synchronize()
    print("dense", n, d, nq, store, path, int(ids.sum()))
    idx.close()
doc_ptr, toks = synth.zipf_postings(20000, 900, 3, mean_len=20)
gp = GpuPostings(*build_csc(doc_ptr, toks, 900), 20000)
qs = synth.bm25_queries(12, 900, 3)
qs[0] = []
print("bm25", gp.search(qs, 10)[0].sum(), gp.score_all(qs[1]).sum(), gp.rank_all(qs[2])[0][:3])
ids = np.stack([np.stack([np.random.default_rng(b * 3 + p).permutation(90)[:30] for p in range(3)]) for b in range(5)]).astype(np.int64)
print("rrf", F.rrf(ids, 10)[0].sum(), "union", F.union(ids, np.zeros(ids.shape, np.float32))[3].sum())
s = torch.randn(4, 6, 20).cuda(); i = torch.randperm(480).reshape(4, 6, 20).cuda()
print("merge", merge_topk(s, i, 20)[0].sum().item())

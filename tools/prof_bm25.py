import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from veritasfi_b200 import synth
from veritasfi_b200.bm25_compat import GpuPostings, build_csc
n_docs, V, nq = 400_000, 262_144, 512
doc_ptr, toks = synth.zipf_postings(n_docs, V, 4000, mean_len=128)
gp = GpuPostings(*build_csc(doc_ptr, toks, V), n_docs)
qs = synth.bm25_queries(nq, V, 4000)
for _ in range(3):
    gp.search(qs, 50)
print("done")

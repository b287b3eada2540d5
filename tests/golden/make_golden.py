"""Generate tests/golden/ensemble_golden.json by running the UNMODIFIED reference classes
(/root/reference/src/utils/{faissRetriever,bm25Retriever,ensembleRetriever}.py) over tests/fixture_world.py.

The reference's third-party dependencies are absent from this image (faiss, bm25s, PyStemmer, langchain_*),
so they are replaced by shims: `faiss` and the scoring half of `bm25s` are backed by the CPU ORACLE (oracle/),
tokenisation by the host tokenizer of veritasfi_b200.bm25_compat, PyStemmer by the oracle's Snowball-English
restatement (oracle/porter2.py).  What the
fixture pins is therefore the reference's own orchestration — depth-2048 dense search, shared seen_ids
de-duplication, bundle gathering, prev/next expansion at 0.72/0.66, title-summary mapping, BM25 k = N then
slice — on top of the canonical arithmetic.  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py
"""
import json
import os
import sys
import tempfile


HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fixture_world as fw  # noqa: E402

REFERENCE = "/root/reference"


from oracle_doubles import install_reference_shims, write_bm25_dir  # noqa: E402


def main():
    install_reference_shims()
    sys.path.insert(0, REFERENCE)
    from src.utils.ensembleRetriever import EnsembleRetriever  # the reference, unmodified

    world = fw.make_world()
    out = {"n_chunks": len(world["metas"]), "cases": []}
    with tempfile.TemporaryDirectory() as tmp:
        write_bm25_dir(world, tmp)
        for cfg in [dict(k=5, enable_expand=True), dict(k=3, faiss_k=6, bm25_k=4, faiss_ts_k=2, enable_expand=False),
                    dict(k=4, faiss_k=0, bm25_k=5, faiss_ts_k=3, enable_expand=True)]:
            chroma, ts = fw.make_collections(world)
            r = EnsembleRetriever(tmp, chroma, ts, cfg["k"], fw.FakeEmbeddings(world),
                                  **{a: b for a, b in cfg.items() if a != "k"})
            for qi, (q, hyde) in enumerate(fw.QUERIES):
                chunks = r.invoke(q, list(hyde))
                out["cases"].append({"cfg": cfg, "query": qi, "chunks": fw.summarize(chunks)})
    with open(os.path.join(HERE, "ensemble_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out["cases"]), "cases;", sum(len(c["chunks"]) for c in out["cases"]), "chunks")

    # BASELINE config C1 at its stated size (~10k chunks x 1024-d, 16 queries, top-10): the reference's EnsembleRetriever
    # with k = 10 (`collections={'zeekr': 10}`, experiments/e2e/qa_e2e_async.py:67) and enable_expand as ragManager.py:104-114
    # builds it, plus its FaissRetriever called with the 16 query strings as one batch.
    world = fw.make_world_c1()
    out = {"n_chunks": len(world["metas"]), "n_titles": len(world["titles"]), "cases": [], "faiss_batch": None}
    with tempfile.TemporaryDirectory() as tmp:
        write_bm25_dir(world, tmp)
        chroma, ts = fw.make_collections(world)
        r = EnsembleRetriever(tmp, chroma, ts, 10, fw.FakeEmbeddings(world), enable_expand=True)
        for qi, (q, hyde) in enumerate(fw.QUERIES_C1):
            out["cases"].append({"query": qi, "chunks": fw.summarize_light(r.invoke(q, list(hyde)))})
        ids, dist = r.faiss_retriever.invoke([q for q, _ in fw.QUERIES_C1], 10)
        out["faiss_batch"] = {"ids": [[int(i) for i in row] for row in ids], "scores": [[float(x).hex() for x in row] for row in dist]}
    with open(os.path.join(HERE, "ensemble_golden_c1.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote C1:", len(out["cases"]), "cases;", sum(len(c["chunks"]) for c in out["cases"]), "chunks of", out["n_chunks"])


if __name__ == "__main__":
    main()

"""A small deterministic retrieval world (BASELINE config C1 in miniature) shared by the golden-vector
generator and the boundary tests: chunk embeddings with document structure (so prev/next expansion and
bundles trigger), title summaries, chunk texts for BM25, duck-typed fakes of the objects the reference's
EnsembleRetriever touches (Chroma collections, the embedding model)."""
from __future__ import annotations

import json
import os
import zlib

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
D = 48
WORDS = ("revenue profit margin vehicle delivery battery factory shanghai listing prospectus shares dividend "
         "merger board audit risk supply chain forecast quarter guidance debt equity cash flow patent").split()


def make_world(seed: int = 4242, n_docs: int = 40, d: int = D, chunks_lo: int = 3, chunks_hi: int = 9):
    D = d          # noqa: N806 (shadows the module constant for the C1-size world)
    rng = np.random.default_rng(seed)
    emb, metas, texts = [], [], []
    titles: list[str] = []
    for doc in range(n_docs):
        u = rng.standard_normal(D)
        n_chunks = int(rng.integers(chunks_lo, chunks_hi))
        title = f"title: report {doc} summary: about {WORDS[doc % len(WORDS)]} and {WORDS[(doc * 7 + 3) % len(WORDS)]}"
        if doc % 5 == 4:
            title = titles[-1]                      # several documents can share one title summary
        titles.append(title)
        bundle = doc if doc % 4 == 1 else None       # every 4th document is a bundle (e.g. a split table)
        for c in range(n_chunks):
            v = 0.85 * u + 0.55 * rng.standard_normal(D)
            emb.append(v)
            words = rng.choice(WORDS, size=int(rng.integers(6, 14)))
            texts.append(f"chunk {doc}-{c} " + " ".join(words))
            metas.append({
                "doc_id": f"doc{doc}-c{c}",
                "bundle_id": bundle,
                "prev_chunk_id": f"doc{doc}-c{c - 1}" if c > 0 else "",
                "next_chunk_id": f"doc{doc}-c{c + 1}" if c + 1 < n_chunks else "",
                "title_summary": title,
                "date_published": f"2024-{1 + doc % 12:02d}-15",
            })
    emb = np.asarray(emb, dtype=np.float32)
    uniq_titles = list(dict.fromkeys(titles))
    ts_emb = np.asarray([_text_vec(t, D) for t in uniq_titles], dtype=np.float32)
    return {"emb": emb, "metas": metas, "texts": texts, "titles": uniq_titles, "ts_emb": ts_emb}


def make_world_c1():
    """BASELINE config C1 at its stated size: ~10k chunks (5 400 documents of 1-3 chunks), 1024-d embeddings, 4 320 distinct
    title summaries — large enough that the reference's calls leave the tiny-shard path: the depth-2048 chunk search runs on
    the exact streaming scorer, the title search (k = 10, one query) on the streaming GEMV, a 16-query batch on the tcgen05
    kernel."""
    return make_world(seed=777, n_docs=5400, d=1024, chunks_lo=1, chunks_hi=4)


def _text_vec(text: str, d: int = D) -> np.ndarray:
    r = np.random.default_rng(zlib.crc32(text.encode()))
    return r.standard_normal(d).astype(np.float32)


class FakeEmbeddings:
    """embed_query(text): 'near:<row>:<salt>' -> a noisy copy of chunk <row>; a title string -> that title's
    vector plus noise; anything else -> a vector seeded by the text."""

    def __init__(self, world):
        self.world = world
        self.d = world["emb"].shape[1]

    def embed_query(self, text: str):
        D = self.d     # noqa: N806
        if text.startswith("near:"):
            _, row, salt = text.split(":")[:3]
            r = np.random.default_rng(int(salt))
            v = self.world["emb"][int(row)] + 0.25 * r.standard_normal(D)
            return v.astype(np.float32).tolist()
        if text in self.world["titles"]:
            v = self.world["ts_emb"][self.world["titles"].index(text)] + 0.1 * _text_vec("salt" + text, D)
            return v.astype(np.float32).tolist()
        return _text_vec(text, D).tolist()


class FakeChroma:
    """The slice of langchain_chroma.Chroma the retrievers use: get(include=...) and get(ids=..., include=...)."""

    def __init__(self, ids, documents, metadatas, embeddings):
        self._ids, self._docs, self._metas, self._emb = list(ids), list(documents), list(metadatas), embeddings
        self._pos = {i: n for n, i in enumerate(self._ids)}
        self.get_calls = 0

    def get(self, ids=None, include=None, **_):
        include = include or ["documents", "metadatas"]
        self.get_calls += 1
        rows = range(len(self._ids)) if ids is None else [self._pos[i] for i in ids if i in self._pos]
        out = {"ids": [self._ids[r] for r in rows]}
        if "documents" in include:
            out["documents"] = [self._docs[r] for r in rows]
        if "metadatas" in include:
            out["metadatas"] = [self._metas[r] for r in rows]
        if "embeddings" in include:
            out["embeddings"] = [self._emb[r].tolist() for r in rows]
        return out


def make_collections(world):
    ids = [m["doc_id"] for m in world["metas"]]
    chroma = FakeChroma(ids, world["texts"], world["metas"], world["emb"])
    ts = FakeChroma([f"ts{i}" for i in range(len(world["titles"]))], world["titles"],
                    [{} for _ in world["titles"]], world["ts_emb"])
    return chroma, ts


QUERIES = [
    ("near:7:1", []),
    ("near:40:2", ["near:41:3", "near:100:4"]),
    ("revenues profits shanghai batteries", []),      # inflected: only a stemmer makes the BM25 path find them
    ("near:120:5", ["near:7:1"]),
    ("near:63:8", ["dividends merged boards"]),
]


# C1: 16 queries (config/example.yaml retriever, top-10); 'near:<row>:<salt>' steers the dense path, the words the sparse one
QUERIES_C1 = [
    ("near:11:1", []), ("near:4000:2", ["near:4001:3"]), ("near:9000:4", ["near:9001:5", "near:17:6", "near:2500:7"]),
    ("revenues profits shanghai batteries", []), ("near:123:8", ["dividends merged boards"]), ("near:7777:9", []),
    ("audited risks of the supply chains", ["near:600:10"]), ("near:5000:11", []), ("near:1:12", ["near:2:13"]),
    ("quarterly guidance on debts and equities", []), ("near:8123:14", []), ("near:3333:15", ["near:3334:16", "near:3335:17"]),
    ("patents listing prospectus", []), ("near:10:18", []), ("near:6500:19", ["forecasting cash flows"]), ("near:9500:20", []),
]


def query_text(i: int, world) -> str:
    """Query i as the pipeline would phrase it: the dense part is steered by the 'near:' token when used as an
    embedding key, the BM25 part sees ordinary words."""
    return QUERIES[i][0]


def summarize(chunks):
    return [[c["retriever"], float(c["score"]).hex(), c["metadata"]["doc_id"], c["bundle_id"], c["page_content"]] for c in chunks]


def summarize_light(chunks):
    """The same without the chunk text (the C1 fixture is large): retriever tag, score bits, doc id, bundle."""
    return [[c["retriever"], float(c["score"]).hex(), c["metadata"]["doc_id"], c["bundle_id"]] for c in chunks]


def load_golden(name: str = "ensemble_golden.json"):
    with open(os.path.join(GOLDEN_DIR, name)) as f:
        return json.load(f)

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


def make_world(seed: int = 4242, n_docs: int = 40):
    rng = np.random.default_rng(seed)
    emb, metas, texts = [], [], []
    titles: list[str] = []
    for doc in range(n_docs):
        u = rng.standard_normal(D)
        n_chunks = int(rng.integers(3, 9))
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
    ts_emb = np.asarray([_text_vec(t) for t in uniq_titles], dtype=np.float32)
    return {"emb": emb, "metas": metas, "texts": texts, "titles": uniq_titles, "ts_emb": ts_emb}


def _text_vec(text: str) -> np.ndarray:
    r = np.random.default_rng(zlib.crc32(text.encode()))
    return r.standard_normal(D).astype(np.float32)


class FakeEmbeddings:
    """embed_query(text): 'near:<row>:<salt>' -> a noisy copy of chunk <row>; a title string -> that title's
    vector plus noise; anything else -> a vector seeded by the text."""

    def __init__(self, world):
        self.world = world

    def embed_query(self, text: str):
        if text.startswith("near:"):
            _, row, salt = text.split(":")
            r = np.random.default_rng(int(salt))
            v = self.world["emb"][int(row)] + 0.25 * r.standard_normal(D)
            return v.astype(np.float32).tolist()
        if text in self.world["titles"]:
            v = self.world["ts_emb"][self.world["titles"].index(text)] + 0.1 * _text_vec("salt" + text)
            return v.astype(np.float32).tolist()
        return _text_vec(text).tolist()


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


def query_text(i: int, world) -> str:
    """Query i as the pipeline would phrase it: the dense part is steered by the 'near:' token when used as an
    embedding key, the BM25 part sees ordinary words."""
    return QUERIES[i][0]


def summarize(chunks):
    return [[c["retriever"], float(c["score"]).hex(), c["metadata"]["doc_id"], c["bundle_id"], c["page_content"]] for c in chunks]


def load_golden():
    with open(os.path.join(GOLDEN_DIR, "ensemble_golden.json")) as f:
        return json.load(f)

"""Host-side mirror of the reference's retriever interface for the hot path — same class names,
constructor arguments, call signatures, return shapes and error behaviour, GPU kernels underneath.

    FaissRetriever.invoke(querys, k) -> (indices, distances)      /root/reference/src/utils/faissRetriever.py:8-38
    BM25Retriever.invoke(query, k)   -> (ids, scores)              /root/reference/src/utils/bm25Retriever.py:22-87
    load_from_chroma_and_save(documents, save_dir)                 /root/reference/src/utils/bm25Retriever.py:10-20
    EnsembleRetriever.invoke(input, hyde_chunks) -> List[Dict]     /root/reference/src/utils/ensembleRetriever.py:16-233

`embedding_fn` needs `.embed_query(str) -> list[float]`; `chroma` / `ts_chroma` need
`.get(include=[...])` and `.get(ids=[...], include=[...])` — the duck-typed surface the reference uses.
The O(N)-per-hit metadata scans of the reference (ensembleRetriever.py:81,145,203) are replaced by maps
built once at construction; the output list is the same.
"""
from __future__ import annotations

import logging
from typing import Any, Dict, List, Optional

import numpy as np

from . import bm25_compat, faiss_compat
from .profiler import profiler

logger = logging.getLogger(__name__)

DENSE_DEPTH = 2048          # ensembleRetriever.py:66
EXPAND_SCORE = 0.72         # ensembleRetriever.py:86
NEIGHBOUR_SCORE = 0.66      # ensembleRetriever.py:93,102
MAX_EXPANDED = 4            # ensembleRetriever.py:89


class _IdentityStemmer:
    """Leaves tokens unstemmed (used by the golden fixtures, which were generated without a stemmer)."""

    def __init__(self, lang: str = "english"):
        self.lang = lang

    def stemWords(self, words):
        return list(words)


def make_stemmer(lang: str = "english"):
    """`Stemmer.Stemmer(lang)` of bm25Retriever.py:14,47.  PyStemmer when the deployment has it, else the native
    Snowball-English routine of this package (csrc/text_host.h behind vfi_stem_english)."""
    try:
        import Stemmer  # PyStemmer, if the deployment has it
        if hasattr(Stemmer, "Stemmer") and not getattr(Stemmer, "_vfi_shim", False):
            return Stemmer.Stemmer(lang)
    except Exception:
        pass
    from .stemmer import Stemmer as NativeStemmer
    return NativeStemmer(lang)


class FaissRetriever:
    def __init__(self, embeddings, embedding_fn, device: int = 0):
        self.embeddings = embedding_fn
        x = np.array(embeddings).astype("float32")
        if x.ndim != 2:
            raise ValueError("embeddings must be a 2-D array-like")
        x = np.ascontiguousarray(x)
        self.index = faiss_compat.IndexFlatIP(x.shape[1], device=device)
        faiss_compat.normalize_L2(x, device=device)
        self.index.add(x)
        logger.info("Building flat IP index with %d vectors of dimension %d", x.shape[0], x.shape[1])

    def invoke(self, querys: list[str], k: int):
        vecs = np.array([self.embeddings.embed_query(q) for q in querys]).astype("float32")
        vecs = np.ascontiguousarray(vecs)
        faiss_compat.normalize_L2(vecs, device=self.index.device)
        distances, indices = self.index.search(vecs, k)
        return indices, distances


def load_from_chroma_and_save(documents, save_dir: str):
    corpus = [doc.page_content for doc in documents]
    doc_ids = [doc.metadata["doc_id"] for doc in documents]
    tokens = bm25_compat.tokenize(corpus, stopwords="english", stemmer=make_stemmer("english"))
    engine = bm25_compat.BM25()
    engine.index(tokens)
    engine.save(save_dir, corpus=doc_ids)
    logging.info("%d documents saved to %s", len(documents), save_dir)


class BM25Retriever:
    def __init__(self, dir_path: str, load_corpus: bool = True, min_score: Optional[float] = None,
                 stemmer: str = "english", device: int = 0):
        self.min_score = min_score
        self._bm25_engine = bm25_compat.BM25.load(dir_path, load_corpus=load_corpus, device=device)
        self._stemmer = make_stemmer(stemmer)
        self.doc_len = len(self._bm25_engine.corpus)

    def invoke(self, query: str, k: int, metadata_filters: Optional[Dict[str, Any]] = None):
        tokens = bm25_compat.tokenize([query], stopwords="english", stemmer=self._stemmer)
        if metadata_filters:
            raise NotImplementedError("Metadata filtering is not supported yet.")
        docs, scores = self._bm25_engine.retrieve(tokens, k=k, return_as="tuple")
        docs, scores = docs[0], scores[0]
        if self.min_score is not None:
            docs = [doc for doc, score in zip(docs, scores) if score >= self.min_score]
        if isinstance(docs, bm25_compat.LazyRow):
            # k beyond the eager head (the reference passes k = num_chunk and reads [:bm25_k], ensembleRetriever.py:189-190):
            # the ids stay lazy too, so the ranks past the head are only produced if somebody reads them
            return _LazyDocIds(docs), scores
        return [doc["id"] for doc in docs], scores


class _LazyDocIds:
    """`[doc["id"] for doc in docs]` of bm25Retriever.py:86 over a lazy row: evaluated on the part that is read."""

    def __init__(self, docs):
        self._docs = docs

    def __len__(self):
        return len(self._docs)

    def __getitem__(self, key):
        got = self._docs[key]
        if isinstance(key, slice):
            return [doc["id"] for doc in got]
        return got["id"]

    def __iter__(self):
        return (doc["id"] for doc in self._docs)


class EnsembleRetriever:
    def __init__(self, bm25_dir: str, chroma, ts_chroma, k: int, embeddings, faiss_k: int = None,
                 bm25_k: int = None, faiss_ts_k: int = None, enable_expand: bool = False, device: int = 0):
        self.embeddings = embeddings
        self.faiss_k = k if faiss_k is None else faiss_k
        self.bm25_k = k if bm25_k is None else bm25_k
        self.faiss_ts_k = k if faiss_ts_k is None else faiss_ts_k
        self.enable_expand = enable_expand
        self.chroma = chroma
        self.bm25_retriever = BM25Retriever(bm25_dir, device=device)
        docs = chroma.get(include=["metadatas", "embeddings"])
        self.faiss_retriever = FaissRetriever(docs["embeddings"], embeddings, device=device)
        ts_docs = ts_chroma.get(include=["documents", "embeddings"])
        self.title_summary_faiss_retriever = FaissRetriever(ts_docs["embeddings"], embeddings, device=device)
        self.chunk_metadata = docs["metadatas"]
        self.docid2idx = {m["doc_id"]: i for i, m in enumerate(self.chunk_metadata)}
        self.num_chunk = len(self.chunk_metadata)
        self.title_summaries = ts_docs["documents"]
        # row maps replacing the per-hit scans; values are in ascending row order like the scans
        self._bundle_rows: dict[Any, list[int]] = {}
        self._title_rows: dict[str, list[int]] = {}
        for i, m in enumerate(self.chunk_metadata):
            b = m.get("bundle_id", None)
            if b is not None:
                self._bundle_rows.setdefault(b, []).append(i)
            self._title_rows.setdefault(m.get("title_summary", ""), []).append(i)

    # -- helpers ---------------------------------------------------------------------------------
    def _claim(self, idx, seen: set) -> list:
        """Mark idx as seen and return the rows it brings along (its whole bundle when it has one)."""
        seen.add(idx)
        meta = self.chunk_metadata[idx]
        b = meta.get("bundle_id", None)
        if b is None:
            return [idx]
        rows = list(self._bundle_rows[b])
        seen.update(rows)
        return rows

    def _emit(self, rows: list, tag: str, score, bundle: int, out: list) -> None:
        doc_ids = [self.chunk_metadata[i]["doc_id"] for i in rows]
        got = self.chroma.get(ids=doc_ids, include=["documents", "metadatas"])
        for text, meta in zip(got["documents"], got["metadatas"]):
            out.append({"retriever": tag, "score": float(score), "page_content": text, "metadata": meta,
                        "bundle_id": bundle})

    def _expand(self, rows: list, meta: dict, score_of: dict, seen: set) -> None:
        prev_id, next_id = meta["prev_chunk_id"], meta["next_chunk_id"]
        while len(rows) < MAX_EXPANDED:
            grew = False
            p = self.docid2idx.get(prev_id, -1) if prev_id != "" else -1
            if p != -1 and score_of.get(p, 0) > NEIGHBOUR_SCORE and p not in seen:
                grew = True
                seen.add(p)
                rows.insert(0, p)
                prev_id = self.chunk_metadata[p]["prev_chunk_id"]
            n = self.docid2idx.get(next_id, -1) if next_id != "" else -1
            if n != -1 and score_of.get(n, 0) > NEIGHBOUR_SCORE and n not in seen:
                grew = True
                seen.add(n)
                rows.append(n)
                next_id = self.chunk_metadata[n]["next_chunk_id"]
            if not grew:
                break

    # -- near-duplicate helpers used by the reranker (ensembleRetriever.py:235-281) ----------------
    def _embed_normalised_index(self, chunks: List[str]):
        from .dense import DenseIndex
        vecs = np.ascontiguousarray(np.array([self.embeddings.embed_query(c) for c in chunks]).astype("float32"))
        device = self.faiss_retriever.index.device
        faiss_compat.normalize_L2(vecs, device=device)
        tmp = DenseIndex(vecs.shape[1], store="f32", device=device)
        tmp.add(vecs)
        return tmp

    def compute_similarity(self, chunks: List[str], selected_indices: List[int], candidate_index: int):
        """Cosine similarity between chunk `candidate_index` and the chunks `selected_indices` (a cuda tensor)."""
        tmp = self._embed_normalised_index(chunks)
        ids = list(selected_indices) + [candidate_index]
        m = tmp.pairwise(ids)
        return m[:-1, -1].clone()

    def compute_similarity_mtx(self, chunks: List[str]):
        """Pairwise cosine similarity of the chunks (a cuda tensor [n, n])."""
        tmp = self._embed_normalised_index(chunks)
        return tmp.pairwise(range(len(chunks)))

    def similarity_mtx_by_row(self, rows: List[int]):
        """The same matrix for chunks already in the index, addressed by row id: no re-embedding."""
        from .dense import DenseIndex  # noqa: F401
        import ctypes as C
        import torch
        from . import _native as N
        idx = self.faiss_retriever.index
        ids = np.ascontiguousarray(rows, dtype=np.int64)
        out = np.empty((len(ids), len(ids)), dtype=np.float32)
        if len(ids):
            N.check(N.load().vfi_index_pairwise(idx._h, ids.ctypes.data_as(C.c_void_p), len(ids), out.ctypes.data_as(C.c_void_p),
                                                N.MEM_HOST, None))
        return torch.from_numpy(out)

    # -- the call the rest of the pipeline makes -------------------------------------------------
    @profiler.profile_function(name="retrieve")
    def invoke(self, input: str, hyde_chunks: list[str]) -> List[Dict]:
        seen: set = set()
        chunks: list = []
        bundle = 0

        if self.faiss_k > 0:
            profiler.start("retrieve_faiss")
            texts = [input] + hyde_chunks
            all_ids, all_scores = self.faiss_retriever.invoke(texts, DENSE_DEPTH)
            for ids, scores in zip(all_ids, all_scores):
                score_of = dict(zip(ids, scores))
                for idx, score in zip(ids[: self.faiss_k], scores[: self.faiss_k]):
                    if idx in seen:
                        continue
                    meta = self.chunk_metadata[idx]
                    rows = self._claim(idx, seen)
                    if score > EXPAND_SCORE and self.enable_expand:
                        self._expand(rows, meta, score_of, seen)
                    self._emit(rows, "FAISS", score, bundle, chunks)
                    bundle += 1
            profiler.end("retrieve_faiss")

        if self.faiss_ts_k > 0:
            profiler.start("retrieve_faiss_ts")
            ts_ids, ts_scores = self.title_summary_faiss_retriever.invoke([input], self.faiss_ts_k)
            for title_idx, score in zip(ts_ids[0], ts_scores[0]):
                title = self.title_summaries[title_idx]
                for idx in self._title_rows.get(title, ()):
                    if idx in seen:
                        continue
                    rows = self._claim(idx, seen)
                    self._emit(rows, "Title Summary", score, bundle, chunks)
                    bundle += 1
            profiler.end("retrieve_faiss_ts")

        if self.bm25_k > 0:
            profiler.start("retrieve_bm25")
            ids, scores = self.bm25_retriever.invoke(input, self.num_chunk)
            for idx, score in zip(ids[: self.bm25_k], scores[: self.bm25_k]):
                if idx in seen:
                    continue
                rows = self._claim(idx, seen)
                self._emit(rows, "BM25", score, bundle, chunks)
                bundle += 1
            profiler.end("retrieve_bm25")

        profiler.add_metric("retrieved_chunks", len(chunks))
        return chunks

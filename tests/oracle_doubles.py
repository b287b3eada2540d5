"""Oracle-backed test doubles of the third-party modules the reference imports (faiss, bm25s, Stemmer,
langchain_*).  Used to run the UNMODIFIED reference files in this container (golden generation) and to
exercise the host-side mirror on CPU.  Test infrastructure only."""
import sys
import types

import numpy as np

from oracle import bm25 as obm, flat_ip, porter2
from veritasfi_b200 import bm25_compat


class OracleIndexFlatIP:
    def __init__(self, d, device=0, store="f32"):
        self.d, self.device, self.x = d, device, np.zeros((0, d), np.float32)

    @property
    def ntotal(self):
        return len(self.x)

    def add(self, x):
        self.x = np.concatenate([self.x, np.ascontiguousarray(x, np.float32)])

    def search(self, q, k):
        return flat_ip.search_exhaustive(q, self.x, k)


def oracle_normalize_L2(x, device=0):
    x[...] = flat_ip.normalize_l2(x)


class OracleBM25(bm25_compat.BM25):
    """bm25s.BM25 whose retrieve() is scored by the oracle (no GPU)."""

    def retrieve(self, query_tokens, corpus=None, k=10, return_as="tuple", **_):
        s = self.scores
        if k > s["num_docs"]:
            raise ValueError("k larger than the number of documents")
        lists = self._query_id_lists(query_tokens)
        ids, scores = obm.retrieve(s["indptr"], s["indices"], s["data"], lists, s["num_docs"], k)
        corpus = corpus if corpus is not None else self.corpus
        docs = np.empty(ids.shape, dtype=object)
        for i in range(ids.shape[0]):
            for j in range(ids.shape[1]):
                docs[i, j] = corpus[int(ids[i, j])]
        return bm25_compat.Results(documents=docs, scores=scores)


class IdentityStemmer:
    def __init__(self, lang="english"):
        pass

    def stemWords(self, ws):
        return list(ws)


class OracleStemmer:
    """PyStemmer's Stemmer.Stemmer('english') backed by the oracle restatement of Snowball English (oracle/porter2.py)."""

    def __init__(self, lang="english"):
        assert lang == "english"

    def stemWord(self, w):
        return porter2.stem(w)

    def stemWords(self, ws):
        return porter2.stem_words(ws)


def install_reference_shims():
    def mod(name, **kw):
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        sys.modules[name] = m
        return m
    mod("faiss", IndexFlatIP=OracleIndexFlatIP, normalize_L2=oracle_normalize_L2)
    mod("bm25s", BM25=OracleBM25, tokenize=bm25_compat.tokenize)
    mod("Stemmer", Stemmer=OracleStemmer)
    mod("langchain_huggingface", HuggingFaceEmbeddings=object)
    mod("langchain_community")
    mod("langchain_community.vectorstores", FAISS=object)
    mod("langchain_chroma", Chroma=object)
    mod("langchain_core")
    mod("langchain_core.documents", Document=object)


def write_bm25_dir(world, path):
    eng = bm25_compat.BM25()
    eng.index(bm25_compat.tokenize(world["texts"], stopwords="english", stemmer=OracleStemmer()))
    eng.save(path, corpus=[m["doc_id"] for m in world["metas"]])



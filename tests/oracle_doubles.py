"""Oracle-backed test doubles of the third-party modules the reference imports (faiss, bm25s, Stemmer,
langchain_*).  Used to run the UNMODIFIED reference files in this container (golden generation) and to
exercise the host-side mirror on CPU.  Test infrastructure only: nothing here imports the product package
(veritasfi_b200) — tokenisation, index construction, the index files and the scoring are the oracle's own
restatements (oracle/bm25.py, oracle/porter2.py, oracle/flat_ip.py)."""
import sys
import types
from collections import namedtuple

import numpy as np

from oracle import bm25 as obm, flat_ip, porter2

Tokenized = namedtuple("Tokenized", ["ids", "vocab"])
Results = namedtuple("Results", ["documents", "scores"])


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


def oracle_tokenize(texts, stopwords="english", stemmer=None, **_):
    if isinstance(texts, str):
        texts = [texts]
    assert stopwords == "english"
    ids, vocab = obm.tokenize(texts, stemmer)
    return Tokenized(ids=ids, vocab=vocab)


class OracleBM25:
    """bm25s.BM25 restated by the oracle: index / save / load / retrieve(k, return_as="tuple")."""

    def __init__(self, **_):
        self.ix = None
        self.corpus = None
        self.vocab_dict = {}

    def index(self, tokens, **_):
        ids, vocab = tokens
        indptr, indices, data = obm.build_index(ids, len(vocab))
        self.ix = dict(indptr=indptr, indices=indices, data=data, n_docs=len(ids))
        self.vocab_dict = dict(vocab)

    def save(self, path, corpus=None, **_):
        obm.save_dir(path, self.ix["indptr"], self.ix["indices"], self.ix["data"], self.vocab_dict, self.ix["n_docs"],
                     corpus if corpus is not None else [])

    @classmethod
    def load(cls, path, load_corpus=False, device=0, **_):
        self = cls()
        got = obm.load_dir(path)
        self.ix = dict(indptr=got["indptr"], indices=got["indices"], data=got["data"], n_docs=got["n_docs"])
        self.vocab_dict = got["vocab"]
        self.corpus = got["corpus"] if load_corpus else None
        return self

    def retrieve(self, query_tokens, corpus=None, k=10, return_as="tuple", **_):
        s = self.ix
        if k > s["n_docs"]:
            raise ValueError("k larger than the number of documents")
        q_ids, q_vocab = query_tokens
        rev = {i: w for w, i in q_vocab.items()}
        lists = [[self.vocab_dict[rev[i]] for i in q if rev[i] in self.vocab_dict] for q in q_ids]
        ids, scores = obm.retrieve(s["indptr"], s["indices"], s["data"], lists, s["n_docs"], k)
        corpus = corpus if corpus is not None else self.corpus
        docs = np.empty(ids.shape, dtype=object)
        for i in range(ids.shape[0]):
            for j in range(ids.shape[1]):
                docs[i, j] = corpus[int(ids[i, j])]
        return Results(documents=docs, scores=scores)


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


def install_reference_shims(langchain_only: bool = False):
    def mod(name, **kw):
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        sys.modules[name] = m
        return m
    if not langchain_only:
        mod("faiss", IndexFlatIP=OracleIndexFlatIP, normalize_L2=oracle_normalize_L2)
        mod("bm25s", BM25=OracleBM25, tokenize=oracle_tokenize)
        mod("Stemmer", Stemmer=OracleStemmer)
    mod("langchain_huggingface", HuggingFaceEmbeddings=object)
    mod("langchain_community")
    mod("langchain_community.vectorstores", FAISS=object)
    mod("langchain_chroma", Chroma=object)
    mod("langchain_core")
    mod("langchain_core.documents", Document=object)


def write_bm25_dir(world, path):
    """The bm25s index directory of the fixture world, built and written by the oracle (what load_from_chroma_and_save,
    bm25Retriever.py:10-20, produces through bm25s)."""
    eng = OracleBM25()
    eng.index(oracle_tokenize(world["texts"], stopwords="english", stemmer=OracleStemmer()))
    eng.save(path, corpus=[m["doc_id"] for m in world["metas"]])

"""CPU tests of the host-side text routines of the sparse path: the native Snowball-English
stemmer and ASCII tokeniser behind the C ABI (no device involved) against the Python restatement in oracle/porter2.py
and against the sample vocabulary published with the algorithm (the only golden vectors that exist for it: PyStemmer is
neither vendored nor pinned by the reference, bm25Retriever.py:8,14)."""
import glob
import itertools
import os
import re

import numpy as np
import pytest

from oracle import porter2
from veritasfi_b200 import bm25_compat
from veritasfi_b200.stemmer import Stemmer, stem_words_native, tokenize_ascii_native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# the two runs of the sample vocabulary shown in the algorithm's description ("consign" ... and "knack" ...)
PUBLISHED = dict(zip(
    """consign consigned consigning consignment consist consisted consistency consistent consistently consisting
    consists consolation consolations consolatory console consoled consoles consolidate consolidated consolidating
    consoling consolingly consols consonant consort consorted consorting conspicuous conspicuously conspiracy
    conspirator conspirators conspire conspired conspiring constable constables constance constancy constant
    knack knackeries knacks knag knave knaves knavish kneaded kneading knee kneel kneeled kneeling kneels knees knell
    knelt knew knick knif knife knight knightly knights knit knits knitted knitting knives knob knobs knock knocked
    knocker knockers knocking knocks knopp knot knots""".split(),
    """consign consign consign consign consist consist consist consist consist consist
    consist consol consol consolatori consol consol consol consolid consolid consolid
    consol consol consol conson consort consort consort conspicu conspicu conspiraci
    conspir conspir conspir conspir conspir constabl constabl constanc constanc constant
    knack knackeri knack knag knave knave knavish knead knead knee kneel kneel kneel kneel knee knell
    knelt knew knick knif knife knight knight knight knit knit knit knit knive knob knob knock knock
    knocker knocker knock knock knopp knot knot""".split()))

# rule-by-rule examples from the algorithm's description (steps 0-5, exceptional forms, R1 prefixes, y -> Y)
RULES = {
    "caresses": "caress", "ponies": "poni", "ties": "tie", "cries": "cri", "gas": "gas", "this": "this", "gaps": "gap",
    "kiwis": "kiwi", "cats": "cat", "luxuriated": "luxuri", "hopping": "hop", "hoping": "hope", "hoped": "hope",
    "hopped": "hop", "plastered": "plaster", "bled": "bled", "motoring": "motor", "sing": "sing", "conflated": "conflat",
    "troubled": "troubl", "sized": "size", "tanned": "tan", "falling": "fall", "hissing": "hiss", "fizzed": "fizz",
    "failing": "fail", "filing": "file", "cry": "cri", "by": "by", "say": "say", "happy": "happi", "relational": "relat",
    "conditional": "condit", "rational": "ration", "national": "nation", "replacement": "replac", "adjustment": "adjust",
    "dependent": "depend", "adoption": "adopt", "homologous": "homolog", "effective": "effect", "bowdlerize": "bowdler",
    "probate": "probat", "rate": "rate", "cease": "ceas", "controll": "control", "roll": "roll",
    "generate": "generat", "generates": "generat", "generated": "generat", "generating": "generat", "general": "general",
    "generally": "general", "generic": "generic", "generous": "generous", "generously": "generous",
    "communism": "communism", "communities": "communiti", "community": "communiti", "communicate": "communic",
    "communication": "communic", "arsenal": "arsenal", "arsenic": "arsenic",
    "sky": "sky", "skies": "sky", "skis": "ski", "dying": "die", "lying": "lie", "tying": "tie", "news": "news",
    "idly": "idl", "gently": "gentl", "ugly": "ugli", "early": "earli", "only": "onli", "singly": "singl",
    "howe": "howe", "atlas": "atlas", "cosmos": "cosmos", "bias": "bias", "andes": "andes",
    "inning": "inning", "innings": "inning", "outing": "outing", "canning": "canning", "herring": "herring",
    "earring": "earring", "proceed": "proceed", "proceeds": "proceed", "proceeded": "proceed", "proceeding": "proceed",
    "exceed": "exceed", "succeed": "succeed",
    "running": "run", "runs": "run", "ran": "ran", "easily": "easili", "fairly": "fair", "university": "univers",
    "universe": "univers", "connection": "connect", "connected": "connect", "connecting": "connect", "argue": "argu",
    "argued": "argu", "argument": "argument", "arguments": "argument", "agreed": "agre", "agreement": "agreement",
    "feed": "feed", "owed": "owe", "controlling": "control", "rolling": "roll", "singing": "sing", "string": "string",
    "meeting": "meet", "fishing": "fish", "bed": "bed", "shed": "shed", "shred": "shred", "youth": "youth", "yes": "yes",
    "saying": "say", "boyish": "boyish", "flying": "fli", "'twas": "twas", "dogs'": "dog", "dog's": "dog",
    "a": "a", "is": "is", "happiness": "happi",
}


def test_oracle_stemmer_reproduces_the_published_vocabulary():
    for word, want in {**PUBLISHED, **RULES}.items():
        assert porter2.stem(word) == want, word


def test_native_stemmer_reproduces_the_published_vocabulary():
    words = list(PUBLISHED) + list(RULES)
    got = stem_words_native(words)
    for w, g in zip(words, got):
        assert g == {**PUBLISHED, **RULES}[w], w


def _word_pool():
    words = set()
    for f in glob.glob(os.path.join(ROOT, "*.md")):
        with open(f, encoding="utf-8") as fh:
            words.update(re.findall(r"[a-z']+", fh.read().lower()))
    stems = ["hop", "hope", "control", "agree", "relate", "nation", "yield", "say", "play", "enjoy", "cry", "try", "tie",
             "die", "ski", "general", "commun", "arsen", "feed", "succeed", "bias", "sens", "vital", "formal", "electric",
             "able", "oper", "happy", "bely", "ow", "ax", "aby", "é", "naïv", "uy", "ye", "yy", "ayy", "x", "qu", "日本"]
    sufs = ["", "s", "es", "ed", "ing", "ingly", "edly", "eed", "eedly", "ly", "li", "ational", "tional", "ization", "izer",
            "ation", "ator", "alism", "aliti", "alli", "fulness", "ousli", "ousness", "iveness", "iviti", "biliti", "bli",
            "ogi", "logi", "fulli", "lessli", "alize", "icate", "iciti", "ical", "ful", "ness", "ative", "al", "ance",
            "ence", "er", "ic", "able", "ible", "ant", "ement", "ment", "ent", "ism", "ate", "iti", "ous", "ive", "ize",
            "ion", "sion", "tion", "e", "l", "ll", "y", "ys", "ies", "ied", "'s", "'", "'s'", "sses", "ss", "us"]
    for a, b in itertools.product(stems, sufs):
        words.add(a + b)
    for a, b, c in itertools.product(stems[:12], sufs[:30], sufs[:12]):
        words.add(a + b + c)
    rng = np.random.default_rng(5)
    letters = np.array(list("aeiouybcdglmnrstwxz'"))
    for _ in range(4000):
        words.add("".join(rng.choice(letters, size=int(rng.integers(1, 12)))))
    return sorted(w for w in words if w)


def test_native_stemmer_equals_the_oracle_word_by_word():
    words = _word_pool()
    assert len(words) > 10000
    got = stem_words_native(words)
    want = porter2.stem_words(words)
    diff = [(w, g, o) for w, g, o in zip(words, got, want) if g != o]
    assert not diff, diff[:10]


def test_stemmer_facade_has_the_pystemmer_surface():
    st = Stemmer("english")
    assert st.stemWord("retrieval") == porter2.stem("retrieval")
    assert st.stemWords(["tables", "figures", "captions"]) == ["tabl", "figur", "caption"]
    assert st.stemWords([]) == []
    with pytest.raises(KeyError):
        Stemmer("klingon")


def test_ascii_tokeniser_equals_the_bm25s_pattern():
    rng = np.random.default_rng(9)
    alphabet = np.array(list("abcXYZ019_ -.,;:'\"()[]\n\t/&%$#@!?+*="))
    pat = re.compile(r"(?u)\b\w\w+\b")
    for _ in range(300):
        text = "".join(rng.choice(alphabet, size=int(rng.integers(0, 80))))
        assert tokenize_ascii_native(text) == pat.findall(text), text
    assert tokenize_ascii_native("naïve café") is None      # non-ASCII: the facade uses the regular expression
    assert bm25_compat._split("naïve café x") == ["naïve", "café"]


def test_tokenize_with_the_native_stemmer_merges_inflections():
    tok = bm25_compat.tokenize(["The batteries are charging", "a charged battery charges"], stopwords="english",
                               stemmer=Stemmer("english"))
    assert set(tok.vocab) == {"batteri", "charg"}
    assert tok.ids == [[tok.vocab["batteri"], tok.vocab["charg"]], [tok.vocab["charg"], tok.vocab["batteri"], tok.vocab["charg"]]]

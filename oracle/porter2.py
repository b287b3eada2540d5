"""ORACLE (test infrastructure, not shipped): the Snowball "english" (Porter2) stemmer restated in plain Python.

Reference call sites (relative to /root/reference/): `Stemmer.Stemmer('english')` at src/utils/bm25Retriever.py:14,47,
handed to `bm25s.tokenize(..., stemmer=...)` at :15,67.  PyStemmer (a wrapper of the Snowball C library) is neither
vendored nor pinned by the reference and is absent from this image, so this follows the published program
`algorithms/english.sbl` of Snowball 2.0-2.2 statement by statement (prelude, mark_regions, Step_1a ... Step_5,
exception1/exception2, postlude), working on a list of characters with a backwards cursor the way the Snowball
runtime does.  Pinned by the sample vocabulary published with the algorithm (tests/test_text.py); the C++ routine
in veritasfi_b200/csrc/text_host.h is written independently (suffix tables on a byte string) and compared with this
one word by word.
"""
from __future__ import annotations

V = set("aeiouy")
V_WXY = V | set("wxY")
VALID_LI = set("cdeghkmnrt")
DOUBLES = {"bb", "dd", "ff", "gg", "mm", "nn", "pp", "rr", "tt"}

EXCEPTION1 = {
    "skis": "ski", "skies": "sky", "dying": "die", "lying": "lie", "tying": "tie",
    "idly": "idl", "gently": "gentl", "ugly": "ugli", "early": "earli", "only": "onli", "singly": "singl",
    "sky": "sky", "news": "news", "howe": "howe", "atlas": "atlas", "cosmos": "cosmos", "bias": "bias", "andes": "andes",
}
EXCEPTION2 = {"inning", "outing", "canning", "herring", "earring", "proceed", "exceed", "succeed"}

STEP2 = [("tional", "tion"), ("enci", "ence"), ("anci", "ance"), ("abli", "able"), ("entli", "ent"), ("izer", "ize"),
         ("ization", "ize"), ("ational", "ate"), ("ation", "ate"), ("ator", "ate"), ("alism", "al"), ("aliti", "al"),
         ("alli", "al"), ("fulness", "ful"), ("ousli", "ous"), ("ousness", "ous"), ("iveness", "ive"), ("iviti", "ive"),
         ("biliti", "ble"), ("bli", "ble"), ("ogi", "<ogi>"), ("fulli", "ful"), ("lessli", "less"), ("li", "<li>")]
STEP3 = [("tional", "tion"), ("ational", "ate"), ("alize", "al"), ("icate", "ic"), ("iciti", "ic"), ("ical", "ic"),
         ("ful", ""), ("ness", ""), ("ative", "<r2>")]
STEP4 = ["al", "ance", "ence", "er", "ic", "able", "ible", "ant", "ement", "ment", "ent", "ism", "ate", "iti", "ous",
         "ive", "ize", "ion"]


def _among(word: list[str], suffixes):
    """the longest suffix of `suffixes` the word ends with (Snowball's `substring` in backward mode), or None"""
    best = None
    text = "".join(word)
    for s in suffixes:
        if text.endswith(s) and (best is None or len(s) > len(best)):
            best = s
    return best


def _gopast(word, pos, want_vowel: bool):
    """forward `gopast v` / `gopast non-v`: index just after the first matching character at or after pos, or None"""
    for i in range(pos, len(word)):
        if (word[i] in V) == want_vowel:
            return i + 1
    return None


def _shortv(word, end) -> bool:
    """backward `shortv` with the cursor at `end`: ( non-v_WXY v non-v ) or ( non-v v atlimit )"""
    if end >= 3 and word[end - 1] not in V_WXY and word[end - 2] in V and word[end - 3] not in V:
        return True
    return end == 2 and word[1] not in V and word[0] in V


def stem(token: str) -> str:
    if token in EXCEPTION1:
        return EXCEPTION1[token]
    if len(token) < 3:                                   # not hop 3
        return token
    w = list(token)
    # prelude
    if w and w[0] == "'":
        del w[0]
    if w and w[0] == "y":
        w[0] = "Y"
    cursor = 0
    while True:                                          # repeat(goto (v ['y']) <-'Y')
        hit = None
        for i in range(cursor, len(w) - 1):
            if w[i] in V and w[i + 1] == "y":
                hit = i
                break
        if hit is None:
            break
        w[hit + 1] = "Y"
        cursor = hit
    # mark_regions
    p1 = p2 = len(w)
    text = "".join(w)
    pos = None
    for prefix in ("gener", "commun", "arsen"):
        if text.startswith(prefix):
            pos = len(prefix)
            break
    if pos is None:
        a = _gopast(w, 0, True)
        pos = _gopast(w, a, False) if a is not None else None
    if pos is not None:
        p1 = pos
        a = _gopast(w, pos, True)
        b = _gopast(w, a, False) if a is not None else None
        if b is not None:
            p2 = b

    def r1(suffix):
        return len(w) - len(suffix) >= p1

    def r2(suffix):
        return len(w) - len(suffix) >= p2

    def cut(n, by=""):
        del w[len(w) - n:]
        w.extend(by)

    # Step_1a
    s = _among(w, ["'", "'s", "'s'"])
    if s:
        cut(len(s))
    s = _among(w, ["sses", "ied", "ies", "s", "us", "ss"])
    if s == "sses":
        cut(4, "ss")
    elif s in ("ied", "ies"):
        cut(3, "i" if len(w) - 3 >= 2 else "ie")          # hop 2 backwards succeeds when two characters precede
    elif s == "s":
        before = len(w) - 1                                # cursor in front of the s
        if before >= 1 and any(c in V for c in w[:before - 1]):   # next, then gopast v
            cut(1)
    if "".join(w) in EXCEPTION2:
        return "".join(w).replace("Y", "y")
    # Step_1b
    s = _among(w, ["eed", "eedly", "ed", "edly", "ing", "ingly"])
    if s in ("eed", "eedly"):
        if r1(s):
            cut(len(s), "ee")
    elif s is not None:
        if any(c in V for c in w[:len(w) - len(s)]):       # test gopast v
            cut(len(s))
            t = _among(w, ["at", "bl", "iz"] + sorted(DOUBLES))
            if t in ("at", "bl", "iz"):
                w.append("e")
            elif t is not None:
                cut(1)
            elif len(w) == p1 and _shortv(w, len(w)):      # atmark p1  test shortv
                w.append("e")
    # Step_1c
    if len(w) >= 1 and w[-1] in ("y", "Y"):
        if len(w) >= 2 and w[-2] not in V and len(w) - 2 > 0:   # non-v, not atlimit
            w[-1] = "i"
    # Step_2
    s = _among(w, [a for a, _ in STEP2])
    if s is not None and r1(s):
        by = dict(STEP2)[s]
        if by == "<ogi>":
            if len(w) > 3 and w[-4] == "l":
                cut(3, "og")
        elif by == "<li>":
            if len(w) > 2 and w[-3] in VALID_LI:
                cut(2)
        else:
            cut(len(s), by)
    # Step_3
    s = _among(w, [a for a, _ in STEP3])
    if s is not None and r1(s):
        by = dict(STEP3)[s]
        if by == "<r2>":
            if r2(s):
                cut(len(s))
        else:
            cut(len(s), by)
    # Step_4
    s = _among(w, STEP4)
    if s is not None and r2(s):
        if s == "ion":
            if len(w) > 3 and w[-4] in ("s", "t"):
                cut(3)
        else:
            cut(len(s))
    # Step_5
    if w and w[-1] == "e":
        if r2("e") or (r1("e") and not _shortv(w, len(w) - 1)):
            cut(1)
    elif w and w[-1] == "l":
        if r2("l") and len(w) >= 2 and w[-2] == "l":
            cut(1)
    return "".join(w).replace("Y", "y")                   # postlude


def stem_words(words):
    return [stem(x) for x in words]

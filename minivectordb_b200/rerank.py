"""Host-side rerank helpers (stay host Python by BASELINE.json's north_star).

* autocut_scores      -- ref minivectordb/vector_database.py:443-464
* hybrid_rerank       -- ref minivectordb/vector_database.py:388-441
* partial_ratio       -- the reference calls thefuzz.fuzz.partial_ratio
  (VDB:5, 411); thefuzz / rapidfuzz are not installable in this image, so the
  published algorithm (best Indel-similarity of the shorter string against
  every equally long window of the longer one, windows allowed to hang over
  both ends) is restated here and used only when the real package is absent.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

try:  # the reference's own dependency, when present
    from thefuzz import fuzz as _fuzz  # type: ignore

    def partial_ratio(a: str, b: str) -> float:
        return _fuzz.partial_ratio(a, b)
except Exception:  # pragma: no cover - exercised in this image
    _fuzz = None

    def _lcs_len(a: str, b: str) -> int:
        """Longest common subsequence length, bit-parallel (Hyyro)."""
        if not a or not b:
            return 0
        masks = {}
        for i, ch in enumerate(a):
            masks[ch] = masks.get(ch, 0) | (1 << i)
        full = (1 << len(a)) - 1
        s = full
        for ch in b:
            m = masks.get(ch, 0)
            u = s & m
            s = ((s + u) | (s - u)) & full
        return len(a) - bin(s).count("1")

    def _ratio(a: str, b: str) -> float:
        total = len(a) + len(b)
        if total == 0:
            return 100.0
        return 100.0 * (2.0 * _lcs_len(a, b)) / total

    def partial_ratio(a: str, b: str) -> float:
        a, b = str(a), str(b)
        if len(a) == len(b) and a != b:
            # equal lengths: either string may be the one that is windowed; the scorer tries both and keeps the better
            return max(_partial_ratio_windows(a, b), _partial_ratio_windows(b, a))
        return _partial_ratio_windows(a, b)

    def _partial_ratio_windows(a: str, b: str) -> float:
        short, long_ = (a, b) if len(a) <= len(b) else (b, a)
        m = len(short)
        if m == 0:
            return 100.0 if len(long_) == 0 else 0.0
        best = 0.0
        # windows of the shorter string's length sliding over the longer one,
        # including partial windows at both ends
        for start in range(-(m - 1), len(long_)):
            lo, hi = max(0, start), min(len(long_), start + m)
            if hi <= lo:
                continue
            r = _ratio(short, long_[lo:hi])
            if r > best:
                best = r
                if best >= 100.0:
                    break
        return int(round(best))


def autocut_scores(score_list: Sequence[float]) -> List[int]:
    """Indices to drop if the largest relative drop between consecutive scores
    exceeds 20 % (cut after it); [] otherwise.  Mirrors VDB:443-464 including
    its division by the previous score."""
    drops = [(score_list[i - 1] - score_list[i]) / score_list[i - 1] for i in range(1, len(score_list))]
    biggest = max(drops)
    if biggest > 0.2:
        return list(range(drops.index(biggest) + 1, len(score_list)))
    return []


def text_hash_scores(vectorizer, query: str, documents: Sequence[str]):
    """Cosine similarity of char 1..6-gram hashing features (VDB:388-408)."""
    if len(documents) == 0:
        return []

    def feats(text):
        return np.asarray(vectorizer.fit_transform([text]).toarray().sum(axis=0), dtype=np.float64)

    qv = feats(query)
    qv = qv / np.linalg.norm(qv)
    return [float(np.dot(qv, dv / np.linalg.norm(dv))) for dv in (feats(doc) for doc in documents)]


def hybrid_rerank(vectorizer, sentences, search_scores, query, k=5, weights=(0.80, 0.15, 0.05)):
    """Weighted blend of vector score, hashed-n-gram cosine and fuzzy partial
    ratio (VDB:413-441).  The reference stacks sentences and scores into ONE
    numpy array, which makes the scores strings and the sort lexicographic
    (VDB:429-432); that observable behaviour is kept.  Any failure returns the
    un-reranked head, as the reference does (VDB:439-441)."""
    try:
        hash_scores = text_hash_scores(vectorizer, query, sentences)
        fuzzy = [partial_ratio(query, doc) for doc in sentences]
        if len(hash_scores) == 0:
            return sentences[:k], search_scores[:k]
        w_search, w_hash, w_fuzzy = weights
        combined = (w_search * np.array(search_scores) + w_hash * np.array(hash_scores)
                    + w_fuzzy * np.array(fuzzy))
        table = np.column_stack((np.array(sentences), np.array(combined)))
        table = table[table[:, 1].argsort()[::-1]]
        out_sentences, out_scores = zip(*table)
        return out_sentences[:k], out_scores[:k]
    except Exception:
        return sentences[:k], search_scores[:k]

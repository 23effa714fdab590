"""Columnar evaluation of MiniVectorDB's mongo-like metadata filters.

The reference evaluates filters with per-row Python loops over sets of row
numbers (ref minivectordb/vector_database.py:157-386) and hands the resulting
set to the scan, which then gathers the admissible rows.  Here the OUTPUT
contract is the same -- the admissible set -- but it is produced as a boolean
column over rows so it can be shipped to the GPU as a bitmask and applied in
the scan's epilogue.  Semantics kept identical to the reference:

* AND filters (`metadata_filter`): dict or list of dicts, every key must match
  (VDB:238-318).  A value that is a dict is an operator clause; ONLY ITS FIRST
  operator is honoured (VDB:243).  Anything else is an equality test.
* OR filters (`or_filters`): dict or list of dicts; the union over all dicts
  AND over all keys inside a dict (VDB:157-236), intersected with the AND
  result (VDB:373-377).  Empty dicts are dropped (VDB:371).
* Exclude filter: dict or list of dicts, equality only, union is removed
  (VDB:320-352).
* Operators $gt $gte $lt $lte $ne $in (VDB:166-173); `$in` is "operand is
  contained in the stored value" (VDB:172); unknown -> ValueError (VDB:175).
* Only rows whose metadata HAS the key can match (inverted index, VDB:260), so
  `$ne` does not match rows lacking the key.

A column caches a typed numpy view (numbers, strings) so the common predicates
are single vectorised comparisons; every other value type falls back to a
Python loop over that column only, with the reference's exact operator calls
(so errors such as comparing incomparable types surface the same way).
"""
from __future__ import annotations

import operator
from typing import Any, Dict, List, Optional

import numpy as np

_OPS = {
    "$gt": operator.gt,
    "$gte": operator.ge,
    "$lt": operator.lt,
    "$lte": operator.le,
    "$ne": operator.ne,
    "$in": lambda stored, operand: operand in stored,
}
_NUM_TYPES = (int, float, np.integer, np.floating)
_EXACT_INT = 1 << 53


def _is_plain_number(v) -> bool:
    return isinstance(v, _NUM_TYPES) and not (isinstance(v, (int, np.integer)) and abs(int(v)) > _EXACT_INT)


class Column:
    """All rows that carry one metadata key: parallel lists (row id, value)."""

    __slots__ = ("rows", "vals", "_rows_np", "_vals_np", "_kind", "_built", "_dev", "_dev_rows",
                 "_post", "_post_plain", "_post_built", "_vocab", "_codes_np")

    def __init__(self):
        self.rows: List[int] = []
        self.vals: List[Any] = []
        self._rows_np = None
        self._vals_np = None
        self._kind = None
        self._built = 0
        self._dev = None        # DeviceColumn mirror (numeric columns only)
        self._dev_rows = 0      # rows [0, _dev_rows) of the database are mirrored
        self._vocab = None      # string columns: value -> code (dictionary encoding) ...
        self._codes_np = None   # ... and the float64 code of every entry (what the device mirror holds)
        self._post = None       # `$in` postings: element -> entry indices, for list / tuple / set values
        self._post_plain = None  # entry indices whose value is anything else (str: substring test, ...)
        self._post_built = 0

    def append(self, row: int, value) -> None:
        self.rows.append(row)
        self.vals.append(value)

    @staticmethod
    def _convert(vals):
        """(kind, typed array or None) of a run of values: "num" (float64), "str" (object array) or
        "obj" (no typed view).  Numbers take numpy's own conversion as the fast path: a numeric result
        dtype proves that every element was a plain number (a str / None / list anywhere yields a
        string or object dtype instead); integers beyond 2^53 are left to the exact Python path."""
        first = vals[0]
        if isinstance(first, _NUM_TYPES):
            try:
                arr = np.asarray(vals)
            except (ValueError, TypeError, OverflowError):
                arr = None
            if arr is not None and arr.ndim == 1 and arr.dtype.kind in "iufb":
                if arr.dtype.kind not in "iu" or (arr.size and np.abs(arr).max() <= _EXACT_INT):
                    return "num", arr.astype(np.float64)
                return "obj", None
        if all(_is_plain_number(v) for v in vals):   # bool is an int: True == 1, as in Python
            return "num", np.asarray(vals, dtype=np.float64)
        if all(isinstance(v, str) for v in vals):
            return "str", np.asarray(vals, dtype=object)
        return "obj", None

    def _typed(self):
        """Typed numpy view of the column, extended INCREMENTALLY: only the entries appended since
        the last call are converted (a database that keeps growing between filtered queries would
        otherwise pay O(rows) Python-object conversion per query)."""
        n = len(self.rows)
        lo = self._built
        if lo == n:
            return self._kind
        tail_rows = np.asarray(self.rows[lo:], dtype=np.int64)
        kind, tail = self._convert(self.vals[lo:])
        if lo and kind != self._kind:
            kind, tail = "obj", None   # the column stopped being homogeneous
        self._rows_np = tail_rows if lo == 0 else np.concatenate((self._rows_np, tail_rows))
        if tail is None:
            self._vals_np = None
        else:
            self._vals_np = tail if lo == 0 else np.concatenate((self._vals_np, tail))
        if kind == "str":
            # dictionary encoding: equality on strings becomes equality on small integers, which numpy
            # does at memory speed and the device predicate kernel can do on its mirror of the codes
            if lo == 0 or self._vocab is None:
                self._vocab, self._codes_np = {}, np.zeros(0, dtype=np.float64)
            vocab = self._vocab
            codes = np.fromiter((vocab.setdefault(v, len(vocab)) for v in tail), dtype=np.float64, count=len(tail))
            self._codes_np = np.concatenate((self._codes_np, codes))
        else:
            self._vocab = self._codes_np = None
        self._kind = kind
        self._built = n
        return kind

    def device_match(self, engine, nrows: int, op: Optional[str], operand):
        """The clause as a device-resident mask (MaskHandle).  Numeric column + numeric operand:
        evaluated by a CUDA kernel on the HBM-resident column (which is extended lazily with the
        rows appended since the last filter); anything else: evaluated here and uploaded."""
        kind = self._typed() if self.rows else None
        numeric = kind == "num" and _is_plain_number(operand) and op != "$in"
        coded = kind == "str" and isinstance(operand, str) and op in (None, "$ne")
        if numeric or coded:
            src = self._vals_np if numeric else self._codes_np   # a column is one or the other for its whole life
            if self._dev is None:
                self._dev, self._dev_rows = engine.column(), 0
            if self._dev_rows < nrows:
                lo = int(np.searchsorted(self._rows_np, self._dev_rows))
                m = nrows - self._dev_rows
                vals = np.zeros(m, dtype=np.float64)
                present = np.zeros(m, dtype=np.uint8)
                idx = self._rows_np[lo:] - self._dev_rows
                keep = idx < m
                vals[idx[keep]] = src[lo:][keep]
                present[idx[keep]] = 1
                self._dev.append(vals, present)
                self._dev_rows = nrows
            # a string nobody stored has no code: -1 equals nothing and differs from everything present
            return self._dev.predicate(op, float(operand) if numeric else float(self._vocab.get(operand, -1)))
        self._dev = None   # the column stopped being purely numeric (or never was)
        return engine.mask_handle(self.match(nrows, op, operand))

    def match(self, nrows: int, op: Optional[str], operand) -> np.ndarray:
        """bool[nrows]: rows of this column whose value satisfies the clause
        (op None = equality)."""
        out = np.zeros(nrows, dtype=bool)
        if not self.rows:
            return out
        kind = self._typed()
        hit = None
        if kind == "num" and _is_plain_number(operand) and op != "$in":
            v, x = self._vals_np, float(operand)
            hit = {None: lambda: v == x, "$gt": lambda: v > x, "$gte": lambda: v >= x, "$lt": lambda: v < x,
                   "$lte": lambda: v <= x, "$ne": lambda: v != x}[op]()
        elif kind == "str" and isinstance(operand, str) and op in (None, "$ne"):
            code = float(self._vocab.get(operand, -1))
            hit = (self._codes_np == code) if op is None else (self._codes_np != code)
        if hit is None and op == "$in":
            hit = self._match_in(operand)
        if hit is None:
            # generic path: the reference's own operator call per stored value
            fn = (lambda stored, x: stored == x) if op is None else _OPS[op]
            hit = np.fromiter((bool(fn(s, operand)) for s in self.vals), dtype=bool, count=len(self.vals))
        out[self._rows_np[hit]] = True
        return out

    def _match_in(self, operand):
        """`operand in stored` (VDB:172) through postings.  Stored lists / tuples / sets of hashable
        elements are indexed element -> entries once (extended incrementally), so a tag filter costs
        O(hits) instead of a Python call per row; `x in [a, b]` compares with ==, which is what a
        dict lookup does for hashable values.  Every other stored value (a str means a SUBSTRING
        test, an int raises TypeError, ...) still goes through the reference's own operator call.
        Returns bool[len(vals)], or None when the operand cannot be looked up (unhashable)."""
        try:
            hash(operand)
        except TypeError:
            return None
        if operand != operand:     # NaN: `in` falls back on identity, a lookup would not
            return None
        if self._post is None:
            self._post, self._post_plain, self._post_built = {}, [], 0
        post, plain = self._post, self._post_plain
        vals = self.vals
        for i in range(self._post_built, len(vals)):
            s = vals[i]
            if type(s) in (list, tuple, set, frozenset):
                try:
                    ok = all(hash(e) is not None and e == e for e in s)   # hashable, and no NaN (identity semantics)
                except TypeError:
                    ok = False
                if ok:
                    for e in s:
                        lst = post.get(e)
                        if lst is None:
                            post[e] = [i]
                        elif lst[-1] != i:
                            lst.append(i)
                    continue
            plain.append(i)
        self._post_built = len(vals)
        hit = np.zeros(len(vals), dtype=bool)
        idx = post.get(operand)
        if idx:
            hit[np.asarray(idx, dtype=np.int64)] = True
        fn = _OPS["$in"]
        for i in plain:
            if fn(vals[i], operand):
                hit[i] = True
        return hit


class FilterIndex:
    """Metadata columns of one database + the three filter combinators."""

    def __init__(self):
        self.columns: Dict[str, Column] = {}

    def clear(self) -> None:
        self.close_device()
        self.columns = {}

    def close_device(self) -> None:
        """Release the HBM mirrors of the columns now (they point into their index, which the
        caller is about to destroy or compact)."""
        for col in self.columns.values():
            dev, col._dev, col._dev_rows = col._dev, None, 0
            if dev is not None:
                dev.close()

    @staticmethod
    def new_column() -> "Column":
        return Column()

    def add_row(self, row: int, metadata: dict) -> None:
        for key, value in metadata.items():
            col = self.columns.get(key)
            if col is None:
                col = self.columns[key] = Column()
            col.append(row, value)

    # -- clause -------------------------------------------------------------
    def _clause(self, nrows: int, key, value) -> np.ndarray:
        if isinstance(value, dict):
            op = next(iter(value))  # only the first operator counts (VDB:243)
            if op not in _OPS:
                raise ValueError(f"Invalid operator: {op}")
            operand = value[op]
        else:
            op, operand = None, value
        col = self.columns.get(key)
        if col is None:
            return np.zeros(nrows, dtype=bool)
        return col.match(nrows, op, operand)

    def _clause_device(self, engine, nrows: int, key, value):
        if isinstance(value, dict):
            op = next(iter(value))
            if op not in _OPS:
                raise ValueError(f"Invalid operator: {op}")
            operand = value[op]
        else:
            op, operand = None, value
        col = self.columns.get(key)
        if col is None:
            return engine.mask_handle(np.zeros(nrows, dtype=bool))   # nothing carries this key (full-size: it may be OR-ed into)
        return col.device_match(engine, nrows, op, operand)

    def admissible_device(self, engine, nrows: int, metadata_filter, exclude_filter, or_filters):
        """Same combinators as `admissible`, evaluated on the device: returns a MaskHandle over
        rows [0, nrows) (tombstones are applied by the engine, not here)."""
        # "no AND filter" is decided BEFORE a dict is wrapped into a list, exactly as the
        # reference does (VDB:356-360): {} starts from every row, [{}] from nothing
        cur = None if metadata_filter else engine.mask_filled(nrows)
        if isinstance(metadata_filter, dict):
            metadata_filter = [metadata_filter]
        if metadata_filter:
            for clause_set in metadata_filter:
                for key, value in clause_set.items():
                    hit = self._clause_device(engine, nrows, key, value)
                    cur = hit if cur is None else cur.iand(hit)
        if or_filters:
            if isinstance(or_filters, dict):
                or_filters = [or_filters]
            or_filters = [f for f in or_filters if f]
            if or_filters:
                union = None
                for clause_set in or_filters:
                    for key, value in clause_set.items():
                        hit = self._clause_device(engine, nrows, key, value)
                        union = hit if union is None else union.ior(hit)
                cur = union if cur is None else cur.iand(union)
        if exclude_filter:
            if isinstance(exclude_filter, dict):
                exclude_filter = [exclude_filter]
            if cur is None:
                raise TypeError("unsupported operand type(s) for -=: 'NoneType' and 'set'")
            for clause_set in exclude_filter:
                for key, value in clause_set.items():
                    col = self.columns.get(key)
                    if col is not None:
                        cur.iandnot(col.device_match(engine, nrows, None, value))   # equality only (VDB:343)
        if cur is None:
            return engine.mask_filled(0)
        return cur

    # -- combinators ----------------------------------------------------------
    def admissible(self, live: np.ndarray, metadata_filter, exclude_filter, or_filters) -> Optional[np.ndarray]:
        """bool[nrows] of admissible rows (already restricted to `live`), or
        None when every live row is admissible (no mask needed)."""
        nrows = live.shape[0]
        cur: Optional[np.ndarray] = None if metadata_filter else live.copy()
        unfiltered = not metadata_filter
        if isinstance(metadata_filter, dict):
            metadata_filter = [metadata_filter]
        if metadata_filter:
            for clause_set in metadata_filter:
                for key, value in clause_set.items():
                    hit = self._clause(nrows, key, value) & live
                    cur = hit if cur is None else (cur & hit)
                    if not cur.any():
                        break
        if or_filters:
            if isinstance(or_filters, dict):
                or_filters = [or_filters]
            or_filters = [f for f in or_filters if f]
            if or_filters:
                union = np.zeros(nrows, dtype=bool)
                for clause_set in or_filters:
                    for key, value in clause_set.items():
                        union |= self._clause(nrows, key, value)
                union &= live
                cur = union if cur is None else (cur & union)
                unfiltered = False
        if exclude_filter:
            if isinstance(exclude_filter, dict):
                exclude_filter = [exclude_filter]
            if cur is None:
                # the reference fails here too (`None -= set`, VDB:348)
                raise TypeError("unsupported operand type(s) for -=: 'NoneType' and 'set'")
            for clause_set in exclude_filter:
                for key, value in clause_set.items():
                    col = self.columns.get(key)
                    if col is not None:
                        cur &= ~col.match(nrows, None, value)  # equality only (VDB:343)
                        unfiltered = False
                    if not cur.any():
                        break
        if cur is None:
            return np.zeros(nrows, dtype=bool)
        if unfiltered:
            return None
        return cur

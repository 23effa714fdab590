"""CPU: columnar filter evaluation details that the drop-in scenarios do not pin."""
import numpy as np

from minivectordb_b200.filters import Column


def _fresh(rows, vals):
    c = Column()
    for r, v in zip(rows, vals):
        c.append(r, v)
    return c


def test_incrementally_typed_column_equals_a_fresh_one():
    """A column converts only the entries appended since its last use; whatever the append /
    query interleaving and however the value types drift (numbers -> strings -> lists), every
    clause must give what a column built in one go gives -- including the TypeError the
    reference's operator call raises for incomparable types."""
    rng = np.random.default_rng(0)
    seqs = {
        "num": [int(v) for v in rng.integers(0, 100, 3000)],
        "float_and_bool": [float(v) for v in rng.random(500)] + [True, False] * 250,
        "str": ["t%d" % v for v in rng.integers(0, 16, 3000)],
        "num_then_str": [1, 2, 3.5] * 500 + ["a", "b"] * 700,
        "str_then_lists": ["a"] * 1000 + [["x", "y"]] * 1000 + ["b"] * 500,
    }
    for name, vals in seqs.items():
        inc = Column()
        i = 0
        for upto in (3, 21, 1500, len(vals)):
            while i < min(len(vals), upto):
                inc.append(2 * i, vals[i])
                i += 1
            ref = _fresh([2 * j for j in range(i)], vals[:i])
            for op, operand in ((None, vals[0]), ("$ne", vals[0]), ("$gt", 40), ("$lte", 0.5), ("$in", "x")):
                got = want = None
                try:
                    got = inc.match(2 * i + 5, op, operand)
                except TypeError:
                    got = "TypeError"
                try:
                    want = ref.match(2 * i + 5, op, operand)
                except TypeError:
                    want = "TypeError"
                if isinstance(got, str) or isinstance(want, str):
                    assert got == want, (name, upto, op)
                else:
                    assert np.array_equal(got, want), (name, upto, op)


def test_in_operator_postings_equal_the_per_row_operator_call():
    """`{"tags": {"$in": "t3"}}` is `operand in stored` per row in the reference (VDB:172).  Lists /
    tuples / sets of hashable elements are served from element -> rows postings; strings (substring
    test), dicts (key test), nested or NaN-carrying lists and unhashable operands keep the exact
    operator call.  Values AND TypeErrors must match the per-row evaluation, also when the column
    keeps growing between queries."""
    from minivectordb_b200.filters import _OPS
    rng = np.random.default_rng(1)

    def generic(c, n, operand):
        out = np.zeros(n, dtype=bool)
        hit = np.fromiter((bool(_OPS["$in"](s, operand)) for s in c.vals), dtype=bool, count=len(c.vals))
        out[np.asarray(c.rows)[hit]] = True
        return out

    def same(c, n, operand):
        try:
            a = c.match(n, "$in", operand)
        except TypeError:
            a = "TypeError"
        try:
            b = generic(c, n, operand)
        except TypeError:
            b = "TypeError"
        if isinstance(a, str) or isinstance(b, str):
            return isinstance(a, str) and isinstance(b, str)
        return np.array_equal(a, b)

    for with_str in (False, True):
        vals = []
        for _ in range(3000):
            r = rng.random()
            if r < 0.6:
                vals.append(["t%d" % rng.integers(0, 16), "t%d" % rng.integers(0, 16)])
            elif r < 0.7:
                vals.append(("x", 1, 2.0, True))
            elif r < 0.8:
                vals.append({"t3", "zz"})
            elif r < 0.88:
                vals.append("t3 and t5 substring" if with_str else ["t3", "t3"])
            elif r < 0.93:
                vals.append([["nested"], "t3"])
            elif r < 0.96:
                vals.append([float("nan"), "t3"])
            else:
                vals.append({"t3": 1})
        c = Column()
        k = 0
        for upto in (10, 700, len(vals)):
            while k < upto:
                c.append(3 * k, vals[k])
                k += 1
            for operand in ("t3", "t15", 1, 1.0, True, "zz", "nested", ("a",), "sub", 2, ["nested"], float("nan")):
                assert same(c, 3 * k + 1, operand), (with_str, upto, operand)


def test_dictionary_coded_string_columns():
    """String equality / $ne run on integer codes (host) and on the device mirror of the codes; the
    answers must be those of comparing the strings, for values seen, never seen, and after growth."""
    rng = np.random.default_rng(2)
    vals = ["t%d" % v for v in rng.integers(0, 9, 4000)]
    c = Column()
    k = 0
    for upto in (5, 900, len(vals)):
        while k < upto:
            c.append(2 * k + 1, vals[k])
            k += 1
        arr = np.asarray(vals[:k], dtype=object)
        for operand in ("t3", "t8", "never", ""):
            for op, want in ((None, arr == operand), ("$ne", arr != operand)):
                out = np.zeros(2 * k + 2, dtype=bool)
                out[2 * np.flatnonzero(want) + 1] = True
                assert np.array_equal(c.match(2 * k + 2, op, operand), out), (upto, op, operand)
    # a non-string operand on a string column keeps the generic comparison (never equal)
    assert not c.match(2 * k + 2, None, 3).any() and c.match(2 * k + 2, "$ne", 3).sum() == k

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

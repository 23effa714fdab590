"""Counter-based synthetic vectors -- numpy twin of the
CUDA generator in csrc/device_utils.cuh (synth_value).  Integer-only, so host
and device produce identical bits without shipping data."""
import numpy as np

DIST_BELL, DIST_UNIFORM = 0, 1


def synth_rows(seed: int, row0: int, n: int, d: int, dist: int = DIST_BELL) -> np.ndarray:
    M = np.uint64
    with np.errstate(over="ignore"):
        rows = np.arange(row0, row0 + n, dtype=np.uint64)[:, None] << M(20)
        z = rows + np.arange(d, dtype=np.uint64)[None, :] + M(seed) * M(0x9E3779B97F4A7C15)
        z = (z ^ (z >> M(30))) * M(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> M(27))) * M(0x94D049BB133111EB)
        z = z ^ (z >> M(31))
    if dist == DIST_BELL:
        s = ((z & M(0xFFFF)) + ((z >> M(16)) & M(0xFFFF)) + ((z >> M(32)) & M(0xFFFF)) + (z >> M(48)))
        return (s.astype(np.int64) - 131070).astype(np.float32) * np.float32(1.0 / 65536.0)
    return (z >> M(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def synth_mask(seed: int, n: int, keep: float = 0.5) -> np.ndarray:
    """bool[n] admissible rows, ~keep fraction (stand-in for a metadata filter such as
    {"value": {"$gt": 49}} over uniform ints; ref tests/test_mongolike_operators.py:41-60)."""
    out = np.empty(n, dtype=bool)
    step = 1 << 20
    for a in range(0, n, step):
        m = min(step, n - a)
        out[a:a + m] = synth_rows(seed, a, m, 1, DIST_UNIFORM)[:, 0] < keep
    return out

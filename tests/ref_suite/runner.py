"""Run the REFERENCE's own test files (read at run time from /root/reference or from the pip-installed
copy under baseline/_ref; never part of this repo's history) against this repo's engine.  Invoked as a subprocess by tests/test_reference_suite.py:

    python tests/ref_suite/runner.py <mode> <workdir> <reference test files...> [-- pytest args]

mode "dropin": `minivectordb.vector_database` / `minivectordb.sharded_vector_database` resolve to the drop-in
               classes of minivectordb_b200 (on the real CUDA engine when a GPU is visible, else on the
               oracle-backed FakeEngine: host logic only);
mode "shim":   the reference's OWN classes are imported unmodified and `faiss` resolves to
               minivectordb_b200.faiss_shim (needs a GPU) -- INTEGRATION.md mode 1.
In both modes `thefuzz` and `minivectordb.embedding_model` are stand-ins (neither is installable here; the
ONNX model blob is absent from the reference checkout): partial_ratio is this repo's restatement and
EmbeddingModel produces bag-of-words hash vectors of the model's dimension.
"""
import os
import sys
import types
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def reference_dir():
    """Where the reference's package + tests can be read from: the checkout in the build container, or
    the copy `pip install --target baseline/_ref` made of it (git-ignored; it travels to the GPU box)."""
    for cand in (os.environ.get("MVDB_REFERENCE_DIR"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "tests")) and os.path.isdir(os.path.join(cand, "minivectordb")):
            return cand
    return None


REF = reference_dir()


class FakeEmbeddingModel:
    """Deterministic bag-of-words embedding: every word is a fixed random direction, a text is the
    normalised sum of its words -- shared words give high cosine similarity, which is all the
    reference's ranking assertions need."""

    def __init__(self, use_quantized_onnx_model=True, alternative_model=None, onnx_model_cpu_core_count=None, **kwargs):
        size = kwargs.get("e5_model_size")
        self.dim = 512 if use_quantized_onnx_model else {"small": 384, "large": 1024}.get(size, 1024)

    # the little "semantics" the reference's ranking assertions rely on ("i like dogs" is closer to "i like
    # animals" than to "i like cars", ref tests/test_vector_database.py:195-218): words of one concept share a direction
    CONCEPTS = {w: "animal" for w in ("dog", "dogs", "animal", "animals", "cat", "cats", "bird", "lion", "panther", "lizard",
                                      "hippo", "dinosaur", "worm", "bug", "mammoth")}

    def _word(self, w):
        v = np.random.default_rng(zlib.crc32(w.encode())).standard_normal(self.dim)
        c = self.CONCEPTS.get(w)
        if c is not None:
            v = 0.5 * v + np.random.default_rng(zlib.crc32(("concept:" + c).encode())).standard_normal(self.dim)
        return v

    def extract_embeddings(self, text):
        v = np.zeros(self.dim, dtype=np.float64)
        for w in str(text).lower().replace(".", " ").replace(",", " ").split():
            v += self._word(w)
        n = np.linalg.norm(v)
        return (v / n if n > 0 else v).astype(np.float32)


def install(mode):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from minivectordb_b200 import _native, rerank
    has_gpu = _native.device_count() > 0
    fuzz = types.ModuleType("thefuzz.fuzz")
    fuzz.partial_ratio = rerank.partial_ratio
    thefuzz = types.ModuleType("thefuzz")
    thefuzz.fuzz = fuzz
    sys.modules["thefuzz"], sys.modules["thefuzz.fuzz"] = thefuzz, fuzz
    emb = types.ModuleType("minivectordb.embedding_model")
    emb.EmbeddingModel = FakeEmbeddingModel
    emb.AlternativeModel = types.SimpleNamespace(small="small", large="large", bgem3="bgem3")
    if mode == "shim":
        if not has_gpu:
            raise SystemExit("mode shim needs a GPU (the shim has no CPU fallback)")
        from minivectordb_b200 import faiss_shim
        faiss_shim.install_as_faiss()
        sys.path.insert(0, REF)
        import minivectordb   # the reference package itself
        sys.modules["minivectordb.embedding_model"] = emb
        minivectordb.embedding_model = emb
        return "reference classes on faiss_shim (CUDA engine)"
    if not has_gpu:
        import minivectordb_b200._store as store
        from fake_engine import FakeEngine
        store.FlatIPEngine = FakeEngine
    from minivectordb_b200.vector_database import VectorDatabase
    from minivectordb_b200.sharded_vector_database import ShardedVectorDatabase
    pkg = types.ModuleType("minivectordb")
    pkg.__path__ = []
    vdb = types.ModuleType("minivectordb.vector_database")
    vdb.VectorDatabase = VectorDatabase
    svdb = types.ModuleType("minivectordb.sharded_vector_database")
    svdb.ShardedVectorDatabase = ShardedVectorDatabase
    pkg.vector_database, pkg.sharded_vector_database, pkg.embedding_model = vdb, svdb, emb
    sys.modules.update({"minivectordb": pkg, "minivectordb.vector_database": vdb,
                        "minivectordb.sharded_vector_database": svdb, "minivectordb.embedding_model": emb})
    return "drop-in classes on " + ("the CUDA engine" if has_gpu else "the oracle-backed FakeEngine (host logic)")


def main():
    mode, workdir = sys.argv[1], sys.argv[2]
    rest = sys.argv[3:]
    extra = []
    if "--" in rest:
        i = rest.index("--")
        rest, extra = rest[:i], rest[i + 1:]
    what = install(mode)
    os.chdir(workdir)   # the reference's tests write db.pkl / shard directories into the cwd
    import pytest
    print(f"[ref_suite] {mode}: {what}", flush=True)
    files = [os.path.join(REF, "tests", f) for f in rest]
    sys.exit(pytest.main(files + ["-q", "-p", "no:cacheprovider", "--rootdir", workdir, "-x"] + extra))


if __name__ == "__main__":
    main()

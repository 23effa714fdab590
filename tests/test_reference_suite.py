"""The REFERENCE's own test files, unmodified, run against this repo's engine.

The files are read at run time from the reference checkout (/root/reference, build container) or from the
copy that `pip install --target baseline/_ref` makes of it (git-ignored, travels to the GPU box; made by
__graft_entry__.build()); they are never part of this repo.  Where neither exists the tests skip.

* CPU (`-m "not gpu"`): drop-in classes on the oracle-backed FakeEngine -- pins the whole HOST side of the
  boundary (API surface, filters, id maps, renumbering, persistence formats, thread safety).
* GPU (`-m gpu`): the same files on the real CUDA engine, twice: through the drop-in classes, and through
  the reference's OWN classes with `faiss` resolved to minivectordb_b200.faiss_shim (INTEGRATION.md mode 1).
"""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "ref_suite"))
import runner  # noqa: E402

REF = runner.reference_dir()
needs_ref = pytest.mark.skipif(REF is None, reason="no reference checkout / baseline/_ref install present")

# need real semantic embeddings (autocut / hybrid-rerank ranking assertions, ref tests/test_vector_database.py:304-323)
SEMANTIC = ["test_similarity_search_with_hybrid_reranking"]
VDB_FILES = ["test_vector_database.py", "test_mongolike_operators.py"]
SVDB_FILES = ["test_sharded_vector_database.py", "test_sharded_mongolike_operators.py"]
MT_FILES = ["test_multithreaded_operations.py", "test_sharded_multithreaded_operations.py"]


def _run(mode, files, tmp_path, timeout=900):
    cmd = [sys.executable, os.path.join(HERE, "ref_suite", "runner.py"), mode, str(tmp_path)] + files
    cmd += ["--", "-k", " and ".join(f"not {name}" for name in SEMANTIC)]
    env = {**os.environ, "PYTHONDONTWRITEBYTECODE": "1"}
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    tail = (p.stdout + p.stderr)[-3000:]
    assert p.returncode == 0, tail
    assert " passed" in p.stdout and " failed" not in p.stdout, tail
    return p.stdout


def _has_gpu():
    from minivectordb_b200 import _native
    return _native.device_count() > 0


# ---- CPU: host logic ------------------------------------------------------------------------------
@needs_ref
def test_reference_vdb_and_filter_tests_pass_on_the_dropin_classes(tmp_path):
    if _has_gpu():
        pytest.skip("covered by the gpu-marked variant on this box")
    assert "27 passed" in _run("dropin", VDB_FILES, tmp_path)      # 25 + 3 tests, one semantic test deselected


@needs_ref
def test_reference_sharded_tests_pass_on_the_dropin_classes(tmp_path):
    if _has_gpu():
        pytest.skip("covered by the gpu-marked variant on this box")
    assert "30 passed" in _run("dropin", SVDB_FILES, tmp_path)     # 28 + 3 tests, one semantic test deselected


@needs_ref
def test_reference_multithreaded_tests_pass_on_the_dropin_classes(tmp_path):
    if _has_gpu():
        pytest.skip("covered by the gpu-marked variant on this box")
    assert "2 passed" in _run("dropin", MT_FILES, tmp_path, timeout=1500)


# ---- GPU: the same files on the real engine -----------------------------------------------------------
@needs_ref
@pytest.mark.gpu
def test_reference_tests_on_gpu_through_the_dropin_classes(tmp_path):
    assert "27 passed" in _run("dropin", VDB_FILES, tmp_path)
    assert "30 passed" in _run("dropin", SVDB_FILES, tmp_path)
    assert "2 passed" in _run("dropin", MT_FILES, tmp_path, timeout=1500)


@needs_ref
@pytest.mark.gpu
def test_reference_classes_unmodified_on_the_faiss_shim(tmp_path):
    """INTEGRATION.md mode 1: the reference's own VectorDatabase / ShardedVectorDatabase, `import faiss`
    resolved to the shim over the C ABI."""
    assert "27 passed" in _run("shim", VDB_FILES, tmp_path, timeout=1500)
    assert "30 passed" in _run("shim", SVDB_FILES, tmp_path, timeout=1500)

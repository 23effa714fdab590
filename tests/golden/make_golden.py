"""Generate tests/golden/*.json by running the REFERENCE's own Python classes
(imported from /root/reference, unmodified) through tests/golden/scenario.py.

The reference delegates its arithmetic to faiss-cpu, which is not installable
here; `faiss` is therefore provided by the oracle's faiss-shaped layer
(oracle.install_as_faiss), `thefuzz` by a stub (import-time only) and
`minivectordb.embedding_model` is never imported.  What these fixtures pin is
the reference's HOST logic around the scan -- filter semantics, row
renumbering on delete, search_k clamping, result mapping -- on top of the
oracle's scan.  Run inside the build container only:

    python tests/golden/make_golden.py
"""
import json
import os
import shutil
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import oracle as O  # noqa: E402

O.install_as_faiss()
fuzz_mod = types.ModuleType("thefuzz")
fuzz_mod.fuzz = types.SimpleNamespace(partial_ratio=lambda a, b: 0)
sys.modules["thefuzz"] = fuzz_mod
sys.path.insert(0, "/root/reference")

from minivectordb.vector_database import VectorDatabase as RefVDB  # noqa: E402
from minivectordb.sharded_vector_database import ShardedVectorDatabase as RefSVDB  # noqa: E402
import scenario  # noqa: E402

tmp = tempfile.mkdtemp()
try:
    rec = scenario.run(RefVDB, storage_file=os.path.join(tmp, "none.pkl"))
    json.dump(rec, open(os.path.join(HERE, "reference_vdb_scenario.json"), "w"))
    rec = scenario.run(RefSVDB, sharded=True, storage_dir=os.path.join(tmp, "shards"), shard_size=64)
    json.dump(rec, open(os.path.join(HERE, "reference_svdb_scenario.json"), "w"))
    # a persisted reference DB, to check that our loader reads the reference's pickles
    emb, meta, ids, deletes, queries = scenario.make_rows()
    db = RefVDB(storage_file=os.path.join(HERE, "reference_db.pkl"))
    if not db.id_map:
        db.store_embeddings_batch(ids[:40], [e for e in emb[:40]], meta[:40])
        db.persist_to_disk()
finally:
    shutil.rmtree(tmp, ignore_errors=True)
print("golden fixtures written to", HERE)

"""Drop-in `VectorDatabase` on the B200 engine.

Same public surface as the reference class (ref minivectordb/
vector_database.py:7-548): constructor argument, method names, argument
meaning, return shapes and error behaviour.  The flat inner-product scan, the
matrix and the delete bookkeeping live on the GPU (see _store.py); filter
parsing and the rerank helpers stay host Python as BASELINE.json's north_star
prescribes.
"""
from __future__ import annotations

import os
import pickle
from collections import defaultdict

import numpy as np
from sklearn.feature_extraction.text import HashingVectorizer

from . import rerank as _rerank
from ._store import GpuStore


class VectorDatabase(GpuStore):
    def __init__(self, storage_file='db.pkl', device: int = 0, scan_shadow: bool = False):
        """`scan_shadow=True` (extension, off by default): single queries stream an int8 shadow of the matrix
        and re-score a rigorous candidate superset in fp32 -- same ids and distances, ~1/4 of the HBM bytes."""
        super().__init__(devices=[device], scan_shadow=scan_shadow)
        # same featuriser the reference builds (VDB:9)
        self.hash_vectorizer = HashingVectorizer(ngram_range=(1, 6), analyzer='char', n_features=64)
        self.storage_file = storage_file
        self._load_database()

    # -- views the reference exposes as plain attributes (VDB:12-16) ----------
    @property
    def id_map(self):
        with self.lock:
            return self._build_views()[0]

    @property
    def inverse_id_map(self):
        with self.lock:
            return self._build_views()[1]

    @property
    def metadata(self):
        with self.lock:
            return self._build_views()[2]

    # -- persistence: the reference's single-pickle format (VDB:28-40, 538-548) --
    def _load_database(self):
        if not os.path.exists(self.storage_file):
            return
        with open(self.storage_file, 'rb') as f:
            data = pickle.load(f)
        emb = data['embeddings']
        if emb is None:
            return
        emb = np.asarray(emb, dtype=np.float32)
        with self.lock:
            self.embedding_size = int(emb.shape[1])
            self._ever_stored = True
            id_map = data['id_map']
            self._append_batch([id_map[row] for row in range(emb.shape[0])], emb, data['metadata'])
            saved = data.get('inverted_index')
            if saved is not None:
                self.inverted_index = defaultdict(set, {k: set(v) for k, v in saved.items()})
            self._flush()  # the reference builds its index at load time (VDB:39-40)

    def persist_to_disk(self):
        with self.lock:   # one lock for the matrix AND the id views: a concurrent store cannot fall between them
            emb = self._materialize_locked() if self._ever_stored else None
            id_map, inverse_id_map, metadata, _ = self._build_views()
            data = {'embeddings': emb, 'metadata': metadata, 'id_map': id_map,
                    'inverse_id_map': inverse_id_map, 'inverted_index': self.inverted_index}
            with open(self.storage_file, 'wb') as f:
                pickle.dump(data, f)

    # -- rows -------------------------------------------------------------------
    def get_vector(self, unique_id):
        with self.lock:
            if unique_id not in self._uid_gid:
                raise ValueError("Unique ID does not exist.")
            return self._row_of_gid(self._uid_gid[unique_id])

    def store_embedding(self, unique_id, embedding, metadata_dict={}):
        with self.lock:
            if unique_id in self._uid_gid:
                raise ValueError("Unique ID already exists.")
            self._append(unique_id, self._as_row(embedding), metadata_dict)

    def store_embeddings_batch(self, unique_ids, embeddings, metadata_dicts=[]):
        with self.lock:
            for uid in unique_ids:
                if uid in self._uid_gid:
                    raise ValueError("Unique ID already exists.")
            if 0 < len(metadata_dicts) < len(unique_ids):
                raise ValueError("Metadata dictionaries must be provided for all unique IDs.")
            if len(metadata_dicts) == 0:
                metadata_dicts = [{} for _ in range(len(unique_ids))]
            block = self._as_rows(embeddings)
            if block.shape[0] != len(unique_ids) or len(metadata_dicts) != len(unique_ids):
                # the reference zips the three lists (VDB:100-107): extra entries are ignored
                m = min(block.shape[0], len(unique_ids), len(metadata_dicts))
                unique_ids, block, metadata_dicts = list(unique_ids)[:m], block[:m], list(metadata_dicts)[:m]
            self._append_batch(unique_ids, block, metadata_dicts)

    def delete_embedding(self, unique_id):
        with self.lock:
            if unique_id not in self._uid_gid:
                raise ValueError("Unique ID does not exist.")
            self._remove(unique_id)

    # -- search -------------------------------------------------------------------
    def find_most_similar(self, embedding, metadata_filter=None, exclude_filter=None, or_filters=None, k=5,
                          autocut=False):
        """Top-k cosine neighbours: (ids, np.float32 scores descending, metadata
        dicts) as tuples; three empty lists when nothing matches (VDB:466-536)."""
        return self._search(embedding, metadata_filter, exclude_filter, or_filters, k, autocut)

    # -- host-side helpers, unchanged semantics -------------------------------------
    def autocut_scores(self, score_list):
        return _rerank.autocut_scores(score_list)

    def hybrid_rerank_results(self, sentences, search_scores, query, k=5, weights=(0.80, 0.15, 0.05)):
        return _rerank.hybrid_rerank(self.hash_vectorizer, sentences, search_scores, query, k=k, weights=weights)

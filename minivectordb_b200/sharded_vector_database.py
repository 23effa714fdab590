"""Drop-in `ShardedVectorDatabase` on the B200 engine.

Public surface of the reference class (ref minivectordb/
sharded_vector_database.py:8-662).  Two different things are called "shard":

* the reference's ON-DISK shards -- `shard_{i}.pkl` files of `shard_size` rows
  each (SVDB:134-204) -- kept in the same format so that existing directories
  load and files written here load in the reference;
* the ROW SHARDS of the scan: with `devices=[0, 1, ...]` the matrix is split
  across several GPUs of one box (rows go to the least-filled device), every
  device scans its part and the per-device top-k are merged (score
  descending, insertion order on exact ties).  The reference has no
  counterpart: its search is one in-memory index over all rows (SVDB:79-84).
"""
from __future__ import annotations

import os
import pickle
from collections import defaultdict

import numpy as np
from sklearn.feature_extraction.text import HashingVectorizer

from . import rerank as _rerank
from ._store import GpuStore


class ShardedVectorDatabase(GpuStore):
    def __init__(self, storage_dir='db_shards', shard_size=5000, devices=None, persist=True, scan_shadow=False):
        super().__init__(devices=devices or [0], scan_shadow=scan_shadow)
        self.hash_vectorizer = HashingVectorizer(ngram_range=(1, 6), analyzer='char', n_features=64)
        self.storage_dir = storage_dir
        self.shard_size = shard_size
        self.persist = persist          # False: skip the per-mutation pickle rewrite (bulk loads)
        self.box_item_map = {}          # disk shard id -> list of unique ids (SVDB:22)
        self.inverse_box_item_map = {}  # unique id -> disk shard id (SVDB:23)
        self._load_database()

    # -- views ---------------------------------------------------------------------
    @property
    def unique_ids(self):
        with self.lock:
            return self._build_views()[3]

    @property
    def inverse_id_map(self):
        with self.lock:
            return self._build_views()[1]

    @property
    def metadata(self):
        with self.lock:
            return self._build_views()[2]

    # -- disk shards (format of SVDB:134-204) ------------------------------------------
    def _shard_path(self, shard_id):
        return os.path.join(self.storage_dir, f'shard_{shard_id}.pkl')

    def _read_shard(self, shard_id):
        path = self._shard_path(shard_id)
        if os.path.exists(path):
            with open(path, 'rb') as f:
                data = pickle.load(f)
            data['inverted_index'] = defaultdict(set, data['inverted_index'])
            return data
        return {'embeddings': np.zeros((0, self.embedding_size), dtype=np.float32), 'metadata': [],
                'unique_ids': [], 'inverted_index': defaultdict(set)}

    def _write_shard(self, shard_id, data):
        out = dict(data)
        out['inverted_index'] = dict(data['inverted_index'])
        with open(self._shard_path(shard_id), 'wb') as f:
            pickle.dump(out, f)

    def _load_database(self):
        if not os.path.exists(self.storage_dir):
            os.makedirs(self.storage_dir)
        files = [f for f in os.listdir(self.storage_dir) if f.endswith('.pkl')]
        files.sort(key=lambda name: int(name.split('_')[1].split('.')[0]))
        with self.lock:
            for name in files:
                shard_id = int(name.split('_')[1].split('.')[0])
                with open(os.path.join(self.storage_dir, name), 'rb') as f:
                    data = pickle.load(f)
                emb = np.asarray(data['embeddings'], dtype=np.float32)
                if emb.ndim == 2 and emb.shape[0] > 0 and self.embedding_size is None:
                    self.embedding_size = int(emb.shape[1])
                self.box_item_map[shard_id] = list(data['unique_ids'])
                for uid in data['unique_ids']:
                    self.inverse_box_item_map[uid] = shard_id
                if len(data['unique_ids']):
                    self._append_batch(data['unique_ids'], emb, data['metadata'])
            if self._n_live:
                self._flush()

    def _next_disk_shard(self):
        for shard_id, items in self.box_item_map.items():
            if len(items) < self.shard_size:
                return shard_id
        return len(self.box_item_map)

    def _persist_rows(self, uids, rows, metas):
        """Assign rows to disk shards (first non-full one, SVDB:98-102) and
        rewrite the touched shard files."""
        groups = defaultdict(list)
        for uid, row, meta in zip(uids, rows, metas):
            shard_id = self._next_disk_shard()
            self.box_item_map.setdefault(shard_id, []).append(uid)
            self.inverse_box_item_map[uid] = shard_id
            groups[shard_id].append((uid, row, meta))
        if not self.persist:
            return
        for shard_id, items in groups.items():
            data = self._read_shard(shard_id)
            data['embeddings'] = np.vstack([data['embeddings']] + [r[None, :] for _, r, _ in items])
            for uid, _, meta in items:
                data['metadata'].append(meta)
                data['unique_ids'].append(uid)
                for key in meta:
                    data['inverted_index'][key].add(uid)
            self._write_shard(shard_id, data)

    def _unpersist_rows(self, uids):
        groups = defaultdict(list)
        for uid in uids:
            groups[self.inverse_box_item_map[uid]].append(uid)
        for shard_id, gone in groups.items():
            gone_set = set(gone)
            self.box_item_map[shard_id] = [u for u in self.box_item_map[shard_id] if u not in gone_set]
            for uid in gone_set:
                del self.inverse_box_item_map[uid]
            if not self.persist:
                continue
            data = self._read_shard(shard_id)
            keep = [i for i, u in enumerate(data['unique_ids']) if u not in gone_set]
            data['embeddings'] = data['embeddings'][keep]
            data['metadata'] = [data['metadata'][i] for i in keep]
            data['unique_ids'] = [data['unique_ids'][i] for i in keep]
            for key in list(data['inverted_index']):
                data['inverted_index'][key] -= gone_set
                if not data['inverted_index'][key]:
                    del data['inverted_index'][key]
            self._write_shard(shard_id, data)

    def _convert_from_non_sharded_db(self, non_sharded_db_object):
        """Migrate every row of a VectorDatabase (SVDB:26-33)."""
        embeddings = np.asarray(non_sharded_db_object.embeddings)
        id_map = non_sharded_db_object.id_map
        unique_ids = [id_map[i] for i in range(len(embeddings))]
        self.store_embeddings_batch(unique_ids, embeddings, list(non_sharded_db_object.metadata))

    # -- rows ------------------------------------------------------------------------------
    def get_vector(self, unique_id):
        with self.lock:
            if unique_id not in self._uid_gid:
                raise ValueError("Unique ID does not exist.")
            if self.persist:
                # the reference answers from the shard file, i.e. the RAW stored row (SVDB:86-96)
                data = self._read_shard(self.inverse_box_item_map[unique_id])
                return data['embeddings'][data['unique_ids'].index(unique_id)]
            return self._row_of_gid(self._uid_gid[unique_id])

    def store_embedding(self, unique_id, embedding, metadata_dict={}):
        with self.lock:
            if unique_id in self._uid_gid:
                raise ValueError("Unique ID already exists.")
            row = self._as_row(embedding)
            self._append(unique_id, row, metadata_dict)
            self._persist_rows([unique_id], [row], [metadata_dict])

    def store_embeddings_batch(self, unique_ids: list, embeddings, metadata_dicts=[]):
        with self.lock:
            if len(unique_ids) != len(embeddings):
                raise ValueError("Number of unique IDs must match number of embeddings.")
            for uid in unique_ids:
                if uid in self._uid_gid:
                    raise ValueError(f"Unique ID {uid} already exists.")
            rows = self._as_rows(embeddings)
            # pad missing metadata with {} (SVDB:259-261) -- on a copy: the reference
            # extends the caller's (default!) list in place, which leaks rows between calls
            metas = list(metadata_dicts) + [{} for _ in range(len(unique_ids) - len(metadata_dicts))]
            self._append_batch(unique_ids, rows, metas[:len(unique_ids)])
            self._persist_rows(unique_ids, rows, metas)

    def delete_embeddings_batch(self, unique_ids):
        with self.lock:
            if not isinstance(unique_ids, list):
                unique_ids = [unique_ids]
            if not unique_ids:
                raise ValueError("No unique IDs provided.")
            if not all(uid in self._uid_gid for uid in unique_ids):
                raise ValueError("One or more unique IDs do not exist.")
            unique_ids = list(dict.fromkeys(unique_ids))
            self._unpersist_rows(unique_ids)
            for uid in unique_ids:
                self._remove(uid)

    # -- search ---------------------------------------------------------------------------
    def find_most_similar(self, embedding, metadata_filter=None, exclude_filter=None, or_filters=None, k=5,
                          autocut=False):
        return self._search(embedding, metadata_filter, exclude_filter, or_filters, k, autocut)

    def autocut_scores(self, score_list):
        return _rerank.autocut_scores(score_list)

    def hybrid_rerank_results(self, sentences, search_scores, query, k=5, weights=(0.80, 0.15, 0.05)):
        return _rerank.hybrid_rerank(self.hash_vectorizer, sentences, search_scores, query, k=k, weights=weights)

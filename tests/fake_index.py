"""Test double for `dvdb_b200.Index` on boxes without a GPU: same facade, arithmetic from the oracle.
Lives under tests/ -- the product never imports it."""
import numpy as np

from oracle import cpu_ref as R


class OracleIndex:
    def __init__(self, space="l2", dim=512, max_elements=1000, store_dtype="f32"):
        self.space, self.dim, self.cap, self.store = space, dim, max_elements, store_dtype
        self.rows, self.labels, self.dead = [], [], set()

    def get_current_count(self):
        return len(self.rows)

    def get_max_elements(self):
        return self.cap

    def resize_index(self, n):
        self.cap = n

    def add_items(self, data, ids):
        data = np.atleast_2d(np.asarray(data, np.float32))
        if len(self.rows) + len(data) > self.cap:
            raise RuntimeError("The number of elements exceeds the specified limit")
        for v, i in zip(data, np.asarray(ids).reshape(-1)):
            self.rows.append(R.prepare_rows(v[None, :], self.space, self.store)[0])
            self.labels.append(int(i))

    def mark_deleted(self, ids):
        self.dead.update(int(i) for i in np.asarray(ids).reshape(-1))

    def knn_query_padded(self, q, k):
        if not self.rows:
            nq = len(np.atleast_2d(q))
            return np.full((nq, k), -1, np.int64), np.full((nq, k), np.inf, np.float32), np.zeros(nq, np.int32)
        return R.knn_exact(q, np.stack(self.rows), np.array(self.labels), k, self.space, deleted=self.dead)

    def save_index(self, path):
        np.savez(path + ".npz", rows=np.array(self.rows, np.float32).reshape(-1, self.dim), labels=np.array(self.labels, np.int64),
                 dead=np.array(sorted(self.dead), np.int64))
        open(path, "wb").write(b"oracle-index")

    def load_index(self, path, max_elements=0):
        z = np.load(path + ".npz")
        self.rows = [r for r in z["rows"]]
        self.labels = z["labels"].tolist()
        self.dead = set(z["dead"].tolist())
        self.cap = max(max_elements, len(self.rows))

    def close(self):
        pass


def factory(space, dim, max_elements, store_dtype, device):
    return OracleIndex(space, dim, max_elements, store_dtype)

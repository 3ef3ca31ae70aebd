"""Data boundary of the path: a column store on disk -> the (N, n_cols) input matrix that
fit() / predict() consume, staged in pinned host memory so that fit()'s batch pipeline can stream it (basemodel._BatchFeeder).

Reference: utils.py:266-278 (`get_aliccp_ctr_df`: one HDF5 dataset per column under the groups `ctr_train` / `ctr_test`, written
by aliccp_dataset_processing.py:237-242), main.py:104-114 (the scenario column `301` is shifted to start from 1),
main.py:323-326 (the per-column dicts handed to fit / predict), main.py:381-396 (one CSV row per run).

HDF5 needs `h5py`, which is NOT in this image (and the reference's files cannot be generated here): `H5ColumnStore` is written
against h5py's API exactly as the reference uses it and is UNTESTED here -- parity unpinned.  The loader itself is exercised
through `NpyDirColumnStore`, which holds the same wire content (one integer / float array per column and group) as .npy files.
Nothing here touches the GPU: this is host-side staging, the compute path stays in the CUDA library.
"""
import os
from datetime import datetime

import numpy as np
import torch

ALICCP_SPARSE = ['101', '121', '122', '124', '125', '126', '127', '128', '129', '205', '206', '207', '210', '216', '508', '509',
                 '702', '853', '301']                                   # reference main.py:99-101 (field order of the model)
_TOPK_COLUMNS = ('10914', '11014', '15014', '12714')                    # history columns stored per top-k (utils.py:273-276)


class NpyDirColumnStore(object):
    """<root>/<group>/<column>.npy, memory-mapped: the same content as the reference's HDF5 file, one array per column."""

    def __init__(self, root):
        self.root = root

    def read(self, group, names, k=3):
        out = {}
        for key in names:
            path = os.path.join(self.root, group, _stored_name(key, k) + '.npy')
            if not os.path.exists(path):
                raise KeyError("column %r of group %r not found under %s" % (key, group, self.root))
            out[key] = np.load(path, mmap_mode='r')
        return out

    def write(self, group, columns):
        os.makedirs(os.path.join(self.root, group), exist_ok=True)
        for key, arr in columns.items():
            np.save(os.path.join(self.root, group, key + '.npy'), np.asarray(arr))


class H5ColumnStore(object):
    """The reference's own container (utils.py:266-278): `h5py.File(path)[group][column][:]`.  Needs h5py (absent from this
    image: untested here)."""

    def __init__(self, path):
        try:
            import h5py
        except ImportError as e:
            raise ImportError("H5ColumnStore needs h5py, which is not installed; convert the file once to a directory of .npy "
                              "columns (NpyDirColumnStore.write) or install h5py") from e
        self._h5py, self.path = h5py, path

    def read(self, group, names, k=3):
        with self._h5py.File(self.path, 'r') as f:
            g = f[group]
            return {key: g[_stored_name(key, k)][:] for key in names}


def _stored_name(key, k):
    return key + '_' + str(int(k)) if key in _TOPK_COLUMNS else key


def load_aliccp(store, group, sparse_features=None, label='click', domain_col='301', k=3):
    """(x dict name -> 1-D array, y) of one split, as reference main.py:104-114,323-326 builds them: the listed columns of
    `group` ('ctr_train' / 'ctr_test'), the scenario column shifted to start from 1 when the stored ids start from 0."""
    sparse_features = list(ALICCP_SPARSE if sparse_features is None else sparse_features)
    cols = store.read(group, [label] + sparse_features, k=k)
    x = {name: np.asarray(cols[name]) for name in sparse_features}
    if domain_col in x and x[domain_col].size and x[domain_col].min() == 0:          # main.py:108-110
        x[domain_col] = x[domain_col] + 1
    return x, np.asarray(cols[label])


def to_input_matrix(model, x, pin=True):
    """The (N, n_cols) matrix of `model` (columns in feature_index order, int32 when every column is integral, else the
    reference's float32) as a torch tensor in pinned host memory: fit(x=matrix, ...) streams it batch by batch."""
    X = torch.from_numpy(np.ascontiguousarray(model._stack_inputs(x)))
    if pin and torch.cuda.is_available():
        X = X.pin_memory()
    return X


def append_result_row(path, tag, test_aucs, test_loss, now=None):
    """One line of the reference's `<data>_results.csv` (main.py:381-396): '<MM-DD-HH-MM>-<tag>,<auc>,...,<loss>'."""
    dt = (now or datetime.now()).strftime('%m-%d-%H-%M')
    row = '%s-%s,' % (dt, tag) + ','.join('%.6f' % a if not isinstance(a, str) else a for a in test_aucs) + ',' + '%.6f' % test_loss
    with open(path, 'a') as f:
        f.write(row + '\n')
    return row

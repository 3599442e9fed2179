"""A numpy-backed stand-in for ``astropy.table.Table`` - just the operations that the
reference's ``SExImageCatalog.execute`` (catalogs.py:823-898) performs on a catalogue -
so that the UNMODIFIED reference can generate the ingest fixtures
(tests/golden/make_ingest_golden.py).  Test infrastructure only."""
from collections import OrderedDict

import numpy as np


class Table(object):
    masked = False

    def __init__(self, data=None, names=None, dtype=None, masked=False):
        self._cols = OrderedDict()
        if isinstance(data, Table):
            for k, v in data._cols.items():
                self._cols[k] = np.array(v)
        elif isinstance(data, dict):
            for k, v in data.items():
                self._cols[k] = np.array(v)
        elif names is not None:
            dtype = dtype or [np.float64] * len(names)
            for k, t in zip(names, dtype):
                self._cols[k] = np.zeros(0, dtype=t)

    @property
    def colnames(self):
        return list(self._cols)

    @property
    def itercols(self):
        return list(self._cols.values())

    def __len__(self):
        return len(next(iter(self._cols.values()))) if self._cols else 0

    def __getitem__(self, key):
        if isinstance(key, str):
            return self._cols[key]
        key = np.asarray(key)
        return Table({k: v[key] for k, v in self._cols.items()})

    def __setitem__(self, key, value):
        self._cols[key] = np.asarray(value)

    def copy(self):
        return Table(self)

    def rename_column(self, old, new):
        self._cols = OrderedDict((new if k == old else k, v) for k, v in self._cols.items())
